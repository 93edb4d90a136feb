"""Plain torch networks used by the examples and tools (no Prior modules: the prior is given to
the sampler's segment table directly).  ResNet20 has the size of the reference's `googleresnet`
(models/google_resnet.py: depth 20, 16/32/64 channels, BatchNorm; 272,282 parameters here)."""
import torch.nn as nn
import torch.nn.functional as F


class Block(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.c1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.b1 = nn.BatchNorm2d(cout)
        self.c2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.b2 = nn.BatchNorm2d(cout)
        self.sc = None if stride == 1 and cin == cout else nn.Conv2d(cin, cout, 1, stride, 0, bias=False)

    def forward(self, x):
        y = F.relu(self.b1(self.c1(x)))
        y = self.b2(self.c2(y))
        return F.relu(y + (x if self.sc is None else self.sc(x)))


class ResNet20(nn.Module):
    def __init__(self, classes=10):
        super().__init__()
        self.c0 = nn.Conv2d(3, 16, 3, 1, 1, bias=False)
        self.b0 = nn.BatchNorm2d(16)
        layers, cin = [], 16
        for cout, stride in ((16, 1), (32, 2), (64, 2)):
            for i in range(3):
                layers.append(Block(cin, cout, stride if i == 0 else 1))
                cin = cout
        self.layers = nn.Sequential(*layers)
        self.fc = nn.Linear(64, classes)

    def forward(self, x):
        x = self.layers(F.relu(self.b0(self.c0(x))))
        return self.fc(F.adaptive_avg_pool2d(x, 1).flatten(1))
