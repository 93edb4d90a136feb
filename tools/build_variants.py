#!/usr/bin/env python
"""Builds of libbnnp.so with other compile-time choices (-D macros of csrc/bnnp_kernels.cu), for
tools/variant_times.py / tools/tune_tiles.py to compare on the GPU box.  No GPU needed.

    python tools/build_variants.py sub1:-DBNNP_TMA_SUBTILES=1 sub2:-DBNNP_TMA_SUBTILES=2 sub4:-DBNNP_TMA_SUBTILES=4
    -> bnn_priors_b200/_lib/tune/{sub1,sub2,sub4}.so     (git-ignored; they travel with gpurun)
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bnn_priors_b200 import build as B  # noqa: E402

TUNE = os.path.join(os.path.dirname(B.OUT), "tune")


def one(spec: str) -> str:
    name, _, rest = spec.partition(":")
    defs, _, src = rest.partition(":")        # an optional third field: another copy of bnnp_kernels.cu (e.g. an older revision)
    defs = [d for d in defs.split(",") if d]
    src = src or B.SRC[0]
    objdir = os.path.join(TUNE, "_obj_" + name)
    os.makedirs(objdir, exist_ok=True)
    nvcc = B.find_nvcc()
    flags = [f for f in B.NVCC_FLAGS if f != "-shared"] + ["-I", B.INCLUDE] + defs
    cmds, objs = [], []
    for part in (0, 1, 2):
        objs.append(os.path.join(objdir, f"k{part}.o"))
        cmds.append([nvcc] + flags + [f"-DBNNP_PART={part}", "-c", src, "-o", objs[-1]])
    objs.append(os.path.join(objdir, "eval.o"))
    cmds.append([nvcc] + flags + ["-c", B.SRC[1], "-o", objs[-1]])
    procs = [subprocess.Popen(c, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for c in cmds]
    for c, p in zip(cmds, procs):
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(" ".join(c) + "\n" + out)
    so = os.path.join(TUNE, name + ".so")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", so] + objs, check=True)
    for o in objs:
        os.remove(o)
    os.rmdir(objdir)
    return so


if __name__ == "__main__":
    os.makedirs(TUNE, exist_ok=True)
    with ThreadPoolExecutor(max_workers=2) as ex:
        for so in ex.map(one, sys.argv[1:]):
            print(so)
