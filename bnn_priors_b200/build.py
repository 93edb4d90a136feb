"""Build libbnnp.so in-tree with nvcc for sm_100a.

    python -m bnn_priors_b200.build [--force] [--verbose]

The library is a plain C-ABI shared object (include/bnnp.h): no torch headers, no
pybind.  It is git-ignored but travels with the tree to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "csrc", "bnnp_kernels.cu"), os.path.join(HERE, "csrc", "bnnp_eval.cu")]
INCLUDE = os.path.join(ROOT, "include")
OUT = os.path.join(HERE, "_lib", "libbnnp.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def up_to_date() -> bool:
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    deps = SRC + [os.path.join(INCLUDE, "bnnp.h"), os.path.join(INCLUDE, "bnnp_eval.h"), os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libbnnp.so.  bnnp_kernels.cu is compiled as three translation units in parallel
    (-DBNNP_PART=0/1/2: the step-kernel instantiations of one noise kind each) next to bnnp_eval.cu,
    then linked: about a third of the time of a single nvcc call."""
    if not force and up_to_date():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    objdir = os.path.join(os.path.dirname(OUT), "_obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = find_nvcc()
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else []) + ["-I", INCLUDE]
    jobs = []
    for part in (0, 1, 2):
        obj = os.path.join(objdir, f"bnnp_kernels_part{part}.o")
        jobs.append((obj, [nvcc] + compile_flags + [f"-DBNNP_PART={part}", "-c", SRC[0], "-o", obj]))
    obj = os.path.join(objdir, "bnnp_eval.o")
    jobs.append((obj, [nvcc] + compile_flags + ["-c", SRC[1], "-o", obj]))
    procs = [(cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)) for _, cmd in jobs]
    failed = False
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + out)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libbnnp.so")
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT] + [o for o, _ in jobs]
    r = subprocess.run(link, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(" ".join(link) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed linking libbnnp.so")
    return OUT


HOST_SRC = os.path.join(HERE, "csrc", "bnnp_host.cpp")
HOST_DIR = os.path.join(HERE, "_lib", "_host")
HOST_OUT = os.path.join(HOST_DIR, "bnnp_host.so")


def build_host(force: bool = False, verbose: bool = False) -> str:
    """Compile the optional host-side helper (csrc/bnnp_host.cpp: the per-step scan of the parameters'
    gradients over the ATen objects instead of Python attribute calls) as a torch C++ extension, in-tree.
    CPU code only; the samplers run without it (Python scan)."""
    if not force and os.path.exists(HOST_OUT) and os.path.getmtime(HOST_OUT) >= os.path.getmtime(HOST_SRC):
        return HOST_OUT
    from torch.utils import cpp_extension
    os.makedirs(HOST_DIR, exist_ok=True)
    cpp_extension.load(name="bnnp_host", sources=[HOST_SRC], build_directory=HOST_DIR, extra_cflags=["-O2"],
                       verbose=verbose)
    if not os.path.exists(HOST_OUT):
        raise RuntimeError("building the host helper left no bnnp_host.so")
    return HOST_OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    try:
        print(build_host(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    except Exception as e:          # optional: the samplers fall back to the Python scan
        print("host helper not built:", e, file=sys.stderr)
