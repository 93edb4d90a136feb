"""The C ABI used from plain C (no Python, no torch): tests/c/abi_consumer.c is compiled with
gcc against include/bnnp.h, linked to libbnnp.so and the CUDA runtime, and run on the GPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "abi_consumer.c")
LIBDIR = os.path.join(ROOT, "bnn_priors_b200", "_lib")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _compile(out):
    cc = shutil.which("gcc") or shutil.which("cc")
    assert cc, "no C compiler"
    cmd = [cc, "-std=c99", "-Wall", "-Werror", SRC, "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           "-L", LIBDIR, "-lbnnp", "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-lm",
           f"-Wl,-rpath,{LIBDIR}", f"-Wl,-rpath,{os.path.join(CUDA, 'lib64')}", "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_header_is_plain_c_and_the_consumer_links(tmp_path):
    "CPU: include/bnnp.h compiles as C99 with -Wall -Werror and every symbol the consumer uses resolves"
    from bnn_priors_b200 import build
    build.build()
    _compile(str(tmp_path / "abi_consumer"))


@pytest.mark.gpu
def test_plain_c_consumer_runs_the_known_answer(tmp_path):
    exe = _compile(str(tmp_path / "abi_consumer"))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi_consumer ok" in r.stdout
