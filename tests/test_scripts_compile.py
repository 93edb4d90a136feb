"""The GPU-only scripts (tools/, examples/) at least parse and byte-compile on the CPU box."""
import glob
import os
import py_compile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tools_and_examples_compile(tmp_path):
    files = sorted(glob.glob(os.path.join(ROOT, "tools", "*.py")) + glob.glob(os.path.join(ROOT, "examples", "*.py")))
    assert len(files) >= 10
    for i, f in enumerate(files):
        py_compile.compile(f, cfile=str(tmp_path / f"{i}.pyc"), doraise=True)
