#!/usr/bin/env python
"""Recipe for `oracle/_ref/`: a snapshot of the UNMODIFIED reference package for the GPU box.

TEST INFRASTRUCTURE ONLY.  `/root/reference` exists in the build container but not on
the GPU box, and the parity definition of this path is "the reference's own runners
(bnn_priors/inference.py, inference_reject.py, experiments/train_bnn.py) drive the
B200 sampler and land on the reference sampler's trajectory".  So the build step
(`__graft_entry__.build()`, which runs where the reference is present) copies the
reference's Python sources -- byte for byte, nothing edited -- into the git-ignored
directory `oracle/_ref/`, which travels to the GPU box with the tree exactly like the
built `libbnnp.so`.  Nothing under `oracle/_ref/` is ever committed, and nothing in
`bnn_priors_b200/` imports it: only `tests/` and `bench.py --impl reference` do.

`pip install --target` is not used: the reference's setup.py lists
`packages=["bnn_priors"]` only, so an install would lack `bnn_priors.mcmc`,
`.prior`, `.models`, `.data`.

    python oracle/make_ref.py [--force]

What is copied: `bnn_priors/**/*.py` (the data files under bnn_priors/data -- 14 MB of
UCI tables -- are not: the tests use synthetic data), `experiments/train_bnn.py`,
`experiments/eval_bnn.py`, and the reference's own sampler tests `testing/test_sgld.py`,
`test_verlet_sgld.py`, `test_hmc.py`, `utils.py` (run on the GPU against the overlay).
A manifest with the sha256 of every file is written next to them.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("BNNP_REFERENCE", "/root/reference")
DEST = os.path.join(HERE, "_ref")

EXTRA_FILES = (
    "experiments/train_bnn.py", "experiments/eval_bnn.py",
    "testing/__init__.py", "testing/utils.py", "testing/test_sgld.py", "testing/test_verlet_sgld.py",
    "testing/test_hmc.py",
)


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def _sources():
    pkg = os.path.join(REFERENCE, "bnn_priors")
    for d, _, files in os.walk(pkg):
        for f in sorted(files):
            if f.endswith(".py"):
                yield os.path.relpath(os.path.join(d, f), REFERENCE)
    for rel in EXTRA_FILES:
        if os.path.exists(os.path.join(REFERENCE, rel)):
            yield rel


def available() -> bool:
    "True if a snapshot is present (on the GPU box: the one the build container made)"
    return os.path.exists(os.path.join(DEST, "MANIFEST.json")) and \
        os.path.exists(os.path.join(DEST, "bnn_priors", "mcmc", "sgld.py"))


def make(force: bool = False) -> str:
    """Create / refresh oracle/_ref from REFERENCE.  No-op (returns the existing path) where the
    reference checkout is absent, e.g. on the GPU box."""
    if not os.path.isdir(os.path.join(REFERENCE, "bnn_priors")):
        return DEST if available() else ""
    manifest = {}
    rels = list(_sources())
    if not force and available():
        try:
            with open(os.path.join(DEST, "MANIFEST.json")) as f:
                old = json.load(f)["files"]
            if set(old) == set(rels) and all(_sha(os.path.join(REFERENCE, r)) == old[r] and
                                             os.path.exists(os.path.join(DEST, r)) for r in rels):
                return DEST
        except Exception:
            pass
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    for rel in rels:
        src, dst = os.path.join(REFERENCE, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REFERENCE, "files": manifest,
                   "note": "unmodified copies; test infrastructure; git-ignored"}, f, indent=1, sort_keys=True)
    return DEST


if __name__ == "__main__":
    out = make(force="--force" in sys.argv)
    print(out or "no reference checkout and no snapshot")
