"""The C-ABI boundary on a machine without a GPU: the library loads, exports every
symbol include/bnnp.h declares, the Python mirror of the header's constants is in
step with it, the host-side planner works and argument validation fails cleanly
(negative code + message) before anything touches CUDA."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from bnn_priors_b200 import _native as N
from bnn_priors_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "bnnp.h")).read()
EVAL_HEADER = open(os.path.join(ROOT, "include", "bnnp_eval.h")).read()


@pytest.fixture(scope="module")
def lib():
    B.build()
    return N.lib()


def _declared_functions():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(bnnp_[a-z_0-9]+)\s*\(", body)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared_functions()
    assert set(names) == set(N.EXPORTS)
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.bnnp_abi_version() == N.ABI_VERSION


def test_eval_header_symbols_constants_and_validation(lib):
    body = re.sub(r"/\*.*?\*/", "", EVAL_HEADER, flags=re.S)
    names = sorted(set(re.findall(r"\b(bnnp_[a-z_0-9]+)\s*\(", body)))
    assert set(names) == set(N.EVAL_EXPORTS)
    for n in names:
        assert getattr(lib, n) is not None
    assert int(re.search(r"#define\s+BNNP_EVAL_ROW\s+(\d+)", body).group(1)) == N.EVAL_ROW
    for block in re.findall(r"enum\s*\{(.*?)\}", body, flags=re.S):
        nxt = 0
        for item in [x.strip() for x in block.split(",") if x.strip()]:
            m = re.match(r"BNNP_([A-Z_0-9]+)\s*(?:=\s*(\d+))?$", item)
            assert m, item
            if m.group(2) is not None:
                nxt = int(m.group(2))
            assert getattr(N, m.group(1)) == nxt, m.group(1)
            nxt += 1
    fields = re.search(r"typedef struct BnnpEvalState \{(.*?)\} BnnpEvalState;", EVAL_HEADER, flags=re.S).group(1)
    fields = re.sub(r"/\*.*?\*/", "", fields, flags=re.S)
    got = [re.sub(r"^[A-Za-z_0-9]+\s*\**", "", d.strip(), count=1).strip() for d in fields.split(";") if d.strip()]
    assert got == [f[0] for f in N.BnnpEvalState._fields_]
    # argument validation happens before any CUDA call
    st = N.BnnpEvalState()
    assert lib.bnnp_eval_batch(C.byref(st), None, 0, None, None, None, 0, 0, 1, 0, None) == -1
    assert b"bad state" in lib.bnnp_eval_last_error()
    for f in ("ens", "lps_lse", "lps_last", "acc_last", "rows"):
        setattr(st, f, 4096)
    st.N, st.C, st.kind = 10, 3, N.EVAL_CATEGORICAL
    assert lib.bnnp_eval_batch(C.byref(st), 4096, 3, None, 4096, None, 0, 8, 4, 0, None) == -1
    assert b"outside the test set" in lib.bnnp_eval_last_error()
    assert lib.bnnp_eval_batch(C.byref(st), 4096, 3, None, None, None, 0, 0, 4, 0, None) == -1
    assert b"need labels" in lib.bnnp_eval_last_error()
    assert lib.bnnp_eval_batch(C.byref(st), 4096, 3, None, 4096, None, 0, 0, 0, 0, None) == 0     # empty batch
    assert lib.bnnp_eval_finish(C.byref(st), 4096, None, 0, 4096, None, None) == -1


def test_python_constants_match_header():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    defines = dict(re.findall(r"#define\s+BNNP_([A-Z_]+)\s+(\d+)\s*$", body, flags=re.M))
    assert int(defines["ABI_VERSION"]) == N.ABI_VERSION
    assert int(defines["SEG_ALIGN"]) == N.SEG_ALIGN
    assert int(defines["THREADS"]) == N.THREADS and int(defines["UNROLL"]) == N.UNROLL
    assert int(defines["NRED"]) == N.NRED and int(defines["STATE_STRIDE"]) == N.STATE_STRIDE
    assert N.CHUNK == N.THREADS * N.UNROLL * 4
    # enums: explicit values and implicit counting
    for block in re.findall(r"enum\s*\{(.*?)\}", body, flags=re.S):
        nxt = 0
        for item in [x.strip() for x in block.split(",") if x.strip()]:
            m = re.match(r"BNNP_([A-Z_0-9]+)\s*(?:=\s*(.+))?$", item)
            assert m, item
            name, val = m.group(1), m.group(2)
            if val is not None:
                val = val.strip()
                sh = re.match(r"1u\s*<<\s*(\d+)", val)
                nxt = (1 << int(sh.group(1))) if sh else int(val)
            if name.startswith("E_"):
                nxt += 1
                continue
            assert getattr(N, name) == nxt, (name, getattr(N, name), nxt)
            nxt += 1


def test_struct_layouts_match_header():
    # BnnpSegment: 3 x 8 + 3 x 4 + 5 x 4 = 56 bytes, natural alignment
    assert N.SEGMENT_DTYPE.itemsize == 56
    assert [N.SEGMENT_DTYPE.fields[k][1] for k in ("off", "numel", "precond", "prior_loc", "prior_kind",
                                                   "first_chunk", "num_chunks", "link")] == [0, 8, 16, 24, 36, 40, 44, 48]
    seg = re.search(r"typedef struct BnnpSegment \{(.*?)\} BnnpSegment;", HEADER, flags=re.S).group(1)
    seg = re.sub(r"/\*.*?\*/", "", seg, flags=re.S)
    seg_names = []
    for decl in seg.split(";"):
        decl = re.sub(r"^[A-Za-z_0-9]+\s*", "", decl.strip(), count=1)
        seg_names += [x.strip() for x in decl.split(",") if x.strip()]
    assert seg_names == list(N.SEGMENT_DTYPE.names)
    epi = re.search(r"typedef struct BnnpEpilogue \{(.*?)\} BnnpEpilogue;", HEADER, flags=re.S).group(1)
    epi = re.sub(r"/\*.*?\*/", "", epi, flags=re.S)
    epi_names = []
    for decl in epi.split(";"):
        decl = re.sub(r"^[A-Za-z_0-9]+\s*", "", decl.strip(), count=1)
        epi_names += [x.strip() for x in decl.split(",") if x.strip()]
    assert epi_names == [f[0] for f in N.BnnpEpilogue._fields_]
    fields = re.search(r"typedef struct BnnpLaunch \{(.*?)\} BnnpLaunch;", HEADER, flags=re.S).group(1)
    fields = re.sub(r"/\*.*?\*/", "", fields, flags=re.S)
    names = []
    for decl in fields.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names += [re.search(r"([A-Za-z_0-9]+)\s*$", x).group(1) for x in decl.split(",")]     # the declarator's name
    assert names == [f[0] for f in N.BnnpLaunch._fields_]
    assert C.sizeof(N.BnnpLaunch) % 8 == 0
    # the block the host writes with one struct.pack_into (nchunks .. pending) is contiguous and unpadded
    assert N.DYN_OFFSET == N.BnnpLaunch.nchunks.offset and N.DYN_OFFSET + N.DYN_STRUCT.size == C.sizeof(N.BnnpLaunch)
    for name, cls in (("BnnpCoef", N.BnnpCoef), ("BnnpControl", N.BnnpControl)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), HEADER, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        got = []
        for decl in body.split(";"):
            if decl.strip():
                got += [re.search(r"([A-Za-z_0-9]+)\s*(\[[A-Z_]+\])?\s*$", x).group(1) for x in decl.split(",")]
        assert got == [f[0] for f in cls._fields_], name
    assert C.sizeof(N.BnnpCoef) == 64 and C.sizeof(N.BnnpControl) == 16 + 64 + 4 * 64
    assert int(re.search(r"#define BNNP_COEF_SLOTS (\d+)", HEADER).group(1)) == N.COEF_SLOTS


def test_plan_layout(lib):
    numel = [10, 4096, 4097, 1, 39200, 31]
    off, first, nch, total, chunks = N.plan_layout(numel)
    chunk_seg = chunks["seg"]
    assert list(nch) == [1, 1, 2, 1, 10, 1]
    assert list(first) == [0, 1, 2, 4, 5, 15]
    assert all(o % N.SEG_ALIGN == 0 for o in off)
    for i in range(len(numel) - 1):
        assert off[i + 1] - off[i] >= numel[i] and off[i + 1] - off[i] < numel[i] + N.SEG_ALIGN
    assert total % N.SEG_ALIGN == 0 and total >= off[-1] + numel[-1]
    assert list(chunk_seg) == [0, 1, 2, 2, 3] + [4] * 10 + [5]
    # BnnpChunk: first float and valid floats of every chunk
    assert list(chunks["rem"][:5]) == [10, 4096, 4096, 1, 1] and chunks["rem"][14] == 39200 - 9 * 4096 and chunks["rem"][15] == 31
    assert list(chunks["fbase"][:5]) == [off[0], off[1], off[2], off[2] + 4096, off[3]]
    assert N.CHUNK_DTYPE.itemsize == 16 and "typedef struct BnnpChunk" in HEADER
    # empty segments are rejected with a message, not a crash
    bad = np.array([4, 0], dtype=np.int64)
    o, f, n = np.zeros(2, np.int64), np.zeros(2, np.int32), np.zeros(2, np.int32)
    rc = lib.bnnp_plan_layout(bad.ctypes.data, 2, o.ctypes.data, f.ctypes.data, n.ctypes.data, None, None, None)
    assert rc == -1 and b"empty" in lib.bnnp_last_error()


def test_launch_validates_arguments_without_a_gpu(lib):
    assert lib.bnnp_launch(None, None) == -1
    a = N.BnnpLaunch()
    assert lib.bnnp_launch(C.byref(a), None) == -1 and b"empty chain" in lib.bnnp_last_error()
    a.nseg, a.nchunks, a.nchunks_total = 1, 1, 1
    assert lib.bnnp_launch(C.byref(a), None) == -1 and b"null table" in lib.bnnp_last_error()
    for f in ("segs", "chunks", "seg_state", "partials", "stamps"):   # chunk_ids may stay null
        setattr(a, f, 4096)
    a.nchunks = 2
    assert lib.bnnp_launch(C.byref(a), None) == -1 and b"does not match the plan" in lib.bnnp_last_error()
    a.nchunks = 1
    a.pending.valid, a.pending.parity, a.parity = 1, 0, 0
    assert lib.bnnp_launch(C.byref(a), None) == -1 and b"overwrite the partial records" in lib.bnnp_last_error()
    a.pending.valid, a.pending.parity, a.pending.flags = 1, 1, N.F_HYPER
    assert lib.bnnp_launch(C.byref(a), None) == -1 and b"bnnp_finalize first" in lib.bnnp_last_error()
    a.pending.valid, a.pending.flags = 0, 0
    assert lib.bnnp_finalize(C.byref(a), None) == 0             # nothing pending: no-op, no CUDA call
    a.P = 4096
    a.op, a.flags = N.OP_SGLD, N.F_HYPER | N.F_READ_P | N.F_LOG_PRIOR
    assert lib.bnnp_launch(C.byref(a), None) == -1 and b"read-only pre-pass" in lib.bnnp_last_error()
    a.op, a.flags = N.OP_REDUCE, N.F_HYPER | N.F_READ_P | N.F_LOG_PRIOR | N.F_WRITE_P
    assert lib.bnnp_launch(C.byref(a), None) == -1 and b"read-only pre-pass" in lib.bnnp_last_error()
    a.op, a.flags, a.P = 0, 0, None
    a.op = 9
    assert lib.bnnp_launch(C.byref(a), None) == -1 and b"bad op" in lib.bnnp_last_error()
    a.op, a.flags = N.OP_SGLD, N.F_READ_P
    assert lib.bnnp_launch(C.byref(a), None) == -1 and b"P is null" in lib.bnnp_last_error()
    a.P, a.flags = 4096 + 4, N.F_READ_P
    assert lib.bnnp_launch(C.byref(a), None) == -2 and b"16-byte" in lib.bnnp_last_error()
    a.P, a.flags = 4096, N.F_READ_P | N.F_SAVE_STATE
    assert lib.bnnp_launch(C.byref(a), None) == -1 and b"SAVE_STATE" in lib.bnnp_last_error()
    a.flags, a.noise = N.F_READ_P, N.NOISE_REPLAY
    assert lib.bnnp_launch(C.byref(a), None) == -1 and b"replay" in lib.bnnp_last_error()
    assert lib.bnnp_rollback(None, None, None, None, None, None, 0, None) == -1
    # capturable mode: control block entry points
    b = N.BnnpLaunch()
    assert lib.bnnp_advance(C.byref(b), None) == -1 and b"no control block" in lib.bnnp_last_error()
    b.ctl, b.coef_slot = 4096, N.COEF_SLOTS
    assert lib.bnnp_advance(C.byref(b), None) == -1 and b"coef_slot" in lib.bnnp_last_error()
    assert lib.bnnp_clear_pending(None, None) == -1
    buf = (C.c_char * 8)()
    assert lib.bnnp_poke(None, buf, 8, None) == -1
    assert lib.bnnp_poke(4096, buf, 6, None) == -1 and b"multiple of 4" in lib.bnnp_last_error()
    assert lib.bnnp_poke(4096, buf, 4000, None) == -1
    assert lib.bnnp_poke(4098, buf, 8, None) == -1


def test_philox_key_matches_the_oracle_specification():
    from oracle import sgmcmc_oracle as O
    for seed, stream in [(0, 0), (1234, 7), (2**63 + 5, 65536 * 3 + 1)]:
        assert N.philox_key(seed, stream) == O.philox_key(seed, stream)


def test_samplers_refuse_cpu_tensors():
    import torch
    from bnn_priors_b200 import mcmc
    for cls, kw in ((mcmc.SGLD, dict(lr=.1, num_data=1)), (mcmc.VerletSGLD, dict(lr=.1, num_data=1)),
                    (mcmc.HMC, dict(lr=.1, num_data=1))):
        with pytest.raises(RuntimeError, match="CUDA tensors only"):
            cls([torch.nn.Parameter(torch.zeros(3))], **kw)
