"""ctypes binding of the C ABI in include/bnnp.h (libbnnp.so, built in-tree by
`bnn_priors_b200.build`).  Nothing here computes: it only describes the structs
and forwards calls.  If the library is missing the import of the sampler classes
fails loudly -- there is no other implementation of the path in this package.
"""
from __future__ import annotations

import ctypes as C
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# BNNP_LIB: load another build of the same ABI (tuning experiments with other tile sizes)
LIB_PATH = os.environ.get("BNNP_LIB") or os.path.join(HERE, "_lib", "libbnnp.so")

# ---- constants of include/bnnp.h (tests/test_abi.py checks them against the header)
ABI_VERSION = 10
SEG_ALIGN = 32
THREADS = 256
UNROLL = 4
CHUNK = THREADS * UNROLL * 4
NRED = 8
STATE_STRIDE = 16

PRIOR_NONE, PRIOR_NORMAL, PRIOR_LAPLACE, PRIOR_STUDENT_T = 0, 1, 2, 3
PRIOR_CAUCHY, PRIOR_GENNORM, PRIOR_LOGNORMAL, PRIOR_UNIFORM, PRIOR_IMPROPER, PRIOR_DOUBLE_GAMMA = 4, 5, 6, 7, 8, 9
PRIOR_HYPER_GAMMA, PRIOR_HYPER_UNIFORM, PRIOR_HYPER_HALFCAUCHY, PRIOR_HYPER_IMPROPER = 10, 11, 12, 13
OP_SGLD, OP_VERLET, OP_HMC, OP_SAMPLE_MOMENTUM, OP_REDUCE = 0, 1, 2, 3, 4
PHASE_INITIAL, PHASE_MID, PHASE_FINAL = 0, 1, 2
NOISE_NONE, NOISE_REPLAY, NOISE_PHILOX = 0, 1, 2

F_READ_P = 1 << 0
F_READ_G = 1 << 1
F_READ_M = 1 << 2
F_WRITE_P = 1 << 3
F_WRITE_M = 1 << 4
F_SAVE_STATE = 1 << 5
F_CALC_METRICS = 1 << 6
F_LOG_PRIOR = 1 << 7
F_CLAMP_GRAD = 1 << 8
F_NOISE_FIRST = 1 << 9
F_MM_PRE_NOISE = 1 << 10
F_UPDATE_SQ = 1 << 11
F_PRIOR_GRAD = 1 << 12
F_ALL_SUMS = 1 << 13
F_HYPER = 1 << 14
F_REVERSE = 1 << 15
F_HYPER_POST = 1 << 16
F_HYPER_CHAIN = 1 << 17

(S_DELTA_ENERGY, S_PREV_NEW_MOM, S_EST_MM, S_EST_PG, S_SUM_GG, S_SUM_MM, S_SQ_MEAN,
 S_LOG_PRIOR, S_GM_OLD, S_GM_NEW, S_MM_OLD, S_MM_NEW, S_NONFINITE, S_LAUNCHES, S_HYPER) = range(15)

# BnnpSegment as a numpy record (the table is built on the host and copied to HBM)
SEGMENT_DTYPE = np.dtype([
    ("off", np.int64), ("numel", np.int64), ("precond", np.float64),
    ("prior_loc", np.float32), ("prior_scale", np.float32), ("prior_df", np.float32),
    ("prior_kind", np.int32), ("first_chunk", np.int32), ("num_chunks", np.int32),
    ("link", np.int32), ("reserved", np.int32)], align=True)
assert SEGMENT_DTYPE.itemsize == 56


# BnnpChunk
CHUNK_DTYPE = np.dtype([("fbase", np.int64), ("rem", np.int32), ("seg", np.int32)], align=True)
assert CHUNK_DTYPE.itemsize == 16


class BnnpEpilogue(C.Structure):
    _fields_ = [
        ("valid", C.c_int32), ("op", C.c_int32), ("phase", C.c_int32), ("flags", C.c_uint32),
        ("parity", C.c_int32), ("reserved", C.c_int32), ("call", C.c_uint64),
        ("c_gm_base", C.c_double), ("curv_base", C.c_double), ("rms_alpha", C.c_double),
        ("inv_num_data", C.c_double),
    ]


class BnnpCoef(C.Structure):
    _fields_ = [("cm", C.c_double), ("cg", C.c_double), ("cn", C.c_double), ("cp", C.c_double),
                ("inv_num_data", C.c_double), ("c_gm_base", C.c_double), ("curv_base", C.c_double),
                ("rms_alpha", C.c_double)]


COEF_SLOTS = 4
SLOT_OTHER = 3          # sample_momentum / reduce / pre-pass launches (slots 0..2: the sampler's phases)


class BnnpControl(C.Structure):
    _fields_ = [("call", C.c_uint64), ("parity", C.c_int32), ("reserved", C.c_int32),
                ("pending", BnnpEpilogue), ("coef", BnnpCoef * COEF_SLOTS)]


class BnnpLaunch(C.Structure):
    _fields_ = [
        ("P", C.c_void_p), ("G", C.c_void_p), ("M", C.c_void_p),
        ("prev_p", C.c_void_p), ("prev_g", C.c_void_p), ("prev_m", C.c_void_p),
        ("replay_noise", C.c_void_p),
        ("segs", C.c_void_p), ("chunks", C.c_void_p), ("chunk_ids", C.c_void_p),
        ("seg_grad", C.c_void_p), ("ctl", C.c_void_p), ("seg_state", C.c_void_p),
        ("partials", C.c_void_p), ("stamps", C.c_void_p),
        ("nseg", C.c_int32), ("nchunks", C.c_int32), ("nchunks_total", C.c_int32), ("parity", C.c_int32),
        ("op", C.c_int32), ("phase", C.c_int32), ("noise", C.c_int32),
        ("flags", C.c_uint32), ("coef_slot", C.c_int32), ("reserved", C.c_int32),
        ("key0", C.c_uint32), ("key1", C.c_uint32),
        ("call", C.c_uint64),
        ("cm", C.c_double), ("cg", C.c_double), ("cn", C.c_double), ("cp", C.c_double),
        ("inv_num_data", C.c_double), ("grad_max", C.c_double),
        ("c_gm_base", C.c_double), ("curv_base", C.c_double), ("rms_alpha", C.c_double),
        ("pending", BnnpEpilogue),
    ]


# The part of BnnpLaunch that changes from launch to launch is contiguous (nchunks .. pending): the
# host writes it with ONE struct.pack_into instead of ~35 ctypes field stores.
#   nchunks nchunks_total parity | op phase noise flags | coef_slot reserved | key0 key1 | call |
#   cm cg cn cp inv_num_data grad_max c_gm_base curv_base rms_alpha |
#   pending: valid op phase flags parity reserved call c_gm_base curv_base rms_alpha inv_num_data
DYN_OFFSET = BnnpLaunch.nchunks.offset
DYN_STRUCT = struct.Struct("<3i3iI2i2IQ9d" + "3iI2iQ4d")
assert DYN_OFFSET + DYN_STRUCT.size == C.sizeof(BnnpLaunch), "BnnpLaunch layout changed: fix DYN_STRUCT"
assert BnnpLaunch.pending.offset == DYN_OFFSET + struct.calcsize("<3i3iI2i2IQ9d")
PENDING_STRUCT = struct.Struct("<3iI2iQ4d")
assert PENDING_STRUCT.size == C.sizeof(BnnpEpilogue)
COEF_STRUCT = struct.Struct("<8d")
assert COEF_STRUCT.size == C.sizeof(BnnpCoef)


# ---- include/bnnp_eval.h
EVAL_CATEGORICAL, EVAL_NORMAL = 0, 1
EV_LP_ENSEMBLE, EV_LP_LAST, EV_ACC_ENSEMBLE, EV_ACC_LAST, EV_LP_ENSEMBLE_CHECK = range(5)
EV_OUT = 8
EVAL_ROW = 5


class BnnpEvalState(C.Structure):
    _fields_ = [
        ("ens", C.c_void_p), ("lps_lse", C.c_void_p), ("lps_last", C.c_void_p), ("acc_last", C.c_void_p),
        ("rows", C.c_void_p), ("N", C.c_int64), ("C", C.c_int32), ("kind", C.c_int32),
    ]


EVAL_EXPORTS = ("bnnp_eval_batch", "bnnp_eval_finish", "bnnp_eval_last_error")

EXPORTS = ("bnnp_abi_version", "bnnp_last_error", "bnnp_device_info", "bnnp_max_ctas_per_sm",
           "bnnp_plan_layout", "bnnp_launch", "bnnp_finalize", "bnnp_advance", "bnnp_clear_pending", "bnnp_poke",
           "bnnp_rollback", "bnnp_probe_stream")


class BnnpError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    """Load libbnnp.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BnnpError(
            f"{LIB_PATH} is missing: build it with `python -m bnn_priors_b200.build` "
            "(nvcc, sm_100a).  bnn_priors_b200 has no fallback implementation.")
    l = C.CDLL(LIB_PATH)
    l.bnnp_abi_version.restype = C.c_int
    l.bnnp_last_error.restype = C.c_char_p
    l.bnnp_device_info.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    l.bnnp_max_ctas_per_sm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    l.bnnp_plan_layout.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.c_void_p]
    l.bnnp_launch.argtypes = [C.POINTER(BnnpLaunch), C.c_void_p]
    l.bnnp_finalize.argtypes = [C.POINTER(BnnpLaunch), C.c_void_p]
    l.bnnp_advance.argtypes = [C.POINTER(BnnpLaunch), C.c_void_p]
    l.bnnp_clear_pending.argtypes = [C.c_void_p, C.c_void_p]
    l.bnnp_poke.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    l.bnnp_rollback.argtypes = [C.c_void_p] * 6 + [C.c_int64, C.c_void_p]
    l.bnnp_probe_stream.argtypes = [C.c_void_p] * 3 + [C.c_int64, C.c_void_p]
    l.bnnp_eval_batch.argtypes = [C.POINTER(BnnpEvalState), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]
    l.bnnp_eval_finish.argtypes = [C.POINTER(BnnpEvalState), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                   C.c_void_p, C.c_void_p]
    l.bnnp_eval_last_error.restype = C.c_char_p
    for name in EXPORTS + EVAL_EXPORTS:
        getattr(l, name)          # AttributeError if a symbol is missing
    if l.bnnp_abi_version() != ABI_VERSION:
        raise BnnpError(f"libbnnp.so has ABI {l.bnnp_abi_version()}, this package expects {ABI_VERSION}; rebuild")
    _lib = l
    return l


_host = False


def host_module():
    """The optional host-side helper (csrc/bnnp_host.cpp, built by `bnn_priors_b200.build.build_host`) or None.
    BNNP_HOST_SCAN=0 disables it (the Python scan in mcmc/_flat.py is the specification and the fallback)."""
    global _host
    if _host is not False:
        return _host
    _host = None
    path = os.path.join(HERE, "_lib", "_host", "bnnp_host.so")
    if os.environ.get("BNNP_HOST_SCAN", "1") != "0" and os.path.exists(path):
        try:
            import importlib.machinery
            import importlib.util
            import torch  # noqa: F401  (the extension links against libtorch)
            loader = importlib.machinery.ExtensionFileLoader("bnnp_host", path)
            spec = importlib.util.spec_from_loader("bnnp_host", loader)
            mod = importlib.util.module_from_spec(spec)
            loader.exec_module(mod)
            _host = mod
        except Exception:           # another torch / python than it was built for: Python scan
            _host = None
    return _host


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise BnnpError(f"{what} failed ({rc}): {lib().bnnp_last_error().decode()}")


def check_eval(rc: int, what: str) -> None:
    if rc != 0:
        raise BnnpError(f"{what} failed ({rc}): {lib().bnnp_eval_last_error().decode()}")


def plan_layout(numels):
    """Host-side layout planner (bnnp_plan_layout).  Returns
    (off[nseg], first_chunk[nseg], num_chunks[nseg], total_elems, chunks[nchunks] as CHUNK_DTYPE records)."""
    numel = np.ascontiguousarray(numels, dtype=np.int64)
    n = int(numel.size)
    off = np.zeros(n, np.int64)
    first = np.zeros(n, np.int32)
    nch = np.zeros(n, np.int32)
    total, chunks = C.c_int64(0), C.c_int32(0)
    l = lib()
    check(l.bnnp_plan_layout(numel.ctypes.data, n, off.ctypes.data, first.ctypes.data, nch.ctypes.data,
                             C.byref(total), C.byref(chunks), None), "bnnp_plan_layout")
    table = np.zeros(max(chunks.value, 1), CHUNK_DTYPE)
    check(l.bnnp_plan_layout(numel.ctypes.data, n, off.ctypes.data, first.ctypes.data, nch.ctypes.data,
                             C.byref(total), C.byref(chunks), table.ctypes.data), "bnnp_plan_layout")
    return off, first, nch, int(total.value), table[:chunks.value]


# ---- Philox key schedule (specified in oracle/sgmcmc_oracle.py:philox_key) -------------
def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def philox_key(seed: int, stream: int):
    k = _splitmix64((seed + stream * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
    return k & 0xFFFFFFFF, k >> 32
