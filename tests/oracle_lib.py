"""ctypes wrapper over oracle/liboracle.so — TEST INFRASTRUCTURE (the CPU restatement of the
reference).  Only tests/, smoke() and bench.py's cpu_baseline / --impl reference legs use it."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
SO = os.path.join(ORACLE_DIR, "liboracle.so")
SO_NATIVE = os.path.join(ORACLE_DIR, "liboracle_native.so")     # same sources, -O3 -march=x86-64-v3 (timing legs)
SO_CPU_CALLER = os.path.join(ORACLE_DIR, "libfgb_cpu_caller.so")  # the product's host code over the oracle's vote

_lib = None
_native = None


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


def load():
    global _lib
    if _lib is not None:
        return _lib
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR)
            if f.endswith((".cpp", ".hpp"))]
    if not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        build()
    lib = C.CDLL(SO)
    d, u8, vp = C.c_double, C.c_uint8, C.c_void_p
    for name, args, res in [
        ("orc_phred_to_ln_error_prob", [u8], d), ("orc_phred_to_ln_correct_prob", [u8], d),
        ("orc_ln_prob_to_phred", [d], u8), ("orc_log1pexp", [d], d),
        ("orc_ln_one_minus_exp", [d], d), ("orc_ln_a_minus_b", [d, d], d),
        ("orc_ln_error_prob_two_trials", [d, d], d), ("orc_ln_sum_exp", [d, d], d),
        ("orc_ln_sum_exp_array", [vp, C.c_size_t], d),
    ]:
        getattr(lib, name).argtypes = args
        getattr(lib, name).restype = res
    lib.orc_builder_call.argtypes = [u8, u8, vp, vp, C.c_size_t, vp, vp, vp, vp]
    lib.orc_builder_call.restype = None
    lib.orc_tables.argtypes = [u8, u8, vp, vp, vp, vp]
    lib.orc_tables.restype = None
    lib.orc_simplex_batch.argtypes = [C.c_uint64, vp, vp, vp, vp, u8, u8, C.c_uint32, u8, vp, vp,
                                      vp, vp, vp, C.c_int]
    lib.orc_simplex_batch.restype = C.c_int
    lib.orc_duplex_combine.argtypes = [vp] * 8 + [C.c_size_t, vp, vp, C.c_long, vp, vp, vp]
    lib.orc_duplex_combine.restype = None
    lib.orc_codec_combine.argtypes = [vp] * 8 + [C.c_size_t] + [vp] * 6
    lib.orc_codec_combine.restype = None
    lib.orc_codec_mask.argtypes = [vp, vp, C.c_size_t, vp, vp, C.c_int, C.c_int, C.c_size_t]
    lib.orc_codec_mask.restype = None
    lib.orc_duplex_job.argtypes = ([vp] * 4 + [C.c_size_t] + [vp] * 4 + [C.c_size_t, vp, vp, C.c_long] +
                                   [vp] * 4)
    lib.orc_duplex_job.restype = C.c_int
    lib.orc_codec_job.argtypes = ([vp] * 4 + [C.c_size_t] + [vp] * 4 + [C.c_size_t, C.c_int, C.c_int,
                                  C.c_size_t, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_double] +
                                  [vp] * 6)
    lib.orc_codec_job.restype = C.c_int
    _lib = lib
    return lib


def builder_call(pre, post, bases: bytes, quals):
    lib = load()
    b = np.frombuffer(bytes(bases), dtype=np.uint8).copy()
    q = np.asarray(quals, dtype=np.uint8).copy()
    assert len(b) == len(q)
    ob, oq = C.c_uint8(), C.c_uint8()
    obs = np.zeros(4, np.uint16)
    ll = np.zeros(4, np.float64)
    lib.orc_builder_call(pre, post, b.ctypes.data, q.ctypes.data, len(b), C.addressof(ob),
                         C.addressof(oq), obs.ctypes.data, ll.ctypes.data)
    return chr(ob.value), oq.value, obs, ll


def tables(pre, post):
    lib = load()
    c = np.zeros(94); e = np.zeros(94); lp = C.c_double(); sq = np.zeros(94, np.uint8)
    lib.orc_tables(pre, post, c.ctypes.data, e.ctypes.data, C.addressof(lp), sq.ctypes.data)
    return c, e, lp.value, sq


def alloc_outputs(batch):
    n = max(batch.n_out, 1)
    return (np.zeros(n, np.uint8), np.zeros(n, np.uint8), np.zeros(n, np.uint16),
            np.zeros(n, np.uint16), np.zeros(max(batch.n_units, 1), np.uint32))


def load_native():
    """The oracle compiled for speed (oracle/Makefile): only orc_simplex_batch is declared."""
    global _native
    if _native is None:
        if not os.path.exists(SO_NATIVE):
            build()
        lib = C.CDLL(SO_NATIVE)
        vp, u8 = C.c_void_p, C.c_uint8
        lib.orc_simplex_batch.argtypes = [C.c_uint64, vp, vp, vp, vp, u8, u8, C.c_uint32, u8, vp, vp, vp, vp, vp, C.c_int]
        lib.orc_simplex_batch.restype = C.c_int
        _native = lib
    return _native


def simplex_batch(batch, pre=45, post=40, min_reads=1, min_cons_q=2, threads=1, outputs=None, native=False):
    """Run the oracle over a PackedBatch; returns (base, qual, depth, errors, cons_len).
    Pass `outputs=alloc_outputs(batch)` to reuse buffers (timing runs)."""
    lib = load_native() if native else load()
    ob, oq, od, oe, cl = outputs if outputs is not None else alloc_outputs(batch)
    rc = lib.orc_simplex_batch(batch.n_units, batch.units.ctypes.data, batch.reads.ctypes.data,
                               batch.bases.ctypes.data, batch.quals.ctypes.data, pre, post,
                               min_reads, min_cons_q, ob.ctypes.data, oq.ctypes.data,
                               od.ctypes.data, oe.ctypes.data, cl.ctypes.data, threads)
    assert rc == 0
    return ob, oq, od, oe, cl[: batch.n_units]
