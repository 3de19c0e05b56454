import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FGB_CALLER_TRACE"] = "1"; os.environ["FGB_SUBMIT_TRACE"] = "1"
import numpy as np
import fgumi_b200 as fg
from tests.test_caller_parity import random_duplex_groups, random_codec_groups
G = 20000
for name, gen, mk in (("duplex", random_duplex_groups, lambda: fg.DuplexConsensusCaller("fgumi", "A", min_reads=(1, 1, 0), n_threads=16)),
                      ("codec", random_codec_groups, lambda: fg.CodecConsensusCaller("fgumi", "A", n_threads=16))):
    base = gen(np.random.default_rng(5), 2000, L=150)
    groups = [base[i % len(base)] for i in range(G)]
    recs = [r for g in groups for r in g]
    blob = np.frombuffer(b"".join(recs), np.uint8)
    off = np.zeros(len(recs) + 1, dtype=np.uint64); off[1:] = np.cumsum([len(r) for r in recs])
    grp = np.zeros(len(groups) + 1, dtype=np.uint64); grp[1:] = np.cumsum([len(g) for g in groups])
    c = mk()
    for rep in range(3):
        sys.stderr.write(f"--- {name} rep {rep}\n"); sys.stderr.flush()
        t0 = time.perf_counter()
        assert c._lib.fgb_caller_add_groups(c._h, blob.ctypes.data, off.ctypes.data, grp.ctypes.data, len(groups)) == 0
        t1 = time.perf_counter()
        import ctypes as C
        data, n, cnt = C.c_void_p(), C.c_uint64(), C.c_uint64()          # the C call itself: no copy into a Python bytes object
        assert c._lib.fgb_caller_flush(c._h, C.byref(data), C.byref(n), C.byref(cnt)) == 0
        t2 = time.perf_counter()
        sys.stderr.write(f"{name} rep {rep}: add {1e3 * (t1 - t0):.1f} ms flush {1e3 * (t2 - t1):.1f} ms, {cnt.value} records, {n.value} bytes\n")
    c.close()
