"""Record-level duplex and CODEC callers end to end (host decode + packing, GPU vote + combine, host record assembly):
input reads per second by host thread count.  usage: python scripts/bench_modes_records.py [groups]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fgumi_b200 as fg
from tests.test_caller_parity import random_duplex_groups, random_codec_groups

G = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
for name, gen, mk in (("duplex", random_duplex_groups, lambda T: fg.DuplexConsensusCaller("fgumi", "A", min_reads=(1, 1, 0), n_threads=T)),
                      ("codec", random_codec_groups, lambda T: fg.CodecConsensusCaller("fgumi", "A", n_threads=T))):
    t0 = time.perf_counter()
    base = gen(np.random.default_rng(5), 2000, L=150)
    groups = [base[i % len(base)] for i in range(G)]           # the same groups over again: throughput, not variety
    recs = [r for g in groups for r in g]
    blob = np.frombuffer(b"".join(recs), np.uint8)
    off = np.zeros(len(recs) + 1, dtype=np.uint64); off[1:] = np.cumsum([len(r) for r in recs])
    grp = np.zeros(len(groups) + 1, dtype=np.uint64); grp[1:] = np.cumsum([len(g) for g in groups])
    print(f"{name}: {len(groups)} groups, {len(recs)} reads ({len(recs) / len(groups):.1f} per group), built in {time.perf_counter() - t0:.1f} s", flush=True)
    for T in (1, 4, 16):
        try:
            c = mk(T)
        except TypeError:
            print(f"{name}: caller takes no n_threads"); break
        best = None
        for rep in range(3):
            t0 = time.perf_counter()
            assert c._lib.fgb_caller_add_groups(c._h, blob.ctypes.data, off.ctypes.data, grp.ctypes.data, len(groups)) == 0
            t1 = time.perf_counter()
            out = c.flush()
            t2 = time.perf_counter()
            if best is None or t2 - t0 < best[0]:
                best = (t2 - t0, t1 - t0, t2 - t1, cnt.value)
        print(f"{name} threads {T:2d}: add {best[1] * 1e3:7.1f} ms, flush {best[2] * 1e3:7.1f} ms, {len(recs) / best[0] / 1e6:.2f} M input reads/s, "
              f"{best[3] / best[0] / 1e6:.3f} M consensus reads/s", flush=True)
        c.close()
