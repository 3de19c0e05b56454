"""Record-level boundary: phase times of one batch (add_groups / flush, FGB_CALLER_TRACE phases) and the leg of
bench.py.  usage: python scripts/bench_records.py [families] [threads]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import fgumi_b200 as fg
from fgumi_b200 import benchlegs

G = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16
lib = fg.lib.load()
if os.environ.get("FGB_BIND_NUMA"):      # like bench.py: run (and first-touch page-locked memory) on the GPU's NUMA node
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    print("numa:", bench.bind_to_gpu_numa(torch, 0))
pinned, rec_off, group_rec, rec_len = benchlegs.make_record_batch(torch, G, 8)
bp, op, gp = pinned.data_ptr(), rec_off.ctypes.data, group_rec.ctypes.data
c = benchlegs._Caller(lib, 0, T)
for rep in range(4):
    if rep == 3:
        os.environ["FGB_CALLER_TRACE"] = "1"
    t0 = time.perf_counter()
    st = lib.fgb_caller_add_groups(c.h, bp, op, gp, G)
    t1 = time.perf_counter()
    data, n, cnt = C.c_void_p(), C.c_uint64(), C.c_uint64()
    st2 = lib.fgb_caller_flush(c.h, C.byref(data), C.byref(n), C.byref(cnt))
    t2 = time.perf_counter()
    assert st == 0 and st2 == 0
    print(f"rep {rep}: add_groups {1e3 * (t1 - t0):.1f} ms, flush {1e3 * (t2 - t1):.1f} ms, {cnt.value} reads, "
          f"{G / (t2 - t0) / 1e6:.2f} M consensus reads/s")
os.environ.pop("FGB_CALLER_TRACE", None)
c.close()
print(benchlegs.records_leg(torch, fg, lib, 0, G, T, steps=4, warmup=2, callers=2))
