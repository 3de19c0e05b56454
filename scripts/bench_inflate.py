"""BGZF inflate on the device (fgb_bgzf_inflate_device, one member per thread) against the host reader (zlib on all
granted CPUs): GB/s of inflated output.  usage: python scripts/bench_inflate.py [MB of records]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fgumi_b200 as fg

MB = int(sys.argv[1]) if len(sys.argv) > 1 else 400
lib = fg.lib.load()
rng = np.random.default_rng(1)
n = MB * 1_000_000 // 337
recs = np.zeros((n, 337), np.uint8)                     # input-BAM-like records: header, name, packed bases, qualities, tags
recs[:, :36] = np.arange(36, dtype=np.uint8) * 7
recs[:, 4:8] = rng.integers(0, 255, size=(n, 4))
recs[:, 36:48] = np.frombuffer(b"read:0000000", np.uint8); recs[:, 41:48] = rng.integers(48, 58, size=(n, 7))
recs[:, 48:123] = rng.integers(0, 256, size=(n, 75))
recs[:, 123:273] = np.clip(np.rint(rng.normal(36, 4, size=(n, 150))), 2, 41)
recs[:, 273:] = np.frombuffer((b"MIZ1234567/A\0RXZACGTACGT-TTGACAGT\0MCZ150M\0RGZA\0" + b"\0" * 64)[:64], np.uint8)
data = recs.reshape(-1)
cap = lib.fgb_bgzf_bound(data.size); comp = np.zeros(cap, np.uint8); cl = C.c_size_t()
assert lib.fgb_bgzf_compress(data.ctypes.data, data.size, 1, 16, 1, comp.ctypes.data, cap, C.byref(cl)) == 0
comp = comp[: cl.value].copy()
nm, total = C.c_uint64(), C.c_uint64()
assert lib.fgb_bgzf_scan_members(comp.ctypes.data, comp.size, None, 0, C.byref(nm), C.byref(total)) == 0
MEMBER = np.dtype([("in_off", "<u8"), ("out_off", "<u8"), ("in_len", "<u4"), ("out_len", "<u4"), ("crc", "<u4"), ("reserved", "<u4")])
members = np.zeros(nm.value, dtype=MEMBER)
t0 = time.perf_counter()
assert lib.fgb_bgzf_scan_members(comp.ctypes.data, comp.size, members.ctypes.data, nm.value, C.byref(nm), C.byref(total)) == 0
scan_ms = (time.perf_counter() - t0) * 1e3
print(f"{data.size / 1e6:.0f} MB of records -> {comp.size / 1e6:.0f} MB compressed (ratio {comp.size / data.size:.3f}), {nm.value} members; host scan {scan_ms:.2f} ms")
dev = "cuda:0"
eng = fg.Engine(0, 45, 40, 1, 2)
d_in = torch.from_numpy(comp).to(dev)
d_m = torch.from_numpy(members.view(np.uint8).reshape(-1).copy()).to(dev)
d_out = torch.empty(total.value + 64, dtype=torch.uint8, device=dev)
d_st = torch.zeros(nm.value, dtype=torch.uint8, device=dev)
for crc in (0, 1):
    ms = []
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        assert lib.fgb_bgzf_inflate_device(eng._h, d_in.data_ptr(), d_m.data_ptr(), nm.value, d_out.data_ptr(), d_st.data_ptr(), crc, None, None) == 0
        e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = min(ms[1:])
    assert not bool(d_st.any())
    print(f"device inflate (crc check {crc}): {t:.2f} ms = {total.value / t / 1e6:.1f} GB/s of output, {comp.size / t / 1e6:.1f} GB/s of input")
assert bool((d_out[: total.value].cpu() == torch.from_numpy(data)).all())
back = np.zeros(total.value, np.uint8); bl = C.c_size_t()
for th in (1, 16):
    best = 1e9
    for _ in range(2):
        t0 = time.perf_counter()
        assert lib.fgb_bgzf_decompress(comp.ctypes.data, comp.size, th, back.ctypes.data, back.size, C.byref(bl)) == 0
        best = min(best, time.perf_counter() - t0)
    print(f"host inflate (zlib, {th} threads): {total.value / best / 1e9:.2f} GB/s of output")
eng.close()
