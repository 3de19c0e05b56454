"""Vote kernel time and counters by family depth (device resident, 1 M families each)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fgumi_b200 as fg
from fgumi_b200 import synth
U = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
minq = int(sys.argv[2]) if len(sys.argv) > 2 else 2
SPECS = sys.argv[3].split(",") if len(sys.argv) > 3 else ("1", "2", "3", "4", "6", "8", "12", "16", "20", "mixed2-20", "zipf1-100")
import json
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
for spec in SPECS:
    if spec.startswith("mixed"):
        depths = np.repeat(np.random.default_rng(1).integers(2, 21, size=U // 2), 2).astype(np.int64)
        if spec.startswith("mixeds"):
            depths = np.sort(depths)
    elif spec.startswith("zipf"):      # BASELINE config 5: P(d) ~ 1/d on 1..100; "zipfg": packed by depth class
        w = 1.0 / np.arange(1, 101)
        depths = np.random.default_rng(2).choice(np.arange(1, 101), size=U // 4, p=w / w.sum()).astype(np.int64)
        if spec.startswith("zipfs"):          # packed in order of depth
            depths = np.sort(depths)
        elif spec.startswith("zipfg"):        # packed by depth class only
            cls = np.where(depths <= 4, 1, np.where(depths >= 24, 2, 0))
            depths = depths[np.argsort(cls, kind="stable")]
    else:
        depths = np.full(U, int(spec), dtype=np.int64)
    tb = synth.device_batch(torch, "cuda:0", depths, 150, 1e-3, seed=7)
    eng = fg.Engine(0, 45, 40, 1, minq)
    out = fg.DeviceColumns(tb.host.n_out, "cuda:0")
    s = torch.cuda.current_stream().cuda_stream
    eng.vote_device(tb, out, s); torch.cuda.synchronize()
    eng.stats_reset()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    for i in range(3):
        eng.vote_device(tb, out, s); ev[i + 1].record()
    torch.cuda.synchronize()
    st = eng.stats()
    ms = min(ev[i].elapsed_time(ev[i + 1]) for i in range(3))
    nu = len(depths)
    nr = int(depths.sum())
    abytes = 2 * nr * 150 + 6 * nu * 150 + 8 * (nr + nu) + 8 * nu
    print(f"depth {spec:>9}: {ms:8.3f} ms  {ms * 1e6 / nu:7.2f} ns/unit  {abytes / ms / 1e6:7.0f} GB/s = {abytes / ms / 1e6 / PEAK:5.3f} of roofline  "
          f"exact/unit {st['exact_positions'] / 3 / nu:8.4f}  tiles {len(tb.host.tiles)} classes {tb.class_tiles}")
    eng.close()
    del tb, out
    torch.cuda.empty_cache()
