"""Device-resident throughput of the duplex (K1 + K2) and CODEC (K1 + K3) pipelines at BASELINE sizes
(configs 3 and 4), with algorithmic bytes per SURVEY §8d.  Reference numbers for DESIGN.md; the
driver's bench line stays the simplex config.  usage: python scripts/bench_modes.py [scale]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fgumi_b200 as fg
from fgumi_b200 import synth

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
DEV, L = "cuda:0", 150
Lo = (L + 7) // 8 * 8
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]


def timed(fn, n=8):
    fn(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    return float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(n)]))


def duplex():
    # the bench leg itself: K1 + K2 as two kernels and with the combine in the vote kernels' epilogue
    from fgumi_b200 import benchlegs
    return benchlegs.duplex_leg(torch, fg, DEV, 0, int(5_000_000 * scale))


def codec():
    M = int(2_000_000 * scale)
    rng = np.random.default_rng(44)
    k = rng.integers(2, 21, size=M)
    depths = np.repeat(k, 2).astype(np.int64)
    tb = synth.device_batch(torch, DEV, depths, L, 1e-3, seed=44)
    eng = fg.Engine(0, 45, 40, 1, 0)
    ss = fg.DeviceColumns(tb.host.n_out, DEV)
    insert = np.clip(np.round(rng.normal(300, 50, size=M)), L, 2 * L).astype(np.int64)
    Lc_pad = (insert + 7) // 8 * 8
    jobs = np.zeros(M, dtype=fg.CODEC_JOB_DTYPE)
    r1n = rng.random(M) < 0.5
    jobs["unit_a"], jobs["unit_b"] = 2 * np.arange(M), 2 * np.arange(M) + 1
    jobs["out_off"][1:] = np.cumsum(Lc_pad)[:-1]
    jobs["len"] = insert
    jobs["rc_a"], jobs["rc_b"], jobs["rc_out"] = r1n, ~r1n, r1n
    jobs["pad_a_left"] = np.where(r1n, insert - L, 0)
    jobs["pad_b_left"] = np.where(~r1n, insert - L, 0)
    tj = torch.from_numpy(jobs.view(np.uint8).reshape(-1)).to(DEV)
    out = fg.DeviceColumns(int(Lc_pad.sum()), DEV)
    st = torch.zeros(M, dtype=torch.uint8, device=DEV)
    dis = torch.zeros(M, dtype=torch.int32, device=DEV); dup = torch.zeros_like(dis)
    cp = fg.lib.FgbCodecParams(-1, -1, 5, 0xFFFFFFFF, 1.0)
    s = torch.cuda.current_stream().cuda_stream
    t1 = timed(lambda: eng.vote_device(tb, ss, s))
    t3 = timed(lambda: eng.codec_combine_device(tb, ss, tj, M, cp, out, st, dis, dup, s))
    R = int(depths.sum())
    k1_bytes = R * 2 * L + 2 * M * 6 * L + 8 * (R + 2 * M) + 16 * M
    k3_bytes = M * 2 * 6 * L + int(insert.sum()) * 6 + 32 * M
    eng.close()
    return {"molecules": M, "k1_ms": t1, "k3_ms": t3, "molecules_per_s": M / ((t1 + t3) * 1e-3),
            "k1_gbs": k1_bytes / t1 / 1e6, "k3_gbs": k3_bytes / t3 / 1e6,
            "k1_frac": k1_bytes / t1 / 1e6 / peak, "k3_frac": k3_bytes / t3 / 1e6 / peak}


def aux_kernels():
    """K0 (PACK8 / BAM4 unpack) and K4 (filter epilogue) on 4 M depth-8 families, device resident."""
    import ctypes as C
    U = int(4_000_000 * scale)
    depths = np.full(U, 8, dtype=np.int64)
    tb = synth.device_batch(torch, DEV, depths, L, 1e-3, seed=42)
    eng = fg.Engine(0, 45, 40, 1, 2)
    lib = eng._lib
    out = fg.DeviceColumns(tb.host.n_out, DEV)
    s = torch.cuda.current_stream().cuda_stream
    eng.vote_device(tb, out, s)
    nb = tb.host.n_bytes
    # BAM4: treat the generated rows as forward-strand raw reads
    nibtab = torch.full((256,), 15, dtype=torch.uint8, device=DEV)
    for i, ch in enumerate(b"=ACMGRSVTWYHKDBN"):
        nibtab[ch] = i
    nib = nibtab[tb.bases[:nb].long()]
    seq4 = ((nib[0::2] << 4) | nib[1::2]).contiguous()
    R = tb.host.n_reads
    rr = np.zeros(R + 1, dtype=fg.RAW_READ_DTYPE)
    rr["src_off"][:R] = np.arange(R, dtype=np.uint64) * np.uint64(Lo)
    rr["raw_len"][:R] = L
    rrd = torch.from_numpy(rr.view(np.uint8).reshape(-1)).to(DEV)
    b2 = torch.zeros(nb + 64, dtype=torch.uint8, device=DEV); q2 = torch.zeros_like(b2)
    raw = fg.lib.FgbRawColumns(nb, seq4.data_ptr(), tb.quals.data_ptr(), rrd.data_ptr(), 10)
    bs = tb.struct()
    def bam4():
        assert lib.fgb_unpack_bam4_device(eng._h, C.byref(bs), C.byref(raw), C.c_void_p(b2.data_ptr()),
                                          C.c_void_p(q2.data_ptr()), C.c_void_p(s)) == 0
    t_b4 = timed(bam4)
    assert torch.equal(b2[:nb].view(-1, Lo)[:, :L], tb.bases[:nb].view(-1, Lo)[:, :L])
    b4_bytes = R * (L // 2 + L + 16 + 8 + 2 * L)
    # K4 on the voted columns
    fp = fg.lib.FgbFilterParams(2, 20, 0.05, 0.2, 30.0, 0.2, 1)
    status = torch.zeros(U, dtype=torch.uint8, device=DEV)
    cs = out.struct()
    def filt():
        assert lib.fgb_filter_simplex_device(eng._h, C.byref(bs), C.byref(cs), C.byref(fp),
                                             C.c_void_p(status.data_ptr()), None, C.c_void_p(s)) == 0
    t_f = timed(filt)
    f_bytes = U * (6 * L + 16 + 1)        # read 4 columns, (masked positions are rare) + status
    eng.close()
    return {"units": U, "bam4_ms": t_b4, "bam4_gbs": b4_bytes / t_b4 / 1e6, "bam4_frac": b4_bytes / t_b4 / 1e6 / peak,
            "filter_ms": t_f, "filter_gbs": f_bytes / t_f / 1e6, "filter_frac": f_bytes / t_f / 1e6 / peak}


print(json.dumps({"aux_kernels": aux_kernels()}))
torch.cuda.empty_cache()
print(json.dumps({"duplex_config3": duplex()}))
torch.cuda.empty_cache()
print(json.dumps({"codec_config4": codec()}))
