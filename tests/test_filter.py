"""Consensus filter as a fused epilogue, single-strand reads.
CPU: pins oracle/record_oracle.py's filter restatement against the reference's tests
(src/lib/commands/filter.rs:1499-1716, 4199-4283).  GPU: the simplex caller with a filter against
oracle caller -> oracle filter."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import record_oracle as R           # noqa: E402
from tests.bam_builder import make_record, parse_records   # noqa: E402


def frec(seq, quals, cD=None, cE=None, cd=None, ce=None):
    """create_filter_test_record, commands/filter.rs:1430-1495"""
    tags = []
    if cD is not None:
        tags.append((b"cD", "C", cD))
    if cE is not None:
        tags.append((b"cE", "f", cE))
    if cd is not None:
        tags.append((b"cd", "Bs", cd))
    if ce is not None:
        tags.append((b"ce", "Bs", ce))
    return bytearray(make_record(name=b"test", flags=4, ref_id=-1, pos=-1, cigar=[], seq=seq, quals=quals, tags=tags))


def test_read_stats():                                # :1499-1551
    assert R.compute_read_stats(frec(b"", [])) == (0, 0.0)
    assert R.compute_read_stats(frec(b"ACGT", [30] * 4)) == (0, 30.0)
    assert R.compute_read_stats(frec(b"ACNTN", [30, 30, 0, 30, 0]))[0] == 2
    assert R.compute_read_stats(frec(b"NNNN", [0] * 4)) == (4, 0.0)
    assert R.compute_read_stats(frec(b"ACGT", [10, 20, 30, 40]))[1] == 25.0


def test_mask_bases():                                # :1553-1633
    r = frec(b"ACGT", [10, 30, 5, 30], cd=[10] * 4, ce=[0] * 4)
    assert R.mask_bases(r, R.FilterThresholds(1, 1.0, 1.0), 20) == 2
    v = R.Rec(bytes(r))
    assert bytes(v.sequence()) == b"NCNT" and list(v.quals()) == [2, 30, 2, 30]
    r = frec(b"ACGT", [30] * 4, cd=[1, 10, 4, 10])
    R.mask_bases(r, R.FilterThresholds(5, 1.0, 1.0), 10)
    assert bytes(R.Rec(bytes(r)).sequence()) == b"NCNT"
    r = frec(b"ACGT", [30] * 4, cd=[10] * 4, ce=[1, 3, 2, 0])
    R.mask_bases(r, R.FilterThresholds(1, 1.0, 0.2), 10)
    assert bytes(R.Rec(bytes(r)).sequence()) == b"ANGT"      # 2/10 == 0.2 is not masked (strictly greater)
    r = frec(b"ACGT", [30] * 4)                              # no per-base tags: depth 0 everywhere
    assert R.mask_bases(r, R.FilterThresholds(1, 1.0, 1.0), None) == 4


def test_filter_read():                               # :1635-1715
    th = R.FilterThresholds(5, 0.1, 0.2)
    assert R.filter_read(R.Rec(bytes(frec(b"ACGT", [30] * 4, cD=10, cE=0.05))).aux(), th) == R.FILTER_PASS
    assert R.filter_read(R.Rec(bytes(frec(b"ACGT", [30] * 4, cD=3, cE=0.05))).aux(), th) == R.FILTER_INSUFFICIENT_READS
    assert R.filter_read(R.Rec(bytes(frec(b"ACGT", [30] * 4, cD=10, cE=0.3))).aux(), th) == R.FILTER_EXCESSIVE_ERROR_RATE
    assert R.filter_read(R.Rec(bytes(frec(b"ACGT", [30] * 4))).aux(), th) == R.FILTER_PASS
    # f32 -> f64 promotion, filter.rs:1418-1441
    assert R.filter_read(R.Rec(bytes(frec(b"A", [30], cD=10, cE=0.099))).aux(), th) == R.FILTER_PASS
    assert R.filter_read(R.Rec(bytes(frec(b"A", [30], cD=10, cE=0.101))).aux(), th) == R.FILTER_EXCESSIVE_ERROR_RATE


def test_no_call_modes():                             # :4199-4283
    r = bytes(frec(b"AANNTTGGCC", [30] * 10, cD=10, cE=0.01))
    assert R.check_no_call_and_quality(r, None, 0.2) and not R.check_no_call_and_quality(r, None, 0.19)
    r = bytes(frec(b"AANNNTTGGC", [30] * 10, cD=10, cE=0.01))
    assert R.check_no_call_and_quality(r, None, 5.0) and R.check_no_call_and_quality(r, None, 3.0)
    assert not R.check_no_call_and_quality(r, None, 2.0)
    assert not R.check_no_call_and_quality(r, 31.0, 5.0) and R.check_no_call_and_quality(r, 30.0, 5.0)


def _filtered_oracle(groups, fopt, vopt_kw, per_base=True):
    from tests import oracle_lib as O
    from tests.test_record_oracle_kat import vote_fn
    caller = R.VanillaCallerOracle("fgumi", "A", R.VanillaOptions(produce_per_base_tags=per_base, **vopt_kw),
                                   vote_fn, O.builder_call)
    stream = bytearray()
    for g in groups:
        d, _ = caller.consensus_reads(g)
        stream += d
    flt = R.SimplexFilterOracle(R.FilterThresholds(fopt["min_reads"], fopt["max_read_error_rate"],
                                                   fopt["max_base_error_rate"]),
                                fopt.get("min_base_quality"), fopt.get("min_mean_base_quality"),
                                fopt.get("max_no_call_fraction", 0.2))
    data, kept = flt.filter_stream(bytes(stream))
    return data, kept, flt


@pytest.mark.gpu
@pytest.mark.parametrize("fopt,per_base", [
    (dict(min_reads=2, max_read_error_rate=0.05, max_base_error_rate=0.2, min_base_quality=20,
          max_no_call_fraction=0.3), True),
    (dict(min_reads=1, max_read_error_rate=1.0, max_base_error_rate=0.34, min_base_quality=None,
          min_mean_base_quality=38.0, max_no_call_fraction=4.0), True),
    (dict(min_reads=3, max_read_error_rate=0.02, max_base_error_rate=0.1, min_base_quality=30,
          max_no_call_fraction=0.9), True),
    (dict(min_reads=1, max_read_error_rate=0.5, max_base_error_rate=0.5, min_base_quality=5,
          max_no_call_fraction=1.0), False),          # no cd/ce arrays: every base has depth 0 for the mask
])
def test_simplex_caller_with_filter_epilogue(fopt, per_base):
    """caller(filter=...) == oracle caller followed by the oracle's `filter` restatement."""
    import fgumi_b200 as fg
    from tests.test_caller_parity import random_groups
    rng = np.random.default_rng(500 + fopt["min_reads"])
    groups = random_groups(rng, 200)
    vkw = dict(min_reads=1, min_consensus_base_quality=2)
    want, kept, flt = _filtered_oracle(groups, fopt, vkw, per_base)
    opts = fg.VanillaUmiConsensusOptions(min_reads=1, min_consensus_base_quality=2, produce_per_base_tags=per_base)
    f = fg.ConsensusFilter(**fopt)
    c = fg.VanillaUmiConsensusCaller("fgumi", "A", opts, filter=f)
    got = c.consensus_reads_batch(groups)
    st = c.statistics()
    c.close()
    assert got.count == kept
    if got.data != want:
        a, b = parse_records(got.data), parse_records(want)
        for i, (x, y) in enumerate(zip(a, b)):
            assert x == y, (i, x, y)
    assert got.data == want
    assert st["filter_records"] == flt.total and st["filter_passed"] == flt.passed
    assert st["filter_bases_masked"] == flt.bases_masked
    assert 0 < flt.passed < flt.total or not per_base


@pytest.mark.gpu
def test_filter_device_statuses():
    """fgb_filter_simplex_device on device-resident columns: every status value is produced and matches
    a straightforward evaluation of the same rules on the host copies."""
    import ctypes as C
    import torch
    import fgumi_b200 as fg
    rng = np.random.default_rng(9)
    units = []
    for i in range(4000):
        depth = int(rng.integers(1, 7))
        L = int(rng.integers(12, 120))
        err = 0.3 if i % 7 == 0 else 0.01
        tmpl = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=L)
        rows = []
        for _ in range(depth):
            b = tmpl.copy()
            m = rng.random(L) < err
            b[m] = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=int(m.sum()))
            q = rng.integers(15 if i % 5 else 2, 41, size=L).astype(np.uint8)
            if i % 11 == 0:
                b[rng.random(L) < 0.4] = ord("N")
            rows.append((b.tobytes(), q.tobytes()))
        units.append(rows)
    batch = fg.pack_source_reads(units, 1)
    eng = fg.Engine(0, 45, 40, 1, 2)
    db = fg.DeviceBatch(batch, "cuda:0")
    out = fg.DeviceColumns(batch.n_out, "cuda:0")
    s = torch.cuda.current_stream().cuda_stream
    eng.vote_device(db, out, s)
    torch.cuda.synchronize()
    pre = out.to_host()
    fp = fg.lib.FgbFilterParams(2, 25, 0.08, 0.25, 44.0, 0.15, 1)
    status = torch.full((batch.n_units,), 77, dtype=torch.uint8, device="cuda:0")
    masked = torch.zeros(batch.n_units, dtype=torch.int32, device="cuda:0")
    b, c = db.struct(), out.struct()
    st = eng._lib.fgb_filter_simplex_device(eng._h, C.byref(b), C.byref(c), C.byref(fp),
                                            C.c_void_p(status.data_ptr()), C.c_void_p(masked.data_ptr()),
                                            C.c_void_p(s))
    assert st == 0
    torch.cuda.synchronize()
    post = out.to_host()
    gs, gm = status.cpu().numpy(), masked.cpu().numpy()
    seen = set()
    for u, sl in enumerate(batch.unit_slices()):
        b0, q0 = pre.base[sl].copy(), pre.qual[sl].copy()
        d, e = pre.depth[sl].astype(np.int64), pre.errors[sl].astype(np.int64)
        with np.errstate(divide="ignore", invalid="ignore"):
            rate = np.where(d > 0, e.astype(np.float64) / np.maximum(d, 1).astype(np.float64), 0.0)
        m = (q0 < 25) | (d < 2) | ((d > 0) & (rate > 0.25))
        newly = int((m & (b0 != ord("N"))).sum())
        b0[m] = ord("N"); q0[m] = 2
        assert np.array_equal(post.base[sl], b0) and np.array_equal(post.qual[sl], q0)
        assert gm[u] == newly
        cD = int(d.max()); td, te = int(d.sum()), int(e.sum())
        cE = np.float32(0) if td == 0 else np.float32(te) / np.float32(td)
        n_n = int((b0 == ord("N")).sum()); non_n = len(b0) - n_n
        mean = float(q0[b0 != ord("N")].astype(np.int64).sum()) / non_n if non_n else 0.0
        if cD < 2: want = 1
        elif float(cE) > 0.08: want = 2
        elif mean < 44.0: want = 3
        elif n_n / len(b0) > 0.15: want = 4
        else: want = 0
        assert gs[u] == want, (u, gs[u], want)
        seen.add(want)
    assert seen == {0, 1, 2, 3, 4}
    stt = eng.stats()
    assert stt["filter_records"] == batch.n_units and stt["filter_passed"] == int((gs == 0).sum())
    assert stt["filter_bases_masked"] == int(gm.sum())
    eng.close()
