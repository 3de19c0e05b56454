"""Duplex consensus filter (duplex arm): pins oracle/record_oracle.py's restatement of
filter_duplex_read / mask_duplex_bases (fgumi-consensus filter.rs:477-557, 702-806) against the
reference's own duplex filter tests (src/lib/commands/filter.rs:2954-3100, 4285-4330) and the rule
table of filter.rs, then checks the C-ABI record filter against the oracle on random records."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import record_oracle as R           # noqa: E402
from tests.bam_builder import make_record       # noqa: E402

T = R.FilterThresholds


def drec(name, bases, quals, aD=None, bD=None, aE=None, bE=None, ad=None, bd=None, ae=None, be=None,
         ac=None, bc=None, cD=None, cE=None, aM=True, ac_array=False):
    """create_duplex_consensus_record, commands/filter.rs:2954-2989 (aD/bD/aM/bM ints, aE/bE floats,
    ad/bd/ae/be i16 arrays), plus the optional ac/bc strings and cD/cE the duplex caller writes."""
    tags = []
    if cD is not None:
        tags.append((b"cD", "i", cD))
    if cE is not None:
        tags.append((b"cE", "f", cE))
    if aD is not None:
        tags.append((b"aD", "i", aD))
    if bD is not None:
        tags.append((b"bD", "i", bD))
    if aM and aD is not None:
        tags.append((b"aM", "i", aD))
    if aM and bD is not None:
        tags.append((b"bM", "i", bD))
    if aE is not None:
        tags.append((b"aE", "f", aE))
    if bE is not None:
        tags.append((b"bE", "f", bE))
    n = len(bases)
    for tag, v in ((b"ad", ad), (b"bd", bd), (b"ae", ae if ae is not None or ad is None else [0] * n),
                   (b"be", be if be is not None or bd is None else [0] * n)):
        if v is not None:
            tags.append((tag, "Bs", list(v)))
    for tag, v in ((b"ac", ac), (b"bc", bc)):
        if v is not None:
            tags.append((tag, "BC" if ac_array else "Z", v))
    return bytearray(make_record(name=name, flags=4, ref_id=-1, pos=-1, cigar=[], seq=bases, quals=quals, tags=tags))


def test_filter_execute_duplex_consensus_port():      # commands/filter.rs:2992-3075
    recs = [drec(b"duplex_pass", b"AAAA", [30] * 4, 10, 10, 0.01, 0.01, [10] * 4, [10] * 4),
            drec(b"duplex_fail_ab", b"AAAA", [30] * 4, 2, 10, 0.01, 0.01, [2] * 4, [10] * 4),
            drec(b"duplex_fail_ba", b"AAAA", [30] * 4, 10, 2, 0.01, 0.01, [10] * 4, [2] * 4)]
    cc = ab = ba = None
    cc, ab, ba = T(5, 0.1, 0.3), T(5, 0.1, 0.3), T(5, 0.1, 0.3)
    o = R.DuplexFilterOracle(cc, ab, ba, min_base_quality=10, max_no_call_fraction=0.5)
    stream = b"".join(R.with_block_size(bytes(r)) for r in recs)
    out, kept = o.filter_stream(stream)
    assert kept == 1 and R.Rec(out[4:]).name == b"duplex_pass"
    # the strand with two reads fails the BA (lenient, worst-strand) tier in both failing records
    assert R.filter_duplex_read(R.Rec(bytes(recs[1])).aux(), cc, ab, ba) == R.FILTER_INSUFFICIENT_READS
    assert R.filter_duplex_read(R.Rec(bytes(recs[2])).aux(), cc, ab, ba) == R.FILTER_INSUFFICIENT_READS


def test_check_duplex_filters_no_call_count_mode():   # commands/filter.rs:4285-4330
    rec = drec(b"r", b"AANNNTTGGC", [30] * 10, 10, 8, 0.01, 0.01)
    th = T(5, 0.1, 0.1)
    aux = R.Rec(bytes(rec)).aux()
    assert R.is_duplex_consensus(aux)
    assert R.filter_duplex_read(aux, th, th, th) == R.FILTER_PASS
    assert R.check_no_call_and_quality(bytes(rec), None, 5.0)
    assert not R.check_no_call_and_quality(bytes(rec), None, 2.0)


def test_filter_duplex_read_tiers():                  # filter.rs:477-557
    aux = lambda **kw: R.Rec(bytes(drec(b"r", b"ACGT", [30] * 4, **kw))).aux()
    cc, ab, ba = T(1, 1.0, 1.0), T(6, 0.0625, 1.0), T(3, 0.25, 1.0)
    f = lambda a: R.filter_duplex_read(a, cc, ab, ba)
    assert f(aux(aD=6, bD=3, aE=0.0625, bE=0.25)) == R.FILTER_PASS            # exactly at every limit
    # the tag is an f32 widened to f64 (filter.rs:1418-1441): 0.05f32 > 0.05f64
    assert R.filter_duplex_read(aux(aD=6, bD=3, aE=0.05, bE=0.05), cc, T(6, 0.05, 1.0), ba) == R.FILTER_EXCESSIVE_ERROR_RATE
    assert f(aux(aD=5, bD=5, aE=0.0, bE=0.0)) == R.FILTER_INSUFFICIENT_READS  # best strand < AB tier
    assert f(aux(aD=9, bD=2, aE=0.0, bE=0.0)) == R.FILTER_INSUFFICIENT_READS  # worst strand < BA tier
    assert f(aux(aD=9, bD=3, aE=0.07, bE=0.08)) == R.FILTER_EXCESSIVE_ERROR_RATE   # best error > AB tier
    assert f(aux(aD=9, bD=3, aE=0.01, bE=0.26)) == R.FILTER_EXCESSIVE_ERROR_RATE   # worst error > BA tier
    assert f(aux(aD=3, bD=9, aE=0.26, bE=0.01)) == R.FILTER_EXCESSIVE_ERROR_RATE   # strand labels do not matter
    # one strand only: the other counts as depth 0; its error stands for both
    assert f(aux(aD=9, aE=0.01)) == R.FILTER_INSUFFICIENT_READS
    assert R.filter_duplex_read(aux(aD=9, aE=0.01), cc, ab, T(0, 0.2, 1.0)) == R.FILTER_PASS
    assert R.filter_duplex_read(aux(bD=9, bE=0.3), cc, T(6, 0.5, 1.0), T(0, 0.2, 1.0)) == R.FILTER_EXCESSIVE_ERROR_RATE
    # aM / bM stand in for missing aD / bD; no strand tags at all passes
    a = R.Rec(bytes(make_record(name=b"r", flags=4, ref_id=-1, pos=-1, cigar=[], seq=b"AC", quals=[30, 30],
                                tags=[(b"aM", "i", 9), (b"bM", "i", 2)]))).aux()
    assert f(a) == R.FILTER_INSUFFICIENT_READS
    assert f(R.Rec(bytes(make_record(name=b"r", flags=4, ref_id=-1, pos=-1, cigar=[], seq=b"AC", quals=[30, 30]))).aux()) == R.FILTER_PASS
    # the CC tier (cD / cE) is checked first
    assert R.filter_duplex_read(aux(aD=9, bD=9, cD=10, cE=0.5), T(11, 1.0, 1.0), ab, ba) == R.FILTER_INSUFFICIENT_READS
    assert R.filter_duplex_read(aux(aD=9, bD=9, cD=10, cE=0.5), T(1, 0.4, 1.0), ab, ba) == R.FILTER_EXCESSIVE_ERROR_RATE


def test_mask_duplex_bases_rules():                   # filter.rs:702-806
    loose = T(0, 1.0, 1.0)

    def run(cc=loose, ab=loose, ba=loose, mbq=None, ss=False, **kw):
        rec = drec(b"r", b"ACGTAC", [30, 30, 30, 5, 30, 30], **kw)
        n = R.mask_duplex_bases(rec, cc, ab, ba, mbq, ss)
        r = R.Rec(bytes(rec))
        return n, bytes(r.sequence()), list(r.quals())
    ad, bd = [10, 4, 2, 10, 0, 10], [10, 10, 1, 10, 9, 10]
    ae, be = [0, 0, 0, 0, 0, 5], [0, 5, 0, 0, 0, 0]
    assert run(ad=ad, bd=bd, ae=ae, be=be) == (0, b"ACGTAC", [30, 30, 30, 5, 30, 30])
    assert run(mbq=10, ad=ad, bd=bd, ae=ae, be=be) == (1, b"ACGNAC", [30, 30, 30, 2, 30, 30])
    assert run(cc=T(4, 1.0, 1.0), ad=ad, bd=bd, ae=ae, be=be)[1] == b"ACNTAC"          # total depth 3 < 4
    assert run(ab=T(10, 1.0, 1.0), ad=ad, bd=bd, ae=ae, be=be)[1] == b"ACNTNC"         # best depth 2 / 9 < 10
    assert run(ba=T(4, 1.0, 1.0), ad=ad, bd=bd, ae=ae, be=be)[1] == b"ACNTNC"          # worst depth 1 / 0 < 4
    assert run(cc=T(0, 1.0, 0.3), ad=ad, bd=bd, ae=ae, be=be)[1] == b"ANGTAC"          # 5/14 > 0.3; 5/20 is not
    assert run(ab=T(0, 1.0, 0.0), ad=ad, bd=bd, ae=ae, be=be)[1] == b"ACGTAC"          # best rate is 0 everywhere
    assert run(ba=T(0, 1.0, 0.45), ad=ad, bd=bd, ae=ae, be=be)[1] == b"ANGTAN"   # worst 0.5 twice
    # missing arrays read as zero depth; short arrays too
    assert run(cc=T(1, 1.0, 1.0))[0] == 6
    assert run(cc=T(1, 1.0, 1.0), ad=[3, 3], bd=[1])[1] == b"ACNNNN"
    # an N already in the read is skipped (not counted)
    rec = drec(b"r", b"ANGT", [30] * 4, ad=[0] * 4, bd=[0] * 4)
    assert R.mask_duplex_bases(rec, T(1, 1.0, 1.0), loose, loose, None, False) == 3
    # single-strand agreement: only where both strands have depth; Z strings or u8 arrays; a short
    # or missing tag reads as N
    kw = dict(ad=[5] * 6, bd=[5, 5, 0, 5, 5, 5])
    assert run(ss=True, ac=b"ACGTAC", bc=b"ACTTAG", **kw)[1] == b"ACGTAN"              # pos 2: bd = 0, skipped
    assert run(ss=True, ac=b"ACGTAC", bc=b"ACG", **kw)[1] == b"ACGNNN"
    assert run(ss=True, ac=b"ACGTAC", **kw)[1] == b"NNGNNN"
    assert run(ss=True, ac=b"ACGTAC", bc=b"ACTTAG", ac_array=True, **kw)[1] == b"ACGTAN"
    assert run(ss=False, ac=b"ACGTAC", bc=b"TTTTTT", **kw)[0] == 0


def test_array_element_types():                       # raw-bam tags.rs:479-497
    """Signed elements clamp at 0 (a depth the caller stored as a wrapped i16 reads as 0)."""
    rec = drec(b"r", b"ACGT", [30] * 4, ad=[-5, 3, 3, 3], bd=[3, 3, -1, 3])
    n = R.mask_duplex_bases(rec, T(4, 1.0, 1.0), T(0, 1.0, 1.0), T(0, 1.0, 1.0), None, False)
    assert n == 2 and bytes(R.Rec(bytes(rec)).sequence()) == b"NCNT"


# ---- the product's host filter (C-ABI, no device) against the oracle -------------------------------
def _random_record(rng):
    n = int(rng.integers(1, 40))
    bases = bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=n, p=[0.23, 0.23, 0.23, 0.23, 0.08]))
    quals = rng.integers(2, 60, size=n).tolist()
    kw = {}
    kind = int(rng.integers(0, 5))
    arr = lambda hi, m=n: rng.integers(-2, hi, size=int(m)).tolist()
    if kind < 4:                                       # duplex record, tags present at random
        if rng.random() < 0.9:
            kw["aD"] = int(rng.integers(-1, 12))
        if rng.random() < 0.8:
            kw["bD"] = int(rng.integers(0, 12))
        kw["aM"] = bool(rng.random() < 0.5)
        if rng.random() < 0.8:
            kw["aE"] = float(np.float32(rng.random() * 0.2))
        if rng.random() < 0.8:
            kw["bE"] = float(np.float32(rng.random() * 0.2))
        m = n if rng.random() < 0.8 else max(1, n - int(rng.integers(0, 4)))
        if rng.random() < 0.9:
            kw["ad"], kw["ae"] = arr(12, m), arr(4, m)
        if rng.random() < 0.9:
            kw["bd"], kw["be"] = arr(12, m), arr(4, m)
        if rng.random() < 0.7:
            k = m if rng.random() < 0.7 else max(1, m - 1)
            kw["ac"] = bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=k))
            kw["bc"] = bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=k)) if rng.random() < 0.5 else kw["ac"]
            kw["ac_array"] = bool(rng.random() < 0.3)
        if rng.random() < 0.7:
            kw["cD"], kw["cE"] = int(rng.integers(0, 25)), float(np.float32(rng.random() * 0.2))
        return drec(b"q", bases, quals, **kw)
    tags = [(b"cD", "i", int(rng.integers(0, 12))), (b"cE", "f", float(np.float32(rng.random() * 0.2))),
            (b"cd", "Bs", arr(12)), (b"ce", "Bs", arr(4))]
    return bytearray(make_record(name=b"s", flags=4, ref_id=-1, pos=-1, cigar=[], seq=bases, quals=quals, tags=tags))


def test_host_record_filter_matches_oracle():
    import fgumi_b200 as fg
    rng = np.random.default_rng(2024)
    seen = set()
    for trial in range(1500):
        mr = sorted(rng.integers(0, 9, size=3).tolist(), reverse=True)            # cc >= ab >= ba
        er = sorted((rng.random(2) * 0.2).tolist())                               # ab <= ba
        br = sorted((rng.random(2) * 0.5).tolist())
        cc_er, cc_br = float(rng.random() * 0.3), float(rng.random() * 0.6)
        mbq = None if rng.random() < 0.4 else int(rng.integers(0, 40))
        mmq = None if rng.random() < 0.5 else float(rng.random() * 45)
        nc = float(rng.random() * 0.6) if rng.random() < 0.7 else float(rng.integers(1, 10))
        ss = bool(rng.random() < 0.5)
        f = fg.DuplexConsensusFilter(mr, (cc_er, er[0], er[1]), (cc_br, br[0], br[1]), mbq, mmq, nc, ss)
        o = R.DuplexFilterOracle(T(mr[0], cc_er, cc_br), T(mr[1], er[0], br[0]), T(mr[2], er[1], br[1]),
                                 mbq, mmq, nc, ss)
        rec = _random_record(rng)
        want = bytearray(rec)
        keep = o.process_record(want)
        got = bytearray(rec)
        status, masked = f.apply(got)
        assert got == want, trial
        assert masked == o.bases_masked, trial
        assert (status == fg.lib.FGB_FILTER_PASS) == keep, (trial, status)
        seen.add(status)
    assert seen == {0, 1, 2, 3, 4}, seen


# ---- duplex caller with the filter enabled (GPU) ---------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("fopt,min_reads,threads", [
    (dict(min_reads=(3, 2, 1), max_read_error_rate=(0.05, 0.05, 0.1), max_base_error_rate=(0.2, 0.1, 0.3),
          min_base_quality=20, max_no_call_fraction=0.3), (1, 1, 1), 1),
    (dict(min_reads=(2,), max_read_error_rate=(0.1,), max_base_error_rate=(0.3,), min_base_quality=None,
          min_mean_base_quality=30.0, max_no_call_fraction=4.0, require_single_strand_agreement=True), (1, 1, 0), 3),
])
def test_duplex_caller_with_filter(fopt, min_reads, threads):
    """DuplexConsensusCaller(filter=...) == oracle duplex caller followed by the oracle's duplex
    `filter` restatement in template mode; counters too."""
    import fgumi_b200 as fg
    from tests import oracle_lib as O
    from tests.bam_builder import parse_records
    from tests.test_caller_parity import random_duplex_groups, duplex_job_fn
    from tests.test_record_oracle_kat import vote_fn
    rng = np.random.default_rng(900 + len(fopt["min_reads"]))
    groups = random_duplex_groups(rng, 220)
    oracle = R.DuplexCallerOracle("fgumi", "A", min_reads=min_reads, per_base=True, cell_tag=b"CB",
                                  vote_fn=vote_fn, builder_fn=O.builder_call, duplex_job_fn=duplex_job_fn)
    stream = bytearray()
    for g in groups:
        stream += oracle.consensus_reads(g)[0]
    f = fg.DuplexConsensusFilter(**fopt)
    three = lambda v: (list(v) + [list(v)[-1]] * 3)[:3]
    mr, er, br = three(fopt["min_reads"]), three(fopt["max_read_error_rate"]), three(fopt["max_base_error_rate"])
    flt = R.DuplexFilterOracle(T(mr[0], er[0], br[0]), T(mr[1], er[1], br[1]), T(mr[2], er[2], br[2]),
                               fopt.get("min_base_quality"), fopt.get("min_mean_base_quality"),
                               fopt["max_no_call_fraction"], fopt.get("require_single_strand_agreement", False))
    want, kept = flt.filter_stream(bytes(stream))
    c = fg.DuplexConsensusCaller("fgumi", "A", min_reads=min_reads, produce_per_base_tags=True, device=0,
                                 cell_tag=b"CB", filter=f, n_threads=threads)
    got = c.consensus_reads_batch(groups)
    st = c.statistics()
    c.close()
    if got.data != want:
        a, b = parse_records(got.data), parse_records(want)
        assert len(a) == len(b), (len(a), len(b))
        for i, (x, y) in enumerate(zip(a, b)):
            assert x == y, (i, x, y)
    assert got.data == want
    assert st["filter_records"] == flt.total and st["filter_passed"] == flt.passed
    assert st["filter_bases_masked"] == flt.bases_masked
    assert 0 < flt.passed < flt.total and flt.bases_masked > 0


def test_int_tag_variants():                          # duplex_caller.rs:4734-4780, raw-bam tags.rs:118-150
    """Every BAM integer type reads as an integer tag; a float or a missing tag does not -- in the
    oracle and in the product's host filter (cD feeds the CC min-reads gate)."""
    import fgumi_b200 as fg
    for typ, val in (("c", 42), ("C", 200), ("s", 1000), ("S", 50000), ("i", 100000), ("I", 200000)):
        rec = bytearray(make_record(name=b"r", flags=4, ref_id=-1, pos=-1, cigar=[], seq=b"AC", quals=[30, 30],
                                    tags=[(b"cD", typ, val)]))
        tags = R._aux_tags(R.Rec(bytes(rec)).aux())
        assert R._find_int(tags, b"cD") == val
        for mr, want in ((val, fg.lib.FGB_FILTER_PASS), (val + 1, fg.lib.FGB_FILTER_INSUFFICIENT_READS)):
            f = fg.DuplexConsensusFilter((mr, 0, 0), (1.0,), (1.0,), None, None, 1.0)
            # per-base masks see depth 0 < min_reads: every base is masked, but the read gate is cD's
            assert R.filter_read(R.Rec(bytes(rec)).aux(), T(mr, 1.0, 1.0)) == want
            status, _ = f.apply(bytearray(rec))
            assert (status == fg.lib.FGB_FILTER_INSUFFICIENT_READS) == (want == fg.lib.FGB_FILTER_INSUFFICIENT_READS)
    rec = make_record(name=b"r", flags=4, ref_id=-1, pos=-1, cigar=[], seq=b"AC", quals=[30, 30], tags=[(b"cD", "f", 1.5)])
    tags = R._aux_tags(R.Rec(bytes(rec)).aux())
    assert R._find_int(tags, b"cD") is None and R._find_int(tags, b"aD") is None
