"""Overlapping-bases pre-pass.
1. pins oracle/record_oracle.py::OverlappingOracle against the reference's own tests
   (crates/fgumi-consensus/src/overlapping.rs:801-1400);
2. compares the product's host pre-pass (fgb_overlap_apply_group, csrc/host/overlap.h) with the
   oracle byte for byte on the same KATs and on random pairs with indels and clips;
3. (gpu) the callers with consensus_call_overlapping_bases=True against oracle pre-pass + oracle caller."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import record_oracle as R           # noqa: E402
from tests.bam_builder import make_record, encode_op, parse_records   # noqa: E402

P, F1, F2, REV, MREV = R.PAIRED, R.FIRST_SEGMENT, R.LAST_SEGMENT, R.REVERSE, R.MATE_REVERSE
M, I, D, N_, S, H = 0, 1, 2, 3, 4, 5
AC, AM, AP = R.AGREE_CONSENSUS, R.AGREE_MAX_QUAL, R.AGREE_PASS_THROUGH
DC, DB, DL = R.DISAGREE_CONSENSUS, R.DISAGREE_MASK_BOTH, R.DISAGREE_MASK_LOWER


def rec(seq, qual, start, cigar, flags=0, name=b"rea", ref_id=0):
    """create_raw_test_record / make_raw_bam, overlapping.rs:727-797 (1-based start)."""
    return make_record(name=name, flags=flags, ref_id=ref_id, pos=start - 1,
                       cigar=[encode_op(k, n) for k, n in cigar], seq=seq, quals=qual)


def both(agree, disagree, r1, r2):
    """Run oracle and product on the same pair; assert identical bytes and stats; return them."""
    import fgumi_b200 as fg
    o = R.OverlappingOracle(agree, disagree)
    a, b = bytearray(r1), bytearray(r2)
    ret = o.call(a, b)
    fl1 = (R.Rec(r1).flags & ~0xC1) | P | F1      # the product API pairs by name + segment flags
    fl2 = (R.Rec(r2).flags & ~0xC1) | P | F2
    p1, p2 = bytearray(r1), bytearray(r2)
    p1[14:16] = fl1.to_bytes(2, "little")
    p2[14:16] = fl2.to_bytes(2, "little")
    got, stats = fg.apply_overlapping_consensus([bytes(p1), bytes(p2)], agree, disagree)
    g1, g2 = bytearray(got[0]), bytearray(got[1])
    g1[14:16] = r1[14:16]
    g2[14:16] = r2[14:16]
    assert bytes(g1) == bytes(a) and bytes(g2) == bytes(b)
    assert stats == o.stats()
    return ret, R.Rec(bytes(a)), R.Rec(bytes(b)), o


Q = lambda v: bytes([v] * 4)
C4 = [(M, 4)]


def test_agreement_strategies():                      # :801-847
    ret, a, b, _ = both(AC, DC, rec(b"ACGT", Q(30), 100, C4), rec(b"ACGT", Q(20), 100, C4))
    assert ret and a.quals()[0] == 50 and b.quals()[0] == 50
    ret, a, b, _ = both(AM, DC, rec(b"ACGT", Q(30), 100, C4), rec(b"ACGT", Q(20), 100, C4))
    assert ret and a.quals()[0] == 30 and b.quals()[0] == 30
    ret, a, b, _ = both(AP, DC, rec(b"ACGT", Q(30), 100, C4), rec(b"ACGT", Q(20), 100, C4))
    assert ret and a.quals()[0] == 30 and b.quals()[0] == 20


def test_disagreement_strategies():                   # :849-990
    ret, a, b, _ = both(AP, DC, rec(b"ACGT", Q(30), 100, C4), rec(b"GCTA", Q(20), 100, C4))
    assert ret and a.sequence()[0] == ord("A") and b.sequence()[0] == ord("A")
    assert a.quals()[0] == 10 and b.quals()[0] == 10
    ret, a, b, _ = both(AP, DB, rec(b"ACGT", Q(30), 100, C4), rec(b"GCTA", Q(20), 100, C4))
    assert a.sequence()[0] == ord("N") and b.sequence()[0] == ord("N") and a.quals()[0] == 2 and b.quals()[0] == 2
    ret, a, b, _ = both(AP, DL, rec(b"ACGT", Q(30), 100, C4), rec(b"GCTA", Q(20), 100, C4))
    assert a.sequence()[0] == ord("A") and b.sequence()[0] == ord("N") and a.quals()[0] == 30 and b.quals()[0] == 2
    ret, a, b, _ = both(AP, DL, rec(b"ACGT", Q(20), 100, C4), rec(b"GCTA", Q(30), 100, C4))
    assert a.sequence()[0] == ord("N") and b.sequence()[0] == ord("G")
    for strat in (DL, DC):                            # equal quality: both masked
        ret, a, b, _ = both(AP, strat, rec(b"ACGT", Q(30), 100, C4), rec(b"GCTA", Q(30), 100, C4))
        assert a.sequence()[0] == ord("N") and b.sequence()[0] == ord("N")
        assert a.quals()[0] == 2 and b.quals()[0] == 2
    ret, a, b, _ = both(AP, DC, rec(b"ACGT", Q(30), 100, C4), rec(b"GCTA", Q(29), 100, C4))   # :1189-1205
    assert a.quals()[0] == 2 and b.quals()[0] == 2 and a.sequence()[0] == ord("A")


def test_no_overlap_unmapped_and_other_reference():   # :992-1048
    assert not both(AC, DC, rec(b"ACGT", Q(30), 100, C4), rec(b"ACGT", Q(20), 200, C4))[0]
    assert not both(AC, DC, rec(b"ACGT", Q(30), 100, C4, flags=R.UNMAPPED), rec(b"ACGT", Q(20), 100, C4))[0]
    assert not both(AC, DC, rec(b"ACGT", Q(30), 100, C4), rec(b"ACGT", Q(20), 100, C4, ref_id=1))[0]


def test_quality_cap_and_stats():                     # :1050-1106
    ret, a, _, _ = both(AC, DC, rec(b"ACGT", Q(50), 100, C4), rec(b"ACGT", Q(50), 100, C4))
    assert a.quals()[0] == 93
    _, _, _, o = both(AC, DC, rec(b"ACGT", Q(30), 100, C4), rec(b"ACGT", Q(20), 100, C4))
    assert o.stats() == (4, 4, 0, 4)
    _, _, _, o = both(AP, DC, rec(b"ACGT", Q(30), 100, C4), rec(b"TGCA", Q(20), 100, C4))
    assert o.stats() == (4, 0, 4, 8)


def test_offsets_clips_and_indels():                  # :1108-1187
    ret, a, _, o = both(AC, DC, rec(b"ACGT", Q(30), 100, C4), rec(b"GTAC", Q(20), 102, C4))
    assert ret and o.stats()[:2] == (2, 2) and list(a.quals()) == [30, 30, 50, 50]
    ret, a, _, o = both(AC, DC, rec(b"NNACGT", bytes([2, 2, 30, 30, 30, 30]), 100, [(S, 2), (M, 4)]),
                        rec(b"ACGT", Q(20), 100, C4))
    assert ret and o.stats()[:2] == (4, 4) and list(a.quals()) == [2, 2, 50, 50, 50, 50]
    ret, a, b, o = both(AC, DC, rec(b"ACTTGG", bytes([30] * 6), 100, [(M, 2), (I, 2), (M, 2)]),
                        rec(b"ACGG", Q(20), 100, C4))
    assert ret and o.stats()[:2] == (4, 4) and list(a.quals()) == [50, 50, 30, 30, 50, 50]
    ret, a, b, o = both(AC, DC, rec(b"ACGG", Q(30), 100, [(M, 2), (D, 2), (M, 2)]),
                        rec(b"ACTTGG", bytes([20] * 6), 100, [(M, 6)]))
    assert ret and o.stats()[:2] == (4, 4) and list(b.quals()) == [50, 50, 20, 20, 50, 50]


def grp(*specs):
    return [rec(b"ACGT", Q(q), 100, C4, flags=fl, name=nm) for nm, fl, q in specs]


def run_group(records):
    import fgumi_b200 as fg
    o = R.OverlappingOracle()
    want = [bytearray(r) for r in records]
    o.apply(want)
    got, stats = fg.apply_overlapping_consensus(records)
    assert [bytes(w) for w in want] == got and stats == o.stats()
    return [R.Rec(g).quals()[0] for g in got], o


def test_apply_group_pairing():                       # :1207-1400
    q, o = run_group(grp((b"rea", P | F1, 30), (b"rea", P | F2, 20)))
    assert q == [50, 50] and o.overlapping_bases > 0
    q, o = run_group(grp((b"rea", P | F1, 30), (b"reb", P | F2, 20)))
    assert q == [30, 20] and o.overlapping_bases == 0
    q, o = run_group(grp((b"rea", P | F2, 20), (b"rea", P | F1, 30)))
    assert q == [50, 50]
    q, o = run_group(grp((b"rea", P | F1, 30), (b"rea", P | F2, 20), (b"rea", P | F1 | R.SUPPLEMENTARY, 10)))
    assert q == [50, 50, 10]
    q, o = run_group(grp((b"rea", P | F1, 30), (b"rea", P | F2, 20), (b"rea", P | F1 | R.SECONDARY, 10)))
    assert q == [50, 50, 10]
    # a second primary R1 with the same name replaces the first in the pairing (map insert semantics)
    q, o = run_group(grp((b"rea", P | F1, 30), (b"rea", P | F2, 20), (b"rea", P | F1, 7)))
    assert q == [30, 27, 27]


def random_cigar(rng, L):
    ops, used = [], 0
    if rng.random() < 0.3:
        n = int(rng.integers(1, 5)); ops.append((S, n)); used += n
    while used < L:
        n = int(min(L - used, rng.integers(3, 25)))
        ops.append((M if rng.random() < 0.8 else (7 if rng.random() < 0.5 else 8), n)); used += n
        if used >= L:
            break
        r = rng.random()
        if r < 0.2:
            ops.append((D, int(rng.integers(1, 4))))
        elif r < 0.3:
            ops.append((N_, int(rng.integers(1, 30))))
        elif r < 0.5:
            n = int(min(L - used, rng.integers(1, 4))); ops.append((I, n)); used += n
    if rng.random() < 0.3 and ops[-1][0] in (M, 7, 8) and ops[-1][1] > 3:
        k, n = ops[-1]
        ops[-1] = (k, n - 2); ops.append((S, 2))
    if rng.random() < 0.1:
        ops = [(H, 3)] + ops
    return ops


def random_overlap_group(rng, n_pairs, L=40):
    recs = []
    for i in range(n_pairs):
        s1 = int(rng.integers(1, 200))
        s2 = s1 + int(rng.integers(-10, L + 15))
        s2 = max(s2, 1)
        for fl, st in ((P | F1 | MREV, s1), (P | F2 | REV, s2)):
            cig = random_cigar(rng, L)
            seq = np.frombuffer(b"ACGTN", np.uint8)[rng.choice(5, size=L, p=[.3, .3, .18, .18, .04])].tobytes()
            qual = rng.integers(2, 60, size=L).astype(np.uint8).tobytes()
            if rng.random() < 0.1:
                qual = bytes([30] * L)
            recs.append(rec(seq, qual, st, cig, flags=fl, name=b"p%d" % i))
    order = rng.permutation(len(recs))
    return [recs[k] for k in order]


@pytest.mark.parametrize("agree,disagree", [(AC, DC), (AM, DB), (AP, DL), (AC, DL)])
def test_product_matches_oracle_on_random_pairs(agree, disagree):
    import fgumi_b200 as fg
    rng = np.random.default_rng(5 + agree * 3 + disagree)
    tot = 0
    for _ in range(60):
        records = random_overlap_group(rng, int(rng.integers(1, 6)))
        o = R.OverlappingOracle(agree, disagree)
        want = [bytearray(r) for r in records]
        o.apply(want)
        got, stats = fg.apply_overlapping_consensus(records, agree, disagree)
        assert [bytes(w) for w in want] == got
        assert stats == o.stats()
        tot += o.overlapping_bases
    assert tot > 2000


@pytest.mark.gpu
def test_simplex_caller_with_overlap_prepass():
    import fgumi_b200 as fg
    from tests import oracle_lib as O
    from tests.test_record_oracle_kat import vote_fn
    from tests.test_caller_parity import random_groups
    rng = np.random.default_rng(77)
    groups = random_groups(rng, 150)
    opts = fg.VanillaUmiConsensusOptions(min_reads=1, min_consensus_base_quality=2)
    c = fg.VanillaUmiConsensusCaller("fgumi", "A", opts, consensus_call_overlapping_bases=True)
    got = c.consensus_reads_batch(groups)
    st = c.statistics()
    c.close()
    ov = R.OverlappingOracle()
    oracle = R.VanillaCallerOracle("fgumi", "A", R.VanillaOptions(min_reads=1, min_consensus_base_quality=2),
                                   vote_fn, O.builder_call)
    want, count = bytearray(), 0
    for g in groups:
        recs = [bytearray(r) for r in g]
        ov.apply(recs)
        d, n = oracle.consensus_reads([bytes(r) for r in recs])
        want += d
        count += n
    assert got.count == count and got.data == bytes(want)
    assert (st["overlapping_bases"], st["overlap_bases_agreeing"], st["overlap_bases_disagreeing"],
            st["overlap_bases_corrected"]) == ov.stats()
    assert ov.overlapping_bases > 0
    # and the pre-pass really changes the output
    c = fg.VanillaUmiConsensusCaller("fgumi", "A", opts)
    plain = c.consensus_reads_batch(groups)
    c.close()
    assert plain.data != got.data


@pytest.mark.gpu
def test_duplex_caller_with_overlap_prepass():
    import fgumi_b200 as fg
    from tests import oracle_lib as O
    from tests.test_record_oracle_kat import vote_fn
    from tests.test_caller_parity import random_duplex_groups, duplex_job_fn
    rng = np.random.default_rng(78)
    groups = random_duplex_groups(rng, 120)
    c = fg.DuplexConsensusCaller("fgumi", "A", min_reads=(1, 1, 0), consensus_call_overlapping_bases=True)
    got = c.consensus_reads_batch(groups)
    st = c.statistics()
    c.close()
    ov = R.OverlappingOracle()
    oracle = R.DuplexCallerOracle("fgumi", "A", min_reads=(1, 1, 0), per_base=True, vote_fn=vote_fn,
                                  builder_fn=O.builder_call, duplex_job_fn=duplex_job_fn)
    want, count = bytearray(), 0
    for g in groups:
        recs = [bytearray(r) for r in g]
        ov.apply(recs)
        d, n = oracle.consensus_reads([bytes(r) for r in recs])
        want += d
        count += n
    assert got.count == count and got.data == bytes(want)
    assert st["overlapping_bases"] == ov.overlapping_bases and ov.overlapping_bases > 0
