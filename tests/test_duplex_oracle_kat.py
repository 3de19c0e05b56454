"""Pins oracle/record_oracle.py::DuplexCallerOracle against the reference's caller-level duplex tests
(crates/fgumi-consensus/src/duplex_caller.rs:3307-3857, ports of fgbio's DuplexConsensusCallerTest)
and its min-reads gate KATs (:4408-4479, :4531-4596).  CPU only."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import record_oracle as R           # noqa: E402
from tests import oracle_lib as O               # noqa: E402
from tests.bam_builder import make_record, encode_op, parse_records   # noqa: E402
from tests.test_record_oracle_kat import vote_fn   # noqa: E402
from tests.test_caller_parity import duplex_job_fn   # noqa: E402

P, F1, F2, REV, MREV = R.PAIRED, R.FIRST_SEGMENT, R.LAST_SEGMENT, R.REVERSE, R.MATE_REVERSE
C10 = [encode_op(0, 10)]
Q20 = [20] * 10


def rec(name, flags, pos, mpos, seq, mi, cigar=C10, extra=()):
    tags = [(b"MI", "Z", mi), (b"RG", "Z", b"A")] + [(t, "Z", v) for t, v in extra]
    return make_record(name=name, flags=flags, ref_id=0, pos=pos, mapq=60, cigar=cigar, mate_ref_id=0,
                       mate_pos=mpos, seq=seq, quals=Q20, tags=tags)


# ab_r1 / ab_r2 / ba_r1 / ba_r2, duplex_caller.rs:3176-3303
def ab_r1(n, seq, mi, **k): return rec(n, P | F1 | MREV, 99, 199, seq, mi, **k)
def ab_r2(n, seq, mi, **k): return rec(n, P | F2 | REV, 199, 99, seq, mi, **k)
def ba_r1(n, seq, mi, **k): return rec(n, P | F1 | REV, 199, 99, seq, mi, **k)
def ba_r2(n, seq, mi, **k): return rec(n, P | F2 | MREV, 99, 199, seq, mi, **k)


def caller(min_reads=(1, 1, 1), cell_tag=None):
    return R.DuplexCallerOracle("consensus", "RG1", min_reads=min_reads, per_base=False, cell_tag=cell_tag,
                                vote_fn=vote_fn, builder_fn=O.builder_call, duplex_job_fn=duplex_job_fn)


A10, C10S = b"A" * 10, b"C" * 10


def molecule(extra_a=(), extra_b=()):
    return [ab_r1(b"q1", A10, b"foo/A", extra=extra_a), ab_r2(b"q1", C10S, b"foo/A", extra=extra_a),
            ba_r1(b"q2", C10S, b"foo/B", extra=extra_b), ba_r2(b"q2", A10, b"foo/B", extra=extra_b)]


def test_fragments_give_nothing():                    # :3307-3362
    frags = [make_record(name=b"frag%d" % i, flags=0, ref_id=0, pos=0, cigar=C10, seq=A10, quals=[30] * 10,
                         tags=[(b"MI", "Z", mi)]) for i, mi in enumerate((b"foo/A", b"foo/B"))]
    assert caller().consensus_reads(frags)[1] == 0


def test_simple_double_stranded_consensus():          # :3368-3405
    data, n = caller().consensus_reads(molecule())
    assert n == 2
    recs = parse_records(data)
    assert [r["flags"] for r in recs] == [0x4D, 0x8D] and recs[0]["name"] == b"consensus:foo"
    assert recs[0]["tags"][b"aD"] == 1 and recs[0]["tags"][b"bD"] == 1


def test_cell_barcode_is_preserved():                 # :3409-3455
    data, n = caller(cell_tag=b"CB").consensus_reads(molecule([(b"CB", b"ACGT")], [(b"CB", b"ACGT")]))
    assert n == 2 and all(r["tags"][b"CB"] == b"ACGT" for r in parse_records(data))


def test_absent_umi_on_either_side():                 # :3459-3551
    data, n = caller().consensus_reads(molecule([(b"RX", b"ACT-")], [(b"RX", b"-ACT")]))
    assert n == 2 and all(r["tags"][b"RX"] == b"ACT-" for r in parse_records(data))
    data, n = caller().consensus_reads(molecule([(b"RX", b"-ACT")], [(b"RX", b"ACT-")]))
    assert n == 2 and all(r["tags"][b"RX"] == b"-ACT" for r in parse_records(data))


def strand(kind, k):
    out = []
    for i in range(1, k + 1):
        nm = b"q%d" % i
        out += [ab_r1(nm, A10, b"foo/A"), ab_r2(nm, C10S, b"foo/A")] if kind == "A" else \
               [ba_r1(nm, C10S, b"foo/B"), ba_r2(nm, A10, b"foo/B")]
    return out


def test_single_strand_consensus():                   # :3554-3705
    data, n = caller((1, 1, 0)).consensus_reads(strand("A", 3))
    assert n == 2 and all(r["tags"][b"aD"] == 3 and r["tags"].get(b"bD", 0) == 0 for r in parse_records(data))
    data, n = caller((1, 1, 0)).consensus_reads(strand("B", 3))
    assert n == 2 and all(r["tags"][b"aD"] == 3 for r in parse_records(data))      # BA sits in the ab slot
    assert caller((1, 1, 1)).consensus_reads(strand("A", 3))[1] == 0


def test_min_reads_hard_filter_after_alignment_filtering():   # :3707-3797
    reads = []
    for i in (1, 2, 3):
        reads += [ab_r1(b"ab%d" % i, A10, b"foo/A"), ab_r2(b"ab%d" % i, C10S, b"foo/A")]
    for i in (4, 5):
        reads += [ba_r1(b"ba%d" % i, C10S, b"foo/B"), ba_r2(b"ba%d" % i, A10, b"foo/B")]
    assert caller((3, 3, 3)).consensus_reads(reads)[1] == 0
    assert caller((2, 2, 2)).consensus_reads(reads)[1] == 2
    cig = [encode_op(0, 5), encode_op(2, 1), encode_op(0, 5)]                    # 5M1D5M
    more = reads + [ba_r1(b"ba6", C10S, b"foo/B", cigar=cig), ba_r2(b"ba6", A10, b"foo/B")]
    assert caller((3, 3, 3)).consensus_reads(more)[1] == 0


def test_min_reads_gates():                           # :4408-4479, :4531-4596
    # has_minimum_number_of_reads / duplex_consensus_has_minimum_reads reduce to the same rule on the
    # pair (num_a, num_b): xy = max, yx = min, total = xy + yx
    ok = lambda mr, na, nb: caller(mr)._min_ok(na, nb)
    assert ok((4, 2, 2), 3, 2) and not ok((6, 2, 2), 3, 2) and not ok((4, 4, 1), 3, 2) and not ok((4, 2, 3), 3, 2)
    assert ok((3, 2, 1), 3, 1) and not ok((3, 2, 2), 3, 1)
    assert ok((8, 5, 3), 6, 4) and not ok((11, 5, 3), 6, 4)
    assert ok((5, 5, 0), 6, 0) and not ok((5, 5, 1), 6, 0)


# ---- the reference's second duplex test module (duplex_caller.rs:4652-6100) -------------------------
def _flag_rec(flags, mi=None, name=b"r"):
    tags = [(b"MI", "Z", mi)] if mi is not None else []
    return R.Rec(make_record(name=name, flags=flags, seq=b"ACGT", quals=[30] * 4, tags=tags))


def test_min_reads_gate_counts_paired_r1_only():       # :5175-5338
    r1, r2, frag = _flag_rec(P | F1), _flag_rec(P | F2), _flag_rec(0)
    D = R.DuplexCallerOracle
    assert D._r1(r1) and not D._r2(r1) and D._r2(r2) and not D._r1(r2) and not D._r1(frag) and not D._r2(frag)

    def has(a, b, total, xy, yx):                      # has_minimum_number_of_reads, :697-749
        n = lambda rs: sum(1 for r in rs if D._r1(r))
        return caller((total, xy, yx))._min_ok(n(a), n(b))
    assert has([r1] * 3, [r1] * 3, 4, 2, 2) and not has([r1] * 3, [r1] * 3, 8, 2, 2)
    assert has([r1] * 3, [], 1, 1, 0) and not has([r1] * 3, [], 1, 1, 1)
    assert has([r1] * 3 + [frag] * 2, [r1] * 3, 6, 3, 3) and not has([r1] * 3 + [frag] * 2, [r1] * 3, 7, 3, 3)
    assert has([r1, r2], [r1], 2, 1, 1) and not has([r1, r2], [r1], 3, 1, 1)


def test_duplex_consensus_has_minimum_reads():         # :5340-5429
    ss = lambda depths: R.SsCons(b"ACG"[:len(depths)], bytes([30] * len(depths)), depths, [0] * len(depths), [])
    d = R.DuplexCons(b"ACG", bytes([60] * 3), [0] * 3, ss([5, 4, 3]), ss([3, 2, 1]))
    assert caller((8, 5, 3))._cons_min_ok(d) and not caller((8, 5, 4))._cons_min_ok(d)
    d = R.DuplexCons(b"A", bytes([30]), [0], ss([5]), None)
    assert caller((1, 1, 0))._cons_min_ok(d) and not caller((1, 1, 1))._cons_min_ok(d)


def test_partition_records_by_strand():                # :5431-5530
    part = R.DuplexCallerOracle.partition_records_by_strand
    recs = [_flag_rec(0, b"UMI1/A", b"r1"), _flag_rec(0, b"UMI1/A", b"r2"), _flag_rec(0, b"UMI1/B", b"r3")]
    base, a, b = part(recs)
    assert base == "UMI1" and [r.name for r in a] == [b"r1", b"r2"] and [r.name for r in b] == [b"r3"]
    assert part([]) == (None, [], [])
    import pytest
    with pytest.raises(ValueError):
        part([_flag_rec(0, b"UMI1")])
    base, a, b = part(recs[:1])
    assert base == "UMI1" and len(a) == 1 and b == []


def test_are_all_same_strand():                        # :4843-4872
    f = R.DuplexCallerOracle.are_all_same_strand
    fwd, rev = _flag_rec(0), _flag_rec(REV)
    assert f([fwd, fwd]) and f([rev, rev]) and not f([fwd, rev]) and f([]) and f([fwd])


def test_duplex_read_into_tags_and_flags():            # :4652-4731, :4783-4841, :4874-5013
    ss = lambda b, q, d, e: R.SsCons(b, bytes(q), d, e, [])
    ab = ss(b"ACGT", [30] * 4, [5] * 4, [0, 1, 0, 1])
    d = R.DuplexCons(b"ACGT", bytes([30] * 4), [0, 1, 0, 1], ab, None)
    o = caller()                                       # per_base off, no cell tag
    r = parse_records(o._record(d, "R1", "UMI123", [], [], True, None))[0]
    assert b"CB" not in r["tags"]
    assert r["tags"][b"bD"] == 0 and r["tags"][b"bM"] == 0 and abs(r["tags"][b"bE"]) < 1e-3
    assert r["tags"][b"aD"] == 5 and r["tags"][b"aM"] == 5
    assert not any(t in r["tags"] for t in (b"ad", b"ae", b"ac", b"bd", b"be", b"bc"))
    flags = {t: parse_records(o._record(d, t, "UMI123", [], [], True, None))[0]["flags"] for t in ("R1", "R2", "Fragment")}
    assert flags["R1"] & P and flags["R1"] & F1 and not flags["R1"] & F2
    assert flags["R2"] & P and flags["R2"] & F2 and not flags["R2"] & F1
    assert not flags["Fragment"] & P
    ab = ss(b"ACGT", [30, 31, 32, 33], [5, 6, 7, 8], [0, 1, 0, 2])
    ba = ss(b"TGCA", [25, 26, 27, 28], [3, 4, 5, 6], [1, 0, 1, 0])
    d = R.DuplexCons(b"ACGT", bytes([30] * 4), [1, 1, 1, 2], ab, ba)
    o = R.DuplexCallerOracle("consensus", "RG1", per_base=True, vote_fn=vote_fn, builder_fn=O.builder_call,
                             duplex_job_fn=duplex_job_fn)
    r = parse_records(o._record(d, "R1", "UMI123", [], [], True, None))[0]
    assert all(t in r["tags"] for t in (b"ad", b"ae", b"ac", b"bd", b"be", b"bc"))
    assert list(r["tags"][b"ad"]) == [5, 6, 7, 8] and r["tags"][b"bc"] == b"TGCA"


def test_b_only_molecule_maps_ba_r2_to_r1():           # :5989-6095
    reads = []
    for i in (1, 2, 3):
        nm = b"q%d" % i
        reads.append(make_record(name=nm, flags=P | F1 | REV, ref_id=0, pos=99, mate_ref_id=0, mate_pos=99,
                                 cigar=C10, seq=b"G" * 10, quals=[30] * 10, tags=[(b"MI", "Z", b"foo/B"), (b"RG", "Z", b"A")]))
        reads.append(make_record(name=nm, flags=P | F2 | MREV, ref_id=0, pos=99, mate_ref_id=0, mate_pos=99,
                                 cigar=C10, seq=b"A" * 10, quals=[30] * 10, tags=[(b"MI", "Z", b"foo/B"), (b"RG", "Z", b"A")]))
    data, n = caller((1, 1, 0)).consensus_reads(reads)
    assert n == 2
    recs = parse_records(data)
    r1 = next(r for r in recs if r["flags"] & F1)
    r2 = next(r for r in recs if r["flags"] & F2)
    assert r1["bases"] == b"A" * 10 and r2["bases"] == b"C" * 10


def test_duplex_consensus_equal_quality_disagreement_and_n():   # :5080-5110, :5112-5140
    ss = lambda b, q, d: R.SsCons(b, bytes(q), d, [0] * len(b), [])
    st, ob, oq, oe = duplex_job_fn(ss(b"A", [30], [5]), ss(b"T", [30], [5]), [])
    assert (ob, list(oq)) == (b"N", [2])
    st, ob, oq, oe = duplex_job_fn(ss(b"NA", [30, 30], [5, 5]), ss(b"AN", [30, 30], [5, 5]), [])
    assert ob == b"NN"
