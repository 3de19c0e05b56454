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
