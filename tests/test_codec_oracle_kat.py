"""Pins oracle/record_oracle.py::CodecCallerOracle (+ the C++ oracle's codec_job) against the
reference's own CODEC integration tests (crates/fgumi-consensus/src/codec_caller.rs:2190-2975,
ports of fgbio's CodecConsensusCallerTest) and its helper KATs (:3103-3190, :3258-3440, :3748-3820,
:3916-4008).  CPU only."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import record_oracle as R           # noqa: E402
from tests import oracle_lib as O               # noqa: E402
from tests.bam_builder import make_record, encode_op, parse_records   # noqa: E402
from tests.test_record_oracle_kat import vote_fn   # noqa: E402

P, F1, F2, REV, MREV = R.PAIRED, R.FIRST_SEGMENT, R.LAST_SEGMENT, R.REVERSE, R.MATE_REVERSE
M, I, D, N_, S, H = 0, 1, 2, 3, 4, 5

# A 400-base stand-in for the reference's REF_BASES constant (codec_caller.rs:1997); the tests
# only depend on both mates being cut from the same template.
REF = np.frombuffer(b"ACGT", np.uint8)[np.random.default_rng(1234).integers(0, 4, size=400)].tobytes()


def codec_job_fn(ss1, ss2, r1_neg, r2_neg, cons_len, o):
    """codec_caller.rs:746-766 through the C++ oracle (codec_job) + the python strand views."""
    L = O.load()
    a8 = lambda x: np.frombuffer(bytes(x), np.uint8).copy()
    a16 = lambda x: np.array(list(x), np.uint16)
    A = [a8(ss1[0]), a8(ss1[1]), a16(ss1[2]), a16(ss1[3])]
    B = [a8(ss2[0]), a8(ss2[1]), a16(ss2[2]), a16(ss2[3])]
    ob, oq = np.zeros(cons_len, np.uint8), np.zeros(cons_len, np.uint8)
    od, oe = np.zeros(cons_len, np.uint16), np.zeros(cons_len, np.uint16)
    nb, nd = C.c_uint64(), C.c_uint64()
    max_dis = (1 << 63) if o.max_dis is None else o.max_dis
    st = L.orc_codec_job(*[x.ctypes.data for x in A], len(ss1[0]), *[x.ctypes.data for x in B], len(ss2[0]),
                         int(r1_neg), int(r2_neg), cons_len,
                         -1 if o.ss_qual is None else o.ss_qual, -1 if o.outer_qual is None else o.outer_qual,
                         o.outer_len, max_dis, o.max_rate,
                         ob.ctypes.data, oq.ctypes.data, od.ctypes.data, oe.ctypes.data,
                         C.addressof(nb), C.addressof(nd))
    _, _, ac, bc = R.codec_strands(ss1, ss2, r1_neg, r2_neg, cons_len)
    return dict(status=st, duplex_bases=nb.value, disagreements=nd.value,
                consensus=(bytes(ob), bytes(oq), list(od), list(oe)), ss_for_ac=ac, ss_for_bc=bc)


def make_codec_oracle(**kw):
    return R.CodecCallerOracle("codec", "RG1", vote_fn=vote_fn, builder_fn=O.builder_call,
                               codec_job_fn=codec_job_fn, **kw)


def create_fr_pair(name, start1, start2, qual, cigar1, cigar2, mi=b"hi", rx=b"ACC-TGA",
                   rev1=False, rev2=True, seq_edit=None, ref_oriented=False):
    """create_fr_pair, codec_caller.rs:2013-2184 (1-based starts, no MC tag).  Like the reference
    helper it stores reverse-strand mates reverse-complemented (:2088-2094), so the two strands of
    its pairs DISAGREE over most of the overlap; `ref_oriented=True` stores SEQ the way an aligner
    does (reference orientation) and gives agreeing strands."""
    def ref_len(cig):
        return sum(n for k, n in cig if k in (M, D, N_, 7, 8))

    def get_seq(start, cig):
        seq, rp = bytearray(), start - 1
        for k, n in cig:
            if k in (M, 7, 8):
                seq += REF[rp:rp + n]
                seq += b"A" * (n - len(REF[rp:rp + n]))
                rp += n
            elif k in (I, S):
                seq += b"A" * n
            elif k in (D, N_):
                rp += n
        return bytes(seq)
    s1, s2 = get_seq(start1, cigar1), get_seq(start2, cigar2)
    if seq_edit:
        s1, s2 = seq_edit(s1, s2)
    if rev1 and not ref_oriented:
        s1 = bytes(R.reverse_complement(s1))
    if rev2 and not ref_oriented:
        s2 = bytes(R.reverse_complement(s2))
    tlen = (start2 + ref_len(cigar2) - start1) if start1 <= start2 else -(start1 + ref_len(cigar1) - start2)
    tags = [(b"MI", "Z", mi)] + ([(b"RX", "Z", rx)] if rx is not None else [])
    enc = lambda cig: [encode_op(k, n) for k, n in cig]
    r1 = make_record(name=name, flags=P | 0x2 | F1 | (REV if rev1 else 0) | (MREV if rev2 else 0), ref_id=0,
                     pos=start1 - 1, cigar=enc(cigar1), seq=s1, quals=[qual] * len(s1), mate_ref_id=0,
                     mate_pos=start2 - 1, tlen=tlen, tags=tags)
    r2 = make_record(name=name, flags=P | 0x2 | F2 | (REV if rev2 else 0) | (MREV if rev1 else 0), ref_id=0,
                     pos=start2 - 1, cigar=enc(cigar2), seq=s2, quals=[qual] * len(s2), mate_ref_id=0,
                     mate_pos=start1 - 1, tlen=-tlen, tags=tags)
    return [r1, r2]


def simple_pair(**kw):
    return create_fr_pair(b"read1", 1, 11, 35, [(M, 30)], [(M, 30)], **kw)


def test_simple_reads_make_a_40bp_consensus():       # :2190-2231, :2476-2520
    o = make_codec_oracle()
    data, n = o.consensus_reads(simple_pair())
    assert n == 1
    rec = parse_records(data)[0]
    assert rec["name"] == b"codec:hi" and len(rec["bases"]) == 40
    assert rec["tags"][b"RX"] == b"ACC-TGA" and rec["flags"] == 0x4
    assert o.consensus_reads_generated == 1 and o.total_input_reads == 2
    # aligner-style SEQ: both strands agree with the template over 11..30, and the single-strand
    # flanks carry the one strand that covers them
    data, n = make_codec_oracle().consensus_reads(simple_pair(ref_oriented=True))
    rec = parse_records(data)[0]
    assert n == 1 and rec["bases"] == REF[:40]
    assert rec["tags"][b"cD"] == 2 and rec["tags"][b"cM"] == 1 and rec["tags"][b"aD"] == 1


def test_r1_deletion_outside_overlap():              # :2235-2269
    data, n = make_codec_oracle().consensus_reads(
        create_fr_pair(b"read1", 1, 13, 35, [(M, 5), (D, 2), (M, 25)], [(M, 30)]))
    assert n == 1 and len(parse_records(data)[0]["bases"]) > 0


def test_rf_pair_is_not_emitted():                   # :2273-2301
    data, n = make_codec_oracle().consensus_reads(
        create_fr_pair(b"read1", 100, 135, 35, [(M, 30)], [(M, 30)], rev1=True, rev2=False))
    assert n == 0 and data == b""


def test_insufficient_reads():                       # :2305-2363
    o = make_codec_oracle(min_reads_per_strand=2)
    assert o.consensus_reads(simple_pair())[1] == 0
    assert o.rejections == {"InsufficientReads": 2}
    assert make_codec_oracle(min_reads_per_strand=1).consensus_reads(simple_pair())[1] == 1


def test_insufficient_overlap():                     # :2367-2430
    assert make_codec_oracle(min_duplex_length=20).consensus_reads(simple_pair())[1] == 1
    o = make_codec_oracle(min_duplex_length=21)
    assert o.consensus_reads(simple_pair())[1] == 0
    assert o.rejections == {"InsufficientOverlap": 2}


def test_unmapped_mate():                            # :2434-2472
    r = simple_pair()
    r2 = bytearray(r[1])
    fl = int.from_bytes(r2[14:16], "little") | R.UNMAPPED
    r2[14:16] = fl.to_bytes(2, "little")
    assert make_codec_oracle().consensus_reads([r[0], bytes(r2)])[1] == 0


def test_high_disagreement_is_dropped():             # :2524-2589
    assert make_codec_oracle(max_dis=100, max_rate=1.0).consensus_reads(simple_pair())[1] == 1
    # the reference's pair agrees everywhere, and still fails max 5 / 5%: single-strand positions
    # count.  Ours must fail the same way when the strands really disagree in the overlap ...
    def edit(s1, s2):
        s2 = bytearray(s2)
        for i in range(8):
            s2[i] = ord("A") if s2[i] != ord("A") else ord("C")
        return s1, bytes(s2)
    o = make_codec_oracle(max_dis=5, max_rate=0.05)
    assert o.consensus_reads(simple_pair(seq_edit=edit, ref_oriented=True))[1] == 0
    assert o.duplex_disagreements == 8 and o.duplex_bases == 20
    o = make_codec_oracle(max_dis=5, max_rate=0.05)
    assert o.consensus_reads(simple_pair(ref_oriented=True))[1] == 1 and o.duplex_disagreements == 0
    # ... and for the unedited pair as the reference test asserts (is_err -> dropped)
    o = make_codec_oracle(max_dis=5, max_rate=0.05)
    got = o.consensus_reads(simple_pair())[1]
    assert got == 0, "reference: strict settings reject the simple pair (codec_caller.rs:2586-2588)"


def test_r2_deletion_outside_overlap_gives_40bp():   # :2593-2631
    data, n = make_codec_oracle().consensus_reads(
        create_fr_pair(b"read1", 1, 11, 35, [(M, 30)], [(M, 25), (D, 5), (M, 5)]))
    assert n == 1 and len(parse_records(data)[0]["bases"]) == 40


def test_soft_clipping_outside_overlap_gives_45bp():  # :2635-2681
    data, n = make_codec_oracle().consensus_reads(
        create_fr_pair(b"read1", 1, 11, 35, [(S, 5), (M, 25)], [(M, 25), (S, 5)]))
    assert n == 1 and len(parse_records(data)[0]["bases"]) == 45


def test_both_soft_clipped_same_end():               # :2685-2721
    data, n = make_codec_oracle().consensus_reads(
        create_fr_pair(b"read1", 1, 1, 35, [(S, 5), (M, 25)], [(S, 5), (M, 25)]))
    assert n == 1 and len(parse_records(data)[0]["bases"]) > 0


def test_chimeric_pair():                            # :2725-2764
    r = create_fr_pair(b"read1", 100, 135, 35, [(M, 30)], [(M, 30)])
    r1, r2 = bytearray(r[0]), bytearray(r[1])
    r1[0:4] = (2).to_bytes(4, "little")              # ref_id
    r2[20:24] = (2).to_bytes(4, "little")            # mate_ref_id
    assert make_codec_oracle().consensus_reads([bytes(r1), bytes(r2)])[1] == 0


def test_r1_end_in_indel_of_r2():                    # :2768-2803
    o = make_codec_oracle()
    assert o.consensus_reads(create_fr_pair(b"read1", 1, 11, 35, [(M, 30)], [(M, 19), (D, 2), (M, 11)]))[1] == 0
    assert o.rejections == {"IndelErrorBetweenStrands": 2}


def test_mask_end_qualities():                       # :2807-2855
    data, n = make_codec_oracle(outer_len=7, outer_qual=5).consensus_reads(
        create_fr_pair(b"read1", 1, 1, 90, [(M, 50)], [(M, 50)]))
    assert n == 1
    q = parse_records(data)[0]["quals"]
    assert all(x <= 5 for x in q[:7]) and all(x <= 5 for x in q[-7:]) and max(q[7:-7]) > 5


def test_mask_single_stranded_regions():             # :2859-2911
    data, n = make_codec_oracle(ss_qual=4).consensus_reads(
        create_fr_pair(b"read1", 1, 20, 90, [(M, 30)], [(M, 30)]))
    assert n == 1
    q = parse_records(data)[0]["quals"]
    assert len(q) == 49 and any(x <= 4 for x in q)       # what the reference test asserts
    # The reference masks only where a strand holds an UPPERCASE 'N' (NO_CALL_BASE, :1195-1198);
    # pad_consensus fills with lowercase 'n' (:980-1023), so padded flanks keep their quality.
    data, n = make_codec_oracle(ss_qual=4).consensus_reads(
        create_fr_pair(b"read1", 1, 20, 90, [(M, 30)], [(M, 30)], ref_oriented=True))
    q = parse_records(data)[0]["quals"]
    assert n == 1 and min(q) > 4
    def with_n(s1, s2):
        return s1[:25] + b"N" + s1[26:], s2
    data, n = make_codec_oracle(ss_qual=4).consensus_reads(
        create_fr_pair(b"read1", 1, 20, 90, [(M, 30)], [(M, 30)], ref_oriented=True, seq_edit=with_n))
    rec = parse_records(data)[0]
    assert rec["bases"][25:26] == b"N" and rec["quals"][25] == 2


def test_fragments_are_counted_and_skipped():        # :564-566
    frag = make_record(name=b"f", flags=0, pos=10, seq=b"ACGTACGT", tags=[(b"MI", "Z", b"hi")])
    o = make_codec_oracle()
    assert o.consensus_reads([frag] + simple_pair())[1] == 1
    assert o.rejections == {"FragmentRead": 1} and o.reads_filtered == 1


# ---- helper KATs ----
def enc(cig):
    return [encode_op(k, n) for k, n in cig]


def test_build_clipped_info():                       # :3103-3190
    f = R.CodecCallerOracle._clipped_info
    fwd = R.Rec(make_record(name=b"r", flags=P | F1, pos=99, cigar=enc([(M, 10)]), seq=b"ACGTACGTAC"))
    ci = f(fwd, 0, 0)
    assert (ci.clip_amount, ci.clipped_seq_len, ci.adjusted_pos, ci.clip_from_start) == (0, 10, 100, False)
    ci = f(fwd, 0, 3)
    assert (ci.clipped_seq_len, ci.adjusted_pos, ci.clip_from_start) == (7, 100, False)
    assert R.reference_length(ci.clipped_cigar) == 7
    rev = R.Rec(make_record(name=b"r", flags=P | F1 | REV, pos=99, cigar=enc([(M, 10)]), seq=b"ACGTACGTAC"))
    ci = f(rev, 0, 3)
    assert (ci.clipped_seq_len, ci.adjusted_pos, ci.clip_from_start) == (7, 103, True)


def test_read_pos_at_ref_pos():                      # :3258-3370, :3916-3950
    f = R.read_pos_at_ref_pos
    assert f(enc([(M, 10)]), 100, 100, False) == 1
    assert f(enc([(M, 10)]), 100, 105, False) == 6
    assert f(enc([(M, 10)]), 100, 109, False) == 10
    assert f(enc([(M, 10)]), 100, 99, False) is None and f(enc([(M, 10)]), 100, 110, False) is None
    cig = enc([(M, 5), (I, 2), (M, 5)])
    assert f(cig, 100, 104, False) == 5 and f(cig, 100, 105, False) == 8
    cig = enc([(M, 5), (D, 2), (M, 5)])
    assert f(cig, 100, 104, False) == 5 and f(cig, 100, 107, False) == 6
    assert f(cig, 100, 105, False) is None and f(cig, 100, 105, True) == 5
    cig = enc([(S, 3), (M, 7)])
    assert f(cig, 100, 100, False) == 4
    cig = enc([(H, 5), (M, 10)])
    assert f(cig, 100, 100, False) == 1


# ---- more helper KATs (create_test_paired_read based) ----
def paired_read(name, seq, qual, first, rev, mate_rev, start, cig):
    """create_test_paired_read, codec_caller.rs:1618-1704: 1-based start, 200 bp insert, MI tag."""
    flags = P | (F1 if first else F2) | (REV if rev else 0) | (MREV if mate_rev else 0)
    ref_len = sum(n for k, n in cig if k in (M, D, N_, 7, 8))
    if rev:
        mate1, tlen = max(start - 200 + ref_len, 1), -200
    else:
        mate1, tlen = max(start + 200 - ref_len, 1), 200
    return make_record(name=name, flags=flags, pos=start - 1, mapq=60, cigar=enc(cig), seq=seq,
                       quals=bytes(qual), mate_ref_id=0, mate_pos=mate1 - 1, tlen=tlen,
                       tags=[(b"MI", "Z", b"UMI123")])


def infos(recs):
    """The test-only wrappers of codec_caller.rs:1477-1530 build ClippedRecordInfo with clip 0."""
    return [R.CodecCallerOracle._clipped_info(R.Rec(b), i, 0) for i, b in enumerate(recs)]


def test_is_fr_pair():                                # :1707-1751
    assert R.is_fr_pair(R.Rec(paired_read(b"read1", b"ACGT", b"####", True, False, True, 100, [(M, 4)])))
    assert R.is_fr_pair(R.Rec(paired_read(b"read1", b"ACGT", b"####", True, True, False, 100, [(M, 4)])))
    assert not R.is_fr_pair(R.Rec(paired_read(b"read1", b"ACGT", b"####", True, False, False, 100, [(M, 4)])))


def test_filter_to_most_common_alignment():           # :1754-1990
    def run(reads):
        o = make_codec_oracle()
        inf = infos(reads)
        kept = o._filter(inf)
        return o, [reads[k.raw_idx] for k in kept]
    mk = lambda nm, cig, seq=b"ACGT", rev=False: paired_read(nm, seq, b"#" * len(seq), True, rev, not rev, 100, cig)
    # three 4M reads and one with a deletion -> the odd one is dropped (:1754-1814)
    o, kept = run([mk(b"r1", [(M, 4)]), mk(b"r2", [(M, 4)]), mk(b"r3", [(M, 4)]),
                   mk(b"r4", [(M, 2), (D, 1), (M, 2)])])
    assert len(kept) == 3 and o.rejections == {"MinorityAlignment": 1}
    # 2 x 4M against 2 x 3M1D1M: the tie goes to the smaller CIGAR, 3 < 4 (:1817-1908)
    o, kept = run([mk(b"r1_4M", [(M, 4)]), mk(b"r2_4M", [(M, 4)]),
                   mk(b"r3_del", [(M, 3), (D, 1), (M, 1)]), mk(b"r4_del", [(M, 3), (D, 1), (M, 1)])])
    assert [R.Rec(k).name for k in kept] == [b"r3_del", b"r4_del"]
    assert o.rejections == {"MinorityAlignment": 2} and o.reads_filtered == 2
    # negative-strand reads compare their REVERSED cigars: 3M1I2M -> 2M1I3M wins (:1910-1990)
    a, b = [(M, 3), (I, 1), (M, 2)], [(M, 2), (I, 1), (M, 3)]
    o, kept = run([mk(b"r1_groupA", a, b"ACGTAC", True), mk(b"r2_groupA", a, b"ACGTAC", True),
                   mk(b"r3_groupB", b, b"ACGTAC", True), mk(b"r4_groupB", b, b"ACGTAC", True)])
    assert [R.Rec(k).name for k in kept] == [b"r1_groupA", b"r2_groupA"]
    assert o.rejections == {"MinorityAlignment": 2}
    # the same two groups on the forward strand: group B's 2M.. is the smaller one
    o, kept = run([mk(b"r1_groupA", a, b"ACGTAC"), mk(b"r2_groupA", a, b"ACGTAC"),
                   mk(b"r3_groupB", b, b"ACGTAC"), mk(b"r4_groupB", b, b"ACGTAC")])
    assert [R.Rec(k).name for k in kept] == [b"r3_groupB", b"r4_groupB"]


def test_to_source_read_for_codec():                  # :2980-3100
    row = R.CodecCallerOracle._source_row
    ci = R.CodecCallerOracle._clipped_info
    fwd = R.Rec(paired_read(b"read1", b"ACGT", [30, 31, 32, 33], True, False, True, 100, [(M, 4)]))
    assert row(fwd, ci(fwd, 0, 0)) == (b"ACGT", bytes([30, 31, 32, 33]))
    rev = R.Rec(paired_read(b"read1", b"ACGT", [30, 31, 32, 33], True, True, False, 100, [(M, 4)]))
    assert row(rev, ci(rev, 1, 0)) == (b"ACGT", bytes([33, 32, 31, 30]))
    rev = R.Rec(paired_read(b"read1", b"AACC", [10, 20, 30, 40], True, True, False, 100, [(M, 4)]))
    assert row(rev, ci(rev, 2, 0)) == (b"GGTT", bytes([40, 30, 20, 10]))
    # clipping at or beyond the read length leaves an empty source read, from either end
    for clip, from_start in ((10, False), (10, True), (4, False)):
        inf = ci(fwd, 0, 0)
        inf.clip_amount, inf.clip_from_start = clip, from_start
        assert row(fwd, inf) == (b"", b"")


def test_check_overlap_phase():                       # :3371-3437, :3952-4000
    f = R.CodecCallerOracle.check_overlap_phase
    a30, q30 = b"A" * 30, [30] * 30
    r1 = paired_read(b"read1", a30, q30, True, False, True, 100, [(M, 30)])
    r2 = paired_read(b"read2", a30, q30, False, True, False, 110, [(M, 30)])
    assert f(*infos([r1, r2]), 110, 129)
    r2 = paired_read(b"read2", a30, q30, False, True, False, 100, [(M, 15), (D, 5), (M, 15)])
    assert f(*infos([r1, r2]), 100, 129) is False      # runs; 1-1 != 30-25
    a50 = b"A" * 50
    r1 = make_record(name=b"r1", flags=0, pos=99, cigar=enc([(M, 50)]), seq=a50, quals=bytes([30] * 50))
    r2 = make_record(name=b"r2", flags=0, pos=119, cigar=enc([(M, 50)]), seq=a50, quals=bytes([30] * 50))
    assert f(*infos([r1, r2]), 120, 149)
    r2 = make_record(name=b"r2", flags=0, pos=119, cigar=enc([(M, 15), (D, 2), (M, 33)]), seq=b"A" * 48,
                     quals=bytes([30] * 48))
    assert not f(*infos([r1, r2]), 120, 149)


def test_compute_codec_consensus_length():            # :3748-3817
    f = R.CodecCallerOracle.compute_consensus_length
    a30, q30 = b"A" * 30, [30] * 30
    pos = paired_read(b"pos", a30, q30, True, False, True, 100, [(M, 30)])
    neg = paired_read(b"neg", a30, q30, False, True, False, 110, [(M, 30)])
    assert f(*infos([pos, neg]), 129) == 40            # 30 + 30 - 20
    pos = paired_read(b"pos", a30, q30, True, False, True, 100, [(M, 15), (D, 5), (M, 15)])
    neg = paired_read(b"neg", a30, q30, False, True, False, 100, [(M, 30)])
    assert f(*infos([pos, neg]), 116) is None          # 116 falls inside the deletion


def test_pad_and_reverse_complement_ss():             # :3439-3523, :3632-3659
    ss = (b"ACGT", bytes([30, 31, 32, 33]), [5, 6, 7, 8], [0, 1, 0, 1])
    p = R._pad_ss(ss, 8, False)
    assert p[0] == b"ACGTnnnn" and p[1][4:] == bytes(4) and p[2][4:] == [0] * 4 and p[3][4:] == [0] * 4
    p = R._pad_ss(ss, 8, True)
    assert p[0] == b"nnnnACGT" and p[1][:4] == bytes(4) and p[2][:4] == [0] * 4 and p[2][4:] == [5, 6, 7, 8]
    assert R._pad_ss(ss, 4, False) == ss and R._pad_ss(ss, 2, False) == ss
    rc = R._rc_ss((b"ACGT", bytes([10, 20, 30, 40]), [1, 2, 3, 4], [5, 6, 7, 8]))
    assert rc == (b"ACGT", bytes([40, 30, 20, 10]), [4, 3, 2, 1], [8, 7, 6, 5])
    assert R._rc_ss((b"AACC", bytes(4), [0] * 4, [0] * 4))[0] == b"GGTT"


def test_mask_consensus_quals_query_based():          # :3526-3575 through the C++ oracle
    L = O.load()
    cons_b = np.frombuffer(b"ACGT", np.uint8).copy()
    cons_q = np.full(4, 30, np.uint8)
    r1 = np.frombuffer(b"NCGT", np.uint8).copy()
    r2 = np.frombuffer(b"ACGN", np.uint8).copy()
    L.orc_codec_mask(cons_b.ctypes.data, cons_q.ctypes.data, 4, r1.ctypes.data, r2.ctypes.data, 10, -1, 5)
    assert cons_q.tolist() == [10, 30, 30, 10]
    # lowercase padding is NOT a no-call for this check (NO_CALL_BASE is b'N', :1195-1198)
    cons_q[:] = 30
    r1 = np.frombuffer(b"nCGT", np.uint8).copy()
    L.orc_codec_mask(cons_b.ctypes.data, cons_q.ctypes.data, 4, r1.ctypes.data, r2.ctypes.data, 10, -1, 5)
    assert cons_q.tolist() == [30, 30, 30, 10]
    # outer-bases masking takes the min with the outer quality on both ends (:1206-1211)
    cons_b = np.frombuffer(b"ACGTACGT", np.uint8).copy()
    cons_q = np.array([30, 3, 30, 30, 30, 30, 30, 30], np.uint8)
    r = np.frombuffer(b"ACGTACGT", np.uint8).copy()
    L.orc_codec_mask(cons_b.ctypes.data, cons_q.ctypes.data, 8, r.ctypes.data, r.ctypes.data, -1, 5, 2)
    assert cons_q.tolist() == [5, 3, 30, 30, 30, 30, 5, 5]
