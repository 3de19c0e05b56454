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
