"""Pins the oracle's restatement of the raw-BAM helpers on the path -- FR-pair detection and the
mate-overlap clip (crates/fgumi-raw-bam/src/overlap.rs), CIGAR arithmetic (cigar.rs) and CIGAR
simplification (noodles_compat.rs) -- against the reference's own unit tests for them -- and, where the product's host code exports the same
helper through the C-ABI (fgb_host_*), runs the same cases against the product.  CPU only."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import record_oracle as R           # noqa: E402
from tests.bam_builder import make_record, encode_op   # noqa: E402

P, F1, F2, REV, MREV, UNM, MUNM = (R.PAIRED, R.FIRST_SEGMENT, R.LAST_SEGMENT, R.REVERSE, R.MATE_REVERSE,
                                   R.UNMAPPED, R.MATE_UNMAPPED)
M, I, D, N_, S, H, PAD, EQ, X = range(9)


def ops(*cig):
    return [encode_op(k, n) for k, n in cig]


class _Oracle:
    """oracle/record_oracle.py"""
    is_fr_pair = staticmethod(R.is_fr_pair)
    num_bases_extending_past_mate = staticmethod(R.num_bases_extending_past_mate)
    clip_cigar_ops = staticmethod(R.clip_cigar_ops)
    read_pos_at_ref_pos = staticmethod(R.read_pos_at_ref_pos)
    simplify_cigar = staticmethod(R.simplify_cigar)


class _Product:
    """The same helpers of the product's host code (fgumi_b200/csrc/host/bam.h) through the C-ABI."""

    @staticmethod
    def _lib():
        import fgumi_b200 as fg
        return fg.lib.load()

    @staticmethod
    def _u32(v):
        import ctypes as C
        return (C.c_uint32 * max(len(v), 1))(*v)

    @classmethod
    def is_fr_pair(cls, r):
        return bool(cls._lib().fgb_host_is_fr_pair(bytes(r.b), len(r.b)))

    @classmethod
    def num_bases_extending_past_mate(cls, r):
        return cls._lib().fgb_host_num_bases_extending_past_mate(bytes(r.b), len(r.b))

    @classmethod
    def clip_cigar_ops(cls, o, clip, from_start):
        import ctypes as C
        out, n, rc = (C.c_uint32 * (len(o) + 2))(), C.c_uint32(), C.c_uint32()
        a = cls._u32(o)
        assert cls._lib().fgb_host_clip_cigar_ops(C.addressof(a), len(o), clip, int(from_start), C.addressof(out),
                                                  C.addressof(n), C.addressof(rc)) == 0
        return list(out[:n.value]), rc.value

    @classmethod
    def read_pos_at_ref_pos(cls, o, start, ref_pos, last):
        import ctypes as C
        out = C.c_uint64()
        a = cls._u32(o)
        ok = cls._lib().fgb_host_read_pos_at_ref_pos(C.addressof(a), len(o), start, ref_pos, int(last), C.addressof(out))
        return out.value if ok else None

    @classmethod
    def simplify_cigar(cls, o):
        import ctypes as C
        kinds, lens, n = (C.c_uint8 * max(len(o), 1))(), (C.c_uint32 * max(len(o), 1))(), C.c_uint32()
        a = cls._u32(o)
        assert cls._lib().fgb_host_simplify_cigar(C.addressof(a), len(o), C.addressof(kinds), C.addressof(lens),
                                                  C.addressof(n)) == 0
        return [(int(kinds[i]), int(lens[i])) for i in range(n.value)]


IMPLS = pytest.mark.parametrize("impl", [_Oracle, _Product], ids=["oracle", "product-host"])


def bam(tid, pos, flag, cigar, seq_len, mate_tid, mate_pos, tlen=0, mc=None):
    """make_bam_bytes / make_bam_bytes_with_tlen, raw-bam testutil.rs:187-259"""
    tags = [(b"MC", "Z", mc)] if mc is not None else []
    return R.Rec(make_record(name=b"rea", flags=flag, ref_id=tid, pos=pos, mapq=0, cigar=cigar, mate_ref_id=mate_tid,
                             mate_pos=mate_pos, tlen=tlen, seq=b"A" * seq_len, quals=[0] * seq_len, tags=tags))


@IMPLS
def test_is_fr_pair_raw(impl):                        # overlap.rs:260-405
    c10 = ops((M, 10))
    R_is_fr_pair = impl.is_fr_pair
    assert not R_is_fr_pair(bam(0, 100, 0, c10, 10, 0, 200))
    assert not R_is_fr_pair(bam(0, 100, P | UNM, c10, 10, 0, 200))
    assert not R_is_fr_pair(bam(0, 100, P | MUNM, c10, 10, -1, -1))
    assert not R_is_fr_pair(bam(0, 100, P | MREV, c10, 10, 1, 200))          # different references
    assert not R_is_fr_pair(bam(0, 100, P, c10, 10, 0, 200))                 # FF
    assert not R_is_fr_pair(bam(0, 100, P | REV | MREV, c10, 10, 0, 200))    # RR
    assert R_is_fr_pair(bam(0, 100, P | MREV, c10, 10, 0, 200, tlen=200))
    assert R_is_fr_pair(bam(0, 100, P | REV, c10, 10, 0, 100, tlen=-10))
    assert not R_is_fr_pair(bam(0, 200, P | MREV, c10, 10, 0, 100, tlen=-100))   # RF


def test_bases_past_and_before_ref_pos():             # overlap.rs:411-535
    past = lambda c, s, t: R._read_pos_at_ref(c, s, t, True)       # compute_bases_past_ref_pos
    before = lambda c, s, t: R._read_pos_at_ref(c, s, t, False)    # compute_bases_before_ref_pos
    c10, ins, dele, sc = ops((M, 10)), ops((M, 5), (I, 3), (M, 5)), ops((M, 5), (D, 3), (M, 5)), ops((S, 3), (M, 10))
    assert [past(c10, 100, t) for t in (105, 100, 109, 110)] == [6, 1, 10, 0]
    assert past(ins, 100, 107) == 11 and past(dele, 100, 106) == 0 and past(sc, 100, 102) == 6
    assert [before(c10, 100, t) for t in (105, 100, 110)] == [5, 0, 0]
    assert before(ins, 100, 107) == 10 and before(dele, 100, 106) == 0 and before(sc, 100, 102) == 5


@IMPLS
def test_num_bases_extending_past_mate_raw(impl):     # overlap.rs:540-780, 826-880
    f = impl.num_bases_extending_past_mate
    c10, c20 = ops((M, 10)), ops((M, 20))
    assert f(bam(0, 100, 0, c10, 10, 0, 200)) == 0
    assert f(bam(0, 100, P | UNM | MREV, c10, 10, 0, 200)) == 0
    assert f(bam(0, 100, P | MUNM | MREV, c10, 10, -1, -1)) == 0
    assert f(bam(0, 100, P, c10, 10, 0, 200, mc=b"10M")) == 0                 # same strand
    assert f(bam(0, 100, P | MREV, c10, 10, 1, 200, mc=b"10M")) == 0          # different references
    assert f(bam(0, 100, P | MREV, c10, 10, 0, 200, tlen=110)) == 0           # no MC tag
    assert f(bam(0, 100, P | MREV, c20, 20, 0, 105, tlen=20, mc=b"10M")) == 5       # forward read past the mate's end
    assert f(bam(0, 100, P | MREV, c10, 10, 0, 200, tlen=110, mc=b"10M")) == 0
    assert f(bam(0, 100, P | REV, c20, 20, 0, 105, mc=b"10M")) == 5                # reverse read before the mate's start
    assert f(bam(0, 200, P | REV, c10, 10, 0, 100, mc=b"10M")) == 0
    assert f(bam(0, 110, P | REV, ops((S, 3), (M, 10)), 13, 0, 105, mc=b"10M")) == 0
    assert f(bam(0, 100, P | MREV, ops((M, 10), (S, 3)), 13, 0, 200, tlen=110, mc=b"10M")) == 0
    # regression (SRR6109273 MI=807): opposite strands but RF orientation -> no clipping
    assert f(bam(0, 11_576_620, P | MREV | F1, ops((M, 145), (S, 124)), 269, 0, 11_576_412, tlen=-28, mc=b"87S182M")) == 0
    assert f(bam(0, 11_576_412, P | REV | F2, ops((S, 87), (M, 182)), 269, 0, 11_576_620, tlen=28, mc=b"145M124S")) == 0


@IMPLS
def test_soft_clip_gaps_reach_into_the_clip(impl):    # overlap.rs:105-134 (the saturating_sub arms), :787-822
    """The leading / trailing soft-clip counters skip hard clips; a gap smaller than the clip leaves
    the difference to be clipped."""
    f = impl.num_bases_extending_past_mate
    # reverse read 3H5S10M starting 2 bases after the mate's unclipped start: 5 - 2 = 3
    assert f(bam(0, 107, P | REV, ops((H, 3), (S, 5), (M, 10)), 15, 0, 105, mc=b"10M")) == 3
    # forward read 10M5S3H ending 2 bases before the mate's unclipped end: 5 - 2 = 3
    assert f(bam(0, 100, P | MREV, ops((M, 10), (S, 5), (H, 3)), 15, 0, 102, tlen=12, mc=b"10M")) == 3


@IMPLS
def test_clip_cigar_ops_raw(impl):                    # cigar.rs:1656-1893
    f = impl.clip_cigar_ops
    assert f(ops((M, 10)), 0, True) == (ops((M, 10)), 0)
    assert f([], 5, True) == ([], 0)
    assert f(ops((S, 5), (M, 10)), 3, True) == (ops((H, 3), (S, 2), (M, 10)), 0)        # upgrade path
    assert f(ops((M, 10), (S, 5)), 3, False) == (ops((M, 10), (S, 2), (H, 3)), 0)
    assert f(ops((M, 10)), 3, True) == (ops((H, 3), (M, 7)), 3)
    assert f(ops((M, 10)), 3, False) == (ops((M, 7), (H, 3)), 0)
    assert f(ops((S, 2), (M, 10)), 5, True) == (ops((H, 5), (M, 7)), 3)                 # past the existing clip
    assert f(ops((M, 10), (S, 2)), 5, False) == (ops((M, 7), (H, 5)), 0)
    assert f(ops((M, 10), (I, 3), (M, 5)), 10, True) == (ops((H, 10), (I, 3), (M, 5)), 10)
    assert f(ops((M, 5), (D, 2), (M, 10)), 5, True) == (ops((H, 5), (M, 10)), 7)        # deletion at the boundary
    assert f(ops((M, 10), (D, 2), (M, 5)), 5, False) == (ops((M, 10), (H, 5)), 0)
    assert f(ops((M, 10)), 4, True) == (ops((H, 4), (M, 6)), 4)
    assert f(ops((M, 10)), 4, False) == (ops((M, 6), (H, 4)), 0)
    assert f(ops((M, 5), (I, 3), (M, 5)), 6, True) == (ops((H, 8), (M, 5)), 5)          # insertion eaten whole
    assert f(ops((EQ, 5), (X, 3)), 4, True) == (ops((H, 4), (EQ, 1), (X, 3)), 4)
    assert f(ops((M, 10)), 10, True) == (ops((H, 10)), 10)
    assert f(ops((M, 10)), 10, False) == (ops((H, 10)), 0)
    cx = ops((S, 3), (M, 10), (I, 2), (M, 5), (S, 4))
    assert f(cx, 8, True) == (ops((H, 8), (M, 5), (I, 2), (M, 5), (S, 4)), 5)
    assert f(cx, 8, False) == (ops((S, 3), (M, 10), (I, 2), (M, 1), (H, 8)), 0)


def test_upgrade_and_edge_clipping_raw():             # cigar.rs:1896-2045
    up = R._upgrade_clipping
    assert up(ops((S, 5), (M, 10)), 3, True) == (ops((H, 3), (S, 2), (M, 10)), 0)
    assert up(ops((H, 2), (S, 5), (M, 10)), 4, True) == (ops((H, 4), (S, 3), (M, 10)), 0)
    assert up(ops((S, 5), (M, 10)), 5, True) == (ops((H, 5), (M, 10)), 0)
    assert up(ops((H, 3), (S, 5), (M, 10)), 3, True) == (ops((H, 3), (S, 5), (M, 10)), 0)
    assert up(ops((M, 10), (S, 5)), 3, False) == (ops((M, 10), (S, 2), (H, 3)), 0)
    assert up(ops((M, 10), (S, 5), (H, 2)), 5, False) == (ops((M, 10), (S, 2), (H, 5)), 0)
    assert up(ops((M, 10), (S, 5)), 5, False) == (ops((M, 10), (H, 5)), 0)
    assert R._clip_start(ops((H, 3), (M, 10)), 2) == (ops((H, 5), (M, 8)), 2)
    assert R._clip_end(ops((M, 10), (H, 3)), 2) == (ops((M, 8), (H, 5)), 0)
    assert R._clip_start(ops((I, 3), (M, 10)), 1) == (ops((H, 3), (M, 10)), 0)
    assert R._clip_end(ops((M, 10), (I, 3)), 1) == (ops((M, 10), (H, 3)), 0)


@IMPLS
def test_read_pos_at_ref_pos_raw(impl):               # cigar.rs:1286-1345, 2050-2086
    f = impl.read_pos_at_ref_pos
    c10 = ops((M, 10))
    assert [f(c10, 100, t, False) for t in (100, 102, 105, 109)] == [1, 3, 6, 10]
    assert f(c10, 100, 99, False) is None and f(c10, 100, 110, False) is None
    dele = ops((M, 5), (D, 3), (M, 5))
    assert f(dele, 100, 106, False) is None and f(dele, 100, 106, True) == 5
    ins = ops((M, 5), (I, 3), (M, 5))
    assert f(ins, 100, 104, False) == 5 and f(ins, 100, 105, False) == 9
    assert f(ops((S, 3), (M, 10)), 100, 100, False) == 4
    assert f(ops((D, 2), (M, 5)), 100, 100, True) == 1           # deletion first: no earlier base, reports 1


def test_mc_string_parsers_and_reference_length():    # cigar.rs:1366-1460
    assert [R._parse_leading_clips(s) for s in ("5S10M", "3H5S10M", "10M", "5S3H")] == [5, 8, 0, 8]
    g = R._parse_ref_len_and_trailing_clips
    assert g("10M5S") == (10, 5) and g("5S10M2I3D5M3S2H") == (18, 5) and g("5S3H") == (0, 0) and g("10=3X") == (13, 0)
    assert R.reference_length(ops((M, 50))) == 50
    assert R.reference_length(ops((M, 10), (D, 3), (M, 5), (N_, 2), (M, 8))) == 28
    assert R.reference_length(ops((M, 10), (I, 5), (M, 10))) == 20


@IMPLS
def test_simplify_cigar_from_raw(impl):               # noodles_compat.rs:294-392
    f = impl.simplify_cigar
    assert f(ops((S, 5), (M, 10), (I, 3), (M, 5), (S, 4))) == [(M, 15), (I, 3), (M, 9)]
    assert f(ops((EQ, 5), (X, 3), (D, 2), (EQ, 4))) == [(M, 8), (D, 2), (M, 4)]
    assert f([]) == []
    assert f(ops((H, 5), (M, 10), (H, 5))) == [(M, 20)]
    assert f(ops((H, 2), (S, 3), (M, 5), (I, 2), (M, 3), (D, 1), (M, 4), (S, 4), (H, 1))) == \
        [(M, 10), (I, 2), (M, 3), (D, 1), (M, 9)]
    assert f(ops((M, 5), (N_, 3), (M, 5))) == [(M, 5), (N_, 3), (M, 5)]
    assert f(ops((S, 10))) == [(M, 10)]


def test_sequence_packing():                          # sequence.rs:530-570, 684-740
    pk = R.pack_sequence
    assert pk(b"") == b"" and pk(b"ACGT") == bytes([0x12, 0x48]) and pk(b"ACG") == bytes([0x12, 0x40])
    assert pk(b"T") == bytes([0x80]) and pk(b"NN") == bytes([0xFF])
    assert pk(b"ACGTACGTACGTACGTA") == bytes([0x12, 0x48] * 4 + [0x10])
    for seq in (b"ACGT", b"ACG", b"NNNN", b"T"):
        r = R.Rec(make_record(name=b"rd", flags=0, ref_id=0, pos=0, cigar=[], seq=seq, quals=[0] * len(seq)))
        assert bytes(r.sequence()) == seq


def _random_cigar(rng, allow_clips=True):
    body = []
    for _ in range(int(rng.integers(1, 7))):
        k = int(rng.choice([M, M, M, I, D, N_, EQ, X]))
        if body and body[-1][0] == k:
            continue
        body.append((k, int(rng.integers(1, 12))))
    if not any(k in (M, EQ, X) for k, _ in body):
        body.append((M, int(rng.integers(1, 12))))
    lead, trail = [], []
    if allow_clips:
        if rng.random() < 0.3:
            lead.append((H, int(rng.integers(1, 6))))
        if rng.random() < 0.4:
            lead.append((S, int(rng.integers(1, 8))))
        if rng.random() < 0.4:
            trail.append((S, int(rng.integers(1, 8))))
        if rng.random() < 0.3:
            trail.append((H, int(rng.integers(1, 6))))
    return lead + body + trail


def test_product_helpers_match_oracle_on_random_cigars():
    """Differential check of the product's host helpers against the oracle over a few thousand random
    CIGARs (clips, indels, skips, =/X) and mate geometries."""
    import numpy as np
    rng = np.random.default_rng(515)
    for trial in range(2500):
        cig = _random_cigar(rng)
        o = ops(*cig)
        qlen = sum(n for k, n in cig if k in (M, I, S, EQ, X))
        for clip in (0, 1, int(rng.integers(0, qlen + 3)), qlen):
            for from_start in (True, False):
                got, want = _Product.clip_cigar_ops(o, clip, from_start), R.clip_cigar_ops(o, clip, from_start)
                assert (list(got[0]), got[1]) == (list(want[0]), want[1]), (cig, clip, from_start)
        start = int(rng.integers(1, 500))
        rlen = R.reference_length(o)
        for ref_pos in (start - 1, start, start + rlen // 2, start + max(rlen - 1, 0), start + rlen):
            for last in (False, True):
                assert _Product.read_pos_at_ref_pos(o, start, ref_pos, last) == R.read_pos_at_ref_pos(o, start, ref_pos, last), (cig, start, ref_pos, last)
        assert _Product.simplify_cigar(o) == R.simplify_cigar(o)
        # mate-overlap clip on a record with this CIGAR
        rev = bool(rng.random() < 0.5)
        pos = int(rng.integers(50, 400))
        mpos = pos + int(rng.integers(-60, 60))
        mate_cig = "".join("%d%s" % (n, "MIDNSHP=X"[k]) for k, n in _random_cigar(rng))
        flag = P | (REV if rev else MREV) | (F1 if rng.random() < 0.5 else F2)
        if rng.random() < 0.1:
            flag ^= MREV                                  # sometimes not an FR pair
        tlen = int(rng.integers(-300, 300))
        r = bam(0, pos, flag, o, qlen, 0, mpos, tlen=tlen, mc=mate_cig.encode() if rng.random() < 0.9 else None)
        assert _Product.is_fr_pair(r) == R.is_fr_pair(r), (cig, pos, mpos, tlen, flag)
        assert _Product.num_bases_extending_past_mate(r) == R.num_bases_extending_past_mate(r), (cig, mate_cig, pos, mpos, tlen, flag)
