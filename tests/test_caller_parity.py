"""GPU parity of the record-level simplex caller (fgb_caller_* — C++ host + CUDA vote) against the
Python record oracle: the ConsensusOutput byte stream and the statistics must be identical."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import record_oracle as R          # noqa: E402
from tests import oracle_lib as O              # noqa: E402
from tests.bam_builder import make_record, encode_op, parse_records   # noqa: E402
from tests.test_record_oracle_kat import vote_fn   # noqa: E402

pytestmark = pytest.mark.gpu

P, F1, F2, REV, MREV = R.PAIRED, R.FIRST_SEGMENT, R.LAST_SEGMENT, R.REVERSE, R.MATE_REVERSE
ACGT = np.frombuffer(b"ACGT", np.uint8)
COMP = bytes.maketrans(b"ACGTN", b"TGCAN")


def random_groups(rng, n_groups, L=60):
    groups = []
    for g in range(n_groups):
        kind = rng.choice(["frag", "pair", "pair", "mixed"])
        depth = int(rng.integers(1, 7))
        insert = int(rng.integers(L // 2, 3 * L))
        tmpl = ACGT[rng.integers(0, 4, size=max(insert, L) + 20)].tobytes()
        umi = b"%d" % g
        rx_len = int(rng.integers(4, 9))
        true_rx = ACGT[rng.integers(0, 4, size=rx_len)].tobytes() + b"-" + ACGT[rng.integers(0, 4, size=rx_len)].tobytes()
        recs = []
        for d in range(depth):
            def mutate(seq):
                s = np.frombuffer(seq, np.uint8).copy()
                m = rng.random(len(s)) < 0.03
                s[m] = ACGT[rng.integers(0, 4, size=int(m.sum()))]
                if rng.random() < 0.1:
                    s[rng.integers(0, len(s))] = ord("N")
                return s.tobytes()

            def quals(n):
                q = rng.integers(2, 41, size=n)
                if rng.random() < 0.2:
                    q[-int(rng.integers(1, 8)):] = 3          # low-quality tail -> masked, trimmed
                return q.astype(np.uint8).tobytes()
            rx = bytearray(true_rx)
            if rng.random() < 0.2:
                rx[int(rng.integers(0, rx_len))] = ACGT[rng.integers(0, 4)]
            tags = [(b"MI", "Z", umi)]
            if rng.random() < 0.9:
                tags.append((b"RX", "Z", bytes(rx)))
            if rng.random() < 0.5:
                tags.append((b"CB", "Z", b"CELL%d" % (g % 3)))
            name = b"g%d_%d" % (g, d)
            if kind == "frag" or (kind == "mixed" and d % 3 == 0):
                rev = rng.random() < 0.3
                seq = mutate(tmpl[:L])
                cig = None
                r = rng.random()
                if r < 0.15:
                    cig = "%dS%dM" % (5, L - 5)
                elif r < 0.25:
                    cig = "%dM1I%dM" % (20, L - 21)
                elif r < 0.30:
                    cig = "%dM2D%dM" % (25, L - 25)
                fl = REV if rev else 0
                if rng.random() < 0.05:
                    fl |= R.SECONDARY
                recs.append(make_record(name=name, flags=fl, pos=1000, seq=seq, quals=quals(L), cigar=cig, tags=tags))
            else:
                # FR pair on a template of length `insert`; short inserts make the reads run past the mate
                l1 = min(L, len(tmpl)); l2 = L
                r1_seq = mutate(tmpl[:l1])
                end = max(insert, L)
                r2_fwd = tmpl[end - l2:end]
                r2_seq = mutate(r2_fwd)                     # stored in reference orientation
                p1, p2 = 1000, 1000 + end - l2
                tl = end if insert >= L else insert
                if insert < L:                               # reads longer than the insert
                    p2 = 1000 - (L - insert) if rng.random() < 0.5 else 1000
                    tl = max(p1 + l1, p2 + l2) - min(p1, p2)
                    if p2 < p1:
                        tl = (p2 + l2) - p1 if (p2 + l2) > p1 else 1
                t1 = tags + [(b"MC", "Z", b"%dM" % l2)]
                t2 = tags + [(b"MC", "Z", b"%dM" % l1)]
                swap = rng.random() < 0.3                    # R1 on the reverse strand
                fa, fb = (F2, F1) if swap else (F1, F2)
                if rng.random() < 0.9:
                    recs.append(make_record(name=name, flags=P | fa | MREV, pos=p1, mate_ref_id=0, mate_pos=p2,
                                            tlen=tl, seq=r1_seq, quals=quals(l1), tags=t1))
                if rng.random() < 0.9:
                    recs.append(make_record(name=name, flags=P | fb | REV, pos=p2, mate_ref_id=0, mate_pos=p1,
                                            tlen=-tl, seq=r2_seq, quals=quals(l2), tags=t2))
        if recs:
            groups.append(recs)
    return groups


@pytest.mark.parametrize("min_reads,min_q,per_base,trim,min_input_q", [(1, 2, True, False, 10), (2, 2, True, False, 10),
                                                                       (2, 40, False, True, 20), (3, 10, True, True, 5)])
def test_simplex_caller_bytes_match_oracle(min_reads, min_q, per_base, trim, min_input_q):
    import fgumi_b200 as fg
    rng = np.random.default_rng(100 + min_reads * 10 + min_q)
    groups = random_groups(rng, 250)
    opt = fg.VanillaUmiConsensusOptions(min_reads=min_reads, min_consensus_base_quality=min_q,
                                        produce_per_base_tags=per_base, trim=trim,
                                        min_input_base_quality=min_input_q)
    caller = fg.VanillaUmiConsensusCaller("fgumi", "A", opt, device=0, cell_tag=b"CB")
    got = caller.consensus_reads_batch(groups)
    gstats = caller.statistics()
    caller.close()

    oracle = R.VanillaCallerOracle("fgumi", "A", R.VanillaOptions(
        min_reads=min_reads, min_consensus_base_quality=min_q, produce_per_base_tags=per_base, trim=trim,
        min_input_base_quality=min_input_q, cell_tag=b"CB"), vote_fn, O.builder_call)
    want, count = bytearray(), 0
    for g in groups:
        d, n = oracle.consensus_reads(g)
        want += d
        count += n
    assert got.count == count
    if got.data != bytes(want):
        a, b = parse_records(got.data), parse_records(bytes(want))
        for i, (x, y) in enumerate(zip(a, b)):
            assert x == y, (i, x, y)
    assert got.data == bytes(want)
    s = oracle.stats
    assert gstats["total_reads"] == s.total_reads and gstats["consensus_reads"] == s.consensus_reads
    assert gstats["filtered_reads"] == s.filtered_reads
    for k in ("InsufficientReads", "SecondaryOrSupplementary", "ZeroLengthAfterTrimming",
              "MinorityAlignment", "OrphanConsensus"):
        assert gstats[k] == s.rejections.get(k, 0), k
    assert count > 50      # the test actually produced consensus reads
    kinds = {r["flags"] for r in parse_records(got.data)}
    if min_reads <= 2:
        assert {0x4, 0x4D, 0x8D} <= kinds     # fragment, R1 and R2 consensus records all present


def test_reference_known_answers_through_the_gpu_caller():
    """vanilla_caller.rs:2083-2112, 2396-2464, 3813-3855 through the product caller."""
    import fgumi_b200 as fg
    opt = fg.VanillaUmiConsensusOptions(min_reads=1, min_consensus_base_quality=0)
    c = fg.VanillaUmiConsensusCaller("consensus", "A", opt)
    mk = lambda n, b, q: make_record(name=n, ref_id=0, pos=99, seq=b, quals=q, tags=[(b"MI", "Z", b"UMI1")])
    out = c.consensus_reads_batch([[mk(b"r1", b"GATTACA", [10] * 7), mk(b"r2", b"GATTACA", [10] * 7)]])
    rec, = parse_records(out.data)
    assert rec["bases"] == b"GATTACA" and all(q > 10 for q in rec["quals"]) and rec["name"] == b"consensus:UMI1"
    c.close()
    opt = fg.VanillaUmiConsensusOptions(min_reads=1, min_input_base_quality=2)
    c = fg.VanillaUmiConsensusCaller("consensus", "A", opt)
    reads = [mk(b"r%d" % i, b"A" * 10, [30] * 10) for i in range(3)] + [mk(b"r4", b"AAAAACAAAA", [30] * 10)]
    rec, = parse_records(c.consensus_reads_batch([reads]).data)
    assert rec["bases"] == b"A" * 10 and rec["tags"][b"cD"] == 4 and rec["tags"][b"cM"] == 4
    assert abs(rec["tags"][b"cE"] - 0.025) < 1e-6 and rec["tags"][b"ce"] == [0] * 5 + [1] + [0] * 4
    c.close()


def test_missing_umi_tag_is_an_error():
    import fgumi_b200 as fg
    c = fg.VanillaUmiConsensusCaller("x", "A", fg.VanillaUmiConsensusOptions(min_reads=1))
    with pytest.raises(fg.lib.FgbError) as ei:
        c.add_group([make_record(seq=b"ACGT")])
    assert ei.value.status == fg.lib.FGB_ERR_MISSING_TAG
    c.close()


# ---------------------------------------------------------------------------------------------------
# duplex
# ---------------------------------------------------------------------------------------------------
def duplex_job_fn(a, b, src_rows):
    """duplex_consensus arms through the C++ oracle."""
    import ctypes as C
    L = O.load()
    arr8 = lambda x: np.frombuffer(bytes(x), np.uint8).copy()
    arr16 = lambda x: np.array(list(x), np.uint16)
    A = [arr8(a.bases), arr8(a.quals), arr16(a.depths), arr16(a.errors)]
    B = [arr8(b.bases), arr8(b.quals), arr16(b.depths), arr16(b.errors)]
    keep = [np.frombuffer(r[0], np.uint8).copy() for r in src_rows]
    ptrs = (C.c_void_p * max(len(keep), 1))(*[k.ctypes.data for k in keep])
    lens = (C.c_size_t * max(len(keep), 1))(*[len(k) for k in keep])
    cap = max(len(a.bases), len(b.bases), 1)
    ob = np.zeros(cap, np.uint8); oq = np.zeros(cap, np.uint8); oe = np.zeros(cap, np.uint16)
    n = C.c_size_t()
    st = L.orc_duplex_job(*[x.ctypes.data for x in A], len(a.bases), *[x.ctypes.data for x in B], len(b.bases),
                          ptrs, lens, len(keep), ob.ctypes.data, oq.ctypes.data, oe.ctypes.data, C.addressof(n))
    return st, bytes(ob[:n.value]), bytes(oq[:n.value]), list(oe[:n.value])


def random_duplex_groups(rng, n_groups, L=50):
    groups = []
    for g in range(n_groups):
        tmpl = ACGT[rng.integers(0, 4, size=3 * L)].tobytes()
        insert = int(rng.integers(L, 3 * L))
        n_ab, n_ba = int(rng.integers(0, 5)), int(rng.integers(0, 5))
        if rng.random() < 0.1:
            n_ba = 0
        rx_a, rx_b = b"AAC-GGT", b"GGT-AAC"
        recs = []
        collide = rng.random() < 0.05
        for strand, n in (("A", n_ab), ("B", n_ba)):
            for d in range(n):
                def mutate(seq, all_n=False):
                    s = np.frombuffer(seq, np.uint8).copy()
                    m = rng.random(len(s)) < 0.04
                    s[m] = ACGT[rng.integers(0, 4, size=int(m.sum()))]
                    if all_n:
                        s[:] = ord("N")
                    return s.tobytes()
                q = lambda: rng.integers(8, 41, size=L).astype(np.uint8).tobytes()
                mi = b"%d/%s" % (g, strand.encode())
                name = b"m%d%s%d" % (g, strand.encode(), d)
                fwd, rev = tmpl[:L], tmpl[insert - L:insert]
                p1, p2 = 500, 500 + insert - L
                # AB: R1 forward at p1, R2 reverse at p2.  BA: R1 reverse at p2, R2 forward at p1.
                r1_fwd = (strand == "A")
                if collide and strand == "B" and d == 0:
                    r1_fwd = True
                dead = (g % 13 == 5 and strand == "B")       # a strand whose bases are all N
                tags = [(b"MI", "Z", mi), (b"RX", "Z", rx_a if strand == "A" else rx_b)]
                if rng.random() < 0.5:
                    tags.append((b"CB", "Z", b"CELL"))
                for which in ("R1", "R2"):
                    if rng.random() < 0.08:
                        continue                               # drop a mate now and then
                    is_fwd = r1_fwd if which == "R1" else not r1_fwd
                    fl = P | (F1 if which == "R1" else F2) | (MREV if is_fwd else REV)
                    seq = mutate(fwd if is_fwd else rev, dead)
                    pos, mpos = (p1, p2) if is_fwd else (p2, p1)
                    recs.append(make_record(name=name, flags=fl, pos=pos, mate_ref_id=0, mate_pos=mpos,
                                            tlen=insert if is_fwd else -insert, seq=seq, quals=q(),
                                            tags=tags + [(b"MC", "Z", b"%dM" % L)]))
        if recs:
            groups.append(recs)
    return groups


@pytest.mark.parametrize("min_reads,per_base", [((1, 1, 0), True), ((1, 1, 1), True), ((3, 2, 1), False),
                                                ((2, 1, 0), True)])
def test_duplex_caller_bytes_match_oracle(min_reads, per_base):
    import fgumi_b200 as fg
    rng = np.random.default_rng(7 + sum(min_reads))
    groups = random_duplex_groups(rng, 220)
    caller = fg.DuplexConsensusCaller("fgumi", "A", min_reads=min_reads, produce_per_base_tags=per_base,
                                      device=0, cell_tag=b"CB")
    got = caller.consensus_reads_batch(groups)
    gstats = caller.statistics()
    caller.close()
    oracle = R.DuplexCallerOracle("fgumi", "A", min_reads=min_reads, per_base=per_base, cell_tag=b"CB",
                                  vote_fn=vote_fn, builder_fn=O.builder_call, duplex_job_fn=duplex_job_fn)
    want, count = bytearray(), 0
    for g in groups:
        d, n = oracle.consensus_reads(g)
        want += d
        count += n
    assert got.count == count
    if got.data != bytes(want):
        a, b = parse_records(got.data), parse_records(bytes(want))
        for i, (x, y) in enumerate(zip(a, b)):
            assert x == y, (i, x, y)
    assert got.data == bytes(want)
    s = oracle.stats
    assert gstats["total_reads"] == s.total_reads and gstats["consensus_reads"] == s.consensus_reads
    assert gstats["filtered_reads"] == s.filtered_reads
    assert gstats["InsufficientReads"] == s.rejections.get("InsufficientReads", 0)
    assert gstats["PotentialCollision"] == s.rejections.get("PotentialCollision", 0)
    assert count >= 40 and s.rejections.get("PotentialCollision", 0) > 0
    recs = parse_records(got.data)
    assert recs[0]["tag_order"][:3] in ([b"MI", b"CB", b"RG"], [b"MI", b"RG", b"aD"])
    if min_reads[2] == 0:
        assert any(b"bc" not in r["tags"] for r in recs)      # single-strand molecules were emitted


def test_duplex_known_answer_end_to_end():
    """duplex_caller.rs:3368-3405: a 4-read molecule (AB pair + BA pair) gives 2 duplex records."""
    import fgumi_b200 as fg
    L = 20
    seq, rc = b"ACGTACGTACGTACGTACGT", b"ACGTACGTACGTACGTACGT"
    def rd(name, mi, first, fwd):
        fl = P | (F1 if first else F2) | (MREV if fwd else REV)
        return make_record(name=name, flags=fl, pos=100 if fwd else 200, mate_ref_id=0, mate_pos=200 if fwd else 100,
                           tlen=120 if fwd else -120, seq=seq, quals=[30] * L,
                           tags=[(b"MI", "Z", mi), (b"RX", "Z", b"ACG-TTA"), (b"MC", "Z", b"20M")])
    grp = [rd(b"a", b"1/A", True, True), rd(b"a", b"1/A", False, False),
           rd(b"b", b"1/B", True, False), rd(b"b", b"1/B", False, True)]
    c = fg.DuplexConsensusCaller("fgumi", "A", min_reads=(1, 1, 1))
    out = c.consensus_reads_batch([grp])
    st = c.statistics()
    c.close()
    recs = parse_records(out.data)
    assert out.count == 2 and [r["flags"] for r in recs] == [0x4D, 0x8D]
    assert recs[0]["name"] == b"fgumi:1" and recs[0]["tags"][b"MI"] == b"1"
    assert recs[0]["tags"][b"aD"] == 1 and recs[0]["tags"][b"bD"] == 1 and recs[0]["tags"][b"cD"] == 2
    assert st["consensus_reads"] == 1 and st["total_reads"] == 4


# ------------------------------------------------------------------------------------------------
# CODEC (codec_caller.rs:531-814)
# ------------------------------------------------------------------------------------------------
def random_codec_groups(rng, n_groups, L=40):
    from tests.test_codec_oracle_kat import create_fr_pair, M, I, D, S
    groups = []
    shapes1 = [[(M, L)], [(S, 4), (M, L - 4)], [(M, 6), (D, 2), (M, L - 6)], [(M, 5), (I, 1), (M, L - 6)]]
    shapes2 = [[(M, L)], [(M, L - 5), (S, 5)], [(M, L - 6), (D, 3), (M, 6)], [(M, L - 8), (I, 2), (M, 6)]]
    for g in range(n_groups):
        start1 = int(rng.integers(1, 150))
        start2 = start1 + int(rng.integers(0, L + 6))        # sometimes past R1's end: no overlap
        c1 = shapes1[int(rng.integers(0, 4))] if rng.random() < 0.4 else shapes1[0]
        c2 = shapes2[int(rng.integers(0, 4))] if rng.random() < 0.4 else shapes2[0]
        r1_rev = rng.random() < 0.4                            # R1 on the negative strand
        n_pairs = int(rng.integers(1, 5))
        noisy = rng.random() < 0.25
        recs = []
        for d in range(n_pairs):
            cc1, cc2 = c1, c2
            if n_pairs > 2 and d == n_pairs - 1 and rng.random() < 0.5:
                cc1 = shapes1[2] if c1 is not shapes1[2] else shapes1[0]   # a minority alignment
            def edit(s1, s2):
                out = []
                for s in (s1, s2):
                    a = np.frombuffer(s, np.uint8).copy()
                    m = rng.random(len(a)) < (0.25 if noisy else 0.03)
                    a[m] = ACGT[rng.integers(0, 4, size=int(m.sum()))]
                    if rng.random() < 0.1:
                        a[int(rng.integers(0, len(a)))] = ord("N")
                    out.append(a.tobytes())
                return out
            mi = None if g % 17 == 3 else b"%d" % g
            if r1_rev:      # R1 is the right-hand, reverse mate
                pair = create_fr_pair(b"p%d_%d" % (g, d), start2, start1, 30, cc2, cc1, mi=mi or b"x",
                                      rx=b"ACG-TTA" if rng.random() < 0.9 else None, rev1=True, rev2=False,
                                      seq_edit=edit, ref_oriented=True)
            else:
                pair = create_fr_pair(b"p%d_%d" % (g, d), start1, start2, 30, cc1, cc2, mi=mi or b"x",
                                      rx=b"ACG-TTA" if rng.random() < 0.9 else None, seq_edit=edit,
                                      ref_oriented=True)
            # random per-base qualities + optional MC / CB tags are appended by re-building the records
            out = []
            for raw in pair:
                r = R.Rec(raw)
                q = rng.integers(5, 41, size=r.l_seq).astype(np.uint8).tobytes()
                raw = bytearray(raw)
                qo = r.seq_offset() + (r.l_seq + 1) // 2
                raw[qo:qo + r.l_seq] = q
                if mi is None:                                  # strip the MI tag (first tag: MI:Z:x\0)
                    t = qo + r.l_seq
                    assert raw[t:t + 3] == b"MIZ"
                    del raw[t:t + 5]
                if rng.random() < 0.5:
                    raw += b"CBZCELL%d\0" % (g % 3)
                out.append(bytes(raw))
            if rng.random() < 0.07:
                out = out[:1]                                    # a lone mate
            recs += out
        if rng.random() < 0.1:
            recs.append(make_record(name=b"frag%d" % g, flags=0, pos=start1, seq=b"ACGTACGTAC",
                                    tags=[(b"MI", "Z", b"%d" % g)]))
        if rng.random() < 0.05:
            recs = list(reversed(recs))
        groups.append(recs)
    return groups


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [
    dict(),
    dict(min_reads_per_strand=2, per_base=True),
    dict(min_duplex_length=20, ss_qual=4, outer_qual=7, outer_len=6, per_base=True),
    dict(max_dis=3, max_rate=0.2, per_base=False),
])
def test_codec_caller_bytes_match_oracle(kw):
    import fgumi_b200 as fg
    from tests.test_codec_oracle_kat import codec_job_fn
    rng = np.random.default_rng(100 + len(kw))
    groups = random_codec_groups(rng, 260)
    caller = fg.CodecConsensusCaller(
        "codec", "RG1", min_reads_per_strand=kw.get("min_reads_per_strand", 1),
        min_duplex_length=kw.get("min_duplex_length", 1), single_strand_qual=kw.get("ss_qual"),
        outer_bases_qual=kw.get("outer_qual"), outer_bases_length=kw.get("outer_len", 5),
        max_duplex_disagreements=kw.get("max_dis"), max_duplex_disagreement_rate=kw.get("max_rate", 1.0),
        produce_per_base_tags=kw.get("per_base", False), device=0, cell_tag=b"CB")
    got = caller.consensus_reads_batch(groups)
    gstats = caller.statistics()
    caller.close()
    oracle = R.CodecCallerOracle("codec", "RG1", vote_fn=vote_fn, builder_fn=O.builder_call,
                                 codec_job_fn=codec_job_fn, cell_tag=b"CB", **kw)
    want, count = bytearray(), 0
    for g in groups:
        d, n = oracle.consensus_reads(g)
        want += d
        count += n
    assert got.count == count
    if got.data != bytes(want):
        a, b = parse_records(got.data), parse_records(bytes(want))
        for i, (x, y) in enumerate(zip(a, b)):
            assert x == y, (i, x, y)
    assert got.data == bytes(want)
    assert gstats["total_reads"] == oracle.total_input_reads
    assert gstats["consensus_reads"] == oracle.consensus_reads_generated
    assert gstats["filtered_reads"] == oracle.reads_filtered
    for k in ("FragmentRead", "InsufficientReads", "MinorityAlignment", "InsufficientOverlap",
              "IndelErrorBetweenStrands"):
        assert gstats[k] == oracle.rejections.get(k, 0), k
    assert gstats["duplex_bases"] == oracle.duplex_bases
    assert gstats["duplex_disagreements"] == oracle.duplex_disagreements
    assert count >= 60
    assert oracle.rejections.get("InsufficientOverlap", 0) > 0 and oracle.rejections.get("FragmentRead", 0) > 0


@pytest.mark.gpu
def test_codec_known_answers_through_the_gpu_caller():
    """codec_caller.rs:2190-2231, 2593-2681, 2768-2803 through fgb_caller_* in CODEC mode."""
    import fgumi_b200 as fg
    from tests.test_codec_oracle_kat import create_fr_pair, simple_pair, REF, M, D, S
    c = fg.CodecConsensusCaller("codec", "RG1")
    out = c.consensus_reads_batch([
        simple_pair(ref_oriented=True),
        create_fr_pair(b"read1", 1, 11, 35, [(M, 30)], [(M, 25), (D, 5), (M, 5)]),
        create_fr_pair(b"read1", 1, 11, 35, [(S, 5), (M, 25)], [(M, 25), (S, 5)]),
        create_fr_pair(b"read1", 1, 11, 35, [(M, 30)], [(M, 19), (D, 2), (M, 11)]),
        create_fr_pair(b"read1", 100, 135, 35, [(M, 30)], [(M, 30)], rev1=True, rev2=False),
    ])
    st = c.statistics()
    c.close()
    recs = parse_records(out.data)
    assert out.count == 3 and [len(r["bases"]) for r in recs] == [40, 40, 45]
    assert recs[0]["bases"] == REF[:40] and recs[0]["name"] == b"codec:hi" and recs[0]["flags"] == 4
    assert recs[0]["tags"][b"RX"] == b"ACC-TGA" and recs[0]["tags"][b"cD"] == 2
    assert st["IndelErrorBetweenStrands"] == 2 and st["total_reads"] == 10 and st["consensus_reads"] == 3


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["simplex", "duplex", "codec"])
def test_threaded_add_groups_is_identical_to_the_sequential_path(mode):
    """fgb_caller_add_groups with n_threads > 1 (per-thread prep state merged in input order, parallel
    record assembly) must produce the same bytes and counters as add_group in a loop."""
    import fgumi_b200 as fg
    rng = np.random.default_rng(321)
    if mode == "simplex":
        groups = random_groups(rng, 700)
        mk = lambda t: fg.VanillaUmiConsensusCaller("fgumi", "A", fg.VanillaUmiConsensusOptions(min_reads=1, min_consensus_base_quality=2),
                                                    cell_tag=b"CB", consensus_call_overlapping_bases=True,
                                                    filter=fg.ConsensusFilter(min_reads=1, max_read_error_rate=0.2, max_base_error_rate=0.3,
                                                                              min_base_quality=10, max_no_call_fraction=0.5),
                                                    n_threads=t)
    elif mode == "duplex":
        groups = random_duplex_groups(rng, 500)
        mk = lambda t: fg.DuplexConsensusCaller("fgumi", "A", min_reads=(1, 1, 0), cell_tag=b"CB", n_threads=t)
    else:
        groups = random_codec_groups(rng, 600)
        mk = lambda t: fg.CodecConsensusCaller("codec", "RG1", produce_per_base_tags=True, cell_tag=b"CB", n_threads=t)
    a = mk(1)
    want = a.consensus_reads_batch(groups)
    sa = a.statistics()
    a.close()
    b = mk(5)
    b.add_groups(groups[:len(groups) // 3])          # two calls: appending to already queued work
    b.add_groups(groups[len(groups) // 3:])
    got = b.flush()
    sb = b.statistics()
    b.close()
    assert got.count == want.count and got.data == want.data
    assert sa == sb
