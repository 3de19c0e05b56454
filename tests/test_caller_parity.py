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
