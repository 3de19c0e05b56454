"""The product's host prep (fgumi_b200/csrc/host/prep.h) through the C-ABI, on the CPU: source-read
preparation + CIGAR filter against the oracle on random MI groups, and the UMI consensus against the
reference's SimpleConsensusCaller tests (simple_umi.rs:257-462).  No GPU needed."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import record_oracle as R           # noqa: E402
from tests import oracle_lib as O               # noqa: E402


def product_source_reads(records, min_q, trim):
    import fgumi_b200 as fg
    lib = fg.lib.load()
    blob = b"".join(records)
    off = np.zeros(len(records) + 1, np.uint64)
    off[1:] = np.cumsum([len(r) for r in records])
    cap = sum(R.Rec(r).l_seq for r in records) + 1
    ob, oq = np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
    row_off, orig = np.zeros(len(records) + 1, np.uint64), np.zeros(max(len(records), 1), np.uint32)
    n_rows, n_min = C.c_uint32(), C.c_uint32()
    buf = np.frombuffer(blob, np.uint8) if blob else np.zeros(1, np.uint8)
    st = lib.fgb_host_source_reads(buf.ctypes.data, off.ctypes.data, len(records), min_q, int(trim), ob.ctypes.data,
                                   oq.ctypes.data, row_off.ctypes.data, orig.ctypes.data, C.addressof(n_rows),
                                   C.addressof(n_min))
    assert st == 0
    rows = [(bytes(ob[int(row_off[r]):int(row_off[r + 1])]), bytes(oq[int(row_off[r]):int(row_off[r + 1])]), int(orig[r]))
            for r in range(n_rows.value)]
    return rows, n_min.value


@pytest.mark.parametrize("min_q,trim", [(10, False), (20, True), (2, False)])
def test_source_reads_and_cigar_filter_match_oracle(min_q, trim):
    from tests.test_caller_parity import random_groups
    rng = np.random.default_rng(4242 + min_q)
    opt = R.VanillaOptions(min_input_base_quality=min_q, trim=trim)
    n_rows = n_rejected = n_clipped = 0
    for group in random_groups(rng, 250):
        recs = [R.Rec(b) for b in group]
        srs = []
        for i, r in enumerate(recs):
            clip = R.num_bases_extending_past_mate(r)
            n_clipped += clip > 0
            sr = R.create_source_read(r, i, clip, opt)
            if sr is not None:
                srs.append(sr)
        kept, minority = R.filter_by_alignment(srs)
        want = [(bytes(s.bases), bytes(s.quals), s.original_idx) for s in kept]
        got, got_minority = product_source_reads(group, min_q, trim)
        assert got == want and got_minority == minority
        n_rows += len(want)
        n_rejected += minority
    assert n_rows > 500 and n_rejected > 0 and n_clipped > 0


def test_consensus_umis_kats():                       # simple_umi.rs:257-462, through the product
    import fgumi_b200 as fg
    lib = fg.lib.load()

    def cu(umis):
        arr = (C.c_char_p * max(len(umis), 1))(*[u.encode() for u in umis])
        out = C.create_string_buffer(256)
        st = lib.fgb_host_consensus_umis(arr, len(umis), out, 256)
        return out.value.decode() if st == 0 else None
    assert cu(["A", "A"]) == "A" and cu(["GATTACA", "GATTACA"]) == "GATTACA"
    assert cu(["A", "C", "G", "T"]) == "N"
    assert cu(["A", "C", "C", "C"]) == "C" and cu(["C", "C", "C", "A"]) == "C"
    assert cu(["GATTACA"] * 3 + ["NNNNNNN"]) == "GATTACA"
    assert cu(["GATT-ACA"] * 3) == "GATT-ACA" and cu(["XGAT", "XGAT"]) == "XGAT" and cu(["GATY", "GATY"]) == "GATY"
    assert cu(["AACC", "CCAA"]) == "NNNN"
    assert cu(["ACGT", "ACGT", "CAGT"]) == "ACGT" and cu(["ACGT"] * 3 + ["ACGG"]) == "ACGT"
    assert cu([]) == "" and cu(["ACGT"]) == "ACGT"
    for bad in (["A", "AC"], ["GATT-ACA", "GATT-ACA", "GATTAACA"], ["GATT-ACA", "GATT+ACA"]):
        assert cu(bad) is None                         # where the reference panics
    # and against the oracle on random UMI sets
    rng = np.random.default_rng(99)
    vote = lambda pre, post, b, q: O.builder_call(pre, post, b, q)[:2]
    for _ in range(300):
        n, ln = int(rng.integers(1, 9)), int(rng.integers(1, 12))
        true = rng.choice(list("ACGT"), size=ln)
        umis = []
        for _k in range(n):
            u = true.copy()
            m = rng.random(ln) < 0.2
            u[m] = rng.choice(list("ACGTNacgt"), size=int(m.sum()))
            umis.append("".join(u))
        assert cu(umis) == R.consensus_umis(list(umis), vote)


def test_malformed_records_do_not_crash_the_host_helpers():
    """Truncated and bit-flipped records through every record-taking host entry point: any status is
    acceptable, a crash or an out-of-bounds read (the test process dying) is not."""
    import fgumi_b200 as fg
    from tests.test_caller_parity import random_groups
    lib = fg.lib.load()
    rng = np.random.default_rng(7)
    fp = fg.DuplexConsensusFilter((2, 1, 1), (0.1,), (0.2,), 10, 20.0, 0.3, True).fill(fg.lib.FgbDuplexFilterParams())
    groups = random_groups(rng, 40)
    n_calls = 0
    for group in groups:
        for rec in group[:3]:
            for _ in range(12):
                b = bytearray(rec)
                mode = int(rng.integers(0, 4))
                if mode == 0:
                    b = b[:int(rng.integers(0, len(b)))]                       # truncate anywhere
                elif mode == 1:
                    for k in rng.integers(0, len(b), size=3):                  # flip bytes (header, cigar, aux)
                        b[int(k)] = int(rng.integers(0, 256))
                elif mode == 2:
                    b[12:14] = int(rng.integers(0, 65536)).to_bytes(2, "little")   # n_cigar_op
                else:
                    b[16:20] = int(rng.integers(0, 1 << 20)).to_bytes(4, "little")   # l_seq
                if len(b) == 0:
                    continue
                buf = (C.c_uint8 * len(b)).from_buffer(b)
                lib.fgb_host_is_fr_pair(C.addressof(buf), len(b))
                lib.fgb_host_num_bases_extending_past_mate(C.addressof(buf), len(b))
                masked, status = C.c_uint32(), C.c_uint8()
                lib.fgb_filter_record(C.addressof(buf), len(b), C.byref(fp), C.addressof(masked), C.addressof(status))
                off = np.array([0, len(b)], np.uint64)
                cap = max(int.from_bytes(bytes(b[16:20]), "little"), 1) if len(b) >= 20 else 1
                if cap <= (1 << 20):
                    ob, oq = np.zeros(cap + 8, np.uint8), np.zeros(cap + 8, np.uint8)
                    ro, oi = np.zeros(2, np.uint64), np.zeros(1, np.uint32)
                    nr, nm = C.c_uint32(), C.c_uint32()
                    lib.fgb_host_source_reads(C.addressof(buf), off.ctypes.data, 1, 10, 0, ob.ctypes.data, oq.ctypes.data,
                                              ro.ctypes.data, oi.ctypes.data, C.addressof(nr), C.addressof(nm))
                n_calls += 1
    assert n_calls > 1000


def _mutations(rng, rec, n):
    out = []
    for _ in range(n):
        b = bytearray(rec)
        mode = int(rng.integers(0, 5))
        if mode == 0:
            b = b[:int(rng.integers(0, len(b)))]
        elif mode == 1:
            for k in rng.integers(0, len(b), size=3):
                b[int(k)] = int(rng.integers(0, 256))
        elif mode == 2:
            b[12:14] = int(rng.integers(0, 65536)).to_bytes(2, "little")
        elif mode == 3:
            b[16:20] = int(rng.integers(0, 1 << 12)).to_bytes(4, "little")
        else:
            b[8] = int(rng.integers(0, 256))                                   # l_read_name
        out.append(bytes(b))
    return out


def test_host_helpers_under_asan(tmp_path):
    """The header-only host helpers (bam.h, prep.h, record_filter.h) compiled with AddressSanitizer and
    UBSan into a small harness (tests/native/host_fuzz.cpp) and run over ~10 k valid, truncated and
    corrupted records held in exact-size heap blocks: any out-of-bounds access aborts the harness."""
    import shutil
    import struct
    import subprocess
    from tests.test_caller_parity import random_groups, random_duplex_groups
    from tests.test_duplex_filter import _random_record
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "host_fuzz"
    r = subprocess.run([gxx, "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                        "-o", str(exe), os.path.join(root, "tests", "native", "host_fuzz.cpp"),
                        os.path.join(root, "fgumi_b200", "csrc", "host_tables.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    rng = np.random.default_rng(11)
    recs = []
    for group in random_groups(rng, 60) + random_duplex_groups(rng, 20):
        for rec in group:
            recs.append(bytes(rec))
            recs.extend(_mutations(rng, rec, 6))
    for _ in range(300):                                   # consensus records with per-base / strand tags
        rec = bytes(_random_record(rng))
        recs.append(rec)
        recs.extend(_mutations(rng, rec, 6))
    path = tmp_path / "records.bin"
    with open(path, "wb") as f:
        for rec in recs:
            f.write(struct.pack("<I", len(rec)) + rec)
    r = subprocess.run([str(exe), str(path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-3000:])
    assert r.stdout.startswith("records %d " % len(recs)) and len(recs) > 5000


def test_simplex_record_builder_matches_oracle():
    """fgb_host_simplex_record (the code the flush's record assembly runs) against the oracle's
    build_consensus_record_into restatement: odd and even lengths, every base code, depths / errors
    beyond i16, integer tag widths, cell and RX tags."""
    import fgumi_b200 as fg
    from tests.bam_builder import make_record
    lib = fg.lib.load()
    rng = np.random.default_rng(808)
    for trial in range(300):
        L = int(rng.integers(1, 200))
        bases = bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=L))
        quals = rng.integers(2, 94, size=L).astype(np.uint8)
        hi = int(rng.choice([5, 200, 40000, 65535]))
        depths = rng.integers(0, hi + 1, size=L).astype(np.uint16)
        errors = np.minimum(rng.integers(0, hi + 1, size=L), depths).astype(np.uint16)
        per_base = bool(rng.random() < 0.7)
        read_type = int(rng.integers(0, 3))
        umi = "%d" % int(rng.integers(0, 10 ** int(rng.integers(1, 9))))
        cell = b"CELL%d" % trial if rng.random() < 0.5 else None
        n_rx = int(rng.integers(0, 6))
        rx_true = "".join(rng.choice(list("ACGT"), size=6)) + "-" + "".join(rng.choice(list("ACGT"), size=6))
        rxs = []
        for _ in range(n_rx):
            u = list(rx_true)
            if rng.random() < 0.3:
                u[int(rng.integers(0, 6))] = str(rng.choice(list("ACGTN")))
            rxs.append("".join(u))
        # oracle: the caller's record builder fed with raw records that carry the RX / cell tags
        opt = R.VanillaOptions(produce_per_base_tags=per_base, cell_tag=b"CB" if cell else None)
        o = R.VanillaCallerOracle("fgumi", "grp", opt, None, O.builder_call)
        raws = [R.Rec(make_record(name=b"r", seq=b"A", quals=[30],
                                  tags=[(b"RX", "Z", x.encode())] + ([(b"CB", "Z", cell)] if cell else []))) for x in rxs]
        if cell and not raws:                         # the cell tag is read from the first source read
            raws = [R.Rec(make_record(name=b"r", seq=b"A", quals=[30], tags=[(b"CB", "Z", cell)]))]
        want = o._record(umi, ("Fragment", "R1", "R2")[read_type], raws, bases, bytes(quals), list(depths), list(errors))
        rx_arr = (C.c_char_p * max(n_rx, 1))(*[x.encode() for x in rxs])
        out = np.zeros(4 * L + 4096, np.uint8)
        n = C.c_size_t()
        b8 = np.frombuffer(bases, np.uint8)
        st = lib.fgb_host_simplex_record(b"fgumi", b"grp", umi.encode(), read_type, int(per_base), b8.ctypes.data,
                                         quals.ctypes.data, depths.ctypes.data, errors.ctypes.data, L,
                                         b"CB" if cell else None, cell, rx_arr, n_rx, out.ctypes.data, len(out),
                                         C.addressof(n))
        assert st == 0
        assert bytes(out[:n.value]) == want, trial


def test_duplex_record_builder_matches_oracle():
    """fgb_host_duplex_record (the code the duplex flush runs per read) against the oracle's
    duplex_read_into restatement: both strands / AB only, per-base tags on and off, depths beyond i16,
    cell tag, RX values from both segments (halves swapped for the other segment)."""
    import fgumi_b200 as fg
    from tests.bam_builder import make_record
    from tests.test_caller_parity import duplex_job_fn
    from tests.test_record_oracle_kat import vote_fn
    lib = fg.lib.load()
    rng = np.random.default_rng(909)
    keep = []                                            # ctypes arrays must outlive the call

    def strand(n, hi):
        b = rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=n).astype(np.uint8)
        q = rng.integers(2, 94, size=n).astype(np.uint8)
        d = rng.integers(0, hi + 1, size=n).astype(np.uint16)
        e = np.minimum(rng.integers(0, hi + 1, size=n), d).astype(np.uint16)
        keep.extend([b, q, d, e])
        sc = fg.lib.FgbStrandColumns(b.ctypes.data, q.ctypes.data, d.ctypes.data, e.ctypes.data, n, 1)
        return sc, R.SsCons(bytes(b), bytes(q), [int(x) for x in d], [int(x) for x in e], [])
    for trial in range(200):
        L = int(rng.integers(1, 180))
        hi = int(rng.choice([5, 300, 40000]))
        per_base = bool(rng.random() < 0.7)
        first = bool(rng.random() < 0.5)
        ab_c, ab_o = strand(L if rng.random() < 0.8 else L + int(rng.integers(1, 5)), hi)
        has_ba = rng.random() < 0.75
        ba_c, ba_o = strand(L, hi) if has_ba else (fg.lib.FgbStrandColumns(None, None, None, None, 0, 0), None)
        bases = rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=L).astype(np.uint8)
        quals = rng.integers(2, 94, size=L).astype(np.uint8)
        errors = rng.integers(0, 50, size=L).astype(np.uint16)
        cell = b"CELL%d" % trial if rng.random() < 0.5 else None
        n_rx = int(rng.integers(0, 6))
        rxs = ["".join(rng.choice(list("ACGT"), size=5)) + "-" + "".join(rng.choice(list("ACGT"), size=5)) for _ in range(n_rx)]
        rx_first = rng.integers(0, 2, size=max(n_rx, 1)).astype(np.uint8)
        o = R.DuplexCallerOracle("fgumi", "grp", per_base=per_base, cell_tag=b"CB" if cell else None, vote_fn=vote_fn,
                                 builder_fn=O.builder_call, duplex_job_fn=duplex_job_fn)
        raws = [R.Rec(make_record(name=b"r", flags=(0x41 if rx_first[i] else 0x81), seq=b"A", quals=[30],
                                  tags=[(b"RX", "Z", rxs[i].encode())])) for i in range(n_rx)]
        d = R.DuplexCons(bytes(bases), bytes(quals), [int(x) for x in errors], ab_o, ba_o)
        want = o._record(d, "R1" if first else "R2", "mol%d" % trial, raws, [], first, cell)
        rx_arr = (C.c_char_p * max(n_rx, 1))(*[x.encode() for x in rxs])
        out = np.zeros(16 * L + 8192, np.uint8)
        n = C.c_size_t()
        st = lib.fgb_host_duplex_record(b"fgumi", b"grp", b"mol%d" % trial, int(first), int(per_base), bases.ctypes.data,
                                        quals.ctypes.data, errors.ctypes.data, L, C.byref(ab_c), C.byref(ba_c),
                                        b"CB" if cell else None, cell, rx_arr, rx_first.ctypes.data, n_rx,
                                        out.ctypes.data, len(out), C.addressof(n))
        assert st == 0
        assert bytes(out[:n.value]) == want, trial


def test_source_reads_match_oracle_on_random_records():
    """make_source_read's fused decode / orientation / mask / clip / strip pass and the CIGAR filter
    against the oracle on records with random CIGARs (clips, indels, =/X), lengths 1-70 (odd and even:
    the two-bases-per-byte tables), every base code, missing qualities, both strands, mate overlaps."""
    from tests.bam_builder import make_record, encode_op
    from tests.test_rawbam_helpers_kat import _random_cigar, M, I, S, EQ, X, P, F1, F2, REV, MREV
    rng = np.random.default_rng(6060)
    codes = np.frombuffer(b"ACGTNRYKMacgtn=", np.uint8)
    n_rows = n_empty = 0
    for trial in range(600):
        min_q, trim = int(rng.choice([2, 10, 25])), bool(rng.random() < 0.4)
        group = []
        for k in range(int(rng.integers(1, 9))):
            cig = _random_cigar(rng)
            qlen = sum(n for kk, n in cig if kk in (M, I, S, EQ, X))
            seq = bytes(rng.choice(codes, size=qlen))
            r = rng.random()
            quals = [0xFF] * qlen if r < 0.05 else rng.integers(0, 45, size=qlen).tolist()
            rev = bool(rng.random() < 0.5)
            pos = int(rng.integers(100, 300))
            flag = P | (REV if rev else MREV) | (F1 if rng.random() < 0.5 else F2)
            tags = [(b"MI", "Z", b"1")]
            if rng.random() < 0.8:
                tags.append((b"MC", "Z", ("%dM" % int(rng.integers(20, 80))).encode()))
            group.append(make_record(name=b"r%d" % k, flags=flag, ref_id=0, pos=pos, mate_ref_id=0,
                                     mate_pos=pos + int(rng.integers(-40, 40)), tlen=int(rng.integers(-200, 200)),
                                     cigar=[encode_op(kk, n) for kk, n in cig], seq=seq, quals=quals, tags=tags))
        opt = R.VanillaOptions(min_input_base_quality=min_q, trim=trim)
        srs = []
        for i, b in enumerate(group):
            rec = R.Rec(b)
            sr = R.create_source_read(rec, i, R.num_bases_extending_past_mate(rec), opt)
            if sr is not None:
                srs.append(sr)
            else:
                n_empty += 1
        kept, minority = R.filter_by_alignment(srs)
        want = [(bytes(s.bases), bytes(s.quals), s.original_idx) for s in kept]
        got, got_minority = product_source_reads(group, min_q, trim)
        assert got == want and got_minority == minority, trial
        n_rows += len(want)
    assert n_rows > 1000 and n_empty > 20
