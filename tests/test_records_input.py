"""FGB_IN_RECORDS (rows built on the device from the BAM records' own bytes): the unpack kernel against a
position-by-position restatement of create_source_read's per-base part (vanilla_caller.rs:893-916), the
whole host-buffer call against the oracle, and the record-level caller at 100 k groups against the product's
host code over the CPU oracle's vote (oracle/libfgb_cpu_caller.so) and the Python record oracle."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import oracle_lib as O              # noqa: E402
from tests.bam_builder import make_record      # noqa: E402

ACGT = np.frombuffer(b"ACGT", np.uint8)
FWD = b"=ACMGRSVTWYHKDBN"
REVC = b"=TGMCRSVAWYHKDBN"


def _random_records(rng, n):
    """Records of assorted lengths / strands / bases (incl. N and IUPAC codes) with a random kept row length."""
    recs, rows = [], []
    for i in range(n):
        L = int(rng.choice([1, 2, 7, 8, 9, 15, 16, 17, 31, 64, 100, 149, 150, 151, 255, 301]))
        seq = ACGT[rng.integers(0, 4, size=L)].copy()
        m = rng.random(L) < 0.05
        seq[m] = np.frombuffer(b"NRYKM", np.uint8)[rng.integers(0, 5, size=int(m.sum()))]
        q = rng.integers(0, 60, size=L).astype(np.uint8)
        rev = bool(rng.random() < 0.5)
        name = b"r%d" % i + b"x" * int(rng.integers(0, 9))       # varies the alignment of the sequence field
        recs.append(make_record(name=name, flags=16 if rev else 0, pos=100, seq=seq.tobytes(), quals=q.tobytes(),
                                tags=[(b"MI", "Z", b"7")] if rng.random() < 0.7 else []))
        final_len = int(rng.integers(1, L + 1))
        rows.append((L, rev, final_len, seq, q))
    return recs, rows


def _expected_row(L, rev, final_len, seq, q, min_q):
    """The record stores the read as sequenced; a reverse-strand row is its reverse complement."""
    code = {c: i for i, c in enumerate(FWD)}
    nib = np.array([code.get(int(b), 15) for b in seq], dtype=np.int64)
    out_b, out_q = np.zeros(final_len, np.uint8), np.zeros(final_len, np.uint8)
    for p in range(final_len):
        i = L - 1 - p if rev else p
        b = (REVC if rev else FWD)[nib[i]]
        qq = int(q[i])
        if qq < min_q:
            b, qq = ord("N"), 2
        out_b[p], out_q[p] = b, qq
    return out_b, out_q


@pytest.mark.gpu
@pytest.mark.parametrize("min_q", [0, 10, 200])
def test_unpack_records_kernel_matches_the_per_base_rule(fg, min_q):
    import torch
    rng = np.random.default_rng(11 + min_q)
    recs, rows = _random_records(rng, 700)
    blob = np.frombuffer(b"".join(recs), np.uint8)
    rec_off = np.concatenate([[0], np.cumsum([len(r) for r in recs])]).astype(np.uint64)
    R = len(recs)
    raw = np.zeros(R, dtype=fg.RAW_READ_DTYPE)
    reads = np.zeros(R + 2, dtype=np.uint64)
    off = 0
    for r, (rec, (L, rev, fl, _, _)) in enumerate(zip(recs, rows)):
        name_len = rec[8]
        n_cig = int.from_bytes(rec[12:14], "little")
        raw["src_off"][r] = int(rec_off[r]) + 32 + name_len + 4 * n_cig
        raw["raw_len"][r] = L
        raw["flags"][r] = 1 if rev else 0
        reads[r] = (off << 16) | fl
        off += (fl + 7) // 8 * 8
    dev = "cuda:0"
    pad = np.zeros(64, np.uint8)
    d_blob = torch.from_numpy(np.concatenate([blob, pad])).to(dev)
    d_raw = torch.from_numpy(raw.view(np.uint8).reshape(-1).copy()).to(dev)
    d_reads = torch.from_numpy(reads.view(np.uint8).copy()).to(dev)
    d_b = torch.full((off + 64,), 0xAA, dtype=torch.uint8, device=dev)
    d_q = torch.full((off + 64,), 0xAA, dtype=torch.uint8, device=dev)
    eng = fg.Engine(0, 45, 40, 1, 2)
    lib = fg.lib.load()
    batch = fg.lib.FgbBatch(0, R, off, 0, 0, None, None, d_reads.data_ptr(), None, None)
    rc = fg.lib.FgbRecordColumns(len(blob) + 64, d_blob.data_ptr(), d_raw.data_ptr(), min_q)
    st = lib.fgb_unpack_records_device(eng._h, C.byref(batch), C.byref(rc), C.c_void_p(d_b.data_ptr()),
                                       C.c_void_p(d_q.data_ptr()), None)
    assert st == 0
    torch.cuda.synchronize()
    assert lib.fgb_wait(eng._h) == 0                       # no span was rejected
    hb, hq = d_b.cpu().numpy(), d_q.cpu().numpy()
    off = 0
    for r, (L, rev, fl, seq, q) in enumerate(rows):
        eb, eq = _expected_row(L, rev, fl, seq, q, min_q)
        padded = (fl + 7) // 8 * 8
        assert np.array_equal(hb[off:off + fl], eb), (r, L, rev, fl)
        assert np.array_equal(hq[off:off + fl], eq), (r, L, rev, fl)
        assert not hb[off + fl:off + padded].any() and not hq[off + fl:off + padded].any()     # zero row padding
        off += padded
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("padded", [True, False])
@pytest.mark.parametrize("min_q", [0, 10])
def test_unpack_bam4_device_both_kernels(fg, padded, min_q):
    """fgb_unpack_bam4_device on spans that start on arbitrary even nibbles: columns whose sizes are multiples of 4
    take the flat word kernel (guarded window loads), any other size the byte-load kernel; both against the per-base
    rule.  Guard bytes around the columns must stay untouched-looking: the kernels never need them."""
    import torch
    rng = np.random.default_rng(21 + min_q)
    _, rows = _random_records(rng, 600)
    code = {c: i for i, c in enumerate(FWD)}
    nibs, quals, raw_list, reads = [], [], [], []
    off_raw, off_row = 0, 0
    for (L, rev, fl, seq, q) in rows:
        gap = 2 * int(rng.integers(0, 4))                     # spans start on any even raw index
        nibs.append(np.zeros(gap, np.uint8)); quals.append(np.zeros(gap, np.uint8))
        off_raw += gap
        raw_list.append((off_raw, L, 1 if rev else 0))
        nibs.append(np.array([code.get(int(b), 15) for b in seq], dtype=np.uint8)); quals.append(q)
        off_raw += L
        if off_raw % 2:
            nibs.append(np.zeros(1, np.uint8)); quals.append(np.zeros(1, np.uint8)); off_raw += 1
        reads.append((off_row << 16) | fl)
        off_row += (fl + 7) // 8 * 8
    want_mod = 0 if padded else 2
    while off_raw % 8 != want_mod:                            # n_raw % 4 and ((n_raw + 1) / 2) % 4 both 0, or not
        nibs.append(np.zeros(2, np.uint8)); quals.append(np.zeros(2, np.uint8)); off_raw += 2
    nib = np.concatenate(nibs); allq = np.concatenate(quals)
    assert nib.size == off_raw and off_raw % 2 == 0
    seq4 = ((nib[0::2] << 4) | nib[1::2]).astype(np.uint8)
    R = len(rows)
    raw = np.zeros(R, dtype=fg.RAW_READ_DTYPE)
    for r, (o, L, f) in enumerate(raw_list):
        raw[r] = (o, L, f)
    rd = np.zeros(R + 2, dtype=np.uint64); rd[:R] = reads
    dev = "cuda:0"
    d_seq = torch.from_numpy(seq4.copy()).to(dev)
    d_q = torch.from_numpy(allq.copy()).to(dev)
    d_raw = torch.from_numpy(raw.view(np.uint8).reshape(-1).copy()).to(dev)
    d_reads = torch.from_numpy(rd.view(np.uint8).copy()).to(dev)
    d_b = torch.full((off_row + 64,), 0xAA, dtype=torch.uint8, device=dev)
    d_qo = torch.full((off_row + 64,), 0xAA, dtype=torch.uint8, device=dev)
    eng = fg.Engine(0, 45, 40, 1, 2)
    lib = fg.lib.load()
    batch = fg.lib.FgbBatch(0, R, off_row, 0, 0, None, None, d_reads.data_ptr(), None, None)
    rc = fg.lib.FgbRawColumns(off_raw, d_seq.data_ptr(), d_q.data_ptr(), d_raw.data_ptr(), min_q)
    st = lib.fgb_unpack_bam4_device(eng._h, C.byref(batch), C.byref(rc), C.c_void_p(d_b.data_ptr()),
                                    C.c_void_p(d_qo.data_ptr()), None)
    assert st == 0
    torch.cuda.synchronize()
    assert lib.fgb_wait(eng._h) == 0
    hb, hq = d_b.cpu().numpy(), d_qo.cpu().numpy()
    off = 0
    for r, (L, rev, fl, seq, q) in enumerate(rows):
        eb, eq = _expected_row(L, rev, fl, seq, q, min_q)
        assert np.array_equal(hb[off:off + fl], eb), (r, L, rev, fl)
        assert np.array_equal(hq[off:off + fl], eq), (r, L, rev, fl)
        off += (fl + 7) // 8 * 8
    assert (hb[off:] == 0xAA).all() and (hq[off:] == 0xAA).all()       # nothing written past the last row
    eng.close()


def _cpu_caller(n_threads=4):
    """The product's host code over the CPU oracle's vote (test infrastructure, oracle/Makefile)."""
    from fgumi_b200 import benchlegs
    O.build()
    os.environ["FGB_CPU_THREADS"] = str(n_threads)
    cpu = C.CDLL(O.SO_CPU_CALLER)
    vp, u64 = C.c_void_p, C.c_uint64
    cpu.fgb_caller_create.argtypes = [C.c_int, C.POINTER(__import__("fgumi_b200").lib.FgbCallerOptions), C.POINTER(vp)]
    cpu.fgb_caller_add_groups.argtypes = [vp, vp, vp, vp, u64]
    cpu.fgb_caller_flush.argtypes = [vp, C.POINTER(vp), C.POINTER(u64), C.POINTER(u64)]
    cpu.fgb_caller_destroy.argtypes = [vp]
    cpu.fgb_caller_last_error.argtypes = [vp, C.c_char_p, C.c_size_t]
    return benchlegs._Caller(cpu, 0, n_threads)


@pytest.mark.gpu
def test_record_level_caller_at_100k_groups(fg):
    """Raw records -> ConsensusOutput on the GPU path (records staged and shipped whole, rows built on the
    device, 8 planning threads) against (a) the same host code over the CPU oracle's vote, every byte of
    100 k groups, and (b) the independent Python record oracle on the first 20 000 groups."""
    from fgumi_b200 import benchlegs, synth
    from oracle import record_oracle as R
    from tests.test_record_oracle_kat import vote_fn
    G = 100_000
    blob, off, grp = synth.record_families(G, 8, 150, 1e-3, seed=17, reverse_fraction=0.5)
    lib = fg.lib.load()
    gpu = benchlegs._Caller(lib, 0, 8)
    p, n, cnt = gpu.process(blob.ctypes.data, off.ctypes.data, grp.ctypes.data, G)
    got = C.string_at(p, n)
    assert cnt == G
    cpu = _cpu_caller(8)
    p2, n2, cnt2 = cpu.process(blob.ctypes.data, off.ctypes.data, grp.ctypes.data, G)
    want = C.string_at(p2, n2)
    assert cnt2 == cnt and len(want) == len(got)
    assert got == want
    # a second batch through the same caller (staging and pinned buffers are reused)
    p, n, cnt = gpu.process(blob.ctypes.data, off.ctypes.data, grp.ctypes.data, G)
    assert C.string_at(p, n) == want
    gpu.close(); cpu.close()
    # independent oracle on a prefix
    K = 20000
    caller = R.VanillaCallerOracle("fgumi", "A", R.VanillaOptions(min_reads=1, min_consensus_base_quality=2), vote_fn,
                                   O.builder_call)
    out = bytearray()
    for g in range(K):
        recs = [blob[int(off[r]):int(off[r + 1])].tobytes() for r in range(int(grp[g]), int(grp[g + 1]))]
        data, _ = caller.consensus_reads(recs)
        out += data
    assert bytes(out) == got[:len(out)]
