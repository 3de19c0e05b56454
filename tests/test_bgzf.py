"""BGZF / BAM framing of the callers' output (host code): every member is a valid gzip
member with the BC extra field and a correct BSIZE, the members decompress to the input, the stream
ends in the standard EOF member, and a header + ConsensusOutput stream parses back as a BAM file."""
import ctypes as C
import gzip
import os
import struct
import sys
import zlib

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.bam_builder import make_record       # noqa: E402

EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def bgzf(data: bytes, level=6, threads=1, eof=True) -> bytes:
    import fgumi_b200 as fg
    lib = fg.lib.load()
    cap = lib.fgb_bgzf_bound(len(data))
    out = np.zeros(cap, np.uint8)
    n = C.c_size_t()
    src = np.frombuffer(data, np.uint8) if data else np.zeros(1, np.uint8)
    assert lib.fgb_bgzf_compress(src.ctypes.data, len(data), level, threads, int(eof), out.ctypes.data, cap, C.addressof(n)) == 0
    return bytes(out[:n.value])


def members(stream: bytes):
    p, out = 0, []
    while p < len(stream):
        assert stream[p:p + 4] == b"\x1f\x8b\x08\x04" and stream[p + 10:p + 16] == b"\x06\x00BC\x02\x00"
        bsize = struct.unpack_from("<H", stream, p + 16)[0] + 1
        block = stream[p:p + bsize]
        crc, isize = struct.unpack_from("<II", block, bsize - 8)
        raw = zlib.decompress(block[18:bsize - 8], -15)
        assert len(raw) == isize and zlib.crc32(raw) == crc and isize <= 0xFF00
        out.append(raw)
        p += bsize
    assert p == len(stream)
    return out


@pytest.mark.parametrize("n,level,threads", [(0, 6, 1), (1, 1, 1), (0xFF00, 6, 2), (0xFF00 + 1, 0, 3), (1_000_003, 5, 4),
                                             (0, 1, 1), (0xFF00, 1, 2), (1_000_003, 1, 4)])
def test_bgzf_members_round_trip(n, level, threads):
    rng = np.random.default_rng(n + level)
    data = (rng.integers(0, 4, size=n).astype(np.uint8) * 17 + rng.integers(0, 2, size=n).astype(np.uint8)).tobytes()
    if n > 100000:                                       # a stretch of incompressible bytes too
        data = data[:500000] + rng.integers(0, 256, size=200000).astype(np.uint8).tobytes() + data[700000:]
    s = bgzf(data, level, threads)
    assert s.endswith(EOF_BLOCK)
    blocks = members(s)
    assert blocks[-1] == b"" and b"".join(blocks) == data
    assert gzip.decompress(s) == data                    # any gzip reader accepts the stream
    assert bgzf(data, level, 1) == s                     # the thread count does not change the bytes
    assert not bgzf(data, level, threads, eof=False).endswith(EOF_BLOCK) or n == 0


def test_bam_file_from_a_consensus_output_stream():
    import fgumi_b200 as fg
    lib = fg.lib.load()
    recs = [make_record(name=b"fgumi:%d" % i, flags=0x4D if i % 2 == 0 else 0x8D, ref_id=-1, pos=-1, cigar=[],
                        seq=b"ACGT" * (5 + i % 7), quals=[30] * (4 * (5 + i % 7)), tags=[(b"MI", "Z", b"%d" % (i // 2))])
            for i in range(5000)]
    stream = b"".join(struct.pack("<I", len(r)) + r for r in recs)       # what fgb_caller_flush returns
    text = b"@HD\tVN:1.6\tSO:unsorted\n@RG\tID:A\tSM:s\n"
    hdr = np.zeros(len(text) + 64, np.uint8)
    n = C.c_size_t()
    assert lib.fgb_bam_header(text, len(text), hdr.ctypes.data, len(hdr), C.addressof(n)) == 0
    bam = bgzf(bytes(hdr[:n.value]) + stream, 6, 3)
    raw = gzip.decompress(bam)
    assert raw[:4] == b"BAM\x01"
    (l_text,) = struct.unpack_from("<i", raw, 4)
    assert raw[8:8 + l_text] == text and struct.unpack_from("<i", raw, 8 + l_text)[0] == 0
    p, got = 12 + l_text, []
    while p < len(raw):
        (bs,) = struct.unpack_from("<I", raw, p)
        got.append(raw[p + 4:p + 4 + bs])
        p += 4 + bs
    assert got == recs


def test_write_bam_helper(tmp_path):
    import fgumi_b200 as fg
    recs = [make_record(name=b"c:%d" % i, flags=4, ref_id=-1, pos=-1, cigar=[], seq=b"ACGTN", quals=[30] * 5) for i in range(10)]
    out = fg.ConsensusOutput(b"".join(struct.pack("<I", len(r)) + r for r in recs), len(recs))
    path = tmp_path / "c.bam"
    fg.write_bam(str(path), b"@HD\tVN:1.6\n", out, level=5, n_threads=2)
    raw = gzip.decompress(open(path, "rb").read())
    assert raw[:4] == b"BAM\x01" and raw.endswith(recs[-1]) and open(path, "rb").read().endswith(EOF_BLOCK)


def test_bgzf_reader_and_record_split_round_trip(tmp_path):
    """The input side of a file-level run: fgb_bgzf_decompress (members on threads, CRC checked), the BAM header
    reader and the in-place record split give back exactly the records that were written."""
    import fgumi_b200 as fg
    from fgumi_b200 import bamio
    rng = np.random.default_rng(5)
    recs = [make_record(name=b"r%d" % i, flags=0, pos=i, seq=bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(1, 200)))),
                        tags=[(b"MI", "Z", b"%d" % (i // 3))]) for i in range(3000)]
    text = b"@HD\tVN:1.6\n@SQ\tSN:chr1\tLN:1000\n@RG\tID:A\tSM:s\n"
    hdr = b"BAM\1" + struct.pack("<I", len(text)) + text + struct.pack("<I", 1) + struct.pack("<I", 5) + b"chr1\0" + struct.pack("<I", 1000)
    data = hdr + b"".join(struct.pack("<I", len(r)) + r for r in recs)
    path = tmp_path / "in.bam"
    path.write_bytes(bgzf(data, level=1, threads=3))
    got_text, bodies, rec_off, tm = bamio.read_bam(str(path), n_threads=3)
    assert got_text == text and len(rec_off) == len(recs) + 1
    assert tm["uncompressed_bytes"] == len(data)
    for i in (0, 1, 17, 2999):
        assert bodies[int(rec_off[i]):int(rec_off[i + 1])].tobytes() == recs[i]
    assert bodies.tobytes() == b"".join(recs)
    groups = bamio.group_by_mi(bodies, rec_off)
    assert len(groups) == 1001 and int(groups[-1]) == 3000 and int(groups[1]) == 3
    # a corrupted member is refused
    lib = fg.lib.load()
    comp = np.frombuffer(path.read_bytes(), np.uint8).copy()
    comp[40] ^= 0x55
    out = np.zeros(len(data) + 16, np.uint8)
    n = C.c_size_t()
    assert lib.fgb_bgzf_decompress(comp.ctypes.data, len(comp), 2, out.ctypes.data, len(out), C.addressof(n)) != 0
    hdr_out = bamio.output_header(text, "A")
    assert hdr_out.startswith(b"@HD\tVN:1.6\tSO:unknown\tGO:query\n@RG\tID:A\tSM:s\n") and b"@PG\tID:fgumi_b200" in hdr_out


def _shapes(rng, kind, m):
    if kind == 0:
        return rng.integers(0, 256, size=m).astype(np.uint8)                       # incompressible
    if kind == 1:
        return np.zeros(m, np.uint8)                                               # maximal matches (258) back to back
    if kind == 2:
        return np.frombuffer((b"ACGTACGTTTGACA-" * (m // 15 + 1))[:m], np.uint8).copy()
    if kind == 3:
        return rng.integers(0, 4, size=m).astype(np.uint8)                         # four symbols: short codes
    if kind == 4:
        return np.repeat(rng.integers(0, 256, size=m // 7 + 1).astype(np.uint8), 7)[:m].copy()
    d = rng.integers(0, 256, size=m).astype(np.uint8)                              # far matches inside noise
    if m > 40000:
        d[35000:35300] = d[0:300]
        d[m // 2:m // 2 + 200] = 7
    return d


def test_builtin_level1_encoder_on_assorted_inputs():
    """Level 1 is the repo's own DEFLATE encoder (csrc/host/fast_deflate.h): zlib and gzip must read every member back,
    whatever the input looks like -- sizes around the encoder's thresholds and the 0xFF00 member boundary, all-literal,
    all-match, few-symbol and far-match data -- and it must never lose to a stored block by more than its header."""
    rng = np.random.default_rng(77)
    sizes = [0, 1, 2, 15, 16, 17, 31, 100, 1000, 0xFF00 - 1, 0xFF00, 0xFF00 + 1, 70_000, 200_000]
    for trial in range(6 * len(sizes)):
        kind, m = trial % 6, sizes[trial // 6]
        data = _shapes(rng, kind, m).tobytes()
        s = bgzf(data, 1, 1 + trial % 3)
        assert b"".join(members(s)) == data, (kind, m)
        blocks = (m + 0xFF00 - 1) // 0xFF00
        assert len(s) <= m + blocks * (18 + 8 + 5) + 28, (kind, m)
    # it does compress what compresses
    text = (b"RGZA\0cDi\x08\0\0\0cMi\x08\0\0\0MIZ1234567\0RXZACGTACGT-TTGACAGT\0" * 3000)
    assert len(bgzf(text, 1)) < len(text) // 10


def test_builtin_crc32_is_zlibs(fg_lib=None):
    """The gzip trailer of a level-1 member carries the encoder's own CRC-32: members() checks it against zlib.crc32
    for every block above; here lengths 0..70 (the slicing-by-8 loop's head and tail) explicitly."""
    rng = np.random.default_rng(5)
    for m in range(0, 71):
        data = rng.integers(0, 256, size=m).astype(np.uint8).tobytes()
        blk = members(bgzf(data, 1, 1, eof=False)) if m else [b""]
        assert blk[0] == data


# ---- the repo's own DEFLATE decoder (csrc/inflate_core.h): host form here, device form in the gpu test below ----
MEMBER_DTYPE = np.dtype([("in_off", "<u8"), ("out_off", "<u8"), ("in_len", "<u4"), ("out_len", "<u4"), ("crc", "<u4"), ("reserved", "<u4")])


def _scan(lib, stream: np.ndarray):
    n, total = C.c_uint64(), C.c_uint64()
    assert lib.fgb_bgzf_scan_members(stream.ctypes.data, stream.size, None, 0, C.byref(n), C.byref(total)) == 0
    members = np.zeros(max(n.value, 1), dtype=MEMBER_DTYPE)
    assert lib.fgb_bgzf_scan_members(stream.ctypes.data, stream.size, members.ctypes.data, n.value, C.byref(n), C.byref(total)) == 0
    return members[: n.value], total.value


def _mixed_stream(rng, sizes=(0, 1, 17, 1000, 0xFF00, 0xFF00 + 5, 70_000, 300_000)):
    """BGZF members written at levels 0 (stored blocks), 1 (the built-in encoder), 6 and 9 (zlib, dynamic and fixed
    blocks) over assorted data; returns (stream bytes, the data)."""
    parts, data = [], []
    for i, m in enumerate(sizes):
        for kind in range(6):
            d = _shapes(rng, kind, m).tobytes()
            parts.append(bgzf(d, (0, 1, 6, 9)[(i + kind) % 4], 1 + kind % 3, eof=False))
            data.append(d)
    return b"".join(parts) + EOF_BLOCK, b"".join(data)


def test_member_table_and_host_decoder_against_zlib():
    import fgumi_b200 as fg
    lib = fg.lib.load()
    rng = np.random.default_rng(91)
    stream_b, data = _mixed_stream(rng)
    stream = np.frombuffer(stream_b, np.uint8)
    members, total = _scan(lib, stream)
    assert total == len(data) and members[-1]["out_len"] == 0            # the EOF member
    # the table against an independent walk of the stream
    p = o = 0
    for mb in members:
        bsize = struct.unpack_from("<H", stream_b, p + 16)[0] + 1
        crc, isize = struct.unpack_from("<II", stream_b, p + bsize - 8)
        assert (mb["in_off"], mb["in_len"], mb["out_off"], mb["out_len"], mb["crc"]) == (p + 18, bsize - 26, o, isize, crc)
        p += bsize; o += isize
    # every member through the host form of the device decoder; guard bytes behind each output
    for mb in members:
        out = np.full(int(mb["out_len"]) + 16, 0xAB, np.uint8)
        st = lib.fgb_host_inflate_member(stream.ctypes.data + int(mb["in_off"]), int(mb["in_len"]), out.ctypes.data, int(mb["out_len"]))
        assert st == 0
        want = data[int(mb["out_off"]):int(mb["out_off"]) + int(mb["out_len"])]
        assert out[: mb["out_len"]].tobytes() == want and (out[mb["out_len"]:] == 0xAB).all()
        assert zlib.crc32(want) == mb["crc"]


def test_host_decoder_on_damaged_members_stays_inside_its_output():
    import fgumi_b200 as fg
    lib = fg.lib.load()
    rng = np.random.default_rng(92)
    stream_b, _ = _mixed_stream(rng, sizes=(1000, 0xFF00))
    stream = np.frombuffer(stream_b, np.uint8).copy()
    members, _ = _scan(lib, stream)
    bad_seen = 0
    for mb in members[:-1]:
        lo, n = int(mb["in_off"]), int(mb["in_len"])
        payload = stream[lo:lo + n].copy()
        for trial in range(6):
            dmg = payload.copy()
            if trial < 4:
                for _ in range(1 + trial):
                    dmg[int(rng.integers(0, n))] ^= np.uint8(1 << int(rng.integers(0, 8)))
            else:
                dmg = dmg[: n // 2 if trial == 4 else max(n - 3, 0)].copy()            # truncated
            out = np.full(int(mb["out_len"]) + 16, 0xAB, np.uint8)
            st = lib.fgb_host_inflate_member(dmg.ctypes.data if dmg.size else None, dmg.size, out.ctypes.data, int(mb["out_len"]))
            assert (out[mb["out_len"]:] == 0xAB).all()
            ok = st == 0 and zlib.crc32(out[: mb["out_len"]].tobytes()) == mb["crc"]
            bad_seen += not ok
    assert bad_seen > 50                                                   # damage is noticed (status or CRC)


def test_reader_with_the_own_decoder_equals_zlibs(monkeypatch):
    import fgumi_b200 as fg
    lib = fg.lib.load()
    rng = np.random.default_rng(93)
    stream_b, data = _mixed_stream(rng, sizes=(17, 0xFF00, 200_000))
    stream = np.frombuffer(stream_b, np.uint8)
    for own in (False, True):
        if own:
            monkeypatch.setenv("FGB_BGZF_OWN_INFLATE", "1")
        out = np.zeros(len(data) + 1, np.uint8)
        n = C.c_size_t()
        assert lib.fgb_bgzf_decompress(stream.ctypes.data, stream.size, 3, out.ctypes.data, out.size, C.byref(n)) == 0
        assert n.value == len(data) and out[: n.value].tobytes() == data


@pytest.mark.gpu
def test_bgzf_members_inflated_on_the_device():
    """fgb_bgzf_inflate_device: a stream of members written at four levels (stored, built-in encoder, zlib dynamic /
    fixed) comes out of the device byte for byte, CRCs checked there; damaged members are flagged one by one and
    touch nothing but their own output range."""
    import torch
    import fgumi_b200 as fg
    lib = fg.lib.load()
    rng = np.random.default_rng(94)
    stream_b, data = _mixed_stream(rng, sizes=(0, 1, 17, 1000, 0xFF00, 0xFF00 + 5, 70_000, 300_000, 2_000_000))
    stream = np.frombuffer(stream_b, np.uint8).copy()
    members, total = _scan(lib, stream)
    dev = "cuda:0"
    eng = fg.Engine(0, 45, 40, 1, 2)
    d_members = torch.from_numpy(members.view(np.uint8).reshape(-1).copy()).to(dev)
    d_out = torch.full((total + 64,), 0xCD, dtype=torch.uint8, device=dev)
    d_st = torch.full((len(members),), 99, dtype=torch.uint8, device=dev)
    d_bad = torch.zeros(1, dtype=torch.int64, device=dev)

    def run(host_stream):
        d_in = torch.from_numpy(host_stream).to(dev)
        d_out.fill_(0xCD); d_st.fill_(99); d_bad.zero_()
        assert lib.fgb_bgzf_inflate_device(eng._h, d_in.data_ptr(), d_members.data_ptr(), len(members), d_out.data_ptr(),
                                           d_st.data_ptr(), 1, d_bad.data_ptr(), None) == 0
        torch.cuda.synchronize()
        return d_out.cpu().numpy(), d_st.cpu().numpy(), int(d_bad.item())

    out, st, bad = run(stream)
    assert bad == 0 and not st.any()
    assert out[:total].tobytes() == data and (out[total:] == 0xCD).all()
    # damage every third member that has a payload; its neighbours must come out whole
    dmg = stream.copy()
    hit = []
    for i, mb in enumerate(members):
        if i % 3 == 0 and mb["in_len"] > 8 and mb["out_len"] > 0:
            for _ in range(3):
                dmg[int(mb["in_off"]) + int(rng.integers(0, mb["in_len"]))] ^= np.uint8(1 << int(rng.integers(0, 8)))
            hit.append(i)
    out, st, bad = run(dmg)
    assert bad == int((st != 0).sum()) and set(np.flatnonzero(st).tolist()) <= set(hit) and bad >= len(hit) * 9 // 10
    for i, mb in enumerate(members):
        if i not in hit:
            a, b = int(mb["out_off"]), int(mb["out_off"] + mb["out_len"])
            assert out[a:b].tobytes() == data[a:b], i
    assert (out[total:] == 0xCD).all()
    eng.close()


def test_encoder_and_decoder_under_sanitizers(tmp_path):
    """tests/native/deflate_fuzz.cpp: fast_deflate.h and inflate_core.h (the device decoder's code) compiled with
    ASan + UBSan against zlib -- encoder blocks read back by zlib and the own decoder, zlib streams of every level and
    strategy read by the own decoder, damaged / truncated / mis-sized streams ending in a status without a stray access."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "deflate_fuzz"
    r = subprocess.run([gxx, "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                        "-o", str(exe), os.path.join(root, "tests", "native", "deflate_fuzz.cpp"), "-lz"],
                       capture_output=True, text=True)
    if r.returncode != 0 and "zlib.h" in r.stderr:
        pytest.skip("zlib headers not available")
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([str(exe), "700"], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1"))
    assert r.returncode == 0 and "deflate_fuzz ok" in r.stdout, (r.stdout[-500:], r.stderr[-3000:])
