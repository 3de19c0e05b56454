"""Record-level caller throughput (host prep + GPU vote + record assembly), single host thread."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fgumi_b200 as fg
from tests.bam_builder import make_record

rng = np.random.default_rng(3)
G, D, L = int(sys.argv[1]) if len(sys.argv) > 1 else 20000, 8, 150
P, F1, F2, REV, MREV = 1, 0x40, 0x80, 0x10, 0x20
t0 = time.perf_counter()
groups = []
acgt = np.frombuffer(b"ACGT", np.uint8)
for g in range(G):
    tmpl = acgt[rng.integers(0, 4, size=400)].tobytes()
    recs = []
    for d in range(D):
        q = bytes(rng.integers(25, 40, size=L).astype(np.uint8))
        tags = [(b"MI", "Z", b"%d" % g), (b"RX", "Z", b"ACGTAC-TTGACA"), (b"MC", "Z", b"150M")]
        recs.append(make_record(name=b"r%d_%d" % (g, d), flags=P | F1 | MREV, pos=1000, mate_ref_id=0, mate_pos=1200,
                                tlen=350, seq=tmpl[:L], quals=q, tags=tags))
        recs.append(make_record(name=b"r%d_%d" % (g, d), flags=P | F2 | REV, pos=1200, mate_ref_id=0, mate_pos=1000,
                                tlen=-350, seq=tmpl[200:200 + L], quals=q, tags=tags))
    groups.append(recs)
print(f"generated {G} groups x {2 * D} reads in {time.perf_counter() - t0:.1f} s")
blobs = []
for recs in groups:
    off = np.zeros(len(recs) + 1, dtype=np.uint64); off[1:] = np.cumsum([len(r) for r in recs])
    blobs.append((np.frombuffer(b"".join(recs), np.uint8), off, len(recs)))
for name, mk in (("simplex", lambda: fg.VanillaUmiConsensusCaller("fgumi", "A", fg.VanillaUmiConsensusOptions(min_reads=1, min_consensus_base_quality=2))),
                 ("simplex+overlap", lambda: fg.VanillaUmiConsensusCaller("fgumi", "A", fg.VanillaUmiConsensusOptions(min_reads=1, min_consensus_base_quality=2), consensus_call_overlapping_bases=True))):
    c = mk()
    lib = c._lib
    for rep in range(2):
        t0 = time.perf_counter()
        for buf, off, n in blobs:
            st = lib.fgb_caller_add_group(c._h, buf.ctypes.data, off.ctypes.data, n)
            assert st == 0
        t1 = time.perf_counter()
        out = c.flush()
        t2 = time.perf_counter()
    nreads = G * 2 * D
    print(f"{name}: add_group {nreads / (t1 - t0) / 1e6:.2f} M reads/s, flush {out.count / (t2 - t1) / 1e6:.3f} M consensus reads/s "
          f"({t2 - t1:.3f} s), total {nreads / (t2 - t0) / 1e6:.2f} M input reads/s")
    c.close()

# fgb_caller_add_groups: one call, n_threads inside the library
recs = [r for g in groups for r in g]
blob = np.frombuffer(b"".join(recs), np.uint8)
off = np.zeros(len(recs) + 1, dtype=np.uint64); off[1:] = np.cumsum([len(r) for r in recs])
grp = np.arange(G + 1, dtype=np.uint64) * np.uint64(2 * D)
for T in (1, 4, 8, 16, 32):
    c = fg.VanillaUmiConsensusCaller("fgumi", "A", fg.VanillaUmiConsensusOptions(min_reads=1, min_consensus_base_quality=2), n_threads=T)
    best = None
    for rep in range(3):
        t0 = time.perf_counter()
        assert c._lib.fgb_caller_add_groups(c._h, blob.ctypes.data, off.ctypes.data, grp.ctypes.data, G) == 0
        t1 = time.perf_counter()
        out = c.flush()
        t2 = time.perf_counter()
        if best is None or t2 - t0 < best[0]:
            best = (t2 - t0, t1 - t0, t2 - t1)
    print(f"n_threads {T:2d}: {G * 2 * D / best[0] / 1e6:6.2f} M input reads/s, {out.count / best[0] / 1e6:6.3f} M consensus reads/s "
          f"(prep {best[1] * 1e3:.0f} ms, flush {best[2] * 1e3:.0f} ms)")
    c.close()
