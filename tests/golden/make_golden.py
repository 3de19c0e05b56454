"""Regenerates the committed golden fixtures from the ORACLE (oracle/ + oracle/record_oracle.py).

The Rust reference cannot be built or run in this environment, so these are not outputs of the fgumi
binary: they freeze what the oracle -- itself pinned by the reference's known-answer tests -- produces
for fixed seeded inputs, so that (a) a change in the oracle shows up as a diff here and (b) the GPU
tests have a second, immutable comparison besides the live oracle.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import fgumi_b200 as fg                                         # noqa: E402  (host-side packing only)
from oracle import record_oracle as R                           # noqa: E402
from tests import oracle_lib as O                               # noqa: E402
from tests.test_record_oracle_kat import vote_fn                # noqa: E402


def column_case(seed, n_units, pre, post, min_reads, min_q):
    rng = np.random.default_rng(seed)
    units = []
    for _ in range(n_units):
        depth = int(rng.integers(1, 11))
        L = int(rng.integers(4, 90))
        tmpl = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=L)
        rows = []
        for _ in range(depth):
            ln = int(rng.integers(max(1, L - 6), L + 1))
            b = tmpl[:ln].copy()
            m = rng.random(ln) < 0.06
            b[m] = rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=int(m.sum()))
            q = rng.integers(2, 46, size=ln).astype(np.uint8)
            rows.append((b.tobytes(), q.tobytes()))
        units.append(rows)
    batch = fg.pack_source_reads(units, min_reads)
    ob, oq, od, oe, cl = O.simplex_batch(batch, pre, post, min_reads, min_q)
    n = batch.n_out
    return dict(bases=batch.bases, quals=batch.quals, reads=batch.reads, units=batch.units.view(np.uint8),
                n=np.array([batch.n_units, batch.n_reads, batch.n_bytes, batch.n_out], np.uint64),
                params=np.array([pre, post, min_reads, min_q], np.uint32),
                out_base=ob[:n], out_qual=oq[:n], out_depth=od[:n], out_errors=oe[:n])


def caller_cases():
    from tests.test_caller_parity import random_groups, random_duplex_groups, random_codec_groups, duplex_job_fn
    from tests.test_codec_oracle_kat import codec_job_fn
    out = {}
    rng = np.random.default_rng(2024)
    specs = {
        "simplex": (random_groups(rng, 40), R.VanillaCallerOracle(
            "fgumi", "A", R.VanillaOptions(min_reads=1, min_consensus_base_quality=2), vote_fn, O.builder_call)),
        "duplex": (random_duplex_groups(rng, 40), R.DuplexCallerOracle(
            "fgumi", "A", min_reads=(1, 1, 0), per_base=True, cell_tag=b"CB", vote_fn=vote_fn,
            builder_fn=O.builder_call, duplex_job_fn=duplex_job_fn)),
        "codec": (random_codec_groups(rng, 60), R.CodecCallerOracle(
            "codec", "RG1", vote_fn=vote_fn, builder_fn=O.builder_call, codec_job_fn=codec_job_fn,
            cell_tag=b"CB", per_base=True)),
    }
    for name, (groups, oracle) in specs.items():
        want = bytearray()
        for g in groups:
            d, _ = oracle.consensus_reads(g)
            want += d
        recs = [r for g in groups for r in g]
        out[name + "_records"] = np.frombuffer(b"".join(recs), np.uint8)
        out[name + "_rec_len"] = np.array([len(r) for r in recs], np.uint32)
        out[name + "_group_len"] = np.array([len(g) for g in groups], np.uint32)
        out[name + "_expected"] = np.frombuffer(bytes(want), np.uint8)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "columns_45_40.npz"), **column_case(11, 400, 45, 40, 1, 2))
    np.savez_compressed(os.path.join(HERE, "columns_30_20_m2.npz"), **column_case(12, 300, 30, 20, 2, 10))
    np.savez_compressed(os.path.join(HERE, "callers.npz"), **caller_cases())
    print("golden fixtures written")
