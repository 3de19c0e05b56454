"""Committed golden fixtures (tests/golden/, written by tests/golden/make_golden.py from the oracle):
CPU -- the live oracle still reproduces them; GPU -- the product reproduces them through the C-ABI."""
import os

import numpy as np
import pytest

import fgumi_b200 as fg
from tests import oracle_lib as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
COLUMN_FILES = ["columns_45_40.npz", "columns_30_20_m2.npz"]


def load_batch(name):
    z = np.load(os.path.join(GOLD, name))
    nu, nr, nb, no = (int(x) for x in z["n"])
    batch = fg.PackedBatch(z["bases"].copy(), z["quals"].copy(), z["reads"].copy(),
                           z["units"].copy().view(fg.UNIT_DTYPE), nu, nr, nb, no)
    return z, batch


@pytest.mark.parametrize("name", COLUMN_FILES)
def test_oracle_reproduces_column_golden(name):
    z, batch = load_batch(name)
    pre, post, min_reads, min_q = (int(x) for x in z["params"])
    ob, oq, od, oe, _ = O.simplex_batch(batch, pre, post, min_reads, min_q)
    n = batch.n_out
    assert np.array_equal(ob[:n], z["out_base"]) and np.array_equal(oq[:n], z["out_qual"])
    assert np.array_equal(od[:n], z["out_depth"]) and np.array_equal(oe[:n], z["out_errors"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", COLUMN_FILES)
def test_gpu_reproduces_column_golden(name):
    z, batch = load_batch(name)
    pre, post, min_reads, min_q = (int(x) for x in z["params"])
    eng = fg.Engine(0, pre, post, min_reads, min_q)
    got = eng.vote(batch)
    eng.close()
    n = batch.n_out
    assert np.array_equal(got.base[:n], z["out_base"]) and np.array_equal(got.qual[:n], z["out_qual"])
    assert np.array_equal(got.depth[:n], z["out_depth"]) and np.array_equal(got.errors[:n], z["out_errors"])


def groups_of(z, mode):
    blob = z[mode + "_records"].tobytes()
    recs, p = [], 0
    for n in z[mode + "_rec_len"]:
        recs.append(blob[p:p + int(n)]); p += int(n)
    groups, k = [], 0
    for n in z[mode + "_group_len"]:
        groups.append(recs[k:k + int(n)]); k += int(n)
    return groups


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["simplex", "duplex", "codec"])
def test_gpu_callers_reproduce_record_golden(mode):
    z = np.load(os.path.join(GOLD, "callers.npz"))
    groups = groups_of(z, mode)
    if mode == "simplex":
        c = fg.VanillaUmiConsensusCaller("fgumi", "A", fg.VanillaUmiConsensusOptions(min_reads=1, min_consensus_base_quality=2))
    elif mode == "duplex":
        c = fg.DuplexConsensusCaller("fgumi", "A", min_reads=(1, 1, 0), produce_per_base_tags=True, cell_tag=b"CB")
    else:
        c = fg.CodecConsensusCaller("codec", "RG1", produce_per_base_tags=True, cell_tag=b"CB")
    got = c.consensus_reads_batch(groups)
    c.close()
    assert got.data == z[mode + "_expected"].tobytes()
    assert got.count > 10
