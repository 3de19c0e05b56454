"""fgb_plan_tiles_jobs (host, no GPU): the tile plan that keeps a duplex job's single-strand units in one tile so the
vote kernels' epilogue can combine them (include/fgumi_b200.h, "since ABI 3").  Checked here: the plan is a valid
tile plan (every unit in exactly one tile, stage limits, class order), attached jobs lie inside their tile, the job
lists are a partition of the attached jobs, and without jobs the plan is fgb_plan_tiles + fgb_sort_tiles_by_class."""
import numpy as np
import pytest

import fgumi_b200 as fg
from fgumi_b200 import lib as L
from fgumi_b200 import synth
from fgumi_b200.engine import plan_tiles, plan_tiles_jobs, sort_tiles_by_class


def _check_plan(host, jobs, tiles, class_tiles, tile_jobs, job_index, n_attached):
    U = host.n_units
    cov = np.zeros(U, int)
    for t in tiles:
        cov[t["unit_begin"]:t["unit_begin"] + t["n_units"]] += 1
    assert (cov == 1).all()
    lib = L.load()
    direct = (tiles["flags"] & 1) != 0
    assert (tiles["byte_len"][~direct] <= lib.fgb_tile_capacity_bytes()).all()
    assert (tiles["n_units"] <= lib.fgb_tile_max_units()).all()
    assert (tiles["n_reads"][~direct] + (tiles["read_begin"][~direct] & 1) <= lib.fgb_tile_max_reads()).all()
    cls = (tiles["flags"] >> 4) & 3
    assert (np.diff(cls.astype(int)) >= 0).all()                      # class order
    assert tuple(int((cls == c).sum()) for c in range(3)) == class_tiles
    # the class is truthful for shallow / deep tiles (the general kernel takes anything)
    rb = host.units["read_begin"].astype(np.int64)
    depth = np.diff(rb)
    for t, c in zip(tiles, cls):
        d = depth[t["unit_begin"]:t["unit_begin"] + t["n_units"]]
        if c == 1:
            assert d.max() <= 4
        if c == 2:
            assert d.min() >= 24
    # job lists
    assert int(tile_jobs["count"].sum()) == n_attached
    begins = np.concatenate([[0], np.cumsum(tile_jobs["count"])[:-1]])
    assert np.array_equal(tile_jobs["begin"], begins)
    listed = job_index[:n_attached]
    assert len(set(listed.tolist())) == n_attached
    cons = host.units["cons_len"]
    for k, t in enumerate(tiles):
        lo, hi = t["unit_begin"], t["unit_begin"] + t["n_units"]
        js = job_index[tile_jobs["begin"][k]:tile_jobs["begin"][k] + tile_jobs["count"][k]]
        assert (np.diff(js.astype(np.int64)) > 0).all()               # ascending inside a tile
        m = 0
        for j in js:
            a, b = int(jobs["unit_a"][j]), int(jobs["unit_b"][j])
            assert lo <= a < hi and lo <= b < hi and not (t["flags"] & 1)
            m = max(m, (min(int(cons[a]), int(cons[b])) + 7) // 8)
        assert tile_jobs["max_items"][k] == m
    # an unlisted job really has its units in different tiles (or in an oversize unit's tile)
    tile_of = np.zeros(U, int)
    for k, t in enumerate(tiles):
        tile_of[t["unit_begin"]:t["unit_begin"] + t["n_units"]] = k
    unlisted = sorted(set(range(len(jobs))) - set(listed.tolist()))
    for j in unlisted:
        a, b = int(jobs["unit_a"][j]), int(jobs["unit_b"][j])
        assert tile_of[a] != tile_of[b] or (tiles["flags"][tile_of[a]] & 1)


def _duplex_jobs(M):
    m = np.arange(M)
    jobs = np.zeros(2 * M, dtype=fg.DUPLEX_JOB_DTYPE)
    jobs["unit_a"][0::2], jobs["unit_b"][0::2] = 4 * m, 4 * m + 3
    jobs["unit_a"][1::2], jobs["unit_b"][1::2] = 4 * m + 1, 4 * m + 2
    jobs["out_off"] = np.arange(2 * M, dtype=np.uint64) * 152
    return jobs


def test_uniform_molecules_are_never_cut():
    M = 3000
    host = synth.make_descriptors(np.full(4 * M, 4, dtype=np.int64), 150, 1)
    jobs = _duplex_jobs(M)
    plan = plan_tiles_jobs(host, jobs)
    _check_plan(host, jobs, *plan)
    tiles, class_tiles, tile_jobs, job_index, n_attached = plan
    assert n_attached == 2 * M and (tiles["n_units"] % 4 == 0).all()
    assert class_tiles == (0, len(tiles), 0)
    assert (tiles["flags"] & 2).all()                                  # still regular tiles


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_mixed_depths_and_stray_jobs(seed):
    rng = np.random.default_rng(seed)
    M = 1500
    depths = rng.integers(1, 45, size=4 * M)
    depths[rng.random(4 * M) < 0.02] = 200                             # oversize units (voted from HBM)
    host = synth.make_descriptors(depths, 150, 1)
    jobs = _duplex_jobs(M)
    extra = np.zeros(200, dtype=fg.DUPLEX_JOB_DTYPE)
    extra["unit_a"] = rng.integers(0, 4 * M, size=200)
    extra["unit_b"] = rng.integers(0, 4 * M, size=200)
    jobs = np.concatenate([jobs, extra])[rng.permutation(2 * M + 200)]
    plan = plan_tiles_jobs(host, jobs)
    _check_plan(host, jobs, *plan)
    assert plan[4] > M                                                 # most molecules fit a stage whole


def test_without_jobs_it_is_the_plain_plan():
    rng = np.random.default_rng(9)
    depths = rng.integers(1, 40, size=5000)
    host = synth.make_descriptors(depths, 150, 1)
    plain = plan_tiles(host).copy()
    want, want_classes = sort_tiles_by_class(plain)
    tiles, class_tiles, tile_jobs, job_index, n_attached = plan_tiles_jobs(host, np.zeros(0, dtype=fg.DUPLEX_JOB_DTYPE))
    assert n_attached == 0 and class_tiles == want_classes
    assert np.array_equal(tiles, want) and not tile_jobs["count"].any()


def test_bad_job_is_refused():
    host = synth.make_descriptors(np.full(8, 4, dtype=np.int64), 150, 1)
    jobs = np.zeros(1, dtype=fg.DUPLEX_JOB_DTYPE)
    jobs["unit_a"], jobs["unit_b"] = 0, 8
    with pytest.raises(L.FgbError):
        plan_tiles_jobs(host, jobs)
