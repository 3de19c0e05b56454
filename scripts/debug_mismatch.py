"""Debug helper (GPU box): isolate failing units of the odd-alphabet test."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fgumi_b200 as fg
from tests import oracle_lib as O
from tests.test_vote_parity import _ragged_units

def run(units):
    batch = fg.pack_source_reads(units, 1)
    eng = fg.Engine(0, 45, 40, 1, 0)
    out = eng.vote(batch); eng.close()
    ob, oq, od, oe, cl = O.simplex_batch(batch, 45, 40, 1, 0, 1)
    bad = []
    for u, sl in enumerate(batch.unit_slices()):
        for p in range(sl.start, sl.stop):
            if (out.base[p], out.qual[p], out.depth[p], out.errors[p]) != (ob[p], oq[p], od[p], oe[p]):
                bad.append((u, p - sl.start, int(out.qual[p]), int(oq[p])))
    return bad, batch

rng = np.random.default_rng(6)
units = _ragged_units(rng, 400, 9, 5, 60, alphabet=b"ACGTNacgtnRYKM.", qlo=0, qhi=255)
bad, _ = run(units)
print("full batch bad:", bad)
for (u, pos, g, r) in bad:
    b1, batch = run([units[u]])
    print("unit", u, "alone bad:", b1, "n_reads", len(units[u]), "lens", [len(x[0]) for x in units[u]])
    if b1:
        for (bb, qq) in units[u]:
            print("   ", bb, list(qq))
        print("   reads desc", [hex(int(x)) for x in batch.reads[:len(units[u])]], batch.units, batch.tiles)
    # neighbours
    b2, _ = run(units[max(0, u - 1): u + 2])
    print("   with neighbours bad:", b2)
