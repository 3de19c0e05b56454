"""MI grouping of an input record stream (ingest side): the oracle's restatement of
MiGroupIterator (src/lib/mi_group.rs:386-470) pinned by the reference's tests (:578-930), and the
product's fgb_host_group_by_mi against it (ported cases + random streams).  CPU only."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import record_oracle as R           # noqa: E402
from tests.bam_builder import make_record       # noqa: E402


def rec(mi=None, cb=None, tag=b"MI"):
    tags = ([(tag, "Z", mi)] if mi is not None else []) + ([(b"CB", "Z", cb)] if cb is not None else [])
    return make_record(name=b"read", flags=4, ref_id=-1, pos=-1, cigar=[], seq=b"ACGT", quals=[30] * 4, tags=tags)


def product(records, tag=b"MI", strip=False, cell=None):
    import fgumi_b200 as fg
    lib = fg.lib.load()
    blob = b"".join(records)
    off = np.zeros(len(records) + 1, np.uint64)
    off[1:] = np.cumsum([len(r) for r in records])
    keep = np.zeros(max(len(records), 1), np.uint8)
    gb = np.zeros(len(records) + 1, np.uint64)
    ng = C.c_uint64()
    buf = np.frombuffer(blob, np.uint8) if blob else np.zeros(1, np.uint8)
    assert lib.fgb_host_group_by_mi(buf.ctypes.data, off.ctypes.data, len(records), tag, int(strip), cell,
                                    keep.ctypes.data, gb.ctypes.data, C.addressof(ng)) == 0
    kept = [i for i in range(len(records)) if keep[i]]
    return [[kept[k] for k in range(int(gb[g]), int(gb[g + 1]))] for g in range(ng.value)]


def both(records, **kw):
    o = R.mi_groups(records, kw.get("tag", b"MI"), kw.get("strip", False), kw.get("cell"))
    assert product(records, **kw) == [idx for _, idx in o]
    return [(k, len(idx)) for k, idx in o]


def test_mi_group_iterator_kats():                    # mi_group.rs:578-760
    assert both([]) == []
    assert both([rec(b"0")] * 3) == [("0", 3)]
    assert both([rec(b"0"), rec(b"0"), rec(b"1"), rec(b"1"), rec(b"1"), rec(b"2")]) == [("0", 2), ("1", 3), ("2", 1)]
    assert both([rec(b"0"), rec(), rec(b"0"), rec(), rec(b"1")]) == [("0", 2), ("1", 1)]     # untagged records skipped
    assert both([rec(b"A", tag=b"RX"), rec(b"A", tag=b"RX"), rec(b"B", tag=b"RX")], tag=b"RX") == [("A", 2), ("B", 1)]
    dup = [rec(b"1/A"), rec(b"1/A"), rec(b"1/B"), rec(b"1/B"), rec(b"2/A"), rec(b"2/B")]
    assert both(dup, strip=True) == [("1", 4), ("2", 2)]
    assert both(dup) == [("1/A", 2), ("1/B", 2), ("2/A", 1), ("2/B", 1)]
    assert R.extract_mi_base("12/A") == "12" and R.extract_mi_base("12/C") == "12/C" and R.extract_mi_base("/B") == ""


def test_cell_tag_composite_keys():                   # mi_group.rs:825-930
    assert both([rec(b"1", b"ACGT")] * 2 + [rec(b"1", b"TGCA")] * 2, cell=b"CB") == [("1\tACGT", 2), ("1\tTGCA", 2)]
    assert both([rec(b"1", b"ACGT"), rec(b"1", b"TGCA")]) == [("1", 2)]
    assert both([rec(b"1"), rec(b"1"), rec(b"1", b"ACGT")], cell=b"CB") == [("1\t", 2), ("1\tACGT", 1)]
    assert both([rec(b"1/A", b"ACGT"), rec(b"1/B", b"ACGT"), rec(b"1/A", b"TGCA")], strip=True, cell=b"CB") == \
        [("1\tACGT", 2), ("1\tTGCA", 1)]


def test_random_streams_and_hand_over_to_a_caller():
    import fgumi_b200 as fg
    rng = np.random.default_rng(8)
    for trial in range(50):
        recs = []
        for g in range(int(rng.integers(1, 30))):
            mi = b"%d" % int(rng.integers(0, 6))
            for _ in range(int(rng.integers(1, 5))):
                r = rng.random()
                recs.append(rec() if r < 0.1 else rec(mi + (b"/A" if r < 0.5 else b"/B"), rng.choice([b"X", b"Y", None])))
        for kw in (dict(), dict(strip=True), dict(cell=b"CB"), dict(strip=True, cell=b"CB")):
            both(recs, **kw)
    # the table feeds fgb_caller_add_groups directly when nothing is skipped
    from tests.test_caller_parity import random_groups
    groups = random_groups(rng, 40)
    flat = [r for g in groups for r in g]
    got = product(flat)
    assert [len(g) for g in got] == [len(g) for g in groups]
    c = fg.VanillaUmiConsensusCaller("fgumi", "A", device=fg.lib.FGB_DEVICE_NONE)
    c.add_groups([[flat[i] for i in g] for g in got])
    one = fg.VanillaUmiConsensusCaller("fgumi", "A", device=fg.lib.FGB_DEVICE_NONE)
    for g in groups:
        one.add_group(g)
    assert c.pending()["units"] == one.pending()["units"]
    c.close(); one.close()
