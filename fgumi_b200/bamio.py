"""File-level run of the consensus path: a grouped BAM file in, a consensus BAM file out.

    BGZF inflate (fgb_bgzf_decompress, members on threads) -> BAM header + record split -> MI grouping
    (fgb_host_group_by_mi) -> record-level caller (fgb_caller_add_groups + fgb_caller_flush: the GPU vote)
    -> output header + ConsensusOutput stream -> BGZF deflate (fgb_bgzf_compress)

What `fgumi simplex -i grouped.bam -o consensus.bam` does around the hot path (commands/simplex.rs,
consensus_runner.rs:120-163 for the header: one @RG, SO:unknown GO:query, @PG appended); everything heavy is
in the library, this module only strings the calls together.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import Dict, Tuple

import numpy as np

from . import lib as _l


def read_bam(path: str, n_threads: int = 4) -> Tuple[bytes, np.ndarray, np.ndarray, Dict[str, float]]:
    """(SAM header text, record bodies uint8[...], rec_off uint64[n + 1], timings)."""
    lib = _l.load()
    t0 = time.perf_counter()
    comp = np.fromfile(path, dtype=np.uint8)
    size = C.c_size_t()
    st = lib.fgb_bgzf_uncompressed_size(comp.ctypes.data, len(comp), C.addressof(size))
    if st != 0:
        raise _l.FgbError(st, "fgb_bgzf_uncompressed_size")
    raw = np.empty(size.value + 16, dtype=np.uint8)
    n = C.c_size_t()
    t1 = time.perf_counter()
    st = lib.fgb_bgzf_decompress(comp.ctypes.data, len(comp), n_threads, raw.ctypes.data, size.value, C.addressof(n))
    if st != 0:
        raise _l.FgbError(st, "fgb_bgzf_decompress")
    t2 = time.perf_counter()
    toff, tlen, roff = C.c_size_t(), C.c_size_t(), C.c_size_t()
    nref = C.c_uint32()
    st = lib.fgb_bam_read_header(raw.ctypes.data, n.value, C.addressof(toff), C.addressof(tlen), C.addressof(nref), C.addressof(roff))
    if st != 0:
        raise _l.FgbError(st, "fgb_bam_read_header")
    text = raw[toff.value:toff.value + tlen.value].tobytes()
    stream = raw[roff.value:n.value]
    cap = max(1, len(stream) // 36)
    rec_off = np.zeros(cap + 1, dtype=np.uint64)
    nrec, used = C.c_uint64(), C.c_size_t()
    st = lib.fgb_bam_split_records(stream.ctypes.data, len(stream), stream.ctypes.data, rec_off.ctypes.data, cap,
                                   C.addressof(nrec), C.addressof(used))       # in place: bodies move down
    if st != 0:
        raise _l.FgbError(st, "fgb_bam_split_records")
    if used.value != len(stream):
        raise ValueError("truncated BAM record stream")
    t3 = time.perf_counter()
    timings = {"read_s": t1 - t0, "inflate_s": t2 - t1, "split_s": t3 - t2, "compressed_bytes": float(len(comp)),
               "uncompressed_bytes": float(n.value)}
    return text, stream[:int(rec_off[nrec.value])], rec_off[:nrec.value + 1], timings


def group_by_mi(bodies: np.ndarray, rec_off: np.ndarray, tag: bytes = b"MI", strip_strand_suffix: bool = False):
    """fgb_host_group_by_mi: group table over the kept records (all of them when every record has the tag)."""
    lib = _l.load()
    n = len(rec_off) - 1
    keep = np.zeros(max(n, 1), dtype=np.uint8)
    gb = np.zeros(n + 2, dtype=np.uint64)
    ng = C.c_uint64()
    st = lib.fgb_host_group_by_mi(bodies.ctypes.data, rec_off.ctypes.data, n, tag, int(strip_strand_suffix), None,
                                  keep.ctypes.data, gb.ctypes.data, C.addressof(ng))
    if st != 0:
        raise _l.FgbError(st, "fgb_host_group_by_mi")
    if int(keep[:n].sum()) != n:
        raise ValueError("records without the MI tag: drop them before calling (the table indexes kept records)")
    return gb[:ng.value + 1]


def output_header(input_text: bytes, read_group_id: str, command_line: str = "fgumi_b200 simplex") -> bytes:
    """consensus_runner.rs:120-163 in spirit: unsorted / query-grouped, one @RG carrying the inputs' SM / LB / PL
    (collapsed when they agree), one @CO, the inputs' @PG chain with ours appended."""
    fields = {}
    pgs = []
    for line in input_text.decode("utf-8", "replace").splitlines():
        if line.startswith("@RG"):
            for f in line.split("\t")[1:]:
                k, _, v = f.partition(":")
                if k in ("SM", "LB", "PL", "PU", "CN", "DS"):
                    fields.setdefault(k, set()).add(v)
        elif line.startswith("@PG"):
            pgs.append(line)
    rg = "@RG\tID:" + read_group_id + "".join("\t%s:%s" % (k, ",".join(sorted(v))) for k, v in sorted(fields.items()))
    out = ["@HD\tVN:1.6\tSO:unknown\tGO:query", rg, "@CO\tconsensus reads called by fgumi_b200"] + pgs
    out.append("@PG\tID:fgumi_b200\tPN:fgumi_b200\tCL:" + command_line)
    return ("\n".join(out) + "\n").encode()


def simplex_file(in_path: str, out_path: str, caller, n_threads: int = 4, level: int = 1) -> Dict[str, float]:
    """One file-level run with an existing VanillaUmiConsensusCaller (its options decide everything else)."""
    lib = _l.load()
    text, bodies, rec_off, tm = read_bam(in_path, n_threads)
    t0 = time.perf_counter()
    groups = group_by_mi(bodies, rec_off)
    t1 = time.perf_counter()
    st = lib.fgb_caller_add_groups(caller._h, bodies.ctypes.data, rec_off.ctypes.data, groups.ctypes.data, len(groups) - 1)
    if st != 0:
        caller._check(st, "fgb_caller_add_groups")
    data, n, cnt = C.c_void_p(), C.c_uint64(), C.c_uint64()
    st = lib.fgb_caller_flush(caller._h, C.byref(data), C.byref(n), C.byref(cnt))
    if st != 0:
        caller._check(st, "fgb_caller_flush")
    t2 = time.perf_counter()
    hdr_text = output_header(text, caller._rg.decode() if hasattr(caller, "_rg") else "A")
    hdr = np.empty(len(hdr_text) + 16, dtype=np.uint8)
    hn = C.c_size_t()
    st = lib.fgb_bam_header(hdr_text, len(hdr_text), hdr.ctypes.data, len(hdr), C.addressof(hn))
    if st != 0:
        raise _l.FgbError(st, "fgb_bam_header")
    payload = np.empty(hn.value + n.value, dtype=np.uint8)
    payload[:hn.value] = hdr[:hn.value]
    if n.value:
        C.memmove(payload.ctypes.data + hn.value, data.value, n.value)
    cap = lib.fgb_bgzf_bound(len(payload))
    comp = np.empty(cap, dtype=np.uint8)
    cn = C.c_size_t()
    t3 = time.perf_counter()
    st = lib.fgb_bgzf_compress(payload.ctypes.data, len(payload), level, n_threads, 1, comp.ctypes.data, cap, C.addressof(cn))
    if st != 0:
        raise _l.FgbError(st, "fgb_bgzf_compress")
    t4 = time.perf_counter()
    comp[:cn.value].tofile(out_path)
    tm.update({"group_s": t1 - t0, "caller_s": t2 - t1, "deflate_s": t4 - t3, "input_records": float(len(rec_off) - 1),
               "groups": float(len(groups) - 1), "consensus_reads": float(cnt.value), "output_bytes": float(len(payload)),
               "output_compressed_bytes": float(cn.value)})
    return tm
