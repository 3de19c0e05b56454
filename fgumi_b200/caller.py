"""Record-level callers over the C-ABI (fgb_caller_*): the host-side mirror of fgumi's
`ConsensusCaller` implementations.  Same names and option meaning as the reference
(crates/fgumi-consensus/src/vanilla_caller.rs:284-341, caller.rs:205-234); groups are queued and
voted in one GPU batch by `flush()`."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

from . import lib as _l
from .engine import VanillaUmiConsensusOptions


class ConsensusOutput:
    """caller.rs:173-178: concatenated `[u32 block_size][BAM record]` bytes + record count."""

    def __init__(self, data: bytes = b"", count: int = 0):
        self.data, self.count = data, count


def bgzf_compress(data: bytes, level: int = 1, n_threads: int = 1, eof: bool = True) -> bytes:
    """BGZF members for a byte stream (fgb_bgzf_compress; host code, zlib bound at call time)."""
    lib = _l.load()
    cap = lib.fgb_bgzf_bound(len(data))
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t()
    src = np.frombuffer(data, dtype=np.uint8) if data else np.zeros(1, dtype=np.uint8)
    st = lib.fgb_bgzf_compress(src.ctypes.data, len(data), level, n_threads, int(eof), out.ctypes.data, cap, C.addressof(n))
    if st != _l.FGB_OK:
        raise _l.FgbError(st, "fgb_bgzf_compress")
    return out[:n.value].tobytes()


def write_bam(path: str, sam_header_text: bytes, output: "ConsensusOutput", level: int = 1, n_threads: int = 1):
    """A BAM file from a caller's ConsensusOutput: header (no reference dictionary -- consensus reads are
    unmapped) + the record stream as it is, BGZF-compressed, EOF member appended."""
    lib = _l.load()
    hdr = np.empty(len(sam_header_text) + 16, dtype=np.uint8)
    n = C.c_size_t()
    st = lib.fgb_bam_header(sam_header_text, len(sam_header_text), hdr.ctypes.data, len(hdr), C.addressof(n))
    if st != _l.FGB_OK:
        raise _l.FgbError(st, "fgb_bam_header")
    with open(path, "wb") as f:
        f.write(bgzf_compress(hdr[:n.value].tobytes() + output.data, level, n_threads, True))


class ConsensusFilter:
    """`fgumi filter` options for single-strand consensus reads (commands/filter.rs:100-160):
    -M min_reads, -E max_read_error_rate, -e max_base_error_rate, -N min_base_quality,
    -q min_mean_base_quality, -n max_no_call_fraction (>= 1.0 = absolute count)."""

    def __init__(self, min_reads: int = 1, max_read_error_rate: float = 0.025,
                 max_base_error_rate: float = 0.1, min_base_quality=None, min_mean_base_quality=None,
                 max_no_call_fraction: float = 0.2):
        self.min_reads, self.max_read_error_rate = min_reads, max_read_error_rate
        self.max_base_error_rate, self.min_base_quality = max_base_error_rate, min_base_quality
        self.min_mean_base_quality, self.max_no_call_fraction = min_mean_base_quality, max_no_call_fraction

    def fill(self, fp: "_l.FgbFilterParams"):
        fp.min_reads = self.min_reads
        fp.min_base_quality = -1 if self.min_base_quality is None else self.min_base_quality
        fp.max_read_error_rate = self.max_read_error_rate
        fp.max_base_error_rate = self.max_base_error_rate
        fp.min_mean_base_quality = -1.0 if self.min_mean_base_quality is None else self.min_mean_base_quality
        fp.max_no_call_fraction = self.max_no_call_fraction
        fp.per_base_tags = 1
        return fp


class DuplexConsensusFilter:
    """`fgumi filter` options for duplex consensus reads: each of min_reads, max_read_error_rate and
    max_base_error_rate takes one to three values for the [duplex, AB, BA] tiers, missing ones filled
    from the last (filter.rs:20-28, 237-330); AB is the stricter strand tier, BA the lenient one."""

    def __init__(self, min_reads=(1,), max_read_error_rate=(0.025,), max_base_error_rate=(0.1,),
                 min_base_quality=None, min_mean_base_quality=None, max_no_call_fraction: float = 0.2,
                 require_single_strand_agreement: bool = False):
        three = lambda v: (list(v) + [list(v)[-1]] * 3)[:3]
        self.min_reads, self.max_read_error_rate = three(min_reads), three(max_read_error_rate)
        self.max_base_error_rate = three(max_base_error_rate)
        self.min_base_quality, self.min_mean_base_quality = min_base_quality, min_mean_base_quality
        self.max_no_call_fraction, self.require_ss = max_no_call_fraction, require_single_strand_agreement

    def fill(self, fp: "_l.FgbDuplexFilterParams"):
        ConsensusFilter(self.min_reads[0], self.max_read_error_rate[0], self.max_base_error_rate[0],
                        self.min_base_quality, self.min_mean_base_quality, self.max_no_call_fraction).fill(fp.cc)
        fp.ab_min_reads, fp.ba_min_reads = self.min_reads[1], self.min_reads[2]
        fp.ab_max_read_error_rate, fp.ba_max_read_error_rate = self.max_read_error_rate[1], self.max_read_error_rate[2]
        fp.ab_max_base_error_rate, fp.ba_max_base_error_rate = self.max_base_error_rate[1], self.max_base_error_rate[2]
        fp.require_ss_agreement = 1 if self.require_ss else 0
        return fp

    def apply(self, record: bytearray):
        """One assembled record through fgb_filter_record (host code): returns (status, newly masked)."""
        lib = _l.load()
        fp = self.fill(_l.FgbDuplexFilterParams())
        buf = (C.c_uint8 * len(record)).from_buffer(record)
        masked, status = C.c_uint32(), C.c_uint8()
        st = lib.fgb_filter_record(C.addressof(buf), len(record), C.byref(fp), C.addressof(masked), C.addressof(status))
        if st != _l.FGB_OK:
            raise _l.FgbError(st, "fgb_filter_record")
        return status.value, masked.value


class _Caller:
    """Shared plumbing over fgb_caller_* (add_group / flush / statistics)."""

    def _create(self, o: "_l.FgbCallerOptions", device: int):
        self._h = C.c_void_p()
        st = self._lib.fgb_caller_create(device, C.byref(o), C.byref(self._h))
        if st != _l.FGB_OK:
            self._h = C.c_void_p()
            raise _l.FgbError(st, "fgb_caller_create")


class VanillaUmiConsensusCaller(_Caller):
    """vanilla_caller.rs:344-455 (new) / :1477-1499 (consensus_reads), batched."""

    def __init__(self, read_name_prefix: str, read_group_id: str,
                 options: VanillaUmiConsensusOptions = VanillaUmiConsensusOptions(), device: int = 0,
                 tag: bytes = b"MI", cell_tag: bytes = b"", consensus_call_overlapping_bases: bool = False,
                 filter: "ConsensusFilter" = None, n_threads: int = 1, track_rejects: bool = False):
        self._lib = _l.load()
        self._prefix = read_name_prefix.encode()
        self._rg = read_group_id.encode()
        o = _l.FgbCallerOptions()
        o.mode = 0
        o.track_rejects = 1 if track_rejects else 0    # vanilla_caller.rs:371-374, 418-424
        o.error_rate_pre_umi = options.error_rate_pre_umi
        o.error_rate_post_umi = options.error_rate_post_umi
        o.min_input_base_quality = options.min_input_base_quality
        o.min_consensus_base_quality = options.min_consensus_base_quality
        o.produce_per_base_tags = 1 if options.produce_per_base_tags else 0
        o.trim = 1 if options.trim else 0
        o.consensus_call_overlapping_bases = 1 if consensus_call_overlapping_bases else 0
        o.n_threads = n_threads
        if filter is not None:           # `fgumi simplex | fgumi filter` in one pass
            o.filter_enabled = 1
            filter.fill(o.filter)
        o.min_reads = options.min_reads
        o.tag = tag
        o.cell_tag = cell_tag if cell_tag else b"\0\0"
        o.read_name_prefix = self._prefix
        o.read_group_id = self._rg
        self._create(o, device)

    def close(self):
        if self._h:
            self._lib.fgb_caller_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st: int, where: str):
        if st != _l.FGB_OK:
            buf = C.create_string_buffer(512)
            self._lib.fgb_caller_last_error(self._h, buf, 512)
            raise _l.FgbError(st, where, buf.value.decode(errors="replace"))

    def add_group(self, records: Sequence[bytes]):
        """Queue one MI group (raw BAM records without the block_size prefix)."""
        if not records:
            return
        blob = b"".join(records)
        off = np.zeros(len(records) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(r) for r in records])
        buf = np.frombuffer(blob, dtype=np.uint8)
        self._check(self._lib.fgb_caller_add_group(self._h, buf.ctypes.data, off.ctypes.data, len(records)),
                    "fgb_caller_add_group")

    def add_groups(self, groups: Sequence[Sequence[bytes]]):
        """Queue many MI groups with one call (fgb_caller_add_groups); the per-group host work runs
        on `n_threads` threads when the caller was created with n_threads > 1."""
        groups = [g for g in groups if g]
        if not groups:
            return
        recs = [r for g in groups for r in g]
        blob = np.frombuffer(b"".join(recs), dtype=np.uint8)
        off = np.zeros(len(recs) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(r) for r in recs])
        grp = np.zeros(len(groups) + 1, dtype=np.uint64)
        grp[1:] = np.cumsum([len(g) for g in groups])
        self._check(self._lib.fgb_caller_add_groups(self._h, blob.ctypes.data, off.ctypes.data, grp.ctypes.data,
                                                    len(groups)), "fgb_caller_add_groups")

    def pending(self) -> Dict[str, object]:
        """What is queued for the next flush (fgb_caller_pending), copied out: per unit the list of
        (bases, quals) source rows and cons_len, plus the duplex / CODEC jobs.  Works on a planning-only
        caller (device=FGB_DEVICE_NONE) too."""
        from .engine import UNIT_DTYPE, DUPLEX_JOB_DTYPE, CODEC_JOB_DTYPE
        b = _l.FgbBatch()
        dj, cj, ndj, ncj = C.c_void_p(), C.c_void_p(), C.c_uint64(), C.c_uint64()
        self._check(self._lib.fgb_caller_pending(self._h, C.byref(b), C.addressof(dj), C.addressof(ndj),
                                                 C.addressof(cj), C.addressof(ncj)), "fgb_caller_pending")
        arr = lambda ptr, n, dt: (np.frombuffer(C.string_at(ptr, int(n) * np.dtype(dt).itemsize), dtype=dt).copy()
                                  if n else np.zeros(0, dt))
        units = arr(b.units, b.n_units, UNIT_DTYPE)
        reads = arr(b.reads, b.n_reads, np.uint64)
        bases, quals = arr(b.bases, b.n_bytes, np.uint8), arr(b.quals, b.n_bytes, np.uint8)
        ends = list(units["read_begin"][1:]) + [int(b.n_reads)]
        out_units = []
        for u in range(len(units)):
            rows = []
            for r in range(int(units["read_begin"][u]), int(ends[u])):
                off, ln = int(reads[r]) >> 16, int(reads[r]) & 0xFFFF
                rows.append((bytes(bases[off:off + ln]), bytes(quals[off:off + ln])))
            out_units.append({"rows": rows, "cons_len": int(units["cons_len"][u])})
        return {"units": out_units, "duplex_jobs": arr(dj.value, ndj.value, DUPLEX_JOB_DTYPE),
                "codec_jobs": arr(cj.value, ncj.value, CODEC_JOB_DTYPE), "n_out": int(b.n_out)}

    def flush(self) -> ConsensusOutput:
        data, n, cnt = C.c_void_p(), C.c_uint64(), C.c_uint64()
        self._check(self._lib.fgb_caller_flush(self._h, C.byref(data), C.byref(n), C.byref(cnt)),
                    "fgb_caller_flush")
        raw = C.string_at(data.value, n.value) if n.value else b""
        return ConsensusOutput(raw, int(cnt.value))

    def take_rejects(self) -> List[bytes]:
        """take_rejected_reads (vanilla_caller.rs:513-516): the raw records rejected since the last take (callers
        created with track_rejects), in reject-site order."""
        data, n, cnt = C.c_void_p(), C.c_uint64(), C.c_uint64()
        self._check(self._lib.fgb_caller_take_rejects(self._h, C.byref(data), C.byref(n), C.byref(cnt)),
                    "fgb_caller_take_rejects")
        raw = C.string_at(data.value, n.value) if n.value else b""
        out, p = [], 0
        while p < len(raw):
            size = int.from_bytes(raw[p:p + 4], "little")
            out.append(raw[p + 4:p + 4 + size])
            p += 4 + size
        assert len(out) == cnt.value
        return out

    def consensus_reads_batch(self, groups: Iterable[Sequence[bytes]]) -> ConsensusOutput:
        for g in groups:
            self.add_group(g)
        return self.flush()

    def statistics(self) -> Dict[str, int]:
        arr = (C.c_uint64 * _l.FGB_NSTATS)()
        self._check(self._lib.fgb_caller_stats(self._h, arr), "fgb_caller_stats")
        d = dict(zip(_l.STAT_NAMES, [int(x) for x in arr]))
        return d


class DuplexConsensusCaller(VanillaUmiConsensusCaller):
    """DuplexConsensusCaller::new (duplex_caller.rs:358-434): min_reads = (total, xy, yx)."""

    def __init__(self, read_name_prefix: str, read_group_id: str, min_reads=(1, 1, 1),
                 error_rate_pre_umi: int = 45, error_rate_post_umi: int = 40,
                 min_input_base_quality: int = 10, produce_per_base_tags: bool = True,
                 trim: bool = False, device: int = 0, cell_tag: bytes = b"",
                 consensus_call_overlapping_bases: bool = False, n_threads: int = 1,
                 filter: "DuplexConsensusFilter" = None):
        self._lib = _l.load()
        self._prefix = read_name_prefix.encode()
        self._rg = read_group_id.encode()
        o = _l.FgbCallerOptions()
        o.mode = 1
        if filter is not None:           # `fgumi duplex | fgumi filter` in one pass
            o.filter_enabled = 1
            filter.fill(o.duplex_filter)
        o.n_threads = n_threads
        o.consensus_call_overlapping_bases = 1 if consensus_call_overlapping_bases else 0
        o.error_rate_pre_umi = error_rate_pre_umi
        o.error_rate_post_umi = error_rate_post_umi
        o.min_input_base_quality = min_input_base_quality
        o.min_consensus_base_quality = 2
        o.produce_per_base_tags = 1 if produce_per_base_tags else 0
        o.trim = 1 if trim else 0
        o.min_reads, o.min_xy_reads, o.min_yx_reads = min_reads
        o.tag = b"MI"
        o.cell_tag = cell_tag if cell_tag else b"\0\0"
        o.read_name_prefix = self._prefix
        o.read_group_id = self._rg
        self._create(o, device)


class CodecConsensusCaller(VanillaUmiConsensusCaller):
    """CodecConsensusCaller::new (codec_caller.rs:306-370) with CodecConsensusOptions (:99-166).
    `max_reads_per_strand` (seeded down-sampling) is not offered; None is the reference default."""

    def __init__(self, read_name_prefix: str, read_group_id: str, min_reads_per_strand: int = 1,
                 min_duplex_length: int = 1, error_rate_pre_umi: int = 45, error_rate_post_umi: int = 40,
                 single_strand_qual=None, outer_bases_qual=None, outer_bases_length: int = 5,
                 max_duplex_disagreements=None, max_duplex_disagreement_rate: float = 1.0,
                 produce_per_base_tags: bool = False, device: int = 0, cell_tag: bytes = b"",
                 n_threads: int = 1):
        self._lib = _l.load()
        self._prefix = read_name_prefix.encode()
        self._rg = read_group_id.encode()
        o = _l.FgbCallerOptions()
        o.mode = 2
        o.n_threads = n_threads
        o.error_rate_pre_umi = error_rate_pre_umi
        o.error_rate_post_umi = error_rate_post_umi
        o.min_input_base_quality = 10            # carried by the options, unused on this path
        o.min_consensus_base_quality = 0
        o.produce_per_base_tags = 1 if produce_per_base_tags else 0
        o.trim = 0
        o.min_reads = min_reads_per_strand
        o.min_duplex_length = min_duplex_length
        o.codec.single_strand_qual = -1 if single_strand_qual is None else single_strand_qual
        o.codec.outer_bases_qual = -1 if outer_bases_qual is None else outer_bases_qual
        o.codec.outer_bases_length = outer_bases_length
        o.codec.max_duplex_disagreements = 0xFFFFFFFF if max_duplex_disagreements is None \
            else max_duplex_disagreements
        o.codec.max_duplex_disagreement_rate = max_duplex_disagreement_rate
        o.tag = b"MI"
        o.cell_tag = cell_tag if cell_tag else b"\0\0"
        o.read_name_prefix = self._prefix
        o.read_group_id = self._rg
        self._create(o, device)


def apply_overlapping_consensus(records: Sequence[bytes], agreement: int = 0, disagreement: int = 0):
    """apply_overlapping_consensus (overlapping.rs:625-667) on one MI group.  Returns the rewritten
    records and (overlapping_bases, bases_agreeing, bases_disagreeing, bases_corrected)."""
    lib = _l.load()
    if not records:
        return [], (0, 0, 0, 0)
    blob = np.frombuffer(b"".join(records), dtype=np.uint8).copy()
    off = np.zeros(len(records) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(r) for r in records])
    stats = np.zeros(4, dtype=np.uint64)
    st = lib.fgb_overlap_apply_group(blob.ctypes.data, off.ctypes.data, len(records), agreement, disagreement,
                                     stats.ctypes.data)
    if st != _l.FGB_OK:
        raise _l.FgbError(st, "fgb_overlap_apply_group")
    out = [blob[int(off[i]):int(off[i + 1])].tobytes() for i in range(len(records))]
    return out, tuple(int(x) for x in stats)
