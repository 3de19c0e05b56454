#!/usr/bin/env python
"""bench.py — consensus reads/sec of the UMI-consensus hot path (BASELINE.json metric).

N=1 workload = BASELINE.json configs[1]: simplex consensus, 10 M families, depth 8, 150 bp,
substitution error rate 1e-3, synthetic (generator: fgumi_b200/synth.py), one B200.
A "step" = one pass of the hot path (ONE kernel launch through fgb_vote_device) over the whole
batch, inputs resident in HBM.  N>1: every rank owns its own 10 M-family shard (families shard
trivially; weak scaling), no data-path collective; the end-of-run device counters are summed with
one NCCL all-reduce.

Keys beyond the base contract:
  roofline      dominant kernel (vote_kernel): algorithmic bytes per launch / CUDA-event time
  cpu_baseline  the CPU oracle (a C++ restatement of fgumi 0.2.0; the Rust reference cannot be
                built here) on a bounded sample, all host threads, rank 0, N=1 only
  e2e           same metric through the host-buffer C-ABI call (fgb_submit/fgb_wait) from pinned
                host memory, H2D + vote + D2H inside the timed region

`--impl reference` times the reference's CPU algorithm (the oracle, all host threads) on a
bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "consensus reads/sec (simplex 150bp, depth-8 families)"
UNIT = "consensus_reads/s"
DEPTH, READ_LEN, ERR = 8, 150, 1e-3
PARAMS = dict(error_rate_pre_umi=45, error_rate_post_umi=40, min_reads=1,
              min_consensus_base_quality=2)


def algorithmic_bytes(n_units: int, n_reads: int, sum_len: int, sum_cons: int) -> int:
    """per unit 2*sum(len) in + 6*cons_len out + 8*(n_reads+1) + 8 index bytes."""
    return 2 * sum_len + 6 * sum_cons + 8 * (n_reads + n_units) + 8 * n_units


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic():
    """dram bytes per launch of vote_kernel from the committed ncu --set full capture, if any."""
    p = os.path.join(ROOT, "profiles", "vote_kernel_traffic.json")
    try:
        with open(p) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    """`nvidia-smi -lms 20` in the background.  It is started before the warm-up (the tool needs a few
    hundred ms to come up -- longer than a short timed region) and every sample is stamped on arrival;
    begin() / end() bracket the timed region and the summary uses the samples that fall inside it
    (all samples under load -- warm-up included -- if none does).  Never raises."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, cmd=None):
        self.samples = []          # (arrival time, line)
        self.proc = None
        self.index = index
        self.t0 = self.t1 = None
        self.cmd = cmd or ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                           "--format=csv,noheader,nounits", "-lms", "20"]

    def start(self):
        try:
            self.proc = subprocess.Popen(self.cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        try:
            for line in self.proc.stdout:
                self.samples.append((time.perf_counter(), line.strip()))
        except Exception:
            pass

    def wait_ready(self, timeout: float = 2.0):
        """Block (bounded) until the tool has delivered its first sample."""
        try:
            t_end = time.perf_counter() + timeout
            while self.proc and not self.samples and time.perf_counter() < t_end and self.proc.poll() is None:
                time.sleep(0.01)
        except Exception:
            pass

    def begin(self):
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        try:
            if self.t1 is None:
                self.t1 = time.perf_counter()
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

            def parse(rows):
                mhz, mx, reasons = [], None, set()
                for _, s in rows:
                    f = [x.strip() for x in s.split(",")]
                    if len(f) < 6:
                        continue
                    try:
                        mhz.append(float(f[0])); mx = float(f[1])
                    except ValueError:
                        continue
                    for nm, v in zip(names, f[2:6]):
                        if v.lower().startswith("active"):
                            reasons.add(nm)
                return mhz, mx, reasons
            rows = list(self.samples)
            inside = [r for r in rows if self.t0 is not None and self.t0 - 0.02 <= r[0] <= self.t1 + 0.02]
            mhz, mx, reasons = parse(inside)
            window = "timed region"
            if not mhz:
                mhz, mx, reasons = parse(rows)
                window = "warm-up + timed region"
            return {"sm_mhz": float(np.median(mhz)) if mhz else None, "sm_max_mhz": mx,
                    "reasons": sorted(reasons), "samples": len(mhz), "window": window}
        except Exception as e:      # the clocks line must never cost the bench line
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling failed: %r" % (e,)]}


def cpu_oracle_rate(n_units: int, threads: int, seed: int = 1234, min_seconds: float = 10.0):
    """Times the CPU oracle (TEST INFRASTRUCTURE used only as the measured baseline) on a sample."""
    import fgumi_b200 as fg
    from fgumi_b200 import synth
    from tests import oracle_lib as O
    b, q = synth.host_pileup(n_units, DEPTH, READ_LEN, ERR, seed=seed)
    batch = fg.pack_uniform(b, q, 1)
    outs = O.alloc_outputs(batch)
    threads, probe = best_thread_count(batch, outs, threads)
    O.simplex_batch(batch, 45, 40, 1, 2, threads, outs)    # warm (page faults, thread start)
    reps, t = 0, time.perf_counter()
    while True:                                            # ~10 s of CPU work
        O.simplex_batch(batch, 45, 40, 1, 2, threads, outs)
        reps += 1
        dt = time.perf_counter() - t
        if dt >= min_seconds or reps >= 2000:
            break
    return n_units * reps / dt, dt, reps, threads, probe


def best_thread_count(batch, outs, max_threads: int):
    """The host may expose more logical CPUs than it schedules well (SMT, shared boxes): probe a few
    thread counts with one pass each and keep the fastest, so the baseline is the CPU at its best."""
    from tests import oracle_lib as O
    cands = sorted({max(1, max_threads >> k) for k in range(0, 4)} | {1})
    O.simplex_batch(batch, 45, 40, 1, 2, max_threads, outs)      # first-touch
    rates = {}
    for th in cands:
        t = time.perf_counter()
        O.simplex_batch(batch, 45, 40, 1, 2, th, outs)
        rates[th] = batch.n_units / (time.perf_counter() - t)
    best = max(rates, key=rates.get)
    return best, {str(k): round(v) for k, v in rates.items()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = int(os.environ.get("FGB_REF_SAMPLE_UNITS", "400000"))
    import fgumi_b200 as fg
    from fgumi_b200 import synth
    from tests import oracle_lib as O
    b, q = synth.host_pileup(n, DEPTH, READ_LEN, ERR, seed=1234)
    batch = fg.pack_uniform(b, q, 1)
    outs = O.alloc_outputs(batch)
    threads, probe = best_thread_count(batch, outs, threads)
    for _ in range(args.warmup):
        O.simplex_batch(batch, 45, 40, 1, 2, threads, outs)
    t = time.perf_counter()
    for _ in range(args.steps):
        O.simplex_batch(batch, 45, 40, 1, 2, threads, outs)
    dt = time.perf_counter() - t
    v = n * args.steps / dt
    sample = (f"{n} families depth {DEPTH} x {READ_LEN} bp per step (same generator as the GPU arm); "
              f"threads probed (families/s): {probe}")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "simplex consensus, depth 8, 150bp, error-rate 1e-3 "
                               "(BASELINE.json configs[1]), bounded sample", "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample + "; oracle = C++ restatement of fgumi 0.2.0 "
                                            "(Rust toolchain absent), std::thread over families"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fgumi_b200", choices=["fgumi_b200", "reference"])
    ap.add_argument("--units", type=int, default=int(os.environ.get("FGB_BENCH_UNITS", "10000000")),
                    help="families per GPU per step (BASELINE config: 10 M)")
    ap.add_argument("--e2e-units", type=int, default=int(os.environ.get("FGB_E2E_UNITS", "1000000")))
    ap.add_argument("--cpu-units", type=int, default=int(os.environ.get("FGB_CPU_UNITS", "400000")))
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    if args.impl == "reference":
        import __graft_entry__ as g
        g.build()
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build()
    import fgumi_b200 as fg
    from fgumi_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: fgumi_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))

    eng = fg.Engine(device=local, **PARAMS)
    U = args.units
    depths = np.full(U, DEPTH, dtype=np.int64)
    tb = synth.device_batch(torch, dev, depths, READ_LEN, ERR, seed=42 + rank, min_reads=1)
    out = fg.DeviceColumns(tb.host.n_out, dev)
    bstruct, cstruct = tb.struct(), out.struct()
    stream = torch.cuda.current_stream().cuda_stream
    import ctypes as C
    lib = fg.lib.load()

    def step():
        st = lib.fgb_vote_device(eng._h, C.byref(bstruct), C.byref(cstruct), C.c_void_p(stream))
        if st != 0:
            raise fg.lib.FgbError(st, "fgb_vote_device")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clk = ClockSampler(local)
    clk.start()
    clk.wait_ready()
    for _ in range(args.warmup):
        step()
    barrier()
    eng.stats_reset()
    launches0 = eng.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    clk.begin()
    torch.cuda.nvtx.range_push("fgb_timed")
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    # K4: the only collective on the path — sum the device counters across ranks (in place)
    ctr = None
    if world > 1:
        class _DevPtr:   # zero-copy view of the engine's u64[FGB_NCOUNTERS] counter block
            __cuda_array_interface__ = {"shape": (fg.lib.FGB_NCOUNTERS,), "typestr": "<i8",
                                        "data": (eng.stats_device_ptr(), False), "version": 2}
        ctr = torch.as_tensor(_DevPtr(), device=dev)
        dist.all_reduce(ctr, op=dist.ReduceOp.SUM)
    barrier()
    torch.cuda.nvtx.range_pop()
    clk.end()
    clocks = clk.stop()
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    launches = eng.launch_count() - launches0
    stats_all = eng.stats()   # after the all-reduce every rank's block holds the global sums

    value = U * world * args.steps / (total_ms * 1e-3)
    # ---- roofline of the dominant kernel (one launch per step) ----
    h = tb.host
    abytes = algorithmic_bytes(h.n_units, h.n_reads, h.n_reads * READ_LEN, h.n_units * READ_LEN)
    k_ms = float(np.mean(per_launch_ms))
    achieved = abytes / (k_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak_gbs()
    tr = recorded_traffic()
    traffic = None
    if tr and tr.get("units") and tr.get("dram_bytes_per_launch"):
        traffic = tr["dram_bytes_per_launch"] * (h.n_units / tr["units"])
    roofline = {"bound": "hbm", "kernel": "vote_kernel", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "algorithmic_bytes_per_launch": abytes, "bytes_per_unit": abytes / h.n_units,
                "kernel_ms": k_ms, "peak_source": peak_src}

    # ---- e2e: host-buffer call, pinned host memory, H2D + vote + D2H inside the timed region ----
    e2e = None
    cpu = None
    if True:
        EU = min(args.e2e_units, U)
        hb = synth.make_descriptors(np.full(EU, DEPTH, dtype=np.int64), READ_LEN, 1)
        nb = hb.n_bytes
        pin = lambda n, dt: torch.empty(n, dtype=dt).pin_memory()
        pb, pq = pin(nb + 16, torch.uint8), pin(nb + 16, torch.uint8)
        pb[:nb].copy_(tb.bases[:nb]); pq[:nb].copy_(tb.quals[:nb])
        torch.cuda.synchronize()
        hb.bases = pb.numpy(); hb.quals = pq.numpy()
        ho = fg.HostColumns(pin(hb.n_out, torch.uint8).numpy(), pin(hb.n_out, torch.uint8).numpy(),
                            pin(hb.n_out, torch.int16).numpy().view(np.uint16),
                            pin(hb.n_out, torch.int16).numpy().view(np.uint16))
        preads = pin(len(hb.reads), torch.int64); preads.numpy()[:] = hb.reads.view(np.int64)
        hb.reads = preads.numpy().view(np.uint64)
        esteps = max(3, min(args.steps, 10))

        def timed_e2e(call, tag):
            for _ in range(args.warmup):
                call(); eng.wait()
            barrier()
            t0 = time.perf_counter()
            torch.cuda.nvtx.range_push(tag)
            for _ in range(esteps):
                call(); eng.wait()
            torch.cuda.synchronize()
            torch.cuda.nvtx.range_pop()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return EU * world * esteps / float(tt.item())

        desc_bytes = hb.n_reads * 8 + (hb.n_units + 1) * 16 + len(hb.tiles) * 32
        d2h = hb.n_out * 6
        # (1) the two-column layout (a base byte and a quality byte per observation)
        v_bytes = timed_e2e(lambda: eng.submit(hb, ho), "fgb_e2e_bytes")
        # (2) PACK8: one byte per observation, expanded on the device in front of the vote.  The
        #     host-side encode is part of source-read preparation, like building the rows themselves.
        packed = fg.pack8_encode(hb.bases[:nb], hb.quals[:nb])
        if packed is not None:
            pp = pin(nb + 16, torch.uint8)
            pp.numpy()[:nb] = packed
            ppn = pp.numpy()
            v_pack = timed_e2e(lambda: eng.submit_pack8(hb, ppn, ho), "fgb_e2e_pack8")
            # (2b) PACK8 in, narrow (u8 depth / errors) out: 4 instead of 6 bytes per position back
            ho8 = fg.HostColumns(ho.base, ho.qual, pin(hb.n_out, torch.uint8).numpy(),
                                 pin(hb.n_out, torch.uint8).numpy())
            v_pack8n = timed_e2e(lambda: eng.submit_ex(hb, ho8, packed=ppn, narrow=True), "fgb_e2e_pack8_u8")
            e2e = {"value": v_pack8n, "unit": UNIT, "h2d_bytes_per_step": int(nb + desc_bytes),
                   "d2h_bytes_per_step": int(hb.n_out * 4), "units_per_step": EU, "steps": esteps,
                   "api": "fgb_submit_ex(FGB_IN_PACK8, FGB_OUT_U8) + fgb_wait (pinned host buffers: "
                          "1 byte per observation in, 4 bytes per consensus position out)",
                   "pack8_u16": {"value": v_pack, "h2d_bytes_per_step": int(nb + desc_bytes),
                                 "d2h_bytes_per_step": int(d2h), "api": "fgb_submit_pack8 + fgb_wait"},
                   "two_column": {"value": v_bytes, "h2d_bytes_per_step": int(2 * nb + desc_bytes),
                                  "d2h_bytes_per_step": int(d2h), "api": "fgb_submit + fgb_wait"}}
            # (3) BAM4: 4-bit sequence + raw qualities, rows built on the device (forward-strand reads;
            #     the generator's masked bases are (N, Q2), so the min-quality mask leaves them as is).
            #     Informative extra leg: a failure here must not cost the line.
            try:
                from fgumi_b200.engine import RawColumns, RAW_READ_DTYPE, _NIBBLE
                Lp = (READ_LEN + 7) // 8 * 8
                nib = _NIBBLE[hb.bases[:nb]]
                seq4 = pin(nb // 2 + 32, torch.uint8)
                seq4.numpy()[:nb // 2] = (nib[0::2] << 4) | nib[1::2]
                rr = pin((hb.n_reads + 1) * 16, torch.uint8)
                rrv = rr.numpy().view(RAW_READ_DTYPE)
                rrv["src_off"][:hb.n_reads] = np.arange(hb.n_reads, dtype=np.uint64) * np.uint64(Lp)
                rrv["raw_len"][:hb.n_reads] = READ_LEN
                rrv["flags"][:] = 0
                rawc = RawColumns(seq4.numpy(), hb.quals, rrv, int(nb), 10)
                v_bam4 = timed_e2e(lambda: eng.submit_bam4(hb, rawc, ho), "fgb_e2e_bam4")
                e2e["bam4"] = {"value": v_bam4,
                               "h2d_bytes_per_step": int(nb + nb // 2 + hb.n_reads * 16 + desc_bytes),
                               "d2h_bytes_per_step": int(d2h),
                               "api": "fgb_submit_bam4 + fgb_wait (4-bit sequence + raw qualities)"}
            except Exception as ex:   # pragma: no cover
                e2e["bam4"] = {"error": repr(ex)[:200]}
        else:
            e2e = {"value": v_bytes, "unit": UNIT, "h2d_bytes_per_step": int(2 * nb + desc_bytes),
                   "d2h_bytes_per_step": int(d2h), "units_per_step": EU, "steps": esteps,
                   "api": "fgb_submit + fgb_wait (pinned host buffers)"}

    if rank == 0 and world == 1 and args.cpu_units > 0:
        threads = os.cpu_count() or 1
        v, dtc, reps, threads, probe = cpu_oracle_rate(args.cpu_units, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "logical_cpus": os.cpu_count(),
               "thread_probe": probe, "kind": "port",
               "sample": f"{reps} passes over {args.cpu_units} families depth {DEPTH} x {READ_LEN} bp, {dtc:.1f} s; "
                         "oracle = C++ restatement of fgumi 0.2.0 (no Rust toolchain), "
                         "std::thread over families"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8 in / u8+u16 out, f64 log-likelihood on the exact path",
            "data": "synthetic",
            "config": {"workload": "simplex consensus, 10M families depth=8, 150bp, error-rate 1e-3 "
                                   "(BASELINE.json configs[1])",
                       "families_per_gpu": U, "depth": DEPTH, "read_len": READ_LEN,
                       "error_rate": ERR, "params": "-1 45 -2 40 -m 10 --min-consensus-base-quality 2 "
                                                    "--min-reads 1, overlapping pre-pass off",
                       "parallelism": f"range-shard x{world}",
                       "l2": "inputs (24 GB/GPU) larger than L2; no flush needed"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks,
            "counters": stats_all,
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
