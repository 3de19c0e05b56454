"""Summarise one `ncu --set full` report (any kernel) as markdown: duration, DRAM traffic and throughput, issue
rates, occupancy, warp-stall breakdown, SASS opcode mix.
usage: python scripts/ncu_kernel_summary.py REPORT.ncu-rep [algorithmic_bytes] > profiles/rNN_<kernel>_ncu.md"""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]
abytes = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def num(name, default=None):
    if name not in m:
        return default
    v, u = m[name]
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return default
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12, "ms": 1e-3, "us": 1e-6, "ns": 1e-9,
             "s": 1, "Gbyte/s": 1e9, "Tbyte/s": 1e12, "Mbyte/s": 1e6}.get(u, 1)
    return x * scale


name = m.get("Kernel Name", ("?", ""))[0]
dur = num("gpu__time_duration.sum")
rd, wr = num("dram__bytes_read.sum", 0.0), num("dram__bytes_write.sum", 0.0)
print(f"# ncu summary: `{name}`\n")
print(f"source: `{rep.split('/')[-1]}` (`ncu --set full --clock-control none --import-source on`, one launch; times under the "
      "profiler are not bench values)\n")
print("| metric | value |\n|---|---|")
print(f"| grid x block | {m.get('Grid Size', ('?',))[0]} x {m.get('Block Size', ('?',))[0]} |")
print(f"| registers / thread | {m.get('Registers Per Thread', ('?',))[0]} |")
print(f"| duration | {dur * 1e3:.3f} ms |")
print(f"| DRAM read + write | {rd / 1e9:.3f} + {wr / 1e9:.3f} GB = {(rd + wr) / dur / 1e9:.0f} GB/s |")
if abytes:
    print(f"| algorithmic bytes | {abytes / 1e9:.3f} GB ({(rd + wr) / abytes:.2f} x as DRAM traffic; {abytes / dur / 1e9:.0f} GB/s algorithmic) |")
for key, label in (("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
                   ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
                   ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
                   ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
                   ("sm__inst_executed.sum", "warp instructions executed"),
                   ("smsp__inst_executed.avg.per_cycle_active", "IPC per SM sub-partition (active)"),
                   ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
                   ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per warp instruction"),
                   ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts")):
    if key in m:
        print(f"| {label} | {m[key][0]} |")
print("\n## warp stall reasons (share of issue-slot samples)\n")
st = {}
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        try:
            st[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(m[h][0].replace(",", ""))
        except ValueError:
            pass
tot = sum(st.values()) or 1.0
print("| reason | warps per issue | share |\n|---|---|---|")
for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]:
    print(f"| {k} | {v:.2f} | {100 * v / tot:.1f} % |")
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(sass)))
try:
    hi = next(i for i, r in enumerate(srows) if "Source" in r and any("Instructions Executed" in c for c in r))
    sh = srows[hi]
    ci, cs = sh.index("Source"), next(i for i, c in enumerate(sh) if c.strip() == "Instructions Executed")
    ops = collections.Counter()
    for r in srows[hi + 1:]:
        if len(r) <= max(ci, cs):
            continue
        try:
            n = float(r[cs].replace(",", ""))
        except ValueError:
            continue
        toks = r[ci].split()
        while toks and toks[0].startswith("@"):
            toks = toks[1:]
        if toks:
            ops[toks[0].split(".")[0]] += n
    tot = sum(ops.values()) or 1.0
    print("\n## SASS opcode mix (warp instructions executed)\n\n| opcode | executed | share |\n|---|---|---|")
    for k, v in ops.most_common(14):
        print(f"| {k} | {v:.0f} | {100 * v / tot:.1f} % |")
except StopIteration:
    pass
