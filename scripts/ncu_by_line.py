"""Aggregate an .ncu-rep's source page by CUDA source line: share of stall samples, of executed
warp instructions, and the average active threads per instruction.
usage: python scripts/ncu_by_line.py report.ncu-rep [top_n]"""
import csv, io, subprocess, sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
cur, hdr, agg = None, None, {}
for r in csv.reader(io.StringIO(txt)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 5 and r[0] == "Line No":
        hdr = r
        i_s, i_i, i_t = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        samp, inst, tinst = int(r[i_s] or 0), int(r[i_i] or 0), int(r[i_t] or 0)
        if samp or inst:
            a = agg.setdefault((cur, int(r[0])), [0, 0, 0, r[1]])
            a[0] += samp; a[1] += inst; a[2] += tinst
tot, toti = sum(a[0] for a in agg.values()) or 1, sum(a[1] for a in agg.values()) or 1
print(f"total samples {tot}  warp instructions {toti}")
key = (lambda x: -x[1][1]) if len(sys.argv) > 3 and sys.argv[3] == "inst" else (lambda x: -x[1][0])
for k, a in sorted(agg.items(), key=key)[:top]:
    print(f"{k[0]}:{k[1]:4d} samp {100 * a[0] / tot:5.1f}% inst {100 * a[1] / toti:5.1f}% thr/inst {a[2] / max(a[1], 1):5.1f} | {a[3].strip()[:88]}")
