"""A/B of the duplex pipeline forms: python scripts/duplex_ab.py [molecules]  (FGUMI_B200_LIB selects the build)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fgumi_b200 as fg
from fgumi_b200 import benchlegs
M = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
r = benchlegs.duplex_leg(torch, fg, "cuda:0", 0, M)
print(json.dumps({"lib": os.path.basename(os.environ.get("FGUMI_B200_LIB", "default")), "molecules": M, "fused_ms": round(r["epilogue"]["ms"], 3),
                  "k1_ms": round(r["two_kernels"]["k1_ms"], 3), "k2_ms": round(r["two_kernels"]["k2_ms"], 3),
                  "fused_frac": round(r["epilogue"]["frac"], 3), "two_frac": round(r["two_kernels"]["frac"], 3), "equal": r["epilogue"]["equals_two_kernel_form"]}))
