"""Debug helper (GPU box): handcrafted pileups with quality bytes >= 128."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fgumi_b200 as fg
from tests import oracle_lib as O

def run(units, mq=0, device_path=False):
    batch = fg.pack_source_reads(units, 1)
    eng = fg.Engine(0, 45, 40, 1, mq)
    if device_path:
        db = fg.DeviceBatch(batch, "cuda:0"); dc = fg.DeviceColumns(batch.n_out, "cuda:0")
        eng.vote_device(db, dc, torch.cuda.current_stream().cuda_stream); torch.cuda.synchronize()
        out = dc.to_host()
    else:
        out = eng.vote(batch)
    st = eng.stats(); eng.close()
    ob, oq, od, oe, cl = O.simplex_batch(batch, 45, 40, 1, mq, 1)
    n = batch.n_out
    return list(out.base[:n]), list(out.qual[:n]), list(oq[:n]), st["exact_positions"]

for q in (23, 93, 100, 127, 128, 129, 200, 255):
    for mq in (0, 2):
        for dp in (False, True):
            u = [[(b"GGGG", bytes([q] * 4))] * 6]
            print("q", q, "mq", mq, "dev" if dp else "host", run(u, mq, dp))
u = [[(b"GGGG", bytes([231, 93, 247, 101])), (b"GGGG", bytes([30, 200, 30, 30])), (b"GGGG", bytes([30] * 4))] * 2]
print("mixed", run(u))
u = [[(b"G", bytes([200]))], [(b"N", bytes([196]))]]
print("single", run(u))
u = [[(b"GN", bytes([30, 196])), (b"G", bytes([30])), (b"G", bytes([30])), (b"G", bytes([30]))]]
print("ragged", run(u))
