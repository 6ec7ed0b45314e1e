"""Ad-hoc timing probe of the polisher summary chain (8 Mbp at 50x, device-resident reads)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pepper_thesis_b200 import synth, polish, device as dev
pb = synth.generate("ont_r9", 8_000_000, 50.0, seed=6)
db = dev.DeviceBatch(pb); torch.cuda.synchronize()
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s = polish.PolishSummary(db); torch.cuda.synchronize(); t1 = time.perf_counter()
    im, pos, ids, regs = s.chunks(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("summary %.2f ms  chunks %.2f ms  rows %d" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, s.n_rows), flush=True)
