"""Ad-hoc timing probe of the summary chain with device-resident inputs (not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pepper_thesis_b200 import synth, device, capi

nreg = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cov = float(sys.argv[2]) if len(sys.argv) > 2 else 50.0
t = time.time()
b = synth.generate("ont_r9", nreg * 100000, cov, seed=1)
print("gen %.1fs regions=%d reads=%d bases=%.1fM ops=%.1fM" % (time.time() - t, b.n_regions, b.n_reads, b.n_bases / 1e6, b.n_ops / 1e6), flush=True)
thr = synth.PROFILES["ont_r9"].thresholds
b.scan_min_qual()                                  # PV_NO_ALLQ=1 keeps the quality loads for an A/B
db = device.DeviceBatch(b)
cap = max(4096, b.total_positions // 50)
ws = device.SummaryWorkspace.for_batch(db, cap)
torch.cuda.synchronize()
lib = capi.load()
for it in range(5):
    lib.pv_profile_reset(); lib.pv_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    device.summary_regions(db, thr, ws)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    lib.pv_profile_enable(0)
    prof = {k: round(v[0], 3) for k, v in capi.profile_collect().items() if v[0] > 0}
    K = int(ws.count.item())
    ab = b.algorithmic_bytes(K)
    print("ctr", ws.ws[:32].view(torch.int32).tolist())
    print("iter %d: %.3f ms  K=%d status=%d  %.1f Mbp/s  %.1f GB/s algorithmic (%.3f of 6536)" % (
        it, ms, K, ws.status(), b.candidate_bp / ms / 1e3, ab / ms / 1e6, ab / ms / 1e6 / 6536), prof, flush=True)
