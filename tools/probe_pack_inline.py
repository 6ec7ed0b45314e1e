"""A/B of the inline transport packing knobs of HotPath.run_host on the bench workload (64 Mbp, 50x ONT R9):
group size, pack threads, CIGAR packed or plain. Prints ms per call (median of 5 after 3 warm-ups)."""
import os, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pepper_thesis_b200 import synth, models, pipeline, device as dev
mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 64
n_regions = int(mbp * 10)
batch = synth.generate("ont_r9", n_regions * 100000 + 1000, 50.0, seed=1, num_regions=n_regions, threads=16)
batch.scan_min_qual(16)
batch.pin_plain(with_quals=False)
thr = synth.PROFILES["ont_r9"].thresholds
model = models.TransducerGRU(26, 1, 256, 28, 3, True).load_state_dict(models.random_variant_state_dict(0))
def run(g, inline, threads=0, cigar=0, ahead=4, taper=True):
    hp = pipeline.HotPath(model, thr, "cuda:0", group_regions=g, pack_inline=inline, pack_threads=threads, taper=taper,
                          pack_cigar=bool(cigar), host_ahead=ahead)
    ts = []
    for it in range(8):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        hp.run_host(batch)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print("group %3d inline %d threads %2d cigar16 %d ahead %d taper %d: %.1f ms (%s) h2d %.2f GB" % (
        g, inline, hp.pack_threads, cigar, ahead, taper, statistics.median(ts[3:]), " ".join("%.0f" % t for t in ts[3:]), hp.last_h2d_bytes / 1e9), flush=True)
    del hp; torch.cuda.empty_cache()
for g in (32, 48, 64):
    for t in (12, 14, 15):
        run(g, True, t, 0, ahead=2)
run(48, True, 14, 1, ahead=2)
