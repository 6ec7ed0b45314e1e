"""Wall time of repeated HotPath.run_host calls (host buffers in, host results out)."""
import gc, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pepper_thesis_b200 import synth, models, pipeline
mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 64
g = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n_regions = int(mbp * 10)
batch = synth.generate("ont_r9", n_regions * 100000 + 1000, 50.0, seed=1, num_regions=n_regions, pinned=True)
batch.pack_wire(pinned=True)
model = models.TransducerGRU(26, 1, 256, 28, 3, True).load_state_dict(models.random_variant_state_dict(0))
hp = pipeline.HotPath(model, synth.PROFILES["ont_r9"].thresholds, "cuda:0", group_regions=g)
for mode in ("gc on", "gc off"):
    if mode == "gc off":
        gc.collect(); gc.disable()
    ts = []
    for it in range(8):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        hp.run_host(batch)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(mode, " ".join("%.1f" % t for t in ts), flush=True)
