"""A/B of the inline transport packing knobs of HotPath.run_host on the bench workload (64 Mbp, 50x ONT R9):
group size, pack threads, CIGAR packed or plain. Prints ms per call (median of 5 after 3 warm-ups)."""
import os, sys, time, statistics, faulthandler
faulthandler.dump_traceback_later(150, exit=True)
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
def make(g, threads, ahead=2, taper=True, inline=True):
    return pipeline.HotPath(model, thr, "cuda:0", group_regions=g, pack_inline=inline, pack_threads=threads, taper=taper, host_ahead=ahead)
configs = [(48, 15, "1"), (48, 15, "0"), (48, 14, "1"), (48, 7, "1"), (48, 7, "0"), (48, 3, "1"), (48, 3, "0"), (48, 1, "1"), (32, 7, "1"), (64, 7, "1")]
hps = {c: make(c[0], c[1]) for c in configs}
res = {c: [] for c in configs}
def call(c):
    os.environ["PV_PACK_MIX"] = c[2]
    hps[c].run_host(batch)
for c in configs:
    for _ in range(4):
        call(c)
for rnd in range(2):
    for c in configs:
        for _ in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            call(c)
            torch.cuda.synchronize(); res[c].append((time.perf_counter() - t0) * 1e3)
for c in configs:
    pk = hps[c]._packer
    print("group %3d threads %2d mix %s: median %.1f ms  (%s)  plain %d packed %d  pack %.2f ns/base wire %.3f ns/B  h2d %.2f GB" % (
        c[0], c[1], c[2], statistics.median(res[c]), " ".join("%.0f" % t for t in res[c]), pk.n_plain, pk.n_packed,
        (pk.pack_s_per_base or 0) * 1e9, (pk.wire_s_per_byte or 0) * 1e9, hps[c].last_h2d_bytes / 1e9), flush=True)
hp0 = make(128, 1, ahead=4, inline=False)
ts = []
for _ in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter(); hp0.run_host(batch); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print("plain upload, groups of 128: median %.1f ms" % statistics.median(ts[3:]))
