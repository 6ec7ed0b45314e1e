"""Timeline of one HotPath.run_host call with inline packing: per group the packing job (host clock), the upload (events on
the copy stream) and the summary kernels (events on the compute stream), all relative to the call's start."""
import os, sys, time, faulthandler
faulthandler.dump_traceback_later(45, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pepper_thesis_b200 import synth, models, pipeline, device as dev
n_regions = 640
batch = synth.generate("ont_r9", n_regions * 100000 + 1000, 50.0, seed=1, num_regions=n_regions, threads=16)
batch.scan_min_qual(16); batch.pin_plain(with_quals=False)
thr = synth.PROFILES["ont_r9"].thresholds
model = models.TransducerGRU(26, 1, 256, 28, 3, True).load_state_dict(models.random_variant_state_dict(0))
g = int(sys.argv[1]) if len(sys.argv) > 1 else 48
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 15
hp = pipeline.HotPath(model, thr, "cuda:0", group_regions=g, pack_inline=True, pack_threads=threads, host_ahead=2)
for _ in range(4):
    hp.run_host(batch)
log = {"pack": [], "up": [], "sum": [], "collect": [], "infer": []}
T0 = [0.0]; E0 = [None]
orig_pack = pipeline._GroupPacker._pack
def pack(self, b, gr, slot):
    t = time.perf_counter(); r = orig_pack(self, b, gr, slot); log["pack"].append((gr, t - T0[0], time.perf_counter() - T0[0])); return r
pipeline._GroupPacker._pack = pack
orig_db = dev.DeviceBatch.__init__
def db_init(self, host, *a, **k):
    s = torch.cuda.current_stream(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t = time.perf_counter(); e0.record(s); orig_db(self, host, *a, **k); e1.record(s)
    log["up"].append((host.n_regions, t - T0[0], time.perf_counter() - T0[0], e0, e1))
dev.DeviceBatch.__init__ = db_init
orig_ls = pipeline.HotPath._launch_summary
def ls(self, db, slot):
    s = torch.cuda.current_stream(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t = time.perf_counter(); e0.record(s); r = orig_ls(self, db, slot); e1.record(s)
    log["sum"].append((db.host.n_regions, t - T0[0], time.perf_counter() - T0[0], e0, e1)); return r
pipeline.HotPath._launch_summary = ls
orig_cs = pipeline.HotPath._collect_summary
def cs(self, h):
    t = time.perf_counter(); r = orig_cs(self, h); log["collect"].append((t - T0[0], time.perf_counter() - T0[0])); return r
pipeline.HotPath._collect_summary = cs
orig_inf = model.infer_windows
def inf(x, **k):
    s = torch.cuda.current_stream(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t = time.perf_counter(); e0.record(s); r = orig_inf(x, **k); e1.record(s)
    log["infer"].append((x.shape[0], t - T0[0], time.perf_counter() - T0[0], e0, e1)); return r
model.infer_windows = inf
torch.cuda.synchronize()
E0[0] = torch.cuda.Event(enable_timing=True); E0[0].record(); torch.cuda.synchronize()
T0[0] = time.perf_counter()
hp.run_host(batch)
torch.cuda.synchronize(); t_end = time.perf_counter() - T0[0]
off = 0.0
print("call: %.1f ms; groups of %d, %d pack threads" % (t_end * 1e3, g, threads))
print("pack jobs (host ms):   " + "  ".join("%d:%.1f-%.1f" % (gr[1] - gr[0], a * 1e3, b * 1e3) for gr, a, b in log["pack"]))
print("uploads (host enqueue | device): " + "  ".join("%d:%.1f|%.1f-%.1f" % (n, a * 1e3, E0[0].elapsed_time(e0), E0[0].elapsed_time(e1)) for n, a, b, e0, e1 in log["up"]))
print("summary (host launch | device):  " + "  ".join("%d:%.1f|%.1f-%.1f" % (n, a * 1e3, E0[0].elapsed_time(e0), E0[0].elapsed_time(e1)) for n, a, b, e0, e1 in log["sum"]))
print("collect waits (host ms):  " + "  ".join("%.1f-%.1f" % (a * 1e3, b * 1e3) for a, b in log["collect"]))
print("inference (host launch | device): " + "  ".join("%d:%.1f|%.1f-%.1f" % (n, a * 1e3, E0[0].elapsed_time(e0), E0[0].elapsed_time(e1)) for n, a, b, e0, e1 in log["infer"]))
