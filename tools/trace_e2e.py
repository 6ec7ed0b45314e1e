"""Timeline of HotPath.run_host: when each group's upload and kernels start / end (CUDA events) and when the host issued them."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pepper_thesis_b200 import synth, models, pipeline, device as dev

mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 64
g = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n_regions = int(mbp * 10)
batch = synth.generate("ont_r9", n_regions * 100000 + 1000, 50.0, seed=1, num_regions=n_regions, pinned=True)
batch.pack_wire(pinned=True)
thr = synth.PROFILES["ont_r9"].thresholds
if os.environ.get("PV_TRACE_QP", "0") == "1":        # qualities as threshold predicates (pv_pack_quals_pred)
    batch.pack_quals_pred(thr.min_snp_baseq, thr.min_indel_baseq, pinned=True)
    batch.quals_packed, batch.qual_bits = None, 0
    if os.environ.get("PV_TRACE_BASES2", "1") == "1":
        batch.pack_bases2(pinned=True)
        batch.bases_patch, batch.read_patch_off = None, None
model = models.TransducerGRU(26, 1, 256, 28, 3, True).load_state_dict(models.random_variant_state_dict(0))
hp = pipeline.HotPath(model, thr, "cuda:0", group_regions=g, taper=os.environ.get("PV_TRACE_TAPER", "1") == "1")
for _ in range(4):
    hp.run_host(batch)
torch.cuda.synchronize()
# instrument
log = []
orig_init = dev.DeviceBatch.__init__
def traced_init(self, *a, **k):
    s = torch.cuda.current_stream()
    e0 = torch.cuda.Event(enable_timing=True); e0.record(s)
    t0 = time.perf_counter()
    orig_init(self, *a, **k)
    t1 = time.perf_counter()
    e1 = torch.cuda.Event(enable_timing=True); e1.record(s)
    log.append(("upload", self.host.n_regions, t0, t1, e0, e1))
dev.DeviceBatch.__init__ = traced_init
orig_unpack = dev.DeviceBatch.unpack
def traced_unpack(self):
    s = torch.cuda.current_stream()
    e0 = torch.cuda.Event(enable_timing=True); e0.record(s)
    t0 = time.perf_counter()
    orig_unpack(self)
    t1 = time.perf_counter()
    e1 = torch.cuda.Event(enable_timing=True); e1.record(s)
    log.append(("unpack", self.host.n_regions, t0, t1, e0, e1))
dev.DeviceBatch.unpack = traced_unpack
orig_sum = hp.summarize
def traced_sum(db):
    s = torch.cuda.current_stream()
    e0 = torch.cuda.Event(enable_timing=True); e0.record(s)
    t0 = time.perf_counter()
    r = orig_sum(db)
    t1 = time.perf_counter()
    e1 = torch.cuda.Event(enable_timing=True); e1.record(s)
    log.append(("summary", db.host.n_regions, t0, t1, e0, e1))
    return r
hp.summarize = traced_sum
for nm in ("_push", "_drain"):
    def wrap(nm):
        f = getattr(hp, nm)
        def g(*a, **k):
            t0 = time.perf_counter(); r = f(*a, **k); t1 = time.perf_counter()
            if (t1 - t0) > 1e-3: log.append((nm, 0, t0, t1, None, None))
            return r
        return g
    setattr(hp, nm, wrap(nm))
orig_inf = model.infer_windows
def traced_inf(*a, **k):
    t0 = time.perf_counter(); r = orig_inf(*a, **k); t1 = time.perf_counter()
    log.append(("infer", int(a[0].shape[0]), t0, t1, None, None))
    return r
model.infer_windows = traced_inf
torch.cuda.synchronize()
base = torch.cuda.Event(enable_timing=True); base.record()
T0 = time.perf_counter()
hp.run_host(batch)
torch.cuda.synchronize()
T1 = time.perf_counter()
print("total %.2f ms" % ((T1 - T0) * 1e3))
for kind, n, t0, t1, e0, e1 in log:
    if e0 is None:
        print("%-8s n=%5d host %7.2f .. %7.2f" % (kind, n, (t0 - T0) * 1e3, (t1 - T0) * 1e3))
    else:
        print("%-8s n=%5d host %7.2f .. %7.2f   gpu %7.2f .. %7.2f" % (kind, n, (t0 - T0) * 1e3, (t1 - T0) * 1e3, base.elapsed_time(e0), base.elapsed_time(e1)))
