"""Ad-hoc timing probe of the LSTM model (device-resident windows)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pepper_thesis_b200 import models, capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
m = models.TransducerGRU().load_state_dict(models.random_variant_state_dict(0))
x = (-torch.randint(0, 50, (n, 33, 26))).to(torch.int16).cuda()
m.infer_windows(x); torch.cuda.synchronize()
lib = capi.load()
for it in range(int(os.environ.get("PV_PROBE_ITERS", "3"))):
    lib.pv_profile_reset(); lib.pv_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m.infer_windows(x); e1.record(); torch.cuda.synchronize()
    lib.pv_profile_enable(0)
    ms = e0.elapsed_time(e1)
    prof = {k: round(v[0], 3) for k, v in capi.profile_collect().items() if v[0] > 0}
    print("n=%d %.3f ms  %.2f Mwin/s  %.1f TFLOP/s  %s" % (n, ms, n / ms / 1e3, n * 161.33e6 / ms / 1e9, prof), flush=True)
