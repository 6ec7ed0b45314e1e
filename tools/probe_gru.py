"""Ad-hoc timing probe of the polisher GRU model (device-resident inputs), per kernel family."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pepper_thesis_b200 import models, capi
ns = [int(v) for v in sys.argv[1:]] or [256, 16384]
m = models.PolisherTransducerGRU().load_state_dict(models.random_polisher_state_dict(0))
lib = capi.load()
for n in ns:
    x = torch.randint(0, 31, (n, 100, 10), dtype=torch.uint8).cuda()
    h0 = torch.zeros(n, 2, 128, device="cuda")
    m(x, h0); torch.cuda.synchronize()
    for it in range(2):
        lib.pv_profile_reset(); lib.pv_profile_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m(x, h0); e1.record(); torch.cuda.synchronize()
        lib.pv_profile_enable(0)
        ms = e0.elapsed_time(e1)
        prof = {k: (round(v[0], 3), v[1]) for k, v in capi.profile_collect().items() if v[0] > 0}
        print("n=%d %.3f ms  %.2f Mwin/s  %.1f TFLOP/s  %s" % (n, ms, n / ms / 1e3, n * 80.44e6 / ms / 1e9, prof), flush=True)
