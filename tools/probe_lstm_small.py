import sys, os, time
sys.path.insert(0, "/root/repo")
import torch
from pepper_thesis_b200 import models
m = models.TransducerGRU().load_state_dict(models.random_variant_state_dict(0))
for n in (256, 512, 1024, 4096):
    x = (-torch.randint(0, 50, (n, 33, 26))).to(torch.int16).cuda()
    outs = []
    for it in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e0.record(); p, a = m.infer_windows(x); e1.record(); torch.cuda.synchronize()
        outs.append((round(e0.elapsed_time(e1), 3), round((time.perf_counter() - t0) * 1e3, 3), p.clone()))
    same = all(torch.equal(outs[0][2], o[2]) for o in outs)
    print("n=%d device ms %s wall ms %s identical=%s" % (n, [o[0] for o in outs], [o[1] for o in outs], same), flush=True)
