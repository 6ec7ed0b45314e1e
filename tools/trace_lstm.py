"""Debug: per-role SM-clock trace of CTA 0 of the last LSTM step kernel (library built with PV_NVCC_FLAGS=-DPV_TRACE)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pepper_thesis_b200 import models, capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
m = models.TransducerGRU().load_state_dict(models.random_variant_state_dict(0))
x = (-torch.randint(0, 50, (n, 33, 26))).to(torch.int16).cuda()
m.infer_windows(x); torch.cuda.synchronize()
lib = capi.load(); lib.pv_trace_ptr.restype = C.c_void_p
ptr = lib.pv_trace_ptr()
buf = torch.zeros(3 * 16 * 8, dtype=torch.int64)
C.cdll.LoadLibrary("libcudart.so").cudaMemcpy(C.c_void_p(buf.data_ptr()), C.c_void_p(ptr), buf.numel() * 8, 2)
t = buf.view(3, 16, 8).numpy()
t0 = t[0, 0, 0]
print("clock cycles relative to producer start (last kernel = last decoder step, CTA 0)")
for it in range(14):
    p, mm, e = t[0, it], t[1, it], t[2, it]
    print("tile %2d  prod[%6d..%6d]  mma[wait %6d start %6d commit %6d]  epi[idle %6d tfull %6d state %6d done %6d]" % (
        it, p[0] - t0, p[1] - t0, mm[0] - t0, mm[1] - t0, mm[2] - t0, e[0] - t0, e[1] - t0, e[2] - t0, e[3] - t0))
