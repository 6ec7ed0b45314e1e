"""Debug: SM-clock trace of CTA 0 of the last GRU recurrence kernel (library built with PV_NVCC_FLAGS=-DPV_TRACE)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pepper_thesis_b200 import models, capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = models.PolisherTransducerGRU().load_state_dict(models.random_polisher_state_dict(0))
x = torch.randint(0, 31, (n, 100, 10), dtype=torch.uint8).cuda()
h0 = torch.zeros(n, 2, 128, device="cuda")
m(x, h0); m(x, h0); torch.cuda.synchronize()
lib = capi.load(); lib.pv_gru_trace_ptr.restype = C.c_void_p
ptr = lib.pv_gru_trace_ptr()
buf = torch.zeros(16 * 16, dtype=torch.int64)
C.cdll.LoadLibrary("libcudart.so").cudaMemcpy(C.c_void_p(buf.data_ptr()), C.c_void_p(ptr), buf.numel() * 8, 2)
t = buf.view(16, 16).numpy()
t0 = t[0, 1]
print("cycles relative to step 0 h_ready (decoder layer, CTA 0)")
print("step  mma[poll hready commitX commitY]  epiX[accfull gxfull tmem done]  epiY[accfull gxfull tmem done]  arrive")
for s in range(14):
    r = t[s] - t0
    print("%2d   %6d %6d %6d %6d   %6d %6d %6d %6d   %6d %6d %6d %6d   %6d" % (s, r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9], r[10], r[11], r[12]))
