"""Device time of the wire-form expansion kernels (160 regions of the bench workload), per form."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pepper_thesis_b200 import synth, capi, device as dev
b = synth.generate("ont_r9", 16000000, 50.0, seed=1)
b.pack_wire(pinned=True, bases_ref=True)
print("bases %d: patches %.4f B/base, quals %.4f B/base, cigar %.4f B/base" % (
    b.n_bases, (b.bases_patch.nbytes + b.read_patch_off.nbytes) / b.n_bases, b.quals_packed.nbytes / b.n_bases, b.cigar16.nbytes / b.n_bases))
db = dev.DeviceBatch(b, defer_unpack=True); torch.cuda.synchronize()
lib = capi.load()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def timed(name, fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): fn()
    e1.record(); torch.cuda.synchronize()
    print("%-12s %.3f ms" % (name, e0.elapsed_time(e1) / 3))
timed("cigar16", lambda: capi.check(lib.pv_unpack_cigar16(C.c_void_p(db.packed_c.data_ptr()), db.t["cigar"].numel(), C.c_void_p(db.t["cigar"].data_ptr()), st)))
timed("bases_ref", lambda: capi.check(lib.pv_unpack_bases_ref(C.byref(db.struct), C.c_void_p(db.t["read_patch_off"].data_ptr()),
                                                              C.c_void_p(db.patches.data_ptr()), C.c_void_p(db.t["bases"].data_ptr()), st)))
timed("quals", lambda: capi.check(lib.pv_unpack_quals(C.c_void_p(db.packed_q.data_ptr()), db.t["quals"].numel(), int(b.qual_bits), C.c_void_p(db.t["quals"].data_ptr()), st)))
