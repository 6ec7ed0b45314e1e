"""TEST INFRASTRUCTURE ONLY -- Python access to the two CPU oracles of the summary path.

  ref_summary(batch, r, thr)   the UNMODIFIED reference C++ (oracle/_ref/pv_ref_oracle, built from
                               /root/reference/pepper_variant/modules/cpp/region_summary.cpp by oracle/Makefile)
  port_summary(batch, r, thr)  the C restatement (oracle/region_summary_port.c)

Both take a packed ReadBatch + region index and return the same dict of numpy arrays:
  position int64[K], depth int32[K], frequency int32[K], alleles list[bytes], images int32/int16 [K,33,26].
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_port = None
_ref = None


def have_ref():
    d = os.path.join(_HERE, "_ref")
    return os.path.isdir(d) and any(f.startswith("pv_ref_oracle") and f.endswith(".so") for f in os.listdir(d))


def ref_module():
    global _ref
    if _ref is None:
        sys.path.insert(0, os.path.join(_HERE, "_ref"))
        import pv_ref_oracle
        _ref = pv_ref_oracle
    return _ref


def port_lib():
    global _port
    if _port is None:
        path = os.path.join(_HERE, "libpv_oracle_port.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle port not built: run `make -C oracle port`")
        _port = C.CDLL(path)
        _port.pv_port_summary_region.restype = C.c_int64
        _port.pv_port_summary_region.argtypes = [C.c_void_p] * 10 + [C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                                 C.c_int64, C.c_int64, C.c_void_p, C.c_int32, C.c_int64, C.c_int64,
                                                 C.c_int64] + [C.c_void_p] * 8
    return _port


def ref_build_reads(batch, r):
    m = ref_module()
    lo, hi = int(batch.region_read_begin[r]), int(batch.region_read_begin[r + 1])
    return m.build_reads(batch.read_pos, batch.read_base_off, batch.read_len, batch.read_cigar_off, batch.read_n_ops,
                         batch.read_flags, batch.read_mapq, batch.bases, batch.quals, batch.cigar, lo, hi)


def ref_run(batch, r, thr, readset, window=32, features=26):
    m = ref_module()
    fo, fl = int(batch.region_ref_off[r]), int(batch.region_ref_len[r])
    contig = batch.contigs[r] if batch.contigs else "c"
    return m.run_region(readset, contig, int(batch.region_ref_start[r]), int(batch.region_ref_end[r]),
                        batch.ref[fo:fo + fl].tobytes(), thr.as_list9(), bool(thr.skip_indels),
                        int(batch.region_cand_start[r]), int(batch.region_cand_end[r]), window, features)


def ref_summary(batch, r, thr):
    return ref_run(batch, r, thr, ref_build_reads(batch, r))


def port_summary(batch, r, thr, capacity=None, want_dense=False):
    lib = port_lib()
    L = int(batch.region_ref_end[r] - batch.region_ref_start[r] + 1)
    cap = capacity if capacity is not None else max(1024, L // 8)
    while True:
        win = np.zeros((cap, 33, 26), np.int16)
        pos = np.zeros(cap, np.int64); dep = np.zeros(cap, np.int32); frq = np.zeros(cap, np.int32)
        al = np.zeros((cap, 64), np.uint8); aln = np.zeros(cap, np.uint8)
        dense = np.zeros((L, 26), np.int32) if want_dense else None
        counts = np.zeros((L, 4), np.int32) if want_dense else None
        t = np.asarray(thr.as_list9(), np.float64)
        fo = int(batch.region_ref_off[r])
        K = lib.pv_port_summary_region(
            batch.read_pos.ctypes.data, batch.read_base_off.ctypes.data, batch.read_len.ctypes.data,
            batch.read_cigar_off.ctypes.data, batch.read_n_ops.ctypes.data, batch.read_flags.ctypes.data,
            batch.read_mapq.ctypes.data, batch.bases.ctypes.data, batch.quals.ctypes.data, batch.cigar.ctypes.data,
            int(batch.region_read_begin[r]), int(batch.region_read_begin[r + 1]),
            batch.ref.ctypes.data + fo, int(batch.region_ref_len[r]), int(batch.region_ref_start[r]),
            int(batch.region_ref_end[r]), t.ctypes.data, int(bool(thr.skip_indels)),
            int(batch.region_cand_start[r]), int(batch.region_cand_end[r]), cap,
            win.ctypes.data, pos.ctypes.data, dep.ctypes.data, frq.ctypes.data, al.ctypes.data, aln.ctypes.data,
            dense.ctypes.data if want_dense else None, counts.ctypes.data if want_dense else None)
        if K < 0:
            raise ValueError("port: bad arguments")
        if K <= cap or capacity is not None:
            break
        cap = int(K)
    K = min(int(K), cap)
    out = dict(position=pos[:K], depth=dep[:K], frequency=frq[:K], images=win[:K],
               alleles=[bytes(al[i, :aln[i]]) for i in range(K)])
    if want_dense:
        out["dense"] = dense
        out["counts"] = counts
    return out
