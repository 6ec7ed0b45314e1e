"""Synthetic regions generated ON THE DEVICE (csrc/synth_device.cu): the device-resident twin of ``synth.generate``.

A group of regions goes straight into a device ``PvReadBatch`` -- bit-identical to what ``synth.generate`` would build
for the same (profile, seed, regions) -- so whole-genome-scale workloads are streamed group by group without ever
existing on the host (BASELINE.json configs[3]). Inputs for bench.py / tests only.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import capi, synth
from .read_batch import PvReadBatchStruct


class _Shape:
    """What HotPath / SummaryWorkspace ask a DeviceBatch's ``host`` for."""

    def __init__(self, n_reads, n_ops, n_regions, min_qual):
        self.n_reads, self.n_ops, self.n_regions, self.min_qual = n_reads, n_ops, n_regions, min_qual


class GeneratedBatch:
    """Duck type of ``device.DeviceBatch`` for batches that were never on the host."""

    def __init__(self):
        self.t = {}
        self.qpatches = None
        self.quals_skipped = False

    def ensure_quals(self):
        raise RuntimeError("a generated batch without qualities cannot re-run with them (generate with quals=True)")

    def record_stream(self, stream):
        for t in self.t.values():
            t.record_stream(stream)


def _cfg(profile, contig_len, coverage, seed, region_size=100000, margin=100, snp_every=1000, indel_every=8000):
    p = synth.PROFILES[profile] if isinstance(profile, str) else profile
    return synth._Cfg(seed, contig_len, region_size, margin, coverage, p.len_median, p.len_sigma, p.len_sd, p.len_min,
                      p.len_max, p.sub, p.ins, p.dele, 0.6, p.qual_lo, p.qual_hi, snp_every, indel_every), p


def generate(profile, contig_len: int, coverage: float, seed: int = 1, first_region: int = 0, num_regions: int = 1,
             device="cuda", quals: bool = False, threads: int | None = None) -> GeneratedBatch:
    """Regions [first_region, first_region + num_regions) of the synthetic contig as a device-resident batch.
    ``quals=False``: the quality array is not generated; the batch carries the profile's smallest quality as its
    ``min_qual`` promise (PvReadBatch.quals == NULL), which the summary accepts when it clears both thresholds."""
    lib = capi.load()
    slib = synth._lib()
    cfg, prof = _cfg(profile, contig_len, coverage, seed)
    dev = torch.device(device)
    threads = threads or min(64, os.cpu_count() or 1)
    if not hasattr(slib, "_pv_dev_ready"):
        slib.pv_synth_region_reads.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
        slib.pv_synth_read_lengths.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
        lib.pv_synth_device_count.argtypes = [C.c_void_p, C.c_int64, C.c_int32] + [C.c_void_p] * 2 + [C.c_int64] + [C.c_void_p] * 3
        lib.pv_synth_device_fill.argtypes = [C.c_void_p, C.c_int64, C.c_int32] + [C.c_void_p] * 2 + [C.c_int64] + [C.c_void_p] * 13
        slib._pv_dev_ready = True
    # host: reads per region, read lengths (the only libm-dependent part), region bounds
    counts = np.zeros(num_regions, np.int64)
    slib.pv_synth_region_reads(C.byref(cfg), first_region, num_regions, counts.ctypes.data)
    read_begin = np.zeros(num_regions + 1, np.int64)
    np.cumsum(counts, out=read_begin[1:])
    n_reads = int(read_begin[-1])
    lens = np.zeros(max(n_reads, 1), np.int32)
    slib.pv_synth_read_lengths(C.byref(cfg), first_region, num_regions, threads, read_begin.ctypes.data, lens.ctypes.data)
    bounds = np.zeros((num_regions, 4), np.int64)
    for i in range(num_regions):
        slib.pv_synth_region_bounds(C.byref(cfg), first_region + i, bounds[i].ctypes.data)
    ref_len = bounds[:, 3] - bounds[:, 2] + 1
    ref_off = np.zeros(num_regions + 1, np.int64)
    np.cumsum(ref_len, out=ref_off[1:])
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    g = GeneratedBatch()
    g.device = dev
    rb_d = torch.from_numpy(read_begin).to(dev)
    len_d = torch.from_numpy(lens).to(dev)
    nb_d = torch.zeros(max(n_reads, 1), dtype=torch.int32, device=dev)
    no_d = torch.zeros(max(n_reads, 1), dtype=torch.int32, device=dev)
    capi.check(lib.pv_synth_device_count(C.byref(cfg), first_region, num_regions, rb_d.data_ptr(), len_d.data_ptr(), n_reads,
                                         nb_d.data_ptr(), no_d.data_ptr(), stream))
    pad = (nb_d[:n_reads].to(torch.int64) + 15) & ~15
    base_off = torch.cumsum(pad, 0) - pad
    op64 = no_d[:n_reads].to(torch.int64)
    op_off = torch.cumsum(op64, 0) - op64
    n_bases = int(pad.sum().item()) if n_reads else 0
    n_ops = int(op64.sum().item()) if n_reads else 0
    t = g.t
    t["read_pos"] = torch.empty(max(n_reads, 1), dtype=torch.int64, device=dev)
    t["read_base_off"] = base_off.contiguous() if n_reads else torch.zeros(1, dtype=torch.int64, device=dev)
    t["read_len"] = torch.empty(max(n_reads, 1), dtype=torch.int32, device=dev)
    t["read_cigar_off"] = op_off.contiguous() if n_reads else torch.zeros(1, dtype=torch.int64, device=dev)
    t["read_n_ops"] = torch.empty(max(n_reads, 1), dtype=torch.int32, device=dev)
    t["read_flags"] = torch.empty(max(n_reads, 1), dtype=torch.uint8, device=dev)
    t["read_mapq"] = torch.empty(max(n_reads, 1), dtype=torch.uint8, device=dev)
    t["bases"] = torch.empty(max(n_bases, 16), dtype=torch.uint8, device=dev)
    if quals:
        t["quals"] = torch.empty(max(n_bases, 16), dtype=torch.uint8, device=dev)
    t["cigar"] = torch.empty(max(n_ops, 1), dtype=torch.int32, device=dev)
    t["ref"] = torch.empty(int(ref_off[-1]), dtype=torch.uint8, device=dev)
    ref_off_d = torch.from_numpy(ref_off[:-1].copy()).to(dev)
    capi.check(lib.pv_synth_device_fill(
        C.byref(cfg), first_region, num_regions, rb_d.data_ptr(), len_d.data_ptr(), n_reads, t["read_base_off"].data_ptr(),
        t["read_cigar_off"].data_ptr(), t["read_pos"].data_ptr(), t["read_len"].data_ptr(), t["read_n_ops"].data_ptr(),
        t["read_flags"].data_ptr(), t["read_mapq"].data_ptr(), t["bases"].data_ptr(),
        t["quals"].data_ptr() if quals else None, t["cigar"].data_ptr(), ref_off_d.data_ptr(), t["ref"].data_ptr(), stream))
    region = {"region_ref_start": bounds[:, 2], "region_ref_end": bounds[:, 3], "region_cand_start": bounds[:, 0],
              "region_cand_end": bounds[:, 1], "region_ref_off": ref_off[:-1], "region_ref_len": ref_len}
    for k, v in region.items():
        t[k] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.int64)).to(dev)
    t["region_read_begin"] = rb_d
    s = PvReadBatchStruct()
    s.n_reads, s.n_bases, s.n_ops, s.n_ref, s.n_regions = n_reads, n_bases, n_ops, int(ref_off[-1]), num_regions
    for name, tensor in t.items():
        setattr(s, name, tensor.data_ptr())
    if not quals:
        s.quals = None
    s.min_qual = int(prof.qual_lo)
    g.struct = s
    g.host = _Shape(n_reads, n_ops, num_regions, int(prof.qual_lo))
    g.region_len = np.ascontiguousarray(ref_len, dtype=np.int64)
    g.total_positions = int(ref_len.sum())
    g.candidate_bp = int((bounds[:, 1] - bounds[:, 0] + 1).sum())
    g.n_bases = n_bases
    g.read_bases = int(nb_d[:n_reads].to(torch.int64).sum().item()) if n_reads else 0
    g._keep = (len_d, nb_d, no_d, ref_off_d)
    return g
