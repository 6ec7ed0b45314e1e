"""Seeded synthetic pileups in the packed layout, via csrc/synth_reads.c.

Presets carry (a) the error/length/quality profile of the sequencing technology and (b) the matching
``generate_summary`` thresholds of the reference's platform presets
(/root/reference/pepper_variant/modules/argparse/SetParameters.py:12-65 R9 Guppy5 SUP, :122-175 R10 Q20,
:176-229 HiFi).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from .read_batch import ReadBatch

_LIB = None


class _Cfg(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("contig_len", C.c_int64), ("region_size", C.c_int64), ("margin", C.c_int64),
                ("coverage", C.c_double), ("len_median", C.c_double), ("len_sigma", C.c_double), ("len_sd", C.c_double),
                ("len_min", C.c_int64), ("len_max", C.c_int64),
                ("sub_rate", C.c_double), ("ins_rate", C.c_double), ("del_rate", C.c_double), ("indel_geom_p", C.c_double),
                ("qual_lo", C.c_int32), ("qual_hi", C.c_int32), ("snp_every", C.c_int32), ("indel_every", C.c_int32)]


class _Sizes(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("n_bases", C.c_int64), ("n_ops", C.c_int64), ("n_ref", C.c_int64)]


def _lib():
    global _LIB
    if _LIB is None:
        from . import nativebuild as build
        path = build.build_synth()
        _LIB = C.CDLL(path)
        _LIB.pv_synth_count.argtypes = [C.POINTER(_Cfg), C.c_int64, C.c_int64, C.c_int, C.POINTER(_Sizes)]
        _LIB.pv_synth_fill.argtypes = [C.POINTER(_Cfg), C.c_int64, C.c_int64, C.c_int] + [C.c_void_p] * 15
        _LIB.pv_synth_region_bounds.argtypes = [C.POINTER(_Cfg), C.c_int64, C.c_void_p]
    return _LIB


@dataclass(frozen=True)
class Thresholds:
    """The ten scalars of RegionalSummaryGenerator::generate_summary (region_summary.h:191-201)."""
    min_snp_baseq: float
    min_indel_baseq: float
    snp_freq: float
    insert_freq: float
    delete_freq: float
    min_coverage: float
    snp_candidate_freq: float
    indel_candidate_freq: float
    candidate_support: float
    skip_indels: bool = False

    def as_list9(self):
        return [float(self.min_snp_baseq), float(self.min_indel_baseq), float(self.snp_freq), float(self.insert_freq),
                float(self.delete_freq), float(self.min_coverage), float(self.snp_candidate_freq),
                float(self.indel_candidate_freq), float(self.candidate_support)]


# SetParameters.py:12-37 / :122-147 / :176-201 (image generation block of each preset)
THRESHOLDS = {
    "ont_r9_guppy5_sup": Thresholds(1, 1, 0.10, 0.15, 0.15, 3, 0.10, 0.10, 2, False),
    "ont_r10_q20": Thresholds(1, 1, 0.10, 0.10, 0.10, 3, 0.10, 0.10, 2, False),
    "hifi": Thresholds(10, 10, 0.10, 0.12, 0.10, 2, 0.10, 0.10, 2, False),
}


@dataclass(frozen=True)
class Profile:
    name: str
    preset: str
    len_median: float
    len_sigma: float      # log-normal sigma; 0 -> normal(len_median, len_sd)
    len_sd: float
    len_min: int
    len_max: int
    sub: float
    ins: float
    dele: float
    qual_lo: int
    qual_hi: int

    @property
    def thresholds(self) -> Thresholds:
        return THRESHOLDS[self.preset]


PROFILES = {
    "ont_r9": Profile("ont_r9", "ont_r9_guppy5_sup", 12000, 0.6, 0, 1000, 100000, 0.020, 0.015, 0.015, 5, 29),
    "ont_r10": Profile("ont_r10", "ont_r10_q20", 12000, 0.6, 0, 1000, 100000, 0.005, 0.0025, 0.0025, 10, 39),
    "hifi": Profile("hifi", "hifi", 15000, 0.0, 2000, 1000, 30000, 0.0005, 0.001, 0.001, 20, 60),
}


def n_regions(contig_len: int, region_size: int = 100000) -> int:
    return (contig_len + region_size - 1) // region_size


def generate(profile: str | Profile, contig_len: int, coverage: float, seed: int = 1, first_region: int = 0,
             num_regions: int | None = None, region_size: int = 100000, margin: int = 100, threads: int | None = None,
             snp_every: int = 1000, indel_every: int = 8000, contig: str = "chrS", pinned: bool = False) -> ReadBatch:
    """Generate regions [first_region, first_region + num_regions) of the synthetic contig as one ReadBatch.

    Region r is the reference's interval r (ImageGenerationUI.py:292-316) with the +-100 bp margin
    (AlignmentSummarizer.py:181-182). ``pinned`` allocates the big arrays in page-locked memory (torch) so
    they can be copied to the GPU asynchronously.
    """
    p = PROFILES[profile] if isinstance(profile, str) else profile
    lib = _lib()
    total = n_regions(contig_len, region_size)
    if num_regions is None:
        num_regions = total - first_region
    if first_region < 0 or first_region + num_regions > total:
        raise ValueError("region range outside the contig")
    threads = threads or min(64, os.cpu_count() or 1)
    cfg = _Cfg(seed, contig_len, region_size, margin, coverage, p.len_median, p.len_sigma, p.len_sd, p.len_min,
               p.len_max, p.sub, p.ins, p.dele, 0.6, p.qual_lo, p.qual_hi, snp_every, indel_every)
    sizes = (_Sizes * num_regions)()
    lib.pv_synth_count(C.byref(cfg), first_region, num_regions, threads, sizes)
    nr = np.array([s.n_reads for s in sizes], np.int64)
    nb = np.array([s.n_bases for s in sizes], np.int64)
    no = np.array([s.n_ops for s in sizes], np.int64)
    nf = np.array([s.n_ref for s in sizes], np.int64)

    def excl(a):
        out = np.zeros(a.shape[0] + 1, np.int64)
        np.cumsum(a, out=out[1:])
        return out

    rb, bb, ob, fb = excl(nr), excl(nb), excl(no), excl(nf)
    R, B, O, Fn = int(rb[-1]), int(bb[-1]), int(ob[-1]), int(fb[-1])

    def alloc(n, dt):
        if pinned:
            import torch
            tdt = {np.uint8: torch.uint8, np.int32: torch.int32, np.int64: torch.int64, np.uint32: torch.int32}[dt]
            t = torch.empty(max(n, 1), dtype=tdt, pin_memory=True)
            a = t.numpy()[:n]
            return a.view(dt) if dt is np.uint32 else a
        return np.empty(n, dt)

    arr = dict(read_pos=alloc(R, np.int64), read_base_off=alloc(R, np.int64), read_len=alloc(R, np.int32),
               read_cigar_off=alloc(R, np.int64), read_n_ops=alloc(R, np.int32), read_flags=alloc(R, np.uint8),
               read_mapq=alloc(R, np.uint8), bases=alloc(B, np.uint8), quals=alloc(B, np.uint8),
               cigar=alloc(O, np.uint32), ref=alloc(Fn, np.uint8))
    order = ["read_pos", "read_base_off", "read_len", "read_cigar_off", "read_n_ops", "read_flags", "read_mapq",
             "bases", "quals", "cigar", "ref"]
    lib.pv_synth_fill(C.byref(cfg), first_region, num_regions, threads,
                      rb.ctypes.data, bb.ctypes.data, ob.ctypes.data, fb.ctypes.data,
                      *[arr[k].ctypes.data for k in order])
    bounds = np.zeros((num_regions, 4), np.int64)
    for i in range(num_regions):
        lib.pv_synth_region_bounds(C.byref(cfg), first_region + i, bounds[i].ctypes.data)
    return ReadBatch(region_ref_start=np.ascontiguousarray(bounds[:, 2]), region_ref_end=np.ascontiguousarray(bounds[:, 3]),
                     region_cand_start=np.ascontiguousarray(bounds[:, 0]), region_cand_end=np.ascontiguousarray(bounds[:, 1]),
                     region_ref_off=np.ascontiguousarray(fb[:-1]), region_ref_len=nf.copy(),
                     region_read_begin=rb, contigs=[contig] * num_regions, **arr)
