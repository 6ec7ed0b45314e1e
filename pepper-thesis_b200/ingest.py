"""BAM / FASTA ingest (libpv_ingest.so, include/pepper_ingest.h): the reference's ``BAM_handler`` / ``FASTA_handler``
(/root/reference/pepper_variant/modules/cpp/bam_handler.cpp, fasta_handler.cpp, bound in pybind_api.h:223-246) without
htslib, plus ``ingest_regions`` = the per-interval body of ``AlignmentSummarizer.create_summary``
(/root/reference/pepper_variant/modules/python/AlignmentSummarizer.py:180-220) for many intervals at once, producing a
packed :class:`ReadBatch` for the CUDA path instead of ``list[type_read]``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence

import numpy as np

from . import nativebuild
from .read_batch import ARRAY_NAMES, PvReadBatchStruct, ReadBatch, _READ_FIELDS, _REGION_FIELDS
from .summarizer import MAX_READS_IN_REGION, RANDOM_SEED, REGION_SAFE_BASES

_LIB = None

EXPORTS = ["pv_ingest_last_error", "pv_bam_open", "pv_bam_close", "pv_bam_n_targets", "pv_bam_target_name",
           "pv_bam_target_len", "pv_bam_sample_names", "pv_fasta_open", "pv_fasta_close", "pv_fasta_n_seq",
           "pv_fasta_seq_name", "pv_fasta_seq_len", "pv_fasta_fetch", "pv_ingest_regions", "pv_bam_get_reads",
           "pv_ingest_view", "pv_ingest_hp_tags", "pv_ingest_pos_end", "pv_ingest_bam_flags", "pv_ingest_query_names", "pv_ingest_select",
           "pv_ingest_free", "pv_ingest_inflated_bytes", "pv_ingest_fast_blocks", "pv_inflate_raw",
           "pv_bam_plan", "pv_bam_plan_comp_bytes", "pv_bam_plan_tid", "pv_bam_plan_load", "pv_bam_plan_n_blocks",
           "pv_bam_plan_n_segments", "pv_bam_plan_inflated_bytes", "pv_bam_plan_tables", "pv_bam_plan_free"]


class _Opt(C.Structure):
    _fields_ = [("include_supplementary", C.c_int32), ("min_mapq", C.c_int32), ("min_baseq", C.c_int32),
                ("safe_bases", C.c_int32), ("threads", C.c_int32), ("_pad", C.c_int32)]


def load():
    global _LIB
    if _LIB is None:
        path = nativebuild.INGEST
        if not os.path.exists(path):
            raise RuntimeError("libpv_ingest.so is not built: run __graft_entry__.build()")
        lib = C.CDLL(path)
        lib.pv_ingest_last_error.restype = C.c_char_p
        lib.pv_bam_open.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]
        lib.pv_bam_close.argtypes = [C.c_void_p]
        lib.pv_bam_n_targets.argtypes = [C.c_void_p]
        lib.pv_bam_target_name.argtypes = [C.c_void_p, C.c_int32]; lib.pv_bam_target_name.restype = C.c_char_p
        lib.pv_bam_target_len.argtypes = [C.c_void_p, C.c_int32]; lib.pv_bam_target_len.restype = C.c_int64
        lib.pv_bam_sample_names.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]; lib.pv_bam_sample_names.restype = C.c_int64
        lib.pv_fasta_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        lib.pv_fasta_close.argtypes = [C.c_void_p]
        lib.pv_fasta_n_seq.argtypes = [C.c_void_p]
        lib.pv_fasta_seq_name.argtypes = [C.c_void_p, C.c_int32]; lib.pv_fasta_seq_name.restype = C.c_char_p
        lib.pv_fasta_seq_len.argtypes = [C.c_void_p, C.c_char_p]; lib.pv_fasta_seq_len.restype = C.c_int64
        lib.pv_fasta_fetch.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_int64, C.c_void_p, C.POINTER(C.c_int64)]
        lib.pv_ingest_regions.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.POINTER(_Opt), C.POINTER(C.c_void_p)]
        lib.pv_bam_get_reads.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_int64, C.POINTER(_Opt), C.POINTER(C.c_void_p)]
        lib.pv_ingest_view.argtypes = [C.c_void_p, C.POINTER(PvReadBatchStruct)]
        lib.pv_ingest_hp_tags.argtypes = [C.c_void_p]; lib.pv_ingest_hp_tags.restype = C.c_void_p
        lib.pv_ingest_pos_end.argtypes = [C.c_void_p]; lib.pv_ingest_pos_end.restype = C.c_void_p
        lib.pv_ingest_bam_flags.argtypes = [C.c_void_p]; lib.pv_ingest_bam_flags.restype = C.c_void_p
        lib.pv_ingest_query_names.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]; lib.pv_ingest_query_names.restype = C.c_void_p
        lib.pv_ingest_select.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]
        lib.pv_ingest_free.argtypes = [C.c_void_p]
        lib.pv_ingest_inflated_bytes.restype = C.c_uint64
        lib.pv_ingest_fast_blocks.restype = C.c_uint64
        lib.pv_inflate_raw.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        lib.pv_bam_plan.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_int64, C.POINTER(C.c_void_p)]
        lib.pv_bam_plan_comp_bytes.argtypes = [C.c_void_p]; lib.pv_bam_plan_comp_bytes.restype = C.c_int64
        lib.pv_bam_plan_tid.argtypes = [C.c_void_p]
        lib.pv_bam_plan_load.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        lib.pv_bam_plan_n_blocks.argtypes = [C.c_void_p]
        lib.pv_bam_plan_n_segments.argtypes = [C.c_void_p]
        lib.pv_bam_plan_inflated_bytes.argtypes = [C.c_void_p]; lib.pv_bam_plan_inflated_bytes.restype = C.c_int64
        lib.pv_bam_plan_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pv_bam_plan_free.argtypes = [C.c_void_p]
        _LIB = lib
    return _LIB


def _check(rc):
    if rc != 0:
        raise RuntimeError("libpv_ingest: %s (code %d)" % (load().pv_ingest_last_error().decode(), rc))


def _copy(ptr, n, dtype):
    dt = np.dtype(dtype)
    if n == 0 or not ptr:
        return np.zeros(0, dt)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * dt.itemsize,)).view(dt).copy()


class IngestedReads:
    """A ReadBatch plus the per-read fields of ``type_read`` that are not on the hot path."""

    def __init__(self, handle, contig):
        lib = load()
        v = PvReadBatchStruct()
        _check(lib.pv_ingest_view(handle, C.byref(v)))
        n, nr = v.n_reads, v.n_regions
        a = {}
        for name, dt in _READ_FIELDS:
            a[name] = _copy(getattr(v, name), n, dt)
        a["bases"] = _copy(v.bases, v.n_bases, np.uint8)
        a["quals"] = _copy(v.quals, v.n_bases, np.uint8)
        a["cigar"] = _copy(v.cigar, v.n_ops, np.uint32)
        for name, dt in _REGION_FIELDS:
            a[name] = _copy(getattr(v, name), nr, dt)
        a["region_read_begin"] = _copy(v.region_read_begin, nr + 1, np.int64)
        a["ref"] = _copy(v.ref, v.n_ref, np.uint8)
        self.batch = ReadBatch(contigs=[contig] * nr, min_qual=int(v.min_qual), **a)
        self.hp_tag = _copy(lib.pv_ingest_hp_tags(handle), n, np.int32)
        self.pos_end = _copy(lib.pv_ingest_pos_end(handle), n, np.int64)
        self.bam_flag = _copy(lib.pv_ingest_bam_flags(handle), n, np.uint16)
        tot = C.c_int64(0)
        p = lib.pv_ingest_query_names(handle, C.byref(tot))
        raw = _copy(p, tot.value, np.uint8).tobytes()
        self.query_names = [s.decode() for s in raw.split(b"\0")[:-1]] if tot.value else []


class BAMHandler:
    """``PEPPER_VARIANT.BAM_handler`` (pybind_api.h:223-232)."""

    def __init__(self, path, index_path=None):
        self._h = C.c_void_p()
        _check(load().pv_bam_open(os.fsencode(path), os.fsencode(index_path) if index_path else None, C.byref(self._h)))
        self.path = path

    def close(self):
        if self._h:
            load().pv_bam_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_chromosome_sequence_names(self) -> List[str]:
        lib = load()
        return [lib.pv_bam_target_name(self._h, i).decode() for i in range(lib.pv_bam_n_targets(self._h))]

    def get_chromosome_sequence_names_with_length(self):
        lib = load()
        return [(lib.pv_bam_target_name(self._h, i).decode(), int(lib.pv_bam_target_len(self._h, i)))
                for i in range(lib.pv_bam_n_targets(self._h))]

    def get_sample_names(self):
        lib = load()
        n = lib.pv_bam_sample_names(self._h, None, 0)
        buf = C.create_string_buffer(int(n) + 1)
        lib.pv_bam_sample_names(self._h, buf, n + 1)
        return set(s for s in buf.value.decode().split("\n") if s)

    def get_reads_packed(self, chromosome, start, stop, include_supplementary, min_mapq=0, min_baseq=0) -> IngestedReads:
        """``get_reads`` (bam_handler.cpp:115-444) as a packed single-span batch (no region/reference fields)."""
        out = C.c_void_p()
        opt = _Opt(int(bool(include_supplementary)), int(min_mapq), int(min_baseq), 0, 1, 0)
        _check(load().pv_bam_get_reads(self._h, chromosome.encode(), int(start), int(stop), C.byref(opt), C.byref(out)))
        try:
            return IngestedReads(out, chromosome)
        finally:
            load().pv_ingest_free(out)


class FASTAHandler:
    """``PEPPER_VARIANT.FASTA_handler`` (pybind_api.h:240-246)."""

    def __init__(self, path):
        self._h = C.c_void_p()
        _check(load().pv_fasta_open(os.fsencode(path), C.byref(self._h)))

    def close(self):
        if self._h:
            load().pv_fasta_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_chromosome_names(self) -> List[str]:
        lib = load()
        return [lib.pv_fasta_seq_name(self._h, i).decode() for i in range(lib.pv_fasta_n_seq(self._h))]

    def get_chromosome_sequence_length(self, name) -> int:
        return int(load().pv_fasta_seq_len(self._h, name.encode()))

    def get_reference_sequence(self, region, start, stop) -> str:
        n = max(0, int(stop) - int(start))
        buf = C.create_string_buffer(n + 1)
        got = C.c_int64(0)
        _check(load().pv_fasta_fetch(self._h, region.encode(), int(start), int(stop), buf, C.byref(got)))
        return buf.raw[:got.value].decode()


def reservoir_indices(total_reads: int, downsample_rate: float) -> np.ndarray | None:
    """Indices kept by the reservoir sample of AlignmentSummarizer.py:191-208 (same RandomState draws), in sample
    order; None when nothing is dropped."""
    allowed = int(min(MAX_READS_IN_REGION, downsample_rate * total_reads))
    if total_reads <= allowed:
        return None
    random = np.random.RandomState(RANDOM_SEED)
    sample = []
    for i in range(total_reads):
        if len(sample) < allowed:
            sample.append(i)
        else:
            j = random.randint(0, i + 1)
            if j < allowed:
                sample[j] = i
    return np.asarray(sample, np.int64)


def ingest_regions(bam: BAMHandler, fasta: FASTAHandler, contig: str, starts: Sequence[int], ends: Sequence[int],
                   include_supplementary=False, min_mapq=0, min_baseq=0, downsample_rate=1.0, threads=0,
                   safe_bases=REGION_SAFE_BASES) -> IngestedReads:
    """Packed batch for intervals ``[starts[i], ends[i]]`` (inclusive candidate intervals) of one contig."""
    lib = load()
    s = np.ascontiguousarray(starts, np.int64)
    e = np.ascontiguousarray(ends, np.int64)
    assert s.shape == e.shape
    opt = _Opt(int(bool(include_supplementary)), int(min_mapq), int(min_baseq), int(safe_bases), int(threads), 0)
    out = C.c_void_p()
    _check(lib.pv_ingest_regions(bam._h, fasta._h, contig.encode(), int(s.shape[0]), s.ctypes.data, e.ctypes.data,
                                 C.byref(opt), C.byref(out)))
    try:
        # reservoir down-sampling stays host-side index arithmetic
        v = PvReadBatchStruct()
        _check(lib.pv_ingest_view(out, C.byref(v)))
        rb = _copy(v.region_read_begin, v.n_regions + 1, np.int64)
        keep, changed = [], False
        for r in range(v.n_regions):
            n = int(rb[r + 1] - rb[r])
            idx = reservoir_indices(n, downsample_rate)
            if idx is None:
                keep.append(np.arange(rb[r], rb[r + 1], dtype=np.int64))
            else:
                keep.append(idx + rb[r]); changed = True
        if changed:
            k = np.ascontiguousarray(np.concatenate(keep)) if keep else np.zeros(0, np.int64)
            sel = C.c_void_p()
            _check(lib.pv_ingest_select(out, k.ctypes.data, int(k.shape[0]), C.byref(sel)))
            lib.pv_ingest_free(out)
            out = sel
        got = IngestedReads(out, contig)
        # the regions' reference is padded with 'N' past the contig end; the stage-3 filter needs the true length to clip its
        # +-10 bp context there like FASTA_handler does (candidate_filter.filter_flags picks this up by default)
        got.batch.region_contig_len = np.full(got.batch.n_regions, int(fasta.get_chromosome_sequence_length(contig)), np.int64)
        return got
    finally:
        lib.pv_ingest_free(out)
