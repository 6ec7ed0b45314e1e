"""ctypes binding of the C-ABI library ``libpepper_b200.so`` (include/pepper_b200.h).

The library has no CPU fallback: every compute call raises ``PvError`` when no sm_100 device is present or the
shared object is missing (``load()`` raises instead of substituting anything).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .read_batch import PvReadBatchStruct, ReadBatch

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PV_LIB_PATH") or os.path.join(PKG, "libpepper_b200.so")   # PV_LIB_PATH: build-variant experiments

PV_WINDOW, PV_FEATURES, PV_ALLELE_BYTES = 33, 26, 64
PV_EOVERFLOW = -4

# every symbol include/pepper_b200.h declares (checked by tests/test_abi.py)
EXPORTS = ["pv_version", "pv_last_error", "pv_device_count", "pv_profile_enable", "pv_profile_collect",
           "pv_profile_reset", "pv_launch_count", "pv_summary_status_offset", "pv_unpack_bases4", "pv_pack_bases4", "pv_unpack_bases2", "pv_pack_bases2", "pv_pack_group", "pv_pack_bases_ref", "pv_unpack_bases_ref", "pv_pack_cigar8", "pv_unpack_cigar8", "pv_pack_quals_pred", "pv_unpack_quals_pred", "pv_min_qual", "pv_unpack_quals", "pv_qual_bits", "pv_pack_quals", "pv_unpack_cigar16", "pv_pack_cigar16", "pv_batch_validate", "pv_synth_device_count", "pv_synth_device_fill", "pv_summary_workspace_bytes",
           "pv_summary_regions", "pv_summary_regions_host", "pv_lstm_create", "pv_lstm_destroy",
           "pv_lstm_workspace_bytes", "pv_lstm_infer", "pv_lstm_infer_host", "pv_gru_create", "pv_gru_destroy",
           "pv_gru_workspace_bytes", "pv_gru_forward", "pv_gru_predict_chunks", "pv_candidate_filter",
           "pv_candidate_filter_host", "pv_polish_workspace_bytes", "pv_polish_count", "pv_polish_emit", "pv_polish_chunks",
           "pv_bam_inflate_blocks", "pv_bam_index_records", "pv_bam_clip_count", "pv_bam_clip_workspace_bytes", "pv_bam_clip_layout",
           "pv_bam_clip_write", "pv_bam_gather_names", "pv_bam_gather_reference"]


class PvError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pepper_b200 error %d: %s" % (code, msg))
        self.code = code


class PvThresholdsStruct(C.Structure):
    _fields_ = [("min_snp_baseq", C.c_double), ("min_indel_baseq", C.c_double), ("snp_freq", C.c_double),
                ("insert_freq", C.c_double), ("delete_freq", C.c_double), ("min_coverage", C.c_double),
                ("snp_candidate_freq", C.c_double), ("indel_candidate_freq", C.c_double),
                ("candidate_support", C.c_double), ("skip_indels", C.c_int32), ("_pad", C.c_int32)]


class PvFilterOptionsStruct(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("snp_p_value", "snp_p_value_in_lc", "insert_p_value", "insert_p_value_in_lc",
                                          "delete_p_value", "delete_p_value_in_lc", "report_snp_above_freq",
                                          "report_indel_above_freq")]


class PvCandidatesStruct(C.Structure):
    _fields_ = [("capacity", C.c_int64), ("windows", C.c_void_p), ("position", C.c_void_p), ("region", C.c_void_p),
                ("depth", C.c_void_p), ("frequency", C.c_void_p), ("allele", C.c_void_p), ("allele_len", C.c_void_p)]


class PvLstmWeightsStruct(C.Structure):
    _fields_ = [("enc_w_ih", C.c_void_p * 2), ("enc_w_hh", C.c_void_p * 2), ("enc_b_ih", C.c_void_p * 2),
                ("enc_b_hh", C.c_void_p * 2), ("dec_w_ih", C.c_void_p * 2), ("dec_w_hh", C.c_void_p * 2),
                ("dec_b_ih", C.c_void_p * 2), ("dec_b_hh", C.c_void_p * 2), ("lin_w", C.c_void_p * 5),
                ("lin_b", C.c_void_p * 5), ("out_w", C.c_void_p), ("out_b", C.c_void_p)]


class PvGruWeightsStruct(C.Structure):
    _fields_ = [("enc_w_ih", C.c_void_p * 2), ("enc_w_hh", C.c_void_p * 2), ("enc_b_ih", C.c_void_p * 2),
                ("enc_b_hh", C.c_void_p * 2), ("dec_w_ih", C.c_void_p * 2), ("dec_w_hh", C.c_void_p * 2),
                ("dec_b_ih", C.c_void_p * 2), ("dec_b_hh", C.c_void_p * 2), ("dense_w", C.c_void_p),
                ("dense_b", C.c_void_p)]


_lib = None


def load() -> C.CDLL:
    """Load libpepper_b200.so (built in-tree by ``build.build_lib``). Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PvError(-2, "libpepper_b200.so is not built (run `python -m pepper_thesis_b200.nativebuild`); "
                              "there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        lib.pv_version.restype = C.c_char_p
        lib.pv_last_error.restype = C.c_char_p
        lib.pv_device_count.restype = C.c_int
        lib.pv_batch_validate.argtypes = [C.POINTER(PvReadBatchStruct)]
        lib.pv_profile_enable.argtypes = [C.c_int]
        lib.pv_profile_enable.restype = None
        lib.pv_profile_collect.argtypes = [C.c_void_p, C.c_void_p]
        lib.pv_profile_reset.restype = None
        lib.pv_launch_count.restype = C.c_int64
        lib.pv_unpack_bases4.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        lib.pv_pack_bases4.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32]
        lib.pv_unpack_bases2.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        lib.pv_pack_bases2.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int32]
        lib.pv_pack_group.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_void_p, C.c_int64,
                                      C.c_void_p, C.POINTER(C.c_int32), C.c_int32]
        lib.pv_pack_bases_ref.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32]
        lib.pv_unpack_bases_ref.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pv_pack_quals_pred.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32]
        lib.pv_unpack_quals_pred.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pv_min_qual.argtypes = [C.c_void_p, C.c_int32]
        lib.pv_min_qual.restype = C.c_int32
        lib.pv_pack_cigar8.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32]
        lib.pv_unpack_cigar8.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pv_unpack_quals.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
        lib.pv_qual_bits.argtypes = [C.c_void_p, C.c_int64, C.c_int32]
        lib.pv_qual_bits.restype = C.c_int32
        lib.pv_pack_quals.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32]
        lib.pv_unpack_cigar16.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        lib.pv_pack_cigar16.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32]
        lib.pv_summary_workspace_bytes.restype = C.c_int64
        lib.pv_summary_workspace_bytes.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int64]
        lib.pv_summary_regions.argtypes = [C.POINTER(PvReadBatchStruct), C.c_void_p, C.c_int64,
                                           C.POINTER(PvThresholdsStruct), C.c_int32, C.c_int32,
                                           C.POINTER(PvCandidatesStruct), C.c_void_p, C.c_void_p, C.c_int64,
                                           C.c_void_p, C.c_void_p]
        lib.pv_summary_regions_host.argtypes = [C.POINTER(PvReadBatchStruct), C.POINTER(PvThresholdsStruct), C.c_int32,
                                                C.c_int32, C.POINTER(PvCandidatesStruct), C.POINTER(C.c_int64),
                                                C.c_void_p]
        if hasattr(lib, "pv_lstm_create"):
            lib.pv_lstm_create.argtypes = [C.POINTER(PvLstmWeightsStruct), C.POINTER(C.c_void_p)]
            lib.pv_lstm_destroy.argtypes = [C.c_void_p]
            lib.pv_lstm_destroy.restype = None
            lib.pv_lstm_workspace_bytes.restype = C.c_int64
            lib.pv_lstm_workspace_bytes.argtypes = [C.c_int64]
            lib.pv_lstm_infer.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int64, C.c_void_p]
            lib.pv_lstm_infer_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
        if hasattr(lib, "pv_gru_create"):
            lib.pv_gru_create.argtypes = [C.POINTER(PvGruWeightsStruct), C.POINTER(C.c_void_p)]
            lib.pv_gru_destroy.argtypes = [C.c_void_p]
            lib.pv_gru_destroy.restype = None
            lib.pv_gru_workspace_bytes.restype = C.c_int64
            lib.pv_gru_workspace_bytes.argtypes = [C.c_int64, C.c_int32]
            lib.pv_gru_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_int64, C.c_void_p]
            lib.pv_gru_predict_chunks.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        if hasattr(lib, "pv_polish_count"):
            lib.pv_polish_workspace_bytes.restype = C.c_int64
            lib.pv_polish_workspace_bytes.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.c_int64]
            lib.pv_polish_count.argtypes = [C.POINTER(PvReadBatchStruct), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                            C.POINTER(C.c_int64), C.c_void_p, C.c_void_p]
            lib.pv_polish_emit.argtypes = [C.POINTER(PvReadBatchStruct), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
            lib.pv_polish_chunks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                             C.c_void_p, C.c_void_p]
        if hasattr(lib, "pv_candidate_filter"):
            lib.pv_candidate_filter.argtypes = [C.c_int64] + [C.c_void_p] * 12 + [C.POINTER(PvFilterOptionsStruct), C.c_void_p, C.c_void_p]
            lib.pv_candidate_filter_host.argtypes = [C.c_int64] + [C.c_void_p] * 7 + [C.c_int32] + [C.c_void_p] * 5 + \
                [C.c_int64, C.POINTER(PvFilterOptionsStruct), C.c_void_p]
        P, I64, I32 = C.c_void_p, C.c_int64, C.c_int32
        lib.pv_bam_inflate_blocks.argtypes = [P, I64, P, I32, P, I64, I32, P, P]
        lib.pv_bam_index_records.argtypes = [P, I64, P, P, I32, P, P, I64, P, P]
        lib.pv_bam_clip_count.argtypes = [P, I64, P, I64, I32, P, P, I32, I32, I32, P, P, P]
        lib.pv_bam_clip_workspace_bytes.argtypes = [I64]; lib.pv_bam_clip_workspace_bytes.restype = I64
        lib.pv_bam_clip_layout.argtypes = [P, I64, P, I64, I32, P, P, I32, I32, I32, P, I64, P, I64, P, P, P, P, P, P, P]
        lib.pv_bam_clip_write.argtypes = [P, I64, P, I64, P, P, P, P] + [P] * 13 + [P, P, P]
        lib.pv_bam_gather_names.argtypes = [P, P, P, P, I64, P, P]
        lib.pv_bam_gather_reference.argtypes = [P, I64, I64, P, P, P, I32, I64, P, P]
        _lib = lib
    return _lib


FAMILIES = ["sum_cigar_prefix", "sum_pileup_tile", "sum_site_alleles", "sum_key_sort", "sum_emit_windows",
            "lstm_input_prep", "lstm_encoder_steps", "lstm_decoder_steps", "lstm_mlp_head", "gru_steps", "gru_misc",
            "candidate_filter", "polish_summary", "gru_gx_gemm", "gru_head", "bam_decode"]


def profile_collect():
    """{family: (total ms, launches)} since the last profile_reset (waits for the recorded events)."""
    ms = (C.c_double * len(FAMILIES))()
    ln = (C.c_int64 * len(FAMILIES))()
    load().pv_profile_collect(ms, ln)
    return {f: (float(ms[i]), int(ln[i])) for i, f in enumerate(FAMILIES)}


def check(rc: int):
    if rc != 0:
        raise PvError(rc, load().pv_last_error().decode("utf-8", "replace"))


def thresholds_struct(thr) -> PvThresholdsStruct:
    return PvThresholdsStruct(float(thr.min_snp_baseq), float(thr.min_indel_baseq), float(thr.snp_freq),
                              float(thr.insert_freq), float(thr.delete_freq), float(thr.min_coverage),
                              float(thr.snp_candidate_freq), float(thr.indel_candidate_freq),
                              float(thr.candidate_support), int(bool(thr.skip_indels)), 0)


class Candidates:
    """Host-side candidate arrays (the SoA form of ``list[CandidateImageSummary]``)."""

    def __init__(self, capacity: int):
        self.capacity = int(capacity)
        self.windows = np.zeros((capacity, PV_WINDOW, PV_FEATURES), np.int16)
        self.position = np.zeros(capacity, np.int64)
        self.region = np.zeros(capacity, np.int32)
        self.depth = np.zeros(capacity, np.int32)
        self.frequency = np.zeros(capacity, np.int32)
        self.allele = np.zeros((capacity, PV_ALLELE_BYTES), np.uint8)
        self.allele_len = np.zeros(capacity, np.uint8)
        self.count = 0

    def as_struct(self) -> PvCandidatesStruct:
        return PvCandidatesStruct(self.capacity, self.windows.ctypes.data, self.position.ctypes.data,
                                  self.region.ctypes.data, self.depth.ctypes.data, self.frequency.ctypes.data,
                                  self.allele.ctypes.data, self.allele_len.ctypes.data)

    def alleles(self):
        return [bytes(self.allele[i, :self.allele_len[i]]) for i in range(self.count)]

    def trimmed(self) -> dict:
        k = self.count
        return dict(position=self.position[:k], region=self.region[:k], depth=self.depth[:k],
                    frequency=self.frequency[:k], images=self.windows[:k], alleles=self.alleles())


def summary_regions_host(batch: ReadBatch, thr, capacity: int | None = None, want_dense: bool = False,
                         window: int = 32, features: int = 26):
    """``pv_summary_regions_host``: host batch in, host candidates out. Grows ``capacity`` on overflow."""
    lib = load()
    cap = int(capacity) if capacity else max(4096, batch.total_positions // 16)
    s = batch.as_struct()
    t = thresholds_struct(thr)
    dense = np.zeros((batch.total_positions, PV_FEATURES), np.int16) if want_dense else None
    while True:
        out = Candidates(cap)
        n = C.c_int64(0)
        rc = lib.pv_summary_regions_host(C.byref(s), C.byref(t), window, features, C.byref(out.as_struct()),
                                         C.byref(n), dense.ctypes.data if want_dense else None)
        if rc == PV_EOVERFLOW and capacity is None:
            cap = max(int(n.value), cap * 4)
            continue
        check(rc)
        out.count = int(n.value)
        return (out, dense) if want_dense else out
