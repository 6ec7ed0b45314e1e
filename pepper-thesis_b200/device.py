"""Device-resident hand-off: torch tensors own the HBM, the C-ABI gets raw pointers and the current CUDA stream.

PyTorch is plumbing only (allocation, streams, H2D/D2H); every kernel is in libpepper_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import capi
from .read_batch import ARRAY_NAMES, PvReadBatchStruct as PvReadBatchStructT, ReadBatch

_TORCH_DT = {np.dtype(np.int64): torch.int64, np.dtype(np.int32): torch.int32, np.dtype(np.uint8): torch.uint8,
             np.dtype(np.uint32): torch.int32}


def _to_torch(a: np.ndarray) -> torch.Tensor:
    if a.dtype == np.uint32:
        a = a.view(np.int32)
    return torch.from_numpy(a)


class DeviceBatch:
    """A ReadBatch whose arrays live in HBM (``PvReadBatch`` with device pointers)."""

    def __init__(self, host: ReadBatch, device: torch.device | str = "cuda", non_blocking: bool = True,
                 defer_unpack: bool = False, skip_quals: bool = False):
        """``skip_quals``: the caller has checked that the batch's ``min_qual`` promise clears both quality thresholds of
        the summary it is going to run (``quals_not_needed``): no kernel then reads a quality, so the quality array --
        half of the plain bytes -- is neither uploaded nor allocated (``PvReadBatch.quals == NULL``). Results are
        identical; the one case the promise cannot decide (a read whose CIGAR runs over its own end) is flagged by the
        kernels (status bit 4) and the caller re-runs after ``ensure_quals()``."""
        self.host = host
        self.device = torch.device(device)
        self.quals_skipped = bool(skip_quals and host.min_qual > 0 and host.quals.size)
        self.t = {}
        self.packed = None          # 4-bit or 2-bit bases
        self.packed_exc = None      # exception list of the 2-bit form
        self.packed_q = None        # bit-packed qualities
        self.packed_c = None        # 16-bit CIGAR words
        self.patches = None         # reference-predicted bases: patch entries (offsets travel with the small arrays)
        self.codes8 = None          # 8-bit CIGAR codes
        self.esc8 = None            # ... and their escape words (offsets travel with the small arrays)
        self.qpatches = None        # quality predicates: patch entries over host.quals_fill (offsets with the small arrays)
        self._unpacked = True
        small = []
        for name in ARRAY_NAMES:
            a = getattr(host, name)
            if name == "bases" and host.bases_patch is not None and a.size:
                # only the bases that differ from their reference-based prediction travel (pv_unpack_bases_ref)
                src = host.bases_patch if host.bases_patch.size else np.zeros(1, np.uint16)
                self.patches = torch.from_numpy(src.view(np.int16)).to(self.device, non_blocking=non_blocking)
                self.t[name] = torch.empty(a.size, dtype=torch.uint8, device=self.device)
                self._unpacked = False
                continue
            if name == "bases" and host.bases2 is not None and a.size:
                self.packed = torch.empty(a.size // 4 + 16, dtype=torch.uint8, device=self.device)
                self.packed[:a.size // 4].copy_(_to_torch(host.bases2), non_blocking=non_blocking)
                if host.base_exceptions.size:
                    self._exc_stage = torch.empty(host.base_exceptions.size, dtype=torch.int64, pin_memory=bool(non_blocking))
                    self._exc_stage.numpy()[:] = host.base_exceptions.view(np.int64)
                    self.packed_exc = self._exc_stage.to(self.device, non_blocking=non_blocking)
                self.t[name] = torch.empty(a.size, dtype=torch.uint8, device=self.device)
                self._unpacked = False
                continue
            if name == "bases" and host.bases4 is not None and a.size:
                # bases travel in the BAM-native 4-bit form and are expanded on the device (pv_unpack_bases4)
                self.packed = _to_torch(host.bases4).to(self.device, non_blocking=non_blocking)
                self.t[name] = torch.empty(a.size, dtype=torch.uint8, device=self.device)
                self._unpacked = False
                continue
            if name == "quals" and self.quals_skipped:
                continue
            if name == "quals" and host.quals_patch is not None and a.size:
                # surrogate qualities that keep every threshold test of the summary (pv_unpack_quals_pred)
                src = host.quals_patch if host.quals_patch.size else np.zeros(1, np.uint16)
                self.qpatches = torch.from_numpy(src.view(np.int16)).to(self.device, non_blocking=non_blocking)
                self.t[name] = torch.empty(a.size + 16, dtype=torch.uint8, device=self.device)[:a.size]
                self._unpacked = False
                continue
            if name == "quals" and host.quals_packed is not None and a.size:
                n_words = (a.size + 31) // 32 * host.qual_bits
                self.packed_q = torch.empty(n_words * 4 + 16, dtype=torch.uint8, device=self.device)
                need = min((a.size * host.qual_bits + 7) // 8, host.quals_packed.shape[0])
                self.packed_q[:need].copy_(_to_torch(host.quals_packed[:need]), non_blocking=non_blocking)
                self.t[name] = torch.empty(a.size + 16, dtype=torch.uint8, device=self.device)[:a.size]
                self._unpacked = False
                continue
            if name == "cigar" and host.cigar8 is not None and a.size:
                self.codes8 = torch.from_numpy(host.cigar8).to(self.device, non_blocking=non_blocking)
                esc = host.cigar_esc if host.cigar_esc.size else np.zeros(1, np.uint32)
                self.esc8 = torch.from_numpy(esc.view(np.int32)).to(self.device, non_blocking=non_blocking)
                self.t[name] = torch.empty(a.size, dtype=torch.int32, device=self.device)
                self._unpacked = False
                continue
            if name == "cigar" and host.cigar16 is not None and a.size:
                self.packed_c = torch.from_numpy(host.cigar16.view(np.int16)).to(self.device, non_blocking=non_blocking)
                self.t[name] = torch.empty(a.size, dtype=torch.int32, device=self.device)
                self._unpacked = False
                continue
            if name not in ("bases", "quals", "cigar", "ref"):
                small.append(name)                       # per-read / per-region arrays: staged together below
                continue
            src = _to_torch(a) if a.size else torch.zeros(1, dtype=_TORCH_DT[np.dtype(a.dtype)])
            self.t[name] = src.to(self.device, non_blocking=non_blocking)
        # the ~14 small arrays (offsets rebuilt per view live in pageable memory, whose "async" copies block the host) go
        # through ONE pinned staging buffer and one copy; the device tensors are views of one allocation
        if self.patches is not None:
            small.append("read_patch_off")
        if self.codes8 is not None:
            small.append("read_esc_off")
        if self.qpatches is not None:
            small.append("read_qpatch_off")
        offs, total = {}, 0
        for name in small:
            offs[name] = total
            total += (max(getattr(host, name).nbytes, 8) + 255) // 256 * 256
        self._stage = torch.empty(max(total, 256), dtype=torch.uint8, pin_memory=bool(non_blocking))
        stage_np = self._stage.numpy()
        for name in small:
            a = getattr(host, name)
            stage_np[offs[name]:offs[name] + a.nbytes] = a.view(np.uint8).reshape(-1)
        self._small_dev = self._stage.to(self.device, non_blocking=non_blocking)
        for name in small:
            a = getattr(host, name)
            dt = _TORCH_DT[np.dtype(a.dtype)]
            self.t[name] = self._small_dev[offs[name]:offs[name] + max(a.nbytes, 8)].view(dt)[:max(a.size, 1) if a.size == 0 else a.size]
        self.region_len = np.ascontiguousarray(host.region_len)
        self.total_positions = int(self.region_len.sum())
        self.struct = host.as_struct({n: (self.t[n].data_ptr() if n in self.t else 0) for n in ARRAY_NAMES})
        if self.qpatches is not None:
            # the device array holds SURROGATE qualities: the fill byte, unless a patch lowers it
            pv = host.quals_patch
            vals = (pv >> 8)[(pv & 255) != 255] if pv.size else pv
            self.struct.min_qual = int(min(host.quals_fill, int(vals.min()))) if vals.size else int(host.quals_fill)
        if not defer_unpack:
            self.unpack()

    def unpack(self):
        """Expand the compact wire forms on the CURRENT stream (a copy stream should only carry copies: the caller runs
        this on the compute stream after waiting for the upload)."""
        if self._unpacked:
            return
        lib = capi.load()
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        if self.codes8 is not None:             # first: the base prediction below walks the CIGAR
            capi.check(lib.pv_unpack_cigar8(C.byref(self.struct), C.c_void_p(self.codes8.data_ptr()),
                                            C.c_void_p(self.t["read_esc_off"].data_ptr()), C.c_void_p(self.esc8.data_ptr()),
                                            C.c_void_p(self.t["cigar"].data_ptr()), st))
        if self.packed_c is not None:
            out = self.t["cigar"]
            capi.check(lib.pv_unpack_cigar16(C.c_void_p(self.packed_c.data_ptr()), out.numel(), C.c_void_p(out.data_ptr()), st))
        if self.patches is not None:
            capi.check(lib.pv_unpack_bases_ref(C.byref(self.struct), C.c_void_p(self.t["read_patch_off"].data_ptr()),
                                               C.c_void_p(self.patches.data_ptr()), C.c_void_p(self.t["bases"].data_ptr()), st))
        if self.packed is not None and self.host.bases2 is not None:
            out = self.t["bases"]
            ne = 0 if self.packed_exc is None else self.packed_exc.numel()
            capi.check(lib.pv_unpack_bases2(C.c_void_p(self.packed.data_ptr()), out.numel(),
                                            C.c_void_p(self.packed_exc.data_ptr() if ne else None), ne, C.c_void_p(out.data_ptr()), st))
        elif self.packed is not None:
            out = self.t["bases"]
            capi.check(lib.pv_unpack_bases4(C.c_void_p(self.packed.data_ptr()), out.numel(), C.c_void_p(out.data_ptr()), st))
        if self.qpatches is not None:
            capi.check(lib.pv_unpack_quals_pred(C.byref(self.struct), int(self.host.quals_fill),
                                                C.c_void_p(self.t["read_qpatch_off"].data_ptr()), C.c_void_p(self.qpatches.data_ptr()),
                                                C.c_void_p(self.t["quals"].data_ptr()), st))
        if self.packed_q is not None:
            out = self.t["quals"]
            capi.check(lib.pv_unpack_quals(C.c_void_p(self.packed_q.data_ptr()), out.numel(), int(self.host.qual_bits),
                                           C.c_void_p(out.data_ptr()), st))
        self._unpacked = True

    def with_min_qual(self, min_qual: int) -> "DeviceBatch":
        """The same device arrays under a different ``min_qual`` promise (0 = none: every quality is loaded and tested)."""
        import copy
        other = copy.copy(self)
        other.struct = type(self.struct).from_buffer_copy(self.struct)
        other.struct.min_qual = int(min_qual)
        return other

    def ensure_quals(self):
        """Uploads the plain qualities of a batch that was created with ``skip_quals`` (blocking; the rare re-run path)."""
        if not self.quals_skipped:
            return
        q = self.host.quals
        self.t["quals"] = torch.from_numpy(np.ascontiguousarray(q)).to(self.device)
        self.struct.quals = self.t["quals"].data_ptr()
        self.quals_skipped = False

    def record_stream(self, stream):
        for t in list(self.t.values()) + [x for x in (self.packed, self.packed_exc, self.packed_q, self.packed_c, self.patches, self.codes8, self.esc8, self.qpatches, self._small_dev) if x is not None]:
            t.record_stream(stream)

    @property
    def h2d_bytes(self) -> int:
        n = int(sum(getattr(self.host, n).nbytes for n in ARRAY_NAMES))
        if self.quals_skipped:
            n -= self.host.quals.nbytes
        if self.patches is not None:
            n -= self.host.bases.nbytes - self.host.bases_patch.nbytes - self.host.read_patch_off.nbytes
        elif self.packed is not None and self.host.bases2 is not None:
            n -= self.host.bases.nbytes - self.host.bases2.nbytes - self.host.base_exceptions.nbytes
        elif self.packed is not None:
            n -= self.host.bases.nbytes - self.host.bases4.nbytes
        if self.qpatches is not None and not self.quals_skipped:
            n -= self.host.quals.nbytes - self.host.quals_patch.nbytes - self.host.read_qpatch_off.nbytes
        if self.packed_q is not None and not self.quals_skipped:
            n -= self.host.quals.nbytes - min((self.host.quals.size * self.host.qual_bits + 7) // 8, self.host.quals_packed.nbytes)
        if self.codes8 is not None:
            n -= self.host.cigar.nbytes - self.host.cigar8.nbytes - self.host.cigar_esc.nbytes - self.host.read_esc_off.nbytes
        elif self.packed_c is not None:
            n -= self.host.cigar.nbytes - self.host.cigar16.nbytes
        return n


class _Shape:
    """What the pipeline asks a batch's ``host`` for when there is no host copy (sizes and the quality promise)."""

    def __init__(self, n_reads, n_bases, n_ops, n_regions, min_qual, contigs, region_contig_len):
        self.n_reads, self.n_bases, self.n_ops, self.n_regions = int(n_reads), int(n_bases), int(n_ops), int(n_regions)
        self.min_qual = int(min_qual)
        self.contigs = contigs
        self.region_contig_len = region_contig_len
        self.quals_pred_thr = None


class DeviceReadBatch:
    """A ``PvReadBatch`` that was BORN on the device (``ingest_gpu``: BAM decoded by kernels): same interface as
    :class:`DeviceBatch` for the summary / pipeline code, no host arrays behind it. ``tensors`` maps every name of
    ``ARRAY_NAMES`` to a device tensor; the region fields are also kept as numpy (``regions``)."""

    def __init__(self, tensors, regions, min_qual, device, contigs=None, region_contig_len=None, h2d_bytes=0):
        self.t = dict(tensors)
        self.device = torch.device(device)
        self.regions = regions                                # dict of numpy arrays: the region_* fields
        self.region_len = np.ascontiguousarray(regions["region_ref_end"] - regions["region_ref_start"] + 1).astype(np.int64)
        self.total_positions = int(self.region_len.sum())
        n_reads = int(self.t["read_pos"].numel()) if "read_pos" in self.t else 0
        self.host = _Shape(n_reads, self.t["bases"].numel(), self.t["cigar"].numel(), len(self.region_len), min_qual,
                           contigs or [], region_contig_len)
        self.quals_skipped = False
        self.qpatches = None
        self._h2d = int(h2d_bytes)
        s = PvReadBatchStructT()
        s.n_reads, s.n_bases, s.n_ops, s.n_ref = n_reads, self.host.n_bases, self.host.n_ops, int(self.t["ref"].numel())
        s.n_regions = self.host.n_regions
        s.min_qual = int(min_qual)
        for name in ARRAY_NAMES:
            setattr(s, name, self.t[name].data_ptr() if self.t[name].numel() else None)
        self.struct = s

    def unpack(self):
        pass

    def ensure_quals(self):
        pass

    def with_min_qual(self, min_qual: int) -> "DeviceReadBatch":
        import copy
        other = copy.copy(self)
        other.struct = type(self.struct).from_buffer_copy(self.struct)
        other.struct.min_qual = int(min_qual)
        return other

    def record_stream(self, stream):
        for t in self.t.values():
            t.record_stream(stream)

    @property
    def h2d_bytes(self) -> int:
        return self._h2d

    def to_host(self) -> ReadBatch:
        """Download (tests, interop)."""
        a = {}
        for name in ARRAY_NAMES:
            v = self.t[name].cpu().numpy()
            a[name] = v.view(np.uint32) if name == "cigar" else v
        n = self.host.n_reads
        for name in ("read_pos", "read_base_off", "read_len", "read_cigar_off", "read_n_ops", "read_flags", "read_mapq"):
            a[name] = a[name][:n]
        return ReadBatch(contigs=list(self.host.contigs), min_qual=self.host.min_qual, region_contig_len=self.host.region_contig_len, **a)


class SummaryWorkspace:
    """Caller-allocated outputs + scratch of ``pv_summary_regions`` (reused across calls of the same shape)."""

    @staticmethod
    def scratch_bytes(n_reads, n_ops, n_regions, total_positions, capacity, max_region_len=0) -> int:
        return int(capi.load().pv_summary_workspace_bytes(int(n_reads), int(n_ops), int(n_regions), int(total_positions),
                                                          int(max_region_len), int(capacity)))

    def __init__(self, n_reads, n_ops, n_regions, total_positions, capacity, device="cuda", want_dense=False,
                 max_region_len=0, min_scratch_bytes=0):
        self.capacity = int(capacity)
        dev = torch.device(device)
        ws_bytes = max(self.scratch_bytes(n_reads, n_ops, n_regions, total_positions, capacity, max_region_len), int(min_scratch_bytes))
        self.ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        self.windows = torch.empty((capacity, capi.PV_WINDOW, capi.PV_FEATURES), dtype=torch.int16, device=dev)
        self.position = torch.empty(capacity, dtype=torch.int64, device=dev)
        self.region = torch.empty(capacity, dtype=torch.int32, device=dev)
        self.depth = torch.empty(capacity, dtype=torch.int32, device=dev)
        self.frequency = torch.empty(capacity, dtype=torch.int32, device=dev)
        self.allele = torch.empty((capacity, capi.PV_ALLELE_BYTES), dtype=torch.uint8, device=dev)
        self.allele_len = torch.empty(capacity, dtype=torch.uint8, device=dev)
        self.count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.dense = torch.empty((total_positions, capi.PV_FEATURES), dtype=torch.int16, device=dev) if want_dense else None
        self.out = capi.PvCandidatesStruct(capacity, self.windows.data_ptr(), self.position.data_ptr(),
                                           self.region.data_ptr(), self.depth.data_ptr(), self.frequency.data_ptr(),
                                           self.allele.data_ptr(), self.allele_len.data_ptr())

    @classmethod
    def for_batch(cls, db: DeviceBatch, capacity: int, want_dense=False) -> "SummaryWorkspace":
        h = db.host
        return cls(h.n_reads, h.n_ops, h.n_regions, db.total_positions, capacity, db.device, want_dense,
                   max_region_len=int(db.region_len.max()) if db.region_len.size else 0)

    def status(self) -> int:
        return int(self.ws[12:16].view(torch.int32).item())


def quals_not_needed(min_qual: int, thr) -> bool:
    """True when a batch promising ``min_qual`` passes every quality test of the summary for these thresholds
    (same rule as pv_summary_regions: min_qual >= ceil(min_snp_baseq) and >= min_indel_baseq)."""
    import math
    t = capi.thresholds_struct(thr)
    return bool(min_qual > 0 and min_qual >= math.ceil(max(0.0, float(t.min_snp_baseq))) and min_qual >= float(t.min_indel_baseq))


def summary_regions(db: DeviceBatch, thr, ws: SummaryWorkspace, window: int = 32, features: int = 26):
    """Launch the summary kernel chain on the current stream (asynchronous; no host sync)."""
    lib = capi.load()
    t = capi.thresholds_struct(thr)
    if db.qpatches is not None and db.host.quals_pred_thr != (float(t.min_snp_baseq), float(t.min_indel_baseq)):
        raise capi.PvError(-1, "the batch carries quality predicates packed for thresholds %s, the summary was asked for %s"
                           % (db.host.quals_pred_thr, (float(t.min_snp_baseq), float(t.min_indel_baseq))))
    stream = torch.cuda.current_stream(db.device).cuda_stream
    rc = lib.pv_summary_regions(C.byref(db.struct), db.region_len.ctypes.data, db.total_positions, C.byref(t), window,
                                features, C.byref(ws.out), ws.count.data_ptr(), ws.ws.data_ptr(), ws.ws.numel(),
                                ws.dense.data_ptr() if ws.dense is not None else None, C.c_void_p(stream))
    capi.check(rc)
