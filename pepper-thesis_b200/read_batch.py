"""Packed SoA read batch -- the host side of ``PvReadBatch`` (include/pepper_b200.h).

Replaces the AoS ``list[type_read]`` the reference hands across pybind
(/root/reference/pepper_variant/modules/cpp/read.h:60-108, cigar.h:30-53): one byte per base, one byte per
quality, BAM-encoded CIGAR words, per-read headers, per-region headers. Reads of a region are contiguous.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

_READ_FIELDS = [("read_pos", np.int64), ("read_base_off", np.int64), ("read_len", np.int32),
                ("read_cigar_off", np.int64), ("read_n_ops", np.int32), ("read_flags", np.uint8),
                ("read_mapq", np.uint8)]
_REGION_FIELDS = [("region_ref_start", np.int64), ("region_ref_end", np.int64), ("region_cand_start", np.int64),
                  ("region_cand_end", np.int64), ("region_ref_off", np.int64), ("region_ref_len", np.int64)]


class PvReadBatchStruct(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("n_bases", C.c_int64), ("n_ops", C.c_int64), ("n_ref", C.c_int64),
                ("n_regions", C.c_int32),
                ("read_pos", C.c_void_p), ("read_base_off", C.c_void_p), ("read_len", C.c_void_p),
                ("read_cigar_off", C.c_void_p), ("read_n_ops", C.c_void_p), ("read_flags", C.c_void_p),
                ("read_mapq", C.c_void_p),
                ("bases", C.c_void_p), ("quals", C.c_void_p), ("cigar", C.c_void_p),
                ("region_ref_start", C.c_void_p), ("region_ref_end", C.c_void_p), ("region_cand_start", C.c_void_p),
                ("region_cand_end", C.c_void_p), ("region_ref_off", C.c_void_p), ("region_ref_len", C.c_void_p),
                ("region_read_begin", C.c_void_p),
                ("ref", C.c_void_p), ("bases4", C.c_void_p),
                ("quals_packed", C.c_void_p), ("qual_bits", C.c_int32), ("min_qual", C.c_int32), ("cigar16", C.c_void_p),
                ("bases2", C.c_void_p), ("base_exceptions", C.c_void_p), ("n_base_exceptions", C.c_int64)]


ARRAY_NAMES = [n for n, _ in _READ_FIELDS] + ["bases", "quals", "cigar"] + [n for n, _ in _REGION_FIELDS] + \
              ["region_read_begin", "ref"]


@dataclass
class ReadBatch:
    """Host-resident packed batch (numpy arrays, C-contiguous)."""
    read_pos: np.ndarray
    read_base_off: np.ndarray
    read_len: np.ndarray
    read_cigar_off: np.ndarray
    read_n_ops: np.ndarray
    read_flags: np.ndarray
    read_mapq: np.ndarray
    bases: np.ndarray
    quals: np.ndarray
    cigar: np.ndarray
    region_ref_start: np.ndarray
    region_ref_end: np.ndarray
    region_cand_start: np.ndarray
    region_cand_end: np.ndarray
    region_ref_off: np.ndarray
    region_ref_len: np.ndarray
    region_read_begin: np.ndarray
    ref: np.ndarray
    contigs: List[str] = field(default_factory=list)
    bases4: Optional[np.ndarray] = None      # optional BAM-native 4-bit packing of `bases` (n_bases / 2 bytes)
    quals_packed: Optional[np.ndarray] = None   # optional dense bit stream of `quals` (qual_bits per quality)
    qual_bits: int = 0
    cigar16: Optional[np.ndarray] = None     # optional low 16 bits of every CIGAR word (all op lengths < 4096)
    bases2: Optional[np.ndarray] = None      # optional 2-bit packing of `bases` (n_bases / 4 bytes) ...
    base_exceptions: Optional[np.ndarray] = None   # ... + uint64 (index << 8 | byte) of every base that is not A/C/G/T
    cigar8: Optional[np.ndarray] = None      # optional 8-bit CIGAR codes (one byte per op) ...
    cigar_esc: Optional[np.ndarray] = None   # ... + the full uint32 words of the ops that do not fit a code byte ...
    read_esc_off: Optional[np.ndarray] = None   # ... of read r at [read_esc_off[r], read_esc_off[r + 1])
    bases_patch: Optional[np.ndarray] = None    # optional reference-predicted form of `bases`: uint16 patch entries ...
    read_patch_off: Optional[np.ndarray] = None  # ... of read r at [read_patch_off[r], read_patch_off[r + 1])
    quals_patch: Optional[np.ndarray] = None    # optional quality-predicate form of `quals` (pv_pack_quals_pred): uint16 patch
    read_qpatch_off: Optional[np.ndarray] = None  # entries of read r at [read_qpatch_off[r], read_qpatch_off[r + 1]) ...
    quals_fill: int = 0                         # ... over this fill byte; valid for the thresholds in quals_pred_thr only
    quals_pred_thr: Optional[tuple] = None      # (min_snp_baseq, min_indel_baseq) the predicate form was packed for
    min_qual: int = 0                           # promise: every quality of every read base is >= this (0 = none); scan_min_qual()
    region_contig_len: Optional[np.ndarray] = None   # int64 [n_regions] length of each region's contig (ingest fills it): the
                                                # stage-3 filter clips its reference context there like FASTA_handler does

    # ---- shape helpers -------------------------------------------------------------------------------------
    @property
    def n_reads(self) -> int:
        return int(self.read_pos.shape[0])

    @property
    def n_regions(self) -> int:
        return int(self.region_ref_start.shape[0])

    @property
    def n_bases(self) -> int:
        return int(self.bases.shape[0])

    @property
    def n_ops(self) -> int:
        return int(self.cigar.shape[0])

    @property
    def region_len(self) -> np.ndarray:
        return (self.region_ref_end - self.region_ref_start + 1).astype(np.int64)

    @property
    def total_positions(self) -> int:
        return int(self.region_len.sum())

    @property
    def candidate_bp(self) -> int:
        """Reference positions of the candidate regions (the unit of the Mbp/s metric)."""
        return int((self.region_cand_end - self.region_cand_start + 1).sum())

    def algorithmic_bytes(self, n_candidates: int) -> int:
        """2*N_bases + 4*N_ops + 32*N_reads + L + K*(33*26*2 + 32)."""
        real_bases = int(self.read_len.astype(np.int64).sum())
        return 2 * real_bases + 4 * self.n_ops + 32 * self.n_reads + self.total_positions + \
            n_candidates * (33 * 26 * 2 + 32)

    def as_struct(self, arrays=None) -> PvReadBatchStruct:
        """ctypes view. ``arrays`` may map names to integer (device) pointers instead of the numpy arrays."""
        s = PvReadBatchStruct()
        s.n_reads, s.n_bases, s.n_ops, s.n_ref = self.n_reads, self.n_bases, self.n_ops, int(self.ref.shape[0])
        s.n_regions = self.n_regions
        s.min_qual = int(self.min_qual)
        for name in ARRAY_NAMES:
            if arrays is not None:
                setattr(s, name, int(arrays[name]))
            else:
                a = getattr(self, name)
                assert a.flags["C_CONTIGUOUS"], name
                setattr(s, name, a.ctypes.data)
        host = arrays is None
        s.bases4 = self.bases4.ctypes.data if (host and self.bases4 is not None) else None
        s.quals_packed = self.quals_packed.ctypes.data if (host and self.quals_packed is not None) else None
        s.qual_bits = int(self.qual_bits) if (host and self.quals_packed is not None) else 0
        s.cigar16 = self.cigar16.ctypes.data if (host and self.cigar16 is not None) else None
        use2 = host and self.bases2 is not None
        s.bases2 = self.bases2.ctypes.data if use2 else None
        s.base_exceptions = self.base_exceptions.ctypes.data if (use2 and self.base_exceptions.size) else None
        s.n_base_exceptions = int(self.base_exceptions.shape[0]) if use2 else 0
        return s

    @staticmethod
    def _host_buffer(nbytes: int, pinned: bool):
        if pinned:
            import torch
            t = torch.empty(max(1, nbytes), dtype=torch.uint8, pin_memory=True)
            return t, t.numpy()[:nbytes]
        return None, np.empty(nbytes, np.uint8)

    def scan_min_qual(self, threads: int = 0) -> "ReadBatch":
        """Sets ``min_qual`` (PvReadBatch.min_qual) from the plain qualities: when it clears both quality thresholds of a
        summary call, the tile kernel never loads a quality. Views made afterwards inherit it (a lower bound stays one)."""
        from . import capi
        import os
        lib = capi.load()
        if self.n_reads:
            st = self.as_struct()
            self.min_qual = int(lib.pv_min_qual(C.byref(st), threads or min(32, os.cpu_count() or 1)))
        return self

    def pack_quals(self, threads: int = 0, pinned: bool = False) -> "ReadBatch":
        """Adds the bit-packed wire form of the qualities (lossless: as many bits as the largest quality needs).
        A batch whose qualities need 8 bits is left alone."""
        from . import capi
        import os
        lib = capi.load()
        threads = threads or min(32, os.cpu_count() or 1)
        if self.n_bases == 0 or self.n_bases % 16:
            return self
        bits = int(lib.pv_qual_bits(self.quals.ctypes.data, self.n_bases, threads))
        if bits >= 8:
            return self
        nbytes = (self.n_bases + 31) // 32 * bits * 4
        self._qualsp_owner, out = self._host_buffer(nbytes, pinned)
        capi.check(lib.pv_pack_quals(self.quals.ctypes.data, self.n_bases, bits, out.ctypes.data, threads))
        self.quals_packed, self.qual_bits = out, bits
        return self

    def pack_cigar16(self, threads: int = 0, pinned: bool = False) -> "ReadBatch":
        """Adds the 16-bit wire form of the CIGAR words when every op length is < 4096 (else leaves the batch alone)."""
        from . import capi
        import os
        lib = capi.load()
        if self.n_ops == 0 or int(self.cigar.max()) >> 16:
            return self
        self._cigar16_owner, out = self._host_buffer(self.n_ops * 2, pinned)
        out = out.view(np.uint16)
        capi.check(lib.pv_pack_cigar16(self.cigar.ctypes.data, self.n_ops, out.ctypes.data, threads or min(32, os.cpu_count() or 1)))
        self.cigar16 = out
        return self

    def pack_cigar8(self, threads: int = 0, pinned: bool = False) -> "ReadBatch":
        """Adds the 8-bit wire form of the CIGAR (pv_pack_cigar8): one code byte per M/I/D op of length 1..64, the full
        word of every other op in a separate escape stream."""
        from . import capi
        import os
        lib = capi.load()
        threads = threads or min(32, os.cpu_count() or 1)
        if self.n_ops == 0 or self.n_reads == 0:
            return self
        st = self.as_struct()
        self._cigar8_owner, codes = self._host_buffer(self.n_ops, pinned)
        off = np.zeros(self.n_reads + 1, np.int64)
        capi.check(lib.pv_pack_cigar8(C.byref(st), codes.ctypes.data, off.ctypes.data, None, 0, threads))
        total = int(off[-1])
        self._cigar_esc_owner, esc = self._host_buffer(max(4, total * 4), pinned)
        esc = esc[:total * 4].view(np.uint32)
        if total:
            capi.check(lib.pv_pack_cigar8(C.byref(st), codes.ctypes.data, off.ctypes.data, esc.ctypes.data, total, threads))
        self.cigar8, self.cigar_esc, self.read_esc_off = codes, esc, off
        return self

    def pack_bases2(self, threads: int = 0, pinned: bool = False) -> "ReadBatch":
        """Adds the 2-bit wire form of the bases + the list of bases that are not upper-case A/C/G/T (any byte value is
        representable). Left alone when the exception list would cost more than 2-bit packing saves over 4 bits."""
        from . import capi
        import os
        lib = capi.load()
        threads = threads or min(32, os.cpu_count() or 1)
        if self.n_bases == 0 or self.n_bases % 16:
            return self
        self._bases2_owner, out = self._host_buffer(self.n_bases // 4, pinned)
        n_exc = C.c_int64(0)
        capi.check(lib.pv_pack_bases2(self.bases.ctypes.data, self.n_bases, out.ctypes.data, None, C.byref(n_exc), threads))
        if n_exc.value * 8 > self.n_bases // 4:
            return self
        exc = np.zeros(n_exc.value, np.uint64)
        if n_exc.value:
            capi.check(lib.pv_pack_bases2(self.bases.ctypes.data, self.n_bases, out.ctypes.data, exc.ctypes.data, C.byref(n_exc), threads))
        self.bases2, self.base_exceptions = out, exc
        return self

    def pack_bases_ref(self, threads: int = 0, pinned: bool = False) -> "ReadBatch":
        """Adds the reference-predicted wire form of the bases (pv_pack_bases_ref): the device rebuilds every read from
        the batch's CIGAR and reference, only the bases that differ from that prediction travel (16-bit patch entries)."""
        from . import capi
        import os
        lib = capi.load()
        threads = threads or min(32, os.cpu_count() or 1)
        if self.n_reads == 0:
            return self
        st = self.as_struct()
        off = np.zeros(self.n_reads + 1, np.int64)
        capi.check(lib.pv_pack_bases_ref(C.byref(st), off.ctypes.data, None, 0, threads))
        total = int(off[-1])
        self._patch_owner, out = self._host_buffer(max(2, total * 2), pinned)
        out = out[:total * 2].view(np.uint16)
        if total:
            capi.check(lib.pv_pack_bases_ref(C.byref(st), off.ctypes.data, out.ctypes.data, total, threads))
        self.bases_patch, self.read_patch_off = out, off
        return self

    def pack_quals_pred(self, min_snp_baseq: float, min_indel_baseq: float, threads: int = 0, pinned: bool = False) -> "ReadBatch":
        """Adds the quality-PREDICATE wire form (pv_pack_quals_pred): a fill byte + per-read patch entries of surrogate
        qualities on which every quality test of the summary (region_summary.cpp:377,393,448-463) has the same outcome as
        on the real qualities -- for these two thresholds only. Summaries / candidates are bit-identical; the qualities
        themselves do not travel. Left alone (lossless forms stay in charge) when a threshold is above 127."""
        from . import capi
        import os
        lib = capi.load()
        threads = threads or min(32, os.cpu_count() or 1)
        if self.n_reads == 0 or self.n_bases % 16 or not (min_snp_baseq <= 127 and min_indel_baseq <= 127):
            return self
        st = self.as_struct()
        off = np.zeros(self.n_reads + 1, np.int64)
        fill = C.c_uint8(0)
        capi.check(lib.pv_pack_quals_pred(C.byref(st), float(min_snp_baseq), float(min_indel_baseq), C.byref(fill),
                                          off.ctypes.data, None, 0, threads))
        total = int(off[-1])
        self._qpatch_owner, out = self._host_buffer(max(2, total * 2), pinned)
        out = out[:total * 2].view(np.uint16)
        if total:
            capi.check(lib.pv_pack_quals_pred(C.byref(st), float(min_snp_baseq), float(min_indel_baseq), C.byref(fill),
                                              off.ctypes.data, out.ctypes.data, total, threads))
        self.quals_patch, self.read_qpatch_off, self.quals_fill = out, off, int(fill.value)
        self.quals_pred_thr = (float(min_snp_baseq), float(min_indel_baseq))
        return self

    def pack_wire(self, threads: int = 0, pinned: bool = False, bases_ref: Optional[bool] = None,
                  quals_pred: Optional[tuple] = None) -> "ReadBatch":
        """All compact (lossless) wire forms: reference-predicted bases when that is the smallest form (~0.6 bits per base
        at ONT error rates), else 2-bit bases + exceptions (else 4-bit when the alphabet allows); bit-packed qualities;
        16-bit CIGAR. ``bases_ref=False`` (or PV_WIRE_BASES_REF=0) keeps the bases in the 2-bit form, whose expansion
        kernel is 6x cheaper (it pays when the device, not the upload, is the bottleneck). ``quals_pred=(min_snp_baseq,
        min_indel_baseq)`` replaces the bit-packed qualities by the quality-predicate form (``pack_quals_pred``: the
        summary's results are unchanged, the qualities themselves are not kept -- opt-in, tied to those thresholds)."""
        from . import capi
        import os
        if bases_ref is None:
            bases_ref = os.environ.get("PV_WIRE_BASES_REF", "1") == "1"
        self.pack_bases2(threads, pinned)
        if bases_ref:
            self.pack_bases_ref(threads, pinned)
        if self.bases_patch is not None:
            ref_bytes = self.bases_patch.nbytes + self.read_patch_off.nbytes
            two_bytes = (self.bases2.nbytes + self.base_exceptions.nbytes) if self.bases2 is not None else self.n_bases
            if ref_bytes < two_bytes:
                self.bases2, self.base_exceptions, self._bases2_owner = None, None, None
            else:
                self.bases_patch, self.read_patch_off, self._patch_owner = None, None, None
        if self.bases2 is None and self.bases_patch is None:
            try:
                self.pack_bases4(threads, pinned)
            except capi.PvError:
                self.bases4 = None
        if quals_pred is not None:
            self.pack_quals_pred(quals_pred[0], quals_pred[1], threads, pinned)
        if self.quals_patch is None:
            self.pack_quals(threads, pinned)
        self.pack_cigar16(threads, pinned)
        if os.environ.get("PV_WIRE_CIGAR8", "1") == "1":
            self.pack_cigar8(threads, pinned)
            if self.cigar8 is not None:
                c8 = self.cigar8.nbytes + self.cigar_esc.nbytes + self.read_esc_off.nbytes
                other = self.cigar16.nbytes if self.cigar16 is not None else self.cigar.nbytes
                if c8 < other:
                    self.cigar16, self._cigar16_owner = None, None
                else:
                    self.cigar8, self.cigar_esc, self.read_esc_off = None, None, None
        return self

    def pin_uploaded(self) -> "ReadBatch":
        """Moves the big arrays the host path actually uploads (the packed wire forms are allocated pinned by
        ``pack_wire(pinned=True)``; this covers the reference and whatever has no packed form) into page-locked memory,
        so every upload is truly asynchronous -- without pinning the plain arrays that never travel."""
        import torch
        names = ["ref"]
        if self.bases2 is None and self.bases4 is None and self.bases_patch is None:
            names.append("bases")
        if self.quals_packed is None and self.quals_patch is None:
            names.append("quals")
        if self.cigar16 is None and self.cigar8 is None:
            names.append("cigar")
        self._pinned_owners = getattr(self, "_pinned_owners", {})
        for name in names:
            a = getattr(self, name)
            if a.size == 0:
                continue
            t = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=True)
            dst = t.numpy().view(a.dtype)
            dst[:] = a
            self._pinned_owners[name] = t
            setattr(self, name, dst)
        return self

    def pin_plain(self, with_quals: bool = True) -> "ReadBatch":
        """Moves the PLAIN big arrays (bases, cigar, ref and -- unless ``with_quals`` is False -- quals) into page-locked
        memory in place, so a host batch in its plain form uploads asynchronously at PCIe speed."""
        import torch
        self._pinned_owners = getattr(self, "_pinned_owners", {})
        for name in ["bases", "cigar", "ref"] + (["quals"] if with_quals else []):
            a = getattr(self, name)
            if a.size == 0 or name in self._pinned_owners:
                continue
            t = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=True)
            dst = t.numpy().view(a.dtype)
            dst[:] = a
            self._pinned_owners[name] = t
            setattr(self, name, dst)
        return self

    def _exceptions_view(self, b_lo: int, b_hi: int):
        if self.bases2 is None:
            return None
        idx = self.base_exceptions >> np.uint64(8)
        lo, hi = np.searchsorted(idx, [b_lo, b_hi])
        e = self.base_exceptions[lo:hi]
        return ((idx[lo:hi] - np.uint64(b_lo)) << np.uint64(8)) | (e & np.uint64(0xff))

    def pack_bases4(self, threads: int = 0, pinned: bool = False) -> "ReadBatch":
        """Adds the 4-bit wire form of the bases (raises if a base is outside the BAM nt16 alphabet)."""
        from . import capi
        import os
        lib = capi.load()
        if pinned:
            import torch
            self._bases4_owner = torch.empty(max(1, self.n_bases // 2), dtype=torch.uint8, pin_memory=True)
            out = self._bases4_owner.numpy()[:self.n_bases // 2]
        else:
            out = np.empty(self.n_bases // 2, np.uint8)
        capi.check(lib.pv_pack_bases4(self.bases.ctypes.data, self.n_bases, out.ctypes.data, threads or min(32, os.cpu_count() or 1)))
        self.bases4 = out
        return self

    def region_range_view(self, r0: int, r1: int) -> "ReadBatch":
        """Regions [r0, r1) as a batch whose big arrays (bases, quals, cigar, ref) are VIEWS of this batch (no copy);
        only the small per-read / per-region offset arrays are rebuilt. Requires the layout pack_regions / synth
        produce: reads, bases, ops and reference of consecutive regions stored consecutively."""
        rb, re_ = int(self.region_read_begin[r0]), int(self.region_read_begin[r1])
        f_lo = int(self.region_ref_off[r0])
        f_hi = int(self.region_ref_off[r1 - 1] + self.region_ref_len[r1 - 1])
        if re_ > rb:
            b_lo = int(self.read_base_off[rb])
            b_hi = int(self.read_base_off[re_ - 1]) + ((int(self.read_len[re_ - 1]) + 15) & ~15)
            b_hi = min(b_hi, self.n_bases)
            c_lo = int(self.read_cigar_off[rb])
            c_hi = int(self.read_cigar_off[re_ - 1] + self.read_n_ops[re_ - 1])
        else:
            b_lo = b_hi = c_lo = c_hi = 0
        return ReadBatch(
            read_pos=self.read_pos[rb:re_], read_base_off=self.read_base_off[rb:re_] - b_lo, read_len=self.read_len[rb:re_],
            read_cigar_off=self.read_cigar_off[rb:re_] - c_lo, read_n_ops=self.read_n_ops[rb:re_],
            read_flags=self.read_flags[rb:re_], read_mapq=self.read_mapq[rb:re_],
            bases=self.bases[b_lo:b_hi], quals=self.quals[b_lo:b_hi], cigar=self.cigar[c_lo:c_hi],
            region_ref_start=self.region_ref_start[r0:r1], region_ref_end=self.region_ref_end[r0:r1],
            region_cand_start=self.region_cand_start[r0:r1], region_cand_end=self.region_cand_end[r0:r1],
            region_ref_off=self.region_ref_off[r0:r1] - f_lo, region_ref_len=self.region_ref_len[r0:r1],
            region_read_begin=self.region_read_begin[r0:r1 + 1] - rb, ref=self.ref[f_lo:f_hi],
            contigs=self.contigs[r0:r1] if self.contigs else [],
            bases4=self.bases4[b_lo // 2:b_hi // 2] if self.bases4 is not None else None,
            # quality i sits at bit i * qual_bits: a 16-aligned base offset is a whole number of bytes
            quals_packed=(self.quals_packed[b_lo * self.qual_bits // 8:min(self.quals_packed.shape[0], ((b_hi + 31) // 32 * 32) * self.qual_bits // 8)]
                          if self.quals_packed is not None and b_lo % 16 == 0 and (b_hi - b_lo) % 16 == 0 else None),
            qual_bits=self.qual_bits if self.quals_packed is not None else 0,
            cigar16=self.cigar16[c_lo:c_hi] if self.cigar16 is not None else None,
            bases2=self.bases2[b_lo // 4:b_hi // 4] if self.bases2 is not None else None,
            base_exceptions=self._exceptions_view(b_lo, b_hi),
            cigar8=self.cigar8[c_lo:c_hi] if self.cigar8 is not None else None,
            cigar_esc=(self.cigar_esc[int(self.read_esc_off[rb]):int(self.read_esc_off[re_])]
                       if self.cigar8 is not None else None),
            read_esc_off=(self.read_esc_off[rb:re_ + 1] - self.read_esc_off[rb] if self.cigar8 is not None else None),
            bases_patch=(self.bases_patch[int(self.read_patch_off[rb]):int(self.read_patch_off[re_])]
                         if self.bases_patch is not None else None),
            read_patch_off=(self.read_patch_off[rb:re_ + 1] - self.read_patch_off[rb]
                            if self.bases_patch is not None else None),
            quals_patch=(self.quals_patch[int(self.read_qpatch_off[rb]):int(self.read_qpatch_off[re_])]
                         if self.quals_patch is not None and (b_hi - b_lo) % 16 == 0 else None),
            read_qpatch_off=(self.read_qpatch_off[rb:re_ + 1] - self.read_qpatch_off[rb]
                             if self.quals_patch is not None and (b_hi - b_lo) % 16 == 0 else None),
            quals_fill=self.quals_fill, quals_pred_thr=self.quals_pred_thr, min_qual=self.min_qual,
            region_contig_len=self.region_contig_len[r0:r1] if self.region_contig_len is not None else None)

    def region_slice(self, r: int) -> "ReadBatch":
        """A single-region batch sharing no offsets with the parent (used for per-region oracle calls)."""
        return select_regions(self, [r])


def select_regions(batch: ReadBatch, regions: Sequence[int]) -> ReadBatch:
    """Re-pack the given regions (in the given order) into a fresh batch."""
    parts = {n: [] for n in ARRAY_NAMES}
    b_cur = o_cur = f_cur = r_cur = 0
    read_begin = [0]
    contigs = []
    for r in regions:
        lo, hi = int(batch.region_read_begin[r]), int(batch.region_read_begin[r + 1])
        L = int(batch.region_ref_len[r])
        fo = int(batch.region_ref_off[r])
        parts["ref"].append(batch.ref[fo:fo + L])
        parts["region_ref_len"].append(np.array([L], np.int64))
        for n in ("region_ref_start", "region_ref_end", "region_cand_start", "region_cand_end"):
            parts[n].append(getattr(batch, n)[r:r + 1])
        parts["region_ref_off"].append(np.array([f_cur], np.int64))
        f_cur += L
        if batch.contigs:
            contigs.append(batch.contigs[r])
        if hi > lo:
            bo = batch.read_base_off[lo:hi]
            ln = batch.read_len[lo:hi].astype(np.int64)
            b_lo, b_hi = int(bo.min()), int((bo + ((ln + 15) & ~15)).max())
            b_hi = min(b_hi, batch.n_bases)
            co = batch.read_cigar_off[lo:hi]
            c_lo, c_hi = int(co.min()), int((co + batch.read_n_ops[lo:hi]).max())
            parts["bases"].append(batch.bases[b_lo:b_hi])
            parts["quals"].append(batch.quals[b_lo:b_hi])
            parts["cigar"].append(batch.cigar[c_lo:c_hi])
            parts["read_base_off"].append(bo - b_lo + b_cur)
            parts["read_cigar_off"].append(co - c_lo + o_cur)
            pad = (-(b_hi - b_lo)) % 16
            if pad:
                parts["bases"].append(np.zeros(pad, np.uint8))
                parts["quals"].append(np.zeros(pad, np.uint8))
            b_cur += b_hi - b_lo + pad
            o_cur += c_hi - c_lo
            for n in ("read_pos", "read_len", "read_n_ops", "read_flags", "read_mapq"):
                parts[n].append(getattr(batch, n)[lo:hi])
        r_cur += hi - lo
        read_begin.append(r_cur)
    dt = dict(_READ_FIELDS + _REGION_FIELDS)
    dt.update(bases=np.uint8, quals=np.uint8, cigar=np.uint32, ref=np.uint8, region_read_begin=np.int64)
    out = {}
    for n in ARRAY_NAMES:
        if n == "region_read_begin":
            out[n] = np.asarray(read_begin, np.int64)
        else:
            out[n] = np.ascontiguousarray(np.concatenate(parts[n]).astype(dt[n], copy=False)) if parts[n] \
                else np.zeros(0, dt[n])
    return ReadBatch(contigs=contigs, **out)


@dataclass
class Region:
    """One call of AlignmentSummarizer.create_summary: the generator ctor args + the candidate interval."""
    contig: str
    ref_start: int
    ref_end: int
    reference: bytes
    cand_start: int
    cand_end: int
    reads: list     # objects with .pos .sequence .base_qualities .cigar_tuples(.cigar_op/.cigar_len) .flags.is_reverse .mapping_quality


def pack_regions(regions: Sequence[Region]) -> ReadBatch:
    """Pack reference-style read objects (``type_read`` duck type, pybind_api.h:208-221) into a ReadBatch."""
    rp, bo, rl, co, no, fl, mq = [], [], [], [], [], [], []
    bases, quals, cig, refs = [], [], [], []
    rs, re_, cs, ce, ro, rln, rb = [], [], [], [], [], [], [0]
    b_cur = o_cur = f_cur = 0
    for reg in regions:
        ref = reg.reference.encode() if isinstance(reg.reference, str) else bytes(reg.reference)
        L = reg.ref_end - reg.ref_start + 1
        if L <= 0:
            raise ValueError("region_end < region_start")
        if len(ref) < L:
            raise ValueError("reference_sequence shorter than region_end - region_start + 1")
        if len(reg.reads) > 32767:
            raise ValueError("more than 32767 reads in a region: window values would not fit int16")
        refs.append(np.frombuffer(ref, np.uint8))
        rs.append(reg.ref_start); re_.append(reg.ref_end); cs.append(reg.cand_start); ce.append(reg.cand_end)
        ro.append(f_cur); rln.append(len(ref)); f_cur += len(ref)
        for rd in reg.reads:
            seq = rd.sequence.encode("latin-1") if isinstance(rd.sequence, str) else bytes(rd.sequence)
            q = np.asarray(rd.base_qualities, dtype=np.int64)
            if q.shape[0] != len(seq):
                raise ValueError("base_qualities and sequence differ in length")
            if q.size and (q.min() < 0 or q.max() > 255):
                raise ValueError("base quality outside [0, 255]")
            ops = rd.cigar_tuples
            words = np.empty(len(ops), np.uint32)
            for k, op in enumerate(ops):
                o, l = (op.cigar_op, op.cigar_len) if hasattr(op, "cigar_op") else op
                if not (0 <= o <= 15) or not (0 <= l < (1 << 28)):
                    raise ValueError("bad CIGAR op")
                words[k] = (l << 4) | o
            n = len(seq)
            pad = (-n) % 16
            rp.append(int(rd.pos)); bo.append(b_cur); rl.append(n); co.append(o_cur); no.append(len(ops))
            fl.append(1 if rd.flags.is_reverse else 0)
            mq.append(min(255, max(0, int(rd.mapping_quality))))
            bases.append(np.frombuffer(seq + b"\0" * pad, np.uint8))
            quals.append(np.concatenate([q.astype(np.uint8), np.zeros(pad, np.uint8)]))
            cig.append(words)
            b_cur += n + pad; o_cur += len(ops)
        rb.append(len(rp))

    def cat(parts, dt):
        return np.ascontiguousarray(np.concatenate(parts).astype(dt, copy=False)) if parts else np.zeros(0, dt)

    return ReadBatch(
        read_pos=np.asarray(rp, np.int64), read_base_off=np.asarray(bo, np.int64), read_len=np.asarray(rl, np.int32),
        read_cigar_off=np.asarray(co, np.int64), read_n_ops=np.asarray(no, np.int32),
        read_flags=np.asarray(fl, np.uint8), read_mapq=np.asarray(mq, np.uint8),
        bases=cat(bases, np.uint8), quals=cat(quals, np.uint8), cigar=cat(cig, np.uint32),
        region_ref_start=np.asarray(rs, np.int64), region_ref_end=np.asarray(re_, np.int64),
        region_cand_start=np.asarray(cs, np.int64), region_cand_end=np.asarray(ce, np.int64),
        region_ref_off=np.asarray(ro, np.int64), region_ref_len=np.asarray(rln, np.int64),
        region_read_begin=np.asarray(rb, np.int64),
        ref=cat(refs, np.uint8), contigs=[r.contig for r in regions])
