"""Polisher pileup summary on the GPU: the reference ``pepper`` module's ``SummaryGenerator``
(/root/reference/pepper/modules/src/pileup_summary/summary_generator.cpp:47-121, 274-306, 370-392, bound as
``PEPPER.SummaryGenerator`` in pepper/modules/headers/pybind_api.h) and ``AlignmentSummarizer.chunk_images``
(/root/reference/pepper/modules/python/AlignmentSummarizer.py:19-56) -- the producer of model M-B's input
(SURVEY.md 8f row 4). Kernels: csrc/polish_summary.cu. PyTorch only allocates."""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import numpy as np
import torch

from . import capi, device as dev
from .read_batch import ReadBatch, Region, pack_regions

IMAGE_HEIGHT = 10        # pepper/modules/python/Options.py:2
SEQ_LENGTH = 1000        # :4
SEQ_OVERLAP = 50         # :5


def chunk_plan(n_rows: int, chunk_size: int = SEQ_LENGTH, chunk_overlap: int = SEQ_OVERLAP) -> List[Tuple[int, int]]:
    """(first row, real rows) of every chunk, the loop of chunk_images (AlignmentSummarizer.py:19-56)."""
    plan = []
    start, end = 0, min(n_rows, chunk_size)
    while True:
        plan.append((start, end - start))
        if end == n_rows:
            break
        start = end - chunk_overlap
        end = min(n_rows, start + chunk_size)
    return plan


class PolishSummary:
    """Summary of every region of a batch: ``image`` uint8 [rows][10], ``genomic_pos`` int64 [rows][2],
    ``region_rows`` int64 [n_regions + 1] (row range of each region) -- all torch tensors in HBM except region_rows."""

    def __init__(self, batch: ReadBatch | dev.DeviceBatch | dev.DeviceReadBatch, device="cuda"):
        # a DeviceReadBatch is a batch born on the device (ingest_gpu: BAM decoded by kernels, safe_bases=0 for the polisher's
        # get_reads(chr, start, end) of pepper/modules/python/AlignmentSummarizer.py:300-305)
        db = batch if isinstance(batch, (dev.DeviceBatch, dev.DeviceReadBatch)) else dev.DeviceBatch(batch, device)
        self.db = db
        lib = capi.load()
        h = db.host
        d = db.device
        stream = C.c_void_p(torch.cuda.current_stream(d).cuda_stream)
        ws_bytes = int(lib.pv_polish_workspace_bytes(h.n_reads, h.n_ops, h.n_regions, db.total_positions))
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=d)
        n_rows = C.c_int64(0)
        self.region_rows = np.zeros(h.n_regions + 1, np.int64)
        capi.check(lib.pv_polish_count(C.byref(db.struct), db.region_len.ctypes.data, db.total_positions, ws.data_ptr(),
                                       ws.numel(), C.byref(n_rows), self.region_rows.ctypes.data, stream))
        n = int(n_rows.value)
        self.n_rows = n
        self.image = torch.empty((n, IMAGE_HEIGHT), dtype=torch.uint8, device=d)
        self.genomic_pos = torch.empty((n, 2), dtype=torch.int64, device=d)
        self.row_region = torch.empty(n, dtype=torch.int32, device=d)
        scratch = torch.empty((max(n, 1), IMAGE_HEIGHT), dtype=torch.int32, device=d)
        capi.check(lib.pv_polish_emit(C.byref(db.struct), db.region_len.ctypes.data, db.total_positions, ws.data_ptr(),
                                      ws.numel(), n, scratch.data_ptr(), self.image.data_ptr(), self.genomic_pos.data_ptr(),
                                      self.row_region.data_ptr(), stream))

    def region(self, r: int):
        lo, hi = int(self.region_rows[r]), int(self.region_rows[r + 1])
        return self.image[lo:hi], self.genomic_pos[lo:hi]

    def chunks(self, chunk_size: int = SEQ_LENGTH, chunk_overlap: int = SEQ_OVERLAP):
        """-> (images uint8 [n_chunks][chunk_size][10], positions int64 [n_chunks][chunk_size][2], chunk_ids, chunk_region):
        chunk_images applied to every region, regions concatenated."""
        # chunk_plan for every region at once: a region of n rows has 1 + ceil(max(0, n - size) / (size - overlap)) chunks,
        # chunk k starts at row k * (size - overlap) and holds min(size, n - start) real rows
        lo = self.region_rows[:-1]
        n = np.diff(self.region_rows)
        step = chunk_size - chunk_overlap
        per = np.where(n > 0, 1 + (np.maximum(0, n - chunk_size) + step - 1) // step, 0)
        regs = np.repeat(np.arange(len(n), dtype=np.int32), per)
        ids = np.arange(int(per.sum()), dtype=np.int64) - np.repeat(np.cumsum(per) - per, per)
        start = ids * step
        rows = lo[regs] + start
        valid = np.minimum(chunk_size, n[regs] - start)
        d = self.image.device
        n_chunks = len(rows)
        images = torch.empty((n_chunks, chunk_size, IMAGE_HEIGHT), dtype=torch.uint8, device=d)
        positions = torch.empty((n_chunks, chunk_size, 2), dtype=torch.int64, device=d)
        if n_chunks:
            tr = torch.from_numpy(np.ascontiguousarray(rows, np.int64)).to(d)
            tv = torch.from_numpy(np.ascontiguousarray(valid, np.int64)).to(d)
            capi.check(capi.load().pv_polish_chunks(self.image.data_ptr(), self.genomic_pos.data_ptr(), tr.data_ptr(), tv.data_ptr(),
                                                    n_chunks, chunk_size, images.data_ptr(), positions.data_ptr(),
                                                    C.c_void_p(torch.cuda.current_stream(d).cuda_stream)))
        return images, positions, ids.astype(np.int64), regs.astype(np.int32)


class SummaryGenerator:
    """Drop-in for ``PEPPER.SummaryGenerator`` (value semantics: ``image`` / ``genomic_pos`` as Python lists)."""

    def __init__(self, reference_sequence, chromosome_name, ref_start, ref_end):
        self.reference_sequence, self.chromosome_name = reference_sequence, chromosome_name
        self.ref_start, self.ref_end = int(ref_start), int(ref_end)
        self.image, self.genomic_pos, self.labels, self.bad_label_positions = [], [], [], []

    def generate_summary(self, reads, start_pos, end_pos):
        if int(start_pos) != self.ref_start or int(end_pos) != self.ref_end:
            raise ValueError("generate_summary(start_pos, end_pos) must equal the constructor's region (as the reference calls it)")
        ref = self.reference_sequence
        ref = ref + "N" * max(0, self.ref_end - self.ref_start + 1 - len(ref))
        b = pack_regions([Region(self.chromosome_name, self.ref_start, self.ref_end, ref, self.ref_start, self.ref_end, list(reads))])
        s = PolishSummary(b)
        self.image = s.image.cpu().numpy().tolist()
        self.genomic_pos = [tuple(p) for p in s.genomic_pos.cpu().numpy().tolist()]

    def generate_train_summary(self, *a, **k):
        raise NotImplementedError("training labels are outside the B200 hot path")
