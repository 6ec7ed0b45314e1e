"""The variant module's LEGACY per-position summary on the GPU: ``PEPPER_VARIANT.SummaryGenerator`` / ``ImageSummary``
(/root/reference/pepper_variant/modules/cpp/summary_generator.h:20-76, summary_generator.cpp:91-169 iterate_over_read,
:330-364 generate_image, :455-487 generate_summary, :491-536 chunk_image; bound at pybind_api.h:24-43). r0.8's
``call_variant`` does not call it (AlignmentSummarizer.py:220 uses RegionalSummaryGenerator), but it is part of the module's
surface.

It is the polisher's generator (pepper/modules/src/pileup_summary/summary_generator.cpp, built as csrc/polish_summary.cu
and pinned bit-exact against that file) with three differences, all on the host side of the same kernels:
  * no ``mapping_quality > 0`` filter (the polisher's :376): every read counts -> the batch is handed over with mapq >= 1;
  * ``ref_image``: the reference base of every row as 0..4 (insert rows: '*' -> 0), :34-41, :472-480;
  * ``chunk_image`` in C++ instead of Python, with ``refs`` and all-zero ``labels`` per chunk.
The per-base ``mapping_quality / 60`` and ``base_quality / 100`` factors the legacy walk computes are never used: every
observation adds exactly 1.0 (:119-120, :140, :156-157), so the sums are counts and the order of the reads cannot change them.
Oracle: the unmodified file compiled as oracle/_ref/pv_ref_legacy (tests/test_legacy_summary_gpu.py)."""
from __future__ import annotations

import numpy as np

from .polish import PolishSummary, chunk_plan
from .read_batch import Region, pack_regions

_REF_CODE = np.zeros(256, np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _REF_CODE[_c] = _REF_CODE[ord(chr(_c).lower())] = _i + 1          # get_reference_feature_index, :34-41 (toupper)


class ImageSummary:
    """summary_generator.h:20-26 (value semantics: nested Python lists, like the pybind conversion of the vectors)."""

    def __init__(self):
        self.images, self.positions, self.refs, self.labels, self.chunk_ids = [], [], [], [], []


class SummaryGenerator:
    """Drop-in for ``PEPPER_VARIANT.SummaryGenerator`` (inference side: generate_summary + chunk_image)."""

    def __init__(self, reference_sequence, chromosome_name, ref_start, ref_end):
        self.reference_sequence, self.chromosome_name = reference_sequence, chromosome_name
        self.ref_start, self.ref_end = int(ref_start), int(ref_end)
        self.image, self.genomic_pos, self.ref_image, self.labels, self.bad_label_positions = [], [], [], [], []
        self.longest_insert_count = {}
        self.summary = None                 # PolishSummary: the image / positions as tensors in HBM
        self._img = self._gp = self._ref = None

    def generate_summary(self, reads, start_pos, end_pos):
        if int(start_pos) != self.ref_start or int(end_pos) != self.ref_end:
            raise ValueError("generate_summary(start_pos, end_pos) must equal the constructor's region (as the reference calls it)")
        ref = self.reference_sequence
        ref = ref + "N" * max(0, self.ref_end - self.ref_start + 1 - len(ref))
        b = pack_regions([Region(self.chromosome_name, self.ref_start, self.ref_end, ref, self.ref_start, self.ref_end, list(reads))])
        b.read_mapq = np.maximum(b.read_mapq, 1).astype(b.read_mapq.dtype)      # no mapping-quality filter in this generator
        s = PolishSummary(b)
        self.summary = s
        self._img = s.image.cpu().numpy()
        self._gp = s.genomic_pos.cpu().numpy()
        rb = np.frombuffer(ref.encode("latin-1"), np.uint8)
        self._ref = np.where(self._gp[:, 1] == 0, _REF_CODE[rb[np.clip(self._gp[:, 0] - self.ref_start, 0, len(rb) - 1)]], 0).astype(np.uint8)
        self.image = self._img.tolist()
        self.genomic_pos = [tuple(p) for p in self._gp.tolist()]
        self.ref_image = self._ref.tolist()
        # :469-475 touch longest_insert_count[i] for every position of the region: the map holds them all
        longest = np.zeros(self.ref_end - self.ref_start + 1, np.int64)
        np.maximum.at(longest, self._gp[:, 0] - self.ref_start, self._gp[:, 1])
        self.longest_insert_count = {self.ref_start + i: int(v) for i, v in enumerate(longest)}

    def chunk_image(self, chunk_size, chunk_overlap, image_height):
        """:491-536: chunks of ``chunk_size`` rows that overlap by ``chunk_overlap``; the last one is padded with zero rows,
        positions (-1, -1), reference 0; labels are all zero."""
        if self._img is None:
            raise RuntimeError("chunk_image before generate_summary")
        out = ImageSummary()
        for cid, (start, n) in enumerate(chunk_plan(len(self._gp), int(chunk_size), int(chunk_overlap))):
            img = np.zeros((chunk_size, image_height), np.uint8)
            pos = np.full((chunk_size, 2), -1, np.int64)
            rf = np.zeros(chunk_size, np.uint8)
            img[:n, :self._img.shape[1]] = self._img[start:start + n]
            pos[:n] = self._gp[start:start + n]
            rf[:n] = self._ref[start:start + n]
            out.images.append(img.tolist())
            out.positions.append([tuple(p) for p in pos.tolist()])
            out.refs.append(rf.tolist())
            out.labels.append([0] * int(chunk_size))
            out.chunk_ids.append(cid)
        return out

    def chunks_device(self, chunk_size, chunk_overlap):
        """The same chunks as tensors in HBM (images uint8 [n][size][10], positions int64 [n][size][2], chunk ids)."""
        images, positions, ids, _ = self.summary.chunks(int(chunk_size), int(chunk_overlap))
        return images, positions, ids

    def generate_train_summary(self, *a, **k):
        raise NotImplementedError("training labels are outside the B200 hot path")

    def chunk_image_train(self, *a, **k):
        raise NotImplementedError("training labels are outside the B200 hot path")
