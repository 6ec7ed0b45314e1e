"""BAM + FASTA -> candidate variants -> VCF with every per-byte step on the device: the library form of
``pepper_variant call_variant`` (/root/reference/pepper_variant/modules/python/CallVariant.py:12-110), i.e. the three stages

    1. ImageGenerationUtils.generate_images   ImageGenerationUI.py:280-330 (intervals of ``region_size`` per contig, interval i
                                              to worker i % threads) -> AlignmentSummarizer.create_summary per interval
    2. run_inference                          RunInference.py / models/predict.py (TransducerGRU over the windows)
    3. process_candidates                     FindCandidates.py:150-190 (find_candidates + VCFWriter)

without the HDF5 files between them: :func:`ingest_gpu.stream_regions_gpu` decodes the BAM on the GPU group by group (the
host share of the next group runs underneath), :class:`pipeline.HotPath` turns every device-born batch into predictions,
``candidate_filter.find_candidates`` applies stage 3's rules on the device and ``vcf_writer.VCFWriter`` writes the five
files. Intervals are dealt to the ranks of a ``torch.distributed`` job exactly as the reference deals them to its worker
processes (interval i -> rank i % world); rank 0 gathers the selected candidates and writes. There is no CPU fallback.
The argparse front end of the reference is out of scope; this is the call a front end would make.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import candidate_filter, ingest
from .pipeline import HotPath, Predictions
from .summarizer import REGION_SAFE_BASES


@dataclass
class CallOptions:
    """The options stages 1-3 read (defaults: the ONT R9 Guppy5 SUP preset of SetParameters.py, region_size of
    pepper_variant.py's call_variant parser)."""
    region_size: int = 100000
    include_supplementary: bool = False
    min_mapq: int = 5
    downsample_rate: float = 1.0
    group_mbp: float = 32.0                     # contig span decoded per device call (a BGZF block takes ~10 ms: large calls amortise that)
    threads: int = 0
    filter: candidate_filter.FilterOptions = field(default_factory=candidate_filter.FilterOptions)
    predictions_hdf: Optional[str] = None       # leave the reference's stage-2 file (DataStorePredict.py:49-66) beside the run:
                                                # one batch per decoded group; rank r of a multi-rank job writes <path>.<r>


def contig_intervals(fasta: ingest.FASTAHandler, regions: Optional[Sequence] = None, region_size: int = 100000) -> List[Tuple[str, int, int]]:
    """ImageGenerationUI.py:289-316: ``regions`` = [(contig, None) | (contig, (start, end))]; every contig span is cut into
    intervals [pos, min(end, pos + region_size)]. None = every sequence of the FASTA."""
    if regions is None:
        regions = [(name, None) for name in fasta.get_chromosome_names()]
    out = []
    for name, span in regions:
        clen = int(fasta.get_chromosome_sequence_length(name))
        if clen <= 0:
            raise RuntimeError("CHROMOSOME NAME NOT PRESENT IN REFERENCE FASTA FILE: %s" % name)
        lo, hi = (0, clen - 1) if span is None else (max(0, int(span[0])), min(int(span[1]), clen - 1))
        for pos in range(lo, hi, region_size):
            out.append((name, max(lo, pos), min(hi, pos + region_size)))
    return out


class _RegionView:
    """What ``find_candidates`` reads of a batch: the region fields and the reference bytes, on the host."""

    def __init__(self, db):
        r = db.regions
        self.n_regions = int(db.host.n_regions)
        self.region_ref_start = np.ascontiguousarray(r["region_ref_start"], np.int64)
        self.region_ref_off = np.ascontiguousarray(r["region_ref_off"], np.int64)
        self.region_ref_len = np.ascontiguousarray(r["region_ref_len"], np.int64)
        self.region_contig_len = db.host.region_contig_len
        self.contigs = list(db.host.contigs)
        self.ref = db.t["ref"].cpu().numpy()


def call_candidates(bam_path: str, fasta_path: str, hot: HotPath, regions: Optional[Sequence] = None,
                    options: Optional[CallOptions] = None, rank: int = 0, world: int = 1, group=None):
    """Stages 1-3 up to the selected candidates: returns ``(contigs, phasing_dict, variant_dict, stats)`` like the
    reference's ``find_candidates`` (on rank 0 of a multi-rank job; the other ranks return None after the gather)."""
    from . import ingest_gpu
    opt = options or CallOptions()
    bam, fasta = ingest.BAMHandler(bam_path), ingest.FASTAHandler(fasta_path)
    intervals = contig_intervals(fasta, regions, opt.region_size)
    mine = [iv for i, iv in enumerate(intervals) if i % world == rank]          # ImageGenerationUI.py:211
    per_group = max(1, int(opt.group_mbp * 1e6 / max(1, opt.region_size)))
    by_contig: Dict[str, List[Tuple[int, int]]] = {}
    for name, s, e in mine:
        by_contig.setdefault(name, []).append((s, e))
    phasing, variant = {}, {}
    contigs: List[str] = []
    stats = dict(intervals=len(mine), candidates=0, reads=0, compressed_bytes=0)
    store, batch_no = None, 0
    if opt.predictions_hdf:
        from . import datastore
        store = datastore.DataStorePredict(opt.predictions_hdf if world == 1 else "%s.%d" % (opt.predictions_hdf, rank), "w")
    for name, ivs in by_contig.items():
        ivs.sort()
        groups = [([s for s, _ in ivs[i:i + per_group]], [e for _, e in ivs[i:i + per_group]]) for i in range(0, len(ivs), per_group)]
        for got in ingest_gpu.stream_regions_gpu(bam, fasta, name, groups, include_supplementary=opt.include_supplementary,
                                                 min_mapq=opt.min_mapq, downsample_rate=opt.downsample_rate, threads=opt.threads,
                                                 safe_bases=REGION_SAFE_BASES, device=hot.device):
            stats["reads"] += int(got.stats["reads"]); stats["compressed_bytes"] += int(got.stats["compressed_bytes"])
            if got.batch.host.n_reads == 0:
                continue
            pred = hot.run_device(got.batch, to_host=True)
            stats["candidates"] += len(pred)
            if store is not None and len(pred):
                datastore.write_prediction_batch(store, batch_no, pred, list(got.batch.host.contigs))
                batch_no += 1
            c, p, v = candidate_filter.find_candidates(pred, _RegionView(got.batch), opt.filter)
            for k, lst in p.items():
                phasing.setdefault(k, []).extend(lst)
            for k, lst in v.items():
                variant.setdefault(k, []).extend(lst)
            for x in c:
                if x not in contigs:
                    contigs.append(x)
    if store is not None:
        store.close()
    return gather_candidates((contigs, phasing, variant, stats), rank, world, group)


def merge_candidates(parts):
    """The ranks' (contigs, phasing_dict, variant_dict, stats) as one result: sites in (contig, position) order, one entry
    per distinct (ref, first alt) per site -- find_candidates' own rule (CandidateFinder.py:561-578), re-applied across
    ranks and groups (an interval boundary can put the same site into two of them; rank order = interval order there)."""
    contigs, phasing, variant = [], {}, {}
    total: Dict[str, int] = {}
    for c, p, v, st in parts:
        for x in c:
            if x not in contigs:
                contigs.append(x)
        for k, lst in p.items():
            phasing.setdefault(k, []).extend(lst)
        for k, lst in v.items():
            variant.setdefault(k, []).extend(lst)
        for k, n in st.items():
            total[k] = total.get(k, 0) + n

    def dedup(d):
        out = {}
        for key in sorted(d):
            seen, keep = [], []
            for c in d[key]:
                ra = (c[3], c[4][0])
                if ra not in seen:
                    seen.append(ra); keep.append(c)
            out[key] = keep
        return out
    return sorted(contigs), dedup(phasing), dedup(variant), total


def gather_candidates(local, rank: int = 0, world: int = 1, group=None):
    """Host-side gather of the ranks' selected candidates to rank 0 (no collective on the data path: the regions are
    independent); the other ranks get None."""
    if world <= 1:
        return merge_candidates([local])
    import torch.distributed as dist
    parts = [None] * world if rank == 0 else None
    dist.gather_object(local, parts, dst=0, group=group)
    return merge_candidates(parts) if rank == 0 else None


def call_variant(bam_path: str, fasta_path: str, hot: HotPath, output_dir: str, sample_name: str = "SAMPLE",
                 regions: Optional[Sequence] = None, options: Optional[CallOptions] = None, vcf_options=None, rank: int = 0,
                 world: int = 1, group=None):
    """BAM + FASTA -> the reference's five VCF files in ``output_dir`` (rank 0 writes). Returns (the five record counts, stats, paths) on rank 0."""
    from .vcf_writer import VCFWriter, VcfOptions
    res = call_candidates(bam_path, fasta_path, hot, regions, options, rank, world, group)
    if res is None:
        return None
    contigs, _phasing, variant, stats = res
    import os
    os.makedirs(output_dir, exist_ok=True)
    if not output_dir.endswith("/"):
        output_dir += "/"                                        # handle_output_directory: the writer concatenates strings
    w = VCFWriter(contigs, fasta_path, sample_name, output_dir, "PEPPER_VARIANT_FULL", "PEPPER_VARIANT_OUTPUT_PEPPER",
                  "PEPPER_VARIANT_OUTPUT_VARIANT_CALLING")          # FindCandidates.py:167-180
    counts = w.write_vcf_records(variant, vcf_options or VcfOptions())
    w.close()
    return counts, stats, w.paths
