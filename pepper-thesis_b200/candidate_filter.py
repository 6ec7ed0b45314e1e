"""Stage 3 of ``call_variant`` without HDF5 and without the process pool: the reference's ``find_candidates`` /
``small_chunk_stitch`` (/root/reference/pepper_variant/modules/python/CandidateFinder.py:357-600) over the SoA
:class:`Predictions` the GPU pipeline produced. The per-candidate decision runs in ``candidate_filter_kernel``
(csrc/candidate_filter.cu); this module only turns the selected records into the reference's tuples and applies its
per-position (ref, alt) de-duplication."""
from __future__ import annotations

import ctypes as C
from collections import defaultdict
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import capi
from .read_batch import ReadBatch

PHASING, VARIANT, IN_REPEAT, GT_SHIFT, DEL_BY_FREQ, BAD_REF = 1, 2, 4, 3, 32, 64
GENOTYPES = ([0, 0], [0, 1], [1, 1])                       # CandidateFinder.py:421-426


@dataclass
class FilterOptions:
    """The option fields small_chunk_stitch reads (defaults: the ONT R9 Guppy5 SUP preset, SetParameters.py:38-65)."""
    snp_p_value: float = 0.1
    snp_p_value_in_lc: float = 0.1
    insert_p_value: float = 0.1
    insert_p_value_in_lc: float = 0.15
    delete_p_value: float = 0.1
    delete_p_value_in_lc: float = 0.1
    report_snp_above_freq: float = 0.0
    report_indel_above_freq: float = 0.0

    def as_struct(self):
        s = capi.PvFilterOptionsStruct()
        for n, _ in s._fields_:
            setattr(s, n, float(getattr(self, n)))
        return s


def filter_flags(pred, batch: ReadBatch, options: FilterOptions, contig_len: Optional[Sequence[int]] = None,
                 region_offset: int = 0) -> np.ndarray:
    """uint8 flag per candidate (PV_FLT_* of include/pepper_b200.h). ``pred.region`` indexes ``batch`` regions after
    subtracting ``region_offset``. ``contig_len`` (per batch region; default: ``batch.region_contig_len``, which the ingest
    fills) clips the context at the contig end. A candidate that reaches the allele-frequency division with depth 0 raises
    ZeroDivisionError like the reference's ``float(allele_frequency) / float(candidate.depth)`` (CandidateFinder.py:477)."""
    n = len(pred)
    flags = np.zeros(n, np.uint8)
    if n == 0:
        return flags
    lib = capi.load()
    region = np.ascontiguousarray(pred.region - region_offset, np.int32)
    if contig_len is None:
        contig_len = batch.region_contig_len
    cl = None if contig_len is None else np.ascontiguousarray(contig_len, np.int64)
    arrs = [np.ascontiguousarray(pred.position, np.int64), region, np.ascontiguousarray(pred.depth, np.int32),
            np.ascontiguousarray(pred.frequency, np.int32), np.ascontiguousarray(pred.allele, np.uint8),
            np.ascontiguousarray(pred.allele_len, np.uint8), np.ascontiguousarray(pred.probs, np.float32)]
    opt = options.as_struct()
    capi.check(lib.pv_candidate_filter_host(
        n, *[a.ctypes.data for a in arrs], batch.n_regions, batch.region_ref_start.ctypes.data,
        batch.region_ref_off.ctypes.data, batch.region_ref_len.ctypes.data, cl.ctypes.data if cl is not None else None,
        batch.ref.ctypes.data, int(batch.ref.shape[0]), C.byref(opt), flags.ctypes.data))
    zero = np.nonzero((np.asarray(pred.depth) <= 0) & ((flags & BAD_REF) == 0))[0]
    if zero.size:
        al = pred.alleles()
        if any(all(ch in b"ACGT" for ch in al[i][1:]) for i in zero):
            raise ZeroDivisionError("float division by zero (candidate with depth 0 at position %d)" % int(pred.position[zero[0]]))
    return flags


def find_candidates(pred, batch: ReadBatch, options: FilterOptions, contig_len=None, region_offset: int = 0):
    """-> (contigs, phasing_dict, variant_calling_dict) exactly as the reference's find_candidates returns them
    (CandidateFinder.py:536-600): dicts keyed by (contig, position) holding the selected-candidate tuples, sorted by
    (contig, position), one entry per distinct (ref, first alt)."""
    flags = filter_flags(pred, batch, options, contig_len, region_offset)
    alleles = pred.alleles()
    phasing, variant = [], []
    for i in np.nonzero(flags & (PHASING | VARIANT))[0]:
        f = int(flags[i])
        r = int(pred.region[i]) - region_offset
        contig = batch.contigs[r] if batch.contigs else ""
        pos = int(pred.position[i])
        ref_base = chr(batch.ref[int(batch.region_ref_off[r]) + pos - int(batch.region_ref_start[r])]).upper()
        g = (f >> GT_SHIFT) & 3
        probs = pred.probs[i]
        depth, freq = int(pred.depth[i]), int(pred.frequency[i])
        allele = alleles[i].decode()
        typ, bases = allele[0], allele[1:]
        if f & PHASING:                                                        # :444-451
            phasing.append((contig, pos, pos + 1, ref_base, [bases], list(GENOTYPES[g]), depth, [freq], probs[g], probs))
        if f & VARIANT:                                                        # :480-519
            non_alt = max(probs[1], probs[2])
            if typ == "3" and not (f & DEL_BY_FREQ):
                ref_allele, alts = bases, [ref_base]
            else:
                ref_allele, alts = ref_base, [bases]
            variant.append((contig, pos, pos + len(ref_allele), ref_allele, alts, list(GENOTYPES[g]), depth, [freq], probs[g],
                            probs, [non_alt], bool(f & IN_REPEAT)))
    phasing.sort(key=lambda x: (x[0], x[1]))
    variant.sort(key=lambda x: (x[0], x[1]))
    p_dict, v_dict = defaultdict(list), defaultdict(list)
    p_seen, v_seen = defaultdict(list), defaultdict(list)
    for c in phasing:                                                          # :561-567
        key, ra = (c[0], c[1]), (c[3], c[4][0])
        if ra in p_seen[key]:
            continue
        p_seen[key].append(ra); p_dict[key].append(c)
    contigs = []
    for c in variant:                                                          # :569-578
        if c[0] not in contigs:
            contigs.append(c[0])
        key, ra = (c[0], c[1]), (c[3], c[4][0])
        if ra in v_seen[key]:
            continue
        v_seen[key].append(ra); v_dict[key].append(c)
    return contigs, p_dict, v_dict
