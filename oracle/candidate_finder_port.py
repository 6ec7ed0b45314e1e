"""TEST INFRASTRUCTURE ONLY: CPU restatement of the per-candidate block of small_chunk_stitch and of find_candidates'
de-duplication (/root/reference/pepper_variant/modules/python/CandidateFinder.py:279-297, 391-529, 536-600). The
reference module itself needs h5py and the compiled PEPPER_VARIANT, so this follows its source text line by line over
plain Python values. PARITY PINNED: oracle/ref_stage3.py runs the UNMODIFIED reference module (stand-ins for those two
imports only) and tests/golden/stage3_filter_seed*.json hold its output; tests/test_candidate_filter.py checks this port
against them everywhere and against the live reference where /root/reference exists."""
from collections import defaultdict

import numpy as np


def repeat_annotation(sequence, kmer_size):                                   # :279-297
    out = [1] * len(sequence)
    for i in range(len(sequence) - (kmer_size - 1)):
        count, end = 0, i + kmer_size - 1
        for j in range(i, len(sequence), kmer_size):
            if sequence[i:i + kmer_size] != sequence[j:j + kmer_size]:
                break
            count += 1
            end = j + kmer_size
        for k in range(i, min(len(sequence), end)):
            out[k] = max(out[k], count)
    return out


def stitch(cands, fetch, options):
    """cands: iterable of (contig, position, depth, [allele str], [frequency], probs[3]); fetch(contig, a, b) -> str
    like FASTA_handler.get_reference_sequence. -> (margin list, deepvariant list)"""
    margin, deepv = [], []
    for contig, position, depth, alleles, freqs, pb in cands:
        ref_base = fetch(contig, position, position + 1).upper()
        up = fetch(contig, position, position + 10).upper()
        down = fetch(contig, max(0, position - 10), position).upper()
        hp = repeat_annotation((down + up).upper(), 1)
        pi = len(down)
        in_repeat = max(hp[max(0, pi - 5):min(len(hp), pi + 4)]) >= 5          # :403-414
        if ref_base not in ["A", "C", "G", "T"]:
            continue
        g = int(np.argmax(pb))
        genotype = [0, 0] if g == 0 else ([0, 1] if g == 1 else [1, 1])
        value = pb[g]
        alts, support = [], []
        for a, f in zip(alleles, freqs):                                       # :431-448
            if any(b not in "ACGT" for b in a[1:]):
                continue
            if a[0] == "1" and g != 0:
                alts.append(a[1:]); support.append(f)
        if alts:
            margin.append((contig, position, position + 1, ref_base, alts, genotype, depth, support, value, pb))
        alts, support, ref_allele, non_alts = [], [], ref_base, []
        for a, f in zip(alleles, freqs):                                       # :462-517
            if any(b not in "ACGT" for b in a[1:]):
                continue
            vaf = float(f) / float(depth)
            non_alt = max(pb[1], pb[2])
            non_alts.append(non_alt)
            t = a[0]
            if t == "1":
                pv, pv_lc, rep = options.snp_p_value, options.snp_p_value_in_lc, options.report_snp_above_freq
            elif t == "2":
                pv, pv_lc, rep = options.insert_p_value, options.insert_p_value_in_lc, options.report_indel_above_freq
            elif t == "3":
                pv, pv_lc, rep = options.delete_p_value, options.delete_p_value_in_lc, options.report_indel_above_freq
            else:
                continue
            by_p = (not in_repeat and non_alt >= pv) or (in_repeat and non_alt >= pv_lc)
            if by_p and t == "3":
                alts.append(ref_allele); ref_allele = a[1:]; support.append(f)
            elif by_p or 0 < rep <= vaf:
                alts.append(a[1:]); support.append(f)
        if alts:
            deepv.append((contig, position, position + len(ref_allele), ref_allele, alts, genotype, depth, support, value, pb,
                          non_alts, in_repeat))
    return margin, deepv


def find_candidates(margin, deepv):                                            # :550-600
    margin = sorted(margin, key=lambda x: (x[0], x[1]))
    deepv = sorted(deepv, key=lambda x: (x[0], x[1]))
    pd, vd, ps, vs = defaultdict(list), defaultdict(list), defaultdict(list), defaultdict(list)
    for c in margin:
        k, ra = (c[0], c[1]), (c[3], c[4][0])
        if ra in ps[k]:
            continue
        ps[k].append(ra); pd[k].append(c)
    contigs = []
    for c in deepv:
        if c[0] not in contigs:
            contigs.append(c[0])
        k, ra = (c[0], c[1]), (c[3], c[4][0])
        if ra in vs[k]:
            continue
        vs[k].append(ra); vd[k].append(c)
    return contigs, pd, vd
