"""TEST INFRASTRUCTURE -- CPU restatement of the site logic of the reference's stage-3 VCF writer.

Follows /root/reference/pepper_variant/modules/python/VcfWriter.py:
  candidate_list_to_variant  :48-139   (multi-allelic merge of the selected candidates of one position)
  write_vcf_records          :141-221  (QUAL, the q cut-offs, which of the five files a record goes to)
pysam is not installed here; oracle/ref_stage3.py runs the UNMODIFIED reference module with a recording stand-in for
pysam's VariantFile and tests/golden/stage3_vcf_seed*.json hold the records it produced (which file, which fields):
PARITY PINNED at the record level (tests/test_vcf_writer.py checks this port and the product writer against them). What
stays open is only the text pysam/htslib render for a float field.
Only tests may import this file.
"""
import math

import numpy as np


def candidate_list_to_variant(candidates, allowed_multiallelics):
    """VcfWriter.py:48-139. ``candidates``: the 12-tuples find_candidates emits for one (contig, position)."""
    candidates = sorted(candidates, key=lambda x: (x[5], x[8]), reverse=True)            # :49
    if len(candidates) > allowed_multiallelics:                                          # :50-51
        candidates = candidates[:allowed_multiallelics]
    max_ref_length, max_ref_allele = 0, ''
    for cand in candidates:                                                              # :56-60
        if len(cand[3]) > max_ref_length:
            max_ref_length, max_ref_allele = len(cand[3]), cand[3]
    normalized = []
    for cand in candidates:                                                              # :62-74
        contig, ref_start, ref_end, ref_allele, alt_allele, genotype, depth, support, gp, predictions, non_alt, in_repeat = cand
        suffix_needed = max_ref_length - len(ref_allele) if len(ref_allele) < max_ref_length else 0
        if suffix_needed > 0:
            suffix_seq = max_ref_allele[-suffix_needed:]
            ref_allele = ref_allele + suffix_seq
            alt_allele = [alt + suffix_seq for alt in alt_allele]
        normalized.append((contig, ref_start, ref_end, ref_allele, alt_allele, genotype, depth, support, gp, predictions, non_alt, in_repeat))
    gt_qual = -1.0
    hp1, hp2 = [], []
    initialized = False
    site = dict(contig='', start=0, end=0, ref='', depth=0, alts=[], supports=[], non_alt=[], in_repeat=False)
    for i, cand in enumerate(normalized):                                                # :95-128
        contig, ref_start, ref_end, ref_allele, alt_allele, genotype, depth, support, gp, predictions, non_alt, in_repeat = cand
        site['in_repeat'] = in_repeat or site['in_repeat']
        predicted = int(np.argmax(predictions))
        if predicted != 0:
            gt_qual = predictions[predicted] if gt_qual < 0 else min(gt_qual, predictions[predicted])
        elif gt_qual < 0:
            gt_qual = max(predictions[1], predictions[2])
        if not initialized:
            site.update(contig=contig, start=ref_start, end=ref_start + len(ref_allele), ref=ref_allele, depth=depth)
            initialized = True
        site['depth'] = min(site['depth'], depth)
        site['alts'].append(alt_allele[0])
        site['supports'].append(support[0])
        site['non_alt'].extend(non_alt)
        if predicted == 1:
            hp1.append(i + 1)
        elif predicted == 2:
            hp1.append(i + 1)
            hp2.append(i + 1)
    if 0 < len(hp1) + len(hp2) <= 2:                                                     # :130-135
        gt = hp1 + hp2
        if len(gt) == 1:
            gt = [0, gt[0]]
    else:
        gt = [0, 0]
    return (site['contig'], site['start'], site['end'], site['ref'], site['alts'], gt, site['depth'], site['supports'],
            gt_qual, site['non_alt'], site['in_repeat'])


def records(variants, options):
    """VcfWriter.py:141-221 without the file objects: yields one dict per written record with the files it goes to."""
    out = []
    last_position = -1
    for contig, position in sorted(variants):                                            # :145
        contig, ref_start, ref_end, ref_seq, alleles, genotype, depth, support, gp, non_alt, in_repeat = \
            candidate_list_to_variant(variants[(contig, position)], options.allowed_multiallelics)
        if len(alleles) <= 0 or ref_start == last_position:                              # :150-153
            continue
        max_alt_len = max(len(ref_seq), max(len(x) for x in alleles))
        last_position = ref_start
        qual = max(1, int(-10 * math.log10(max(0.000000001, 1.0 - gp))))                 # :157
        is_snp = max_alt_len == 1
        if is_snp:                                                                       # :161-172
            failed = qual <= (options.snp_q_cutoff_in_lc if in_repeat else options.snp_q_cutoff)
        else:
            failed = qual <= (options.indel_q_cutoff_in_lc if in_repeat else options.indel_q_cutoff)
        regenotype = genotype == [0, 0] or failed                                        # :176-178
        vafs = [round(ad / max(1, depth), 3) for ad in support]                          # :180
        files = ['full'] + ((['variant_calling_snp' if is_snp else 'variant_calling_indel', 'variant_calling'])
                            if regenotype else ['pepper'])                               # :204-219
        out.append(dict(contig=str(contig), start=ref_start, stop=ref_end, qual=qual,
                        filter='refCall' if genotype == [0, 0] else 'PASS', alleles=(ref_seq,) + tuple(alleles),
                        GT=genotype, AP=list(non_alt), GQ=qual, DP=depth, AD=list(support), VAF=vafs,
                        REP='1' if in_repeat else '0', files=files))
    return out
