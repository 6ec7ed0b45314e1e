"""Per-region driver: drop-in for the inference branch of the reference's ``AlignmentSummarizer.create_summary``
(/root/reference/pepper_variant/modules/python/AlignmentSummarizer.py:68-242, branch :180-240).

Same constructor, same ``create_summary(options, bed_list, thread_id)`` call, same +-100 bp margin, the same
reservoir down-sampling with the fixed seed, and the same ten threshold options -- but the generator it constructs is
this package's ``PEPPER_VARIANT.RegionalSummaryGenerator`` (CUDA path). ``bam_handler`` / ``fasta_handler`` are the
caller's objects (anything with ``get_reads(chrom, start, stop, include_supplementary, min_mapq, min_baseq)`` returning
``type_read``-like objects and ``get_reference_sequence(chrom, start, stop)``): BAM/FASTA decoding is outside the hot path.
"""
from __future__ import annotations

import numpy as np

REGION_SAFE_BASES = 100          # Options.py:2  ConsensCandidateFinder.REGION_SAFE_BASES
MAX_READS_IN_REGION = 5000       # Options.py:98 AlingerOptions.MAX_READS_IN_REGION
RANDOM_SEED = 2719747673         # Options.py:99
CANDIDATE_WINDOW_SIZE = 32       # Options.py:8
IMAGE_HEIGHT = 26                # Options.py:6


def _module():
    from .build import PEPPER_VARIANT      # pepper-thesis_b200/build/PEPPER_VARIANT*.so (built by build.build_pymod)
    return PEPPER_VARIANT


def reservoir_downsample(all_reads, downsample_rate):
    """AlignmentSummarizer.py:191-208, including np.random.RandomState(2719747673).randint draws."""
    total_reads = len(all_reads)
    total_allowed_reads = int(min(MAX_READS_IN_REGION, downsample_rate * total_reads))
    if total_reads > total_allowed_reads:
        random = np.random.RandomState(RANDOM_SEED)
        sample = []
        for i, read in enumerate(all_reads):
            if len(sample) < total_allowed_reads:
                sample.append(read)
            else:
                j = random.randint(0, i + 1)
                if j < total_allowed_reads:
                    sample[j] = read
        all_reads = sample
    return all_reads


class AlignmentSummarizer:
    def __init__(self, bam_handler, fasta_handler, chromosome_name, region_start, region_end):
        self.bam_handler = bam_handler
        self.fasta_handler = fasta_handler
        self.chromosome_name = chromosome_name
        self.region_start_position = region_start
        self.region_end_position = region_end

    def create_summary(self, options, bed_list=None, thread_id=0):
        if getattr(options, "train_mode", False):
            raise NotImplementedError("train_mode is outside the B200 hot path (inference only)")
        region_start = max(0, self.region_start_position - REGION_SAFE_BASES)
        region_end = self.region_end_position + REGION_SAFE_BASES
        all_reads = self.bam_handler.get_reads(self.chromosome_name, region_start, region_end,
                                               options.include_supplementary, options.min_mapq, options.min_snp_baseq)
        all_reads = reservoir_downsample(all_reads, options.downsample_rate)
        if len(all_reads) == 0:
            return None
        ref_seq = self.fasta_handler.get_reference_sequence(self.chromosome_name, region_start, region_end + 1)
        pv = _module()
        regional_summary = pv.RegionalSummaryGenerator(self.chromosome_name, region_start, region_end, ref_seq)
        regional_summary.generate_max_insert_summary(all_reads)
        candidate_image_summary = regional_summary.generate_summary(
            all_reads, options.min_snp_baseq, options.min_indel_baseq, options.snp_frequency, options.insert_frequency,
            options.delete_frequency, options.min_coverage_threshold, options.snp_candidate_frequency_threshold,
            options.indel_candidate_frequency_threshold, options.candidate_support_threshold, options.skip_indels,
            self.region_start_position, self.region_end_position, CANDIDATE_WINDOW_SIZE, IMAGE_HEIGHT, False)
        all_candidate_images = []
        all_candidate_images.extend(candidate_image_summary)
        return all_candidate_images
