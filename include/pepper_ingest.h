/*
 * pepper_ingest.h -- C ABI of libpv_ingest.so: BAM / FASTA ingest without htslib (BGZF over zlib, BAI bin query,
 * .fai random access), emitting the packed read batch of pepper_b200.h directly.
 *
 * This is SURVEY.md section 8(f) row 1: the caller either side of the hot path. Each entry point cites the
 * reference interface it replaces (paths relative to /root/reference):
 *
 *   pv_bam_open / pv_bam_close          BAM_handler::BAM_handler / ~BAM_handler   pepper_variant/modules/cpp/bam_handler.cpp:7-29, 446-451
 *   pv_bam_n_targets / _target_name /
 *   _target_len                         get_chromosome_sequence_names(_with_length)  bam_handler.cpp:85-113
 *   pv_bam_sample_names                 get_sample_names                          bam_handler.cpp:31-56
 *   pv_bam_get_reads                    BAM_handler::get_reads                    bam_handler.cpp:115-444
 *                                       (region clipping of every read, flag / mapq filters, soft clips and inserts
 *                                       kept only behind an anchor, HP tag), as bound in pybind_api.h:223-232
 *   pv_fasta_open / _close / _seq_len /
 *   _n_seq / _seq_name / _fetch         FASTA_handler                             pepper_variant/modules/cpp/fasta_handler.cpp:7-56
 *   pv_ingest_regions                   the per-interval body of AlignmentSummarizer.create_summary up to the
 *                                       generate_summary call                     pepper_variant/modules/python/AlignmentSummarizer.py:180-220
 *                                       for MANY intervals at once (one worker thread per interval, like
 *                                       ImageGenerationUI.py:211,326), output = one PvReadBatch
 *
 * Host only (no CUDA). All functions return PV_OK (0) or a negative PV_E* code; pv_ingest_last_error() has the text.
 * Unlike the reference (exit(EXIT_FAILURE) inside native code, bam_handler.cpp:9-26) errors are returned.
 */
#ifndef PEPPER_INGEST_H
#define PEPPER_INGEST_H

#include <stdint.h>
#include "pepper_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct PvBamFile PvBamFile;
typedef struct PvFastaFile PvFastaFile;
typedef struct PvIngestBatch PvIngestBatch;

const char* pv_ingest_last_error(void);

/* bai_path NULL -> "<bam_path>.bai", then "<bam_path without .bam>.bai" */
int pv_bam_open(const char* bam_path, const char* bai_path, PvBamFile** out);
void pv_bam_close(PvBamFile* f);
int32_t pv_bam_n_targets(const PvBamFile* f);
const char* pv_bam_target_name(const PvBamFile* f, int32_t i);
int64_t pv_bam_target_len(const PvBamFile* f, int32_t i);
/* '\n'-separated, sorted, unique SM values of the @RG lines; returns the byte length (excluding the NUL) */
int64_t pv_bam_sample_names(const PvBamFile* f, char* out, int64_t out_cap);

int pv_fasta_open(const char* fasta_path, PvFastaFile** out);   /* needs "<fasta_path>.fai" */
void pv_fasta_close(PvFastaFile* f);
int32_t pv_fasta_n_seq(const PvFastaFile* f);
const char* pv_fasta_seq_name(const PvFastaFile* f, int32_t i);
int64_t pv_fasta_seq_len(const PvFastaFile* f, const char* name);      /* -1 when absent */
/* bases [start, stop) of `name`, upper-cased, clipped to the sequence; *out_len = bytes written (out holds stop-start) */
int pv_fasta_fetch(const PvFastaFile* f, const char* name, int64_t start, int64_t stop, char* out, int64_t* out_len);

/* Options of get_reads / create_summary */
typedef struct PvIngestOptions {
    int32_t include_supplementary;   /* options.include_supplementary */
    int32_t min_mapq;                /* options.min_mapq */
    int32_t min_baseq;               /* options.min_snp_baseq as passed to get_reads (only feeds bad_indicies there) */
    int32_t safe_bases;              /* ConsensCandidateFinder.REGION_SAFE_BASES = 100 (Options.py:2) */
    int32_t threads;                 /* worker threads over intervals; <= 0 -> hardware concurrency */
    int32_t _pad;
} PvIngestOptions;

/*
 * Reads of n_regions intervals [region_start[i], region_end[i]] (inclusive, the candidate interval of
 * generate_summary) of one contig. Per interval: reads of [max(0, start - safe), end + safe] through get_reads
 * semantics, the reference sequence of that span (+1 base, AlignmentSummarizer.py:214-216; padded with 'N' past the
 * contig end where the reference would read out of bounds), region_ref_start/end = the padded span,
 * region_cand_start/end = the interval. Reads keep file order. No down-sampling here: the reservoir sample of
 * AlignmentSummarizer.py:191-208 is drawn on indices by the host code and applied with pv_ingest_select().
 */
int pv_ingest_regions(PvBamFile* bam, const PvFastaFile* fasta, const char* contig, int32_t n_regions,
                      const int64_t* region_start, const int64_t* region_end, const PvIngestOptions* opt,
                      PvIngestBatch** out);
/* Same read extraction for one span, no reference: get_reads(contig, start, stop, ...) */
int pv_bam_get_reads(PvBamFile* bam, const char* contig, int64_t start, int64_t stop, const PvIngestOptions* opt,
                     PvIngestBatch** out);
/* borrowed view of the batch's host arrays (valid until pv_ingest_free) */
int pv_ingest_view(const PvIngestBatch* b, PvReadBatch* view);
const int32_t* pv_ingest_hp_tags(const PvIngestBatch* b);        /* type_read.hp_tag per read */
const int64_t* pv_ingest_pos_end(const PvIngestBatch* b);        /* type_read.pos_end per read */
const uint16_t* pv_ingest_bam_flags(const PvIngestBatch* b);     /* BAM FLAG per read (type_read.flags, bam_handler.cpp:72-87) */
const char* pv_ingest_query_names(const PvIngestBatch* b, int64_t* total_bytes);   /* NUL-separated, read order */
/* keep reads keep_idx[0..n_keep) (global read indices, grouped by region in region order); new batch */
int pv_ingest_select(const PvIngestBatch* b, const int64_t* keep_idx, int64_t n_keep, PvIngestBatch** out);
void pv_ingest_free(PvIngestBatch* b);
/* diagnostics: uncompressed BGZF bytes inflated by this library so far (all handles, all threads) */
uint64_t pv_ingest_inflated_bytes(void);
/* ... and how many BGZF blocks were decoded by the library's own DEFLATE decoder (csrc/fast_inflate.h; every block is
 * confirmed by its CRC-32, zlib takes over otherwise; PV_INGEST_ZLIB_ONLY=1 switches the decoder off) */
uint64_t pv_ingest_fast_blocks(void);
/* The decoder itself, for tests: raw DEFLATE (RFC 1951) of exactly n_out bytes; PV_OK or PV_EINVAL. */
int pv_inflate_raw(const uint8_t* in, int64_t n_in, uint8_t* out, int64_t n_out);

/* ---- host plan of the device-side decode (pv_bam_inflate_blocks ... pv_bam_clip_write of pepper_b200.h) ----
 * pv_bam_plan: the BAI query of get_reads(contig, start, stop) (bam_handler.cpp:127-130) -> the byte ranges of the file that
 * can hold a record overlapping [start, stop). pv_bam_plan_load reads them (threaded pread) into the caller's buffer of
 * pv_bam_plan_comp_bytes bytes -- page-locked, so the upload runs at PCIe speed -- and walks the BGZF headers there:
 * block table (payload offset in that buffer, sizes, CRC, position in the concatenated inflated stream) and the chain
 * segments (record-aligned entry points: chunk starts + the linear index's offsets). */
typedef struct PvBamPlan PvBamPlan;
int pv_bam_plan(const PvBamFile* bam, const char* contig, int64_t start, int64_t stop, PvBamPlan** out);
int64_t pv_bam_plan_comp_bytes(const PvBamPlan* p);
int32_t pv_bam_plan_tid(const PvBamPlan* p);
int pv_bam_plan_load(PvBamPlan* p, uint8_t* comp_out, int32_t threads);
int32_t pv_bam_plan_n_blocks(const PvBamPlan* p);
int32_t pv_bam_plan_n_segments(const PvBamPlan* p);
int64_t pv_bam_plan_inflated_bytes(const PvBamPlan* p);
int pv_bam_plan_tables(const PvBamPlan* p, PvBgzfBlock* blocks_out, int64_t* seg_begin_out, int64_t* seg_end_out);
void pv_bam_plan_free(PvBamPlan* p);

#ifdef __cplusplus
}
#endif
#endif
