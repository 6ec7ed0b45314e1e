/*
 * pepper_b200.h -- C-ABI of the B200-native PEPPER hot path (libpepper_b200.so).
 *
 * Plain pointers and sizes only; no C++/torch types. Every entry point returns 0 on success and a negative
 * PV_E* code on failure; pv_last_error() then holds a message (thread-local). Nothing here ever falls back
 * to the CPU: without a CUDA device every compute entry point fails with PV_ENODEVICE.
 *
 * What each entry point replaces in the reference (/root/reference, paths relative to it):
 *
 *   pv_summary_*            RegionalSummaryGenerator::{ctor, generate_max_insert_summary, generate_summary}
 *                           pepper_variant/modules/cpp/region_summary.cpp:9-17, 69-96, 568-916 as bound in
 *                           pepper_variant/modules/cpp/pybind_api.h:55-62 and called from
 *                           pepper_variant/modules/python/AlignmentSummarizer.py:220-238
 *   PvReadBatch             type_read / CigarOp (read.h:60-108, cigar.h:30-53): packed SoA instead of AoS
 *   PvCandidates            vector<CandidateImageSummary> (region_summary.h:88-111)
 *   pv_lstm_*               pepper_variant TransducerGRU.forward (modules/python/models/simple_model.py:48-82)
 *                           as driven by predict_distributed_gpu.py:57-69 (weights: ModelHander.py:18-44)
 *   pv_gru_*                pepper (polisher) TransducerGRU.forward (pepper/modules/python/models/simple_model.py:27-42)
 *                           and the sliding-window loop predict_distributed_gpu.py:63-96
 */
#ifndef PEPPER_B200_H
#define PEPPER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PV_OK           0
#define PV_EINVAL      -1   /* bad argument / inconsistent batch */
#define PV_ENODEVICE   -2   /* no CUDA device / wrong architecture (needs sm_100) */
#define PV_ECUDA       -3   /* CUDA runtime error (message has the detail) */
#define PV_EOVERFLOW   -4   /* more candidates than the caller-provided capacity (count is still returned) */
#define PV_ENOMEM      -5

#define PV_WINDOW       33  /* candidate_window_size + 1 rows (Options.py:8, region_summary.cpp:831) */
#define PV_FEATURES     26  /* ImageSizeOptions.IMAGE_HEIGHT (Options.py:6) */
#define PV_ALLELE_BYTES 64  /* "<type digit><bases>", at most 61 bytes (region_summary.cpp:461,511), NUL padded */

/* CIGAR op codes (cigar.h:15-28 == BAM). cigar[] entries are BAM-encoded: (length << 4) | op. */
enum { PV_CIGAR_MATCH = 0, PV_CIGAR_INS = 1, PV_CIGAR_DEL = 2, PV_CIGAR_REF_SKIP = 3, PV_CIGAR_SOFT_CLIP = 4,
       PV_CIGAR_HARD_CLIP = 5, PV_CIGAR_PAD = 6, PV_CIGAR_EQUAL = 7, PV_CIGAR_DIFF = 8 };

/*
 * Packed read batch: the reads of n_regions regions, grouped by region (region_read_begin[r] ..
 * region_read_begin[r+1]). All arrays live in the SAME memory space (all host for *_host entry points, all
 * device for the others). Invariants checked by pv_batch_validate():
 *   read_base_off[i] % 16 == 0; read_base_off[i] + read_len[i] <= n_bases; n_bases % 16 == 0 when a packed wire form
 *   is used; cigar ranges inside n_ops;
 *   region_ref_end >= region_ref_start; region_ref_len[r] >= ref_end-ref_start+1; region_ref_off[r] + region_ref_len[r] <= n_ref;
 *   reads per region <= 32767 (window values must fit int16).
 */
typedef struct PvReadBatch {
    int64_t n_reads, n_bases, n_ops, n_ref;
    int32_t n_regions;
    /* per read */
    const int64_t*  read_pos;         /* type_read.pos: reference position of the first aligned base */
    const int64_t*  read_base_off;    /* offset of the read's first base in bases[]/quals[] (16-byte aligned) */
    const int32_t*  read_len;         /* type_read.sequence.length() */
    const int64_t*  read_cigar_off;   /* offset of the read's first op in cigar[] */
    const int32_t*  read_n_ops;
    const uint8_t*  read_flags;       /* bit0: type_read.flags.is_reverse */
    const uint8_t*  read_mapq;        /* type_read.mapping_quality clipped to [0,255]; 0 = skipped (region_summary.cpp:619) */
    /* per base / per op */
    const uint8_t*  bases;            /* type_read.sequence bytes as given (any byte value) */
    const uint8_t*  quals;            /* type_read.base_qualities (0..255) */
    const uint32_t* cigar;
    /* per region */
    const int64_t*  region_ref_start; /* RegionalSummaryGenerator ctor region_start (inclusive) */
    const int64_t*  region_ref_end;   /* region_end (inclusive) */
    const int64_t*  region_cand_start;/* generate_summary candidate_region_start */
    const int64_t*  region_cand_end;  /* candidate_region_end (inclusive) */
    const int64_t*  region_ref_off;   /* offset of the region's reference_sequence in ref[] */
    const int64_t*  region_ref_len;   /* reference_sequence.length() (>= ref_end-ref_start+1; deletion alleles are
                                         reference substrings truncated at this length, region_summary.cpp:500) */
    const int64_t*  region_read_begin;/* n_regions + 1 entries */
    const uint8_t*  ref;              /* reference bytes as given; region r owns region_ref_len[r] bytes */
    /* Optional wire format for HOST batches: the BAM-native 4-bit packing of `bases` (nt16 codes of
     * "=ACMGRSVTWYHKDBN", base i in byte i/2, even i in the high nibble; n_bases/2 bytes). When non-NULL the host entry
     * points upload this instead of `bases` (which may then be NULL) and expand it on the device (pv_unpack_bases4);
     * only valid when every base is one of those 16 upper-case letters (what BAM_handler::get_reads produces,
     * bam_handler.cpp:213). Ignored (may be NULL) for device-resident batches. */
    const uint8_t*  bases4;
    /* Optional wire formats of `quals` and `cigar` for HOST batches (same rules as bases4: uploaded instead of the plain
     * arrays, expanded on the device, lossless):
     *   quals_packed: dense little-endian bit stream, quality i in bits [i*qual_bits, (i+1)*qual_bits), qual_bits in
     *                 1..7 = bits of the largest quality of the batch; ceil(n_bases/32)*qual_bits*4 bytes (pv_pack_quals).
     *   cigar16:      the low 16 bits of every CIGAR word; only valid when every op length is < 4096 (pv_pack_cigar16). */
    const uint8_t*  quals_packed;
    int32_t         qual_bits;
    /* Promise about `quals` (any batch, host or device): every quality of every read base [base_off, base_off + len) is
     * >= min_qual; 0 = no promise. A batch whose min_qual clears both quality thresholds of a summary call passes every
     * quality test of region_summary.cpp (:377, :393, :448-463), so the tile kernel skips the quality array altogether.
     * pv_min_qual computes it for a host batch. A wrong promise gives wrong candidates. */
    int32_t         min_qual;
    const uint16_t* cigar16;
    /*   bases2:      2 bits per base (A C G T = 0 1 2 3, base i in byte i/4 at bits 2*(i%4)); every base that is not an
     *                upper-case A/C/G/T is listed in base_exceptions as (index << 8) | byte (ascending index). The 0
     *                padding behind a read is not listed (it decodes to 'A'; no kernel looks at it). Takes precedence
     *                over bases4 (pv_pack_bases2). */
    const uint8_t*  bases2;
    const uint64_t* base_exceptions;
    int64_t         n_base_exceptions;
} PvReadBatch;

/* The ten scalars of generate_summary (region_summary.h:191-201), same order, same double compares. */
typedef struct PvThresholds {
    double min_snp_baseq, min_indel_baseq;
    double snp_freq, insert_freq, delete_freq;
    double min_coverage;
    double snp_candidate_freq, indel_candidate_freq;
    double candidate_support;
    int32_t skip_indels;
    int32_t _pad;
} PvThresholds;

/*
 * Candidate output, caller-allocated for `capacity` candidates, ordered like the reference emits them
 * (region ascending, position ascending, allele in std::set<std::string> order; region_summary.cpp:669-670).
 */
typedef struct PvCandidates {
    int64_t  capacity;
    int16_t* windows;     /* [capacity][33][26] CandidateImageSummary.image_matrix (values as C++ int, fit int16) */
    int64_t* position;    /* CandidateImageSummary.position */
    int32_t* region;      /* index of the region in the batch (gives .contig) */
    int32_t* depth;       /* CandidateImageSummary.depth (<=125) */
    int32_t* frequency;   /* CandidateImageSummary.candidate_frequency[0] (<=125) */
    uint8_t* allele;      /* [capacity][PV_ALLELE_BYTES] CandidateImageSummary.candidates[0] */
    uint8_t* allele_len;  /* strlen of the above */
} PvCandidates;

const char* pv_version(void);
const char* pv_last_error(void);
/* number of usable CUDA devices (sm_100); 0 when none. Never fails. */
int pv_device_count(void);

/* Launch accounting and optional CUDA-event profiling of the library's own kernels (used by bench.py for
 * `gpu_launches` and the per-kernel roofline). Families, in order: 0 summary per-tile work lists (CIGAR spans), 1 summary pileup tile,
 * 2 summary site alleles, 3 summary key sort, 4 summary window emit, 5 LSTM input prep, 6 LSTM encoder steps,
 * 7 LSTM decoder steps, 8 LSTM MLP + head, 9 GRU steps, 10 GRU misc, 11 candidate filter, 12 polisher summary,
 * 13 GRU input-projection GEMMs, 14 GRU head, 15 BAM decode (pv_bam_*). */
#define PV_PROFILE_FAMILIES 16
void pv_profile_enable(int on);
int pv_profile_collect(double* ms_by_family, int64_t* launches_by_family);
void pv_profile_reset(void);
int64_t pv_launch_count(void);

/* byte offset, inside the summary workspace, of the int32 status word of the last pv_summary_regions call
 * (bit0 site scratch overflow, bit1 allele-event scratch overflow, bit2 candidate capacity overflow -- all three cured by a
 * larger out->capacity --, bit3 internal, bit4 the batch came without its quality array (quals == NULL, allowed when min_qual
 * clears both thresholds) but a read whose CIGAR runs over its own end needed one: re-run with the qualities, bit5 internal:
 * work-list scratch exhausted). */
int pv_summary_status_offset(void);

/* Device: expand 4-bit packed bases (see PvReadBatch.bases4) into one byte per base; n_bases must be a multiple of 16. */
int pv_unpack_bases4(const uint8_t* packed_dev, int64_t n_bases, uint8_t* bases_dev, void* stream);
/* Device: expand a bit-packed quality stream (PvReadBatch.quals_packed) into one byte per quality; n_bases % 16 == 0;
 * packed_dev must hold ceil(n_bases/32)*qual_bits*4 readable bytes. */
int pv_unpack_quals(const uint8_t* packed_dev, int64_t n_bases, int32_t qual_bits, uint8_t* quals_dev, void* stream);
/* Host: bits needed for the largest quality (1..8; 8 = not worth packing) and the packing itself
 * (packed_host: ceil(n_bases/32)*qual_bits*4 bytes). */
int32_t pv_qual_bits(const uint8_t* quals_host, int64_t n_bases, int32_t threads);
int pv_pack_quals(const uint8_t* quals_host, int64_t n_bases, int32_t qual_bits, uint8_t* packed_host, int32_t threads);
/* Device: zero-extend 16-bit CIGAR words; Host: truncate them (PV_EINVAL if an op length is >= 4096). */
int pv_unpack_cigar16(const uint16_t* packed_dev, int64_t n_ops, uint32_t* cigar_dev, void* stream);
int pv_pack_cigar16(const uint32_t* cigar_host, int64_t n_ops, uint16_t* packed_host, int32_t threads);
/* Device: expand 2-bit bases and apply the exception list; n_bases % 16 == 0. */
int pv_unpack_bases2(const uint8_t* packed_dev, int64_t n_bases, const uint64_t* exceptions_dev, int64_t n_exceptions,
                     uint8_t* bases_dev, void* stream);
/* Host: 2-bit packing. packed_host: n_bases/4 bytes. exceptions_host may be NULL to only count; else it must hold
 * *n_exceptions entries (the count of a previous call). */
int pv_pack_bases2(const uint8_t* bases_host, int64_t n_bases, uint8_t* packed_host, uint64_t* exceptions_host,
                   int64_t* n_exceptions, int32_t threads);
/* Reference-predicted bases (wire form "bases_ref" of the host path; csrc/wire_ref.cu). Every base of a read is predicted
 * from the batch's own CIGAR and reference (an aligned M/=/X base inside the region's reference_sequence = that reference
 * byte, anything else = 'A'); only bases that differ from their prediction travel, as 16-bit entries per read (low byte
 * s: 0..254 = the patched base sits s bases behind the cursor, 255 = move the cursor 255 bases; high byte = the base).
 * Lossless for any byte value. The reference has no counterpart (type_read.sequence is a std::string, read.h:60-108).
 * Host: read_patch_off_host [n_reads + 1]. First call with patches_host == NULL fills it (entry counts as an exclusive
 * prefix; the total is read_patch_off_host[n_reads]); second call writes the entries. Needs the plain bases, cigar, ref.
 * Device: rebuilds bases_dev[read_base_off[r] .. + read_len[r]) of every read (+ zero padding to the 16-byte boundary) from
 * a batch whose cigar, ref and per-read / per-region arrays are device pointers. */
int pv_pack_bases_ref(const PvReadBatch* host_batch, int64_t* read_patch_off_host, uint16_t* patches_host,
                      int64_t patch_capacity, int32_t threads);
int pv_unpack_bases_ref(const PvReadBatch* batch_dev_ptrs, const int64_t* read_patch_off_dev, const uint16_t* patches_dev,
                        uint8_t* bases_dev, void* stream);
/* 8-bit CIGAR (wire form "cigar8"): one code byte c per op, indexed like cigar[] -- bit 0 = 0: M of length (c >> 1) + 1
 * (1..128); bit 0 = 1: bits 2:1 = 0 I / 1 D of length (c >> 3) + 1 (1..32), 3 = escaped: the op's full 32-bit word is the
 * next entry of the escape stream (other op types, longer or empty ops). Escapes of read r sit at [read_esc_off[r], read_esc_off[r + 1]). Lossless.
 * Host: first call with escapes_host == NULL writes the code bytes and fills read_esc_off_host [n_reads + 1]; second call
 * writes the escape words. Device: rebuilds cigar_dev[] (needs read_cigar_off / read_n_ops on the device). */
int pv_pack_cigar8(const PvReadBatch* host_batch, uint8_t* codes_host, int64_t* read_esc_off_host, uint32_t* escapes_host,
                   int64_t escape_capacity, int32_t threads);
int pv_unpack_cigar8(const PvReadBatch* batch_dev_ptrs, const uint8_t* codes_dev, const int64_t* read_esc_off_dev,
                     const uint32_t* escapes_dev, uint32_t* cigar_dev, void* stream);
/* Quality predicates (wire form "quals_pred"; csrc/wire_ref.cu). The summary reads a base quality only through
 * `q >= min_snp_baseq` (region_summary.cpp:377,:393) and, per insert, `sum q[anchor .. anchor+len] >= min_indel_baseq *
 * (len + 1)` with `q[anchor] < min_snp_baseq` (:448-463). With the thresholds known at pack time the device quality array
 * is rebuilt as one fill byte + a per-read patch list (entry format of bases_ref) of SURROGATE values wherever the fill
 * value would change one of those predicates (irregular inserts -- chained, behind a deletion/clip, over the read's end --
 * carry their real qualities). Every predicate evaluates identically on the surrogate array, so summaries and candidates are
 * bit-identical; the qualities themselves are NOT recoverable, and the form is tied to the two thresholds it was packed
 * for (both <= 127, else PV_EINVAL). Two-call protocol like pv_pack_bases_ref; *fill_out receives the fill byte.
 * Device: needs read_base_off / read_len on the device; n_bases % 16 == 0. */
int pv_pack_quals_pred(const PvReadBatch* host_batch, double min_snp_baseq, double min_indel_baseq, uint8_t* fill_out,
                       int64_t* read_patch_off_host, uint16_t* patches_host, int64_t patch_capacity, int32_t threads);
int pv_unpack_quals_pred(const PvReadBatch* batch_dev_ptrs, int32_t fill, const int64_t* read_patch_off_dev,
                         const uint16_t* patches_dev, uint8_t* quals_dev, void* stream);
/* Host: the smallest quality over all read bases of a host batch (padding excluded); 255 for a batch without bases. */
int32_t pv_min_qual(const PvReadBatch* host_batch, int32_t threads);
/* Host: pack ASCII bases into the 4-bit form; returns PV_EINVAL if a byte is outside the nt16 alphabet (0 pads map to '='). */
int pv_pack_bases4(const uint8_t* bases_host, int64_t n_bases, uint8_t* packed_host, int32_t threads);
/* Host: the bases2 + cigar16 forms of one GROUP of regions in one streaming pass per array (csrc/host_pack.cpp: AVX2 when the
 * CPU has it, non-temporal stores, `threads` slices) -- fast enough to run inside the end-to-end path while the group before
 * is on the wire (pipeline.HotPath.run_host(pack_inline=True)). The plain arrays are what the reference hands over
 * (type_read.sequence / cigar_tuples, read.h:60-108); the compact forms are a transport detail, expanded on the device by
 * pv_unpack_bases2 / pv_unpack_cigar16. n_bases % 4 == 0. exceptions_host holds up to exception_capacity entries;
 * *n_exceptions receives the number found (larger than the capacity = nothing was written: call again with room).
 * *cigar_fits = 0 when an op length is >= 4096 (cigar16_host is then not usable: upload the plain words). Output identical to
 * pv_pack_bases2 / pv_pack_cigar16. */
int pv_pack_group(const uint8_t* bases_host, int64_t n_bases, uint8_t* bases2_host, uint64_t* exceptions_host,
                  int64_t exception_capacity, int64_t* n_exceptions, const uint32_t* cigar_host, int64_t n_ops,
                  uint16_t* cigar16_host, int32_t* cigar_fits, int32_t threads);

/* Host-side consistency check of a HOST-resident batch. */
int pv_batch_validate(const PvReadBatch* host_batch);

/* ---------------------------------------------------------------------------------------------------------
 * Summary: device-resident interface. `stream` is a cudaStream_t passed as void*.
 * pv_summary_workspace_bytes: bytes of scratch pv_summary_regions needs for this batch shape (max_region_len = the longest
 *   region_ref_end - region_ref_start + 1 of the batch; it bounds the per-tile work lists: a read touches at most every
 *   tile of its own region. <= 0 means "unknown": total_positions is assumed).
 * pv_summary_regions: everything asynchronous on `stream`; *n_candidates_dev (device int64) receives the
 *   number of candidates found. If it exceeds out->capacity an unspecified subset of `capacity` candidates is
 *   stored (still sorted); the host wrapper reports PV_EOVERFLOW and the caller retries with a larger capacity.
 * dense_image_dev (optional, may be NULL): int16 [total_positions][26] clamped image_matrix of every region
 *   back to back (debug / parity hook for region_summary.cpp:598-654).
 * ------------------------------------------------------------------------------------------------------- */
int64_t pv_summary_workspace_bytes(int64_t n_reads, int64_t n_ops, int32_t n_regions, int64_t total_positions,
                                   int64_t max_region_len, int64_t capacity);
int pv_summary_regions(const PvReadBatch* batch_dev_ptrs, const int64_t* region_len_host, int64_t total_positions,
                       const PvThresholds* thr, int32_t window, int32_t features,
                       const PvCandidates* out_dev_ptrs, int64_t* n_candidates_dev,
                       void* workspace_dev, int64_t workspace_bytes, int16_t* dense_image_dev, void* stream);

/* Summary: host-buffer convenience (what the pybind RegionalSummaryGenerator and bench e2e call).
 * Copies the batch to the current device, runs pv_summary_regions, copies the candidates back.
 * out_host arrays are host memory with `capacity` slots; *n_candidates receives the true count. */
int pv_summary_regions_host(const PvReadBatch* host_batch, const PvThresholds* thr, int32_t window, int32_t features,
                            const PvCandidates* out_host, int64_t* n_candidates, int16_t* dense_image_host);

/* ---------------------------------------------------------------------------------------------------------
 * Model M-A: pepper_variant TransducerGRU (2x biLSTM-256 + 5x SELU MLP + softmax(3)).
 * Weights are the fp32 tensors of the checkpoint's model_state_dict (train_distributed.py:36-42), row-major,
 * PyTorch layouts; [2] = {forward, reverse}.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct PvLstmWeights {
    const float* enc_w_ih[2];   /* [1024][26]  encoder.weight_ih_l0(_reverse), gate rows i,f,g,o */
    const float* enc_w_hh[2];   /* [1024][256] */
    const float* enc_b_ih[2];   /* [1024] */
    const float* enc_b_hh[2];   /* [1024] */
    const float* dec_w_ih[2];   /* [1024][512] */
    const float* dec_w_hh[2];   /* [1024][256] */
    const float* dec_b_ih[2];
    const float* dec_b_hh[2];
    const float* lin_w[5];      /* linear_1 [512][16896], linear_2..5 [512][512] */
    const float* lin_b[5];      /* [512] */
    const float* out_w;         /* output_layer_type [3][512] */
    const float* out_b;         /* [3] */
} PvLstmWeights;

typedef struct PvLstmModel PvLstmModel;
int pv_lstm_create(const PvLstmWeights* host_weights, PvLstmModel** model);
void pv_lstm_destroy(PvLstmModel* model);
int64_t pv_lstm_workspace_bytes(int64_t max_windows);
/* windows_dev int16 [n][33][26] (4-byte aligned); probs_dev float [n][3] (softmax); argmax_dev uint8 [n] (may be NULL).
 * wrap_int8 != 0 reproduces the reference pipeline's int8 HDF5 round trip (DataStore.py:68,
 * dataloader_predict.py:90): the value fed to the network is (int8_t)window value. */
int pv_lstm_infer(PvLstmModel* model, const int16_t* windows_dev, int64_t n, int32_t wrap_int8,
                  float* probs_dev, uint8_t* argmax_dev, void* workspace_dev, int64_t workspace_bytes, void* stream);
/* host-buffer convenience: H2D, infer (in chunks), D2H */
int pv_lstm_infer_host(PvLstmModel* model, const int16_t* windows_host, int64_t n, int32_t wrap_int8,
                       float* probs_host, uint8_t* argmax_host);

/* ---------------------------------------------------------------------------------------------------------
 * Model M-B: pepper (polisher) TransducerGRU (2x biGRU-128 + Linear(256,5)), T = 100 positions, 10 features.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct PvGruWeights {
    const float* enc_w_ih[2];   /* [384][10]  gru_encoder.weight_ih_l0(_reverse), gate rows r,z,n */
    const float* enc_w_hh[2];   /* [384][128] */
    const float* enc_b_ih[2];   /* [384] */
    const float* enc_b_hh[2];
    const float* dec_w_ih[2];   /* [384][256] */
    const float* dec_w_hh[2];   /* [384][128] */
    const float* dec_b_ih[2];
    const float* dec_b_hh[2];
    const float* dense_w;       /* dense1 [5][256] */
    const float* dense_b;       /* [5] */
} PvGruWeights;

typedef struct PvGruModel PvGruModel;
int pv_gru_create(const PvGruWeights* host_weights, PvGruModel** model);
void pv_gru_destroy(PvGruModel* model);
int64_t pv_gru_workspace_bytes(int64_t max_batch, int32_t seq_len);
/* One TransducerGRU.forward: images_dev uint8 [n][seq_len][10] (raw 0..254 counts cast to float, no
 * normalisation; predict_distributed_gpu.py:60), hidden_dev float [n][2][128] read as h0 and overwritten with
 * the decoder's h_n; logits_dev float [n][seq_len][5]. */
int pv_gru_forward(PvGruModel* model, const uint8_t* images_dev, int64_t n, int32_t seq_len, float* hidden_dev,
                   float* logits_dev, void* workspace_dev, int64_t workspace_bytes, void* stream);
/* The polisher's chunk loop (predict_distributed_gpu.py:63-96): images_dev uint8 [n][chunk_len][10]; windows of
 * `window` positions every `stride`; hidden carried between windows; softmax summed over overlaps into
 * prob_sum_dev float [n][chunk_len][5]; labels_dev uint8 [n][chunk_len] = argmax. */
int pv_gru_predict_chunks(PvGruModel* model, const uint8_t* images_dev, int64_t n, int32_t chunk_len, int32_t window,
                          int32_t stride, float* prob_sum_dev, uint8_t* labels_dev, void* workspace_dev,
                          int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Stage 3 ("next" row 3 of SURVEY.md 8f): the per-candidate decision of small_chunk_stitch
 * (pepper_variant/modules/python/CandidateFinder.py:391-529) on the device.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct PvFilterOptions {          /* options.* of find_candidates (CandidateFinder.py:480-515) */
    double snp_p_value, snp_p_value_in_lc;
    double insert_p_value, insert_p_value_in_lc;
    double delete_p_value, delete_p_value_in_lc;
    double report_snp_above_freq, report_indel_above_freq;
} PvFilterOptions;
/* flags per candidate */
#define PV_FLT_PHASING    1    /* goes to selected_candidate_list_margin (:444-451) */
#define PV_FLT_VARIANT    2    /* goes to selected_candidate_list_deepvariant (:480-519) */
#define PV_FLT_IN_REPEAT  4    /* candidate_in_repeat: homopolymer run >= 5 within [-5, +4) of the site (:399-414) */
#define PV_FLT_GT_SHIFT   3    /* bits 3-4: argmax of the class probabilities (0 hom-ref, 1 het, 2 hom-alt; :419-426) */
#define PV_FLT_DEL_BY_FREQ 32  /* a delete kept by the frequency rule: alt = the allele string, ref stays one base (:513-515) */
#define PV_FLT_BAD_REF    64   /* reference base not A/C/G/T: skipped (:416-417) */
/* All pointers device memory; candidate arrays as in PvCandidates (+ probs float [n][3] from pv_lstm_infer); the region
 * arrays / ref are those of the device PvReadBatch the candidates came from; region_contig_len (may be NULL) clips
 * the reference context at the contig end like FASTA_handler does. */
int pv_candidate_filter(int64_t n, const int64_t* position, const int32_t* region, const int32_t* depth,
                        const int32_t* frequency, const uint8_t* allele, const uint8_t* allele_len, const float* probs,
                        const int64_t* region_ref_start, const int64_t* region_ref_off, const int64_t* region_ref_len,
                        const int64_t* region_contig_len, const uint8_t* ref, const PvFilterOptions* opt,
                        uint8_t* flags, void* stream);
/* same with host arrays (copies in, runs the kernel, copies the flags out) */
int pv_candidate_filter_host(int64_t n, const int64_t* position, const int32_t* region, const int32_t* depth,
                             const int32_t* frequency, const uint8_t* allele, const uint8_t* allele_len,
                             const float* probs, int32_t n_regions, const int64_t* region_ref_start,
                             const int64_t* region_ref_off, const int64_t* region_ref_len,
                             const int64_t* region_contig_len, const uint8_t* ref, int64_t n_ref,
                             const PvFilterOptions* opt, uint8_t* flags);

/* ---------------------------------------------------------------------------------------------------------
 * Polisher pileup summary ("next" row 4 of SURVEY.md 8f): SummaryGenerator::generate_summary of the `pepper` module
 * (pepper/modules/src/pileup_summary/summary_generator.cpp:47-121, 274-306, 370-392) and chunk_images
 * (pepper/modules/python/AlignmentSummarizer.py:19-56): the producer of model M-B's uint8 [chunk][1000][10] input.
 * Regions of the batch: region_ref_start/end = the generator's ref_start/ref_end = start_pos/end_pos. All device
 * memory except the *_host arguments.
 * ------------------------------------------------------------------------------------------------------- */
int64_t pv_polish_workspace_bytes(int64_t n_reads, int64_t n_ops, int32_t n_regions, int64_t total_positions);
/* counts + row layout; writes the number of output rows (positions + inserted columns) to *n_rows_host and, when
 * given, the first row of every region (n_regions + 1 entries). Synchronises the stream. */
int pv_polish_count(const PvReadBatch* batch_dev, const int64_t* region_len_host, int64_t total_positions,
                    void* workspace_dev, int64_t workspace_bytes, int64_t* n_rows_host, int64_t* region_rows_host,
                    void* stream);
/* image_dev uint8 [n_rows][10] (SummaryGenerator.image), gpos_dev int64 [n_rows][2] (genomic_pos: position, insert
 * index), row_region_dev int32 [n_rows]; ins_scratch_dev uint32 [n_rows][10] scratch. Same workspace as pv_polish_count. */
int pv_polish_emit(const PvReadBatch* batch_dev, const int64_t* region_len_host, int64_t total_positions,
                   void* workspace_dev, int64_t workspace_bytes, int64_t n_rows, uint32_t* ins_scratch_dev,
                   uint8_t* image_dev, int64_t* gpos_dev, int32_t* row_region_dev, void* stream);
/* chunk c = rows [chunk_row[c], chunk_row[c] + chunk_valid[c]) padded with zeros / (-1,-1) to chunk_size rows */
int pv_polish_chunks(const uint8_t* image_dev, const int64_t* gpos_dev, const int64_t* chunk_row_dev,
                     const int64_t* chunk_valid_dev, int64_t n_chunks, int32_t chunk_size, uint8_t* out_images_dev,
                     int64_t* out_positions_dev, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Synthetic inputs generated on the device (bench / tests only; csrc/synth_device.cu): the device twin of the seeded host
 * generator csrc/synth_reads.c, bit-identical to it, so that whole-genome-scale workloads are
 * streamed group by group without existing on the host. cfg = the host generator's PvSynthConfig. Count pass: bases and
 * CIGAR ops of every read (read lengths come from the host: pv_synth_read_lengths of libpv_synth.so); fill pass: the arrays
 * of a device PvReadBatch at the offsets the caller derived from the counts (bases padded to 16 per read).
 * ------------------------------------------------------------------------------------------------------- */
int pv_synth_device_count(const void* cfg, int64_t first_region, int32_t n_regions, const int64_t* read_begin_dev,
                          const int32_t* read_len_dev, int64_t n_reads, int32_t* n_bases_dev, int32_t* n_ops_dev, void* stream);
int pv_synth_device_fill(const void* cfg, int64_t first_region, int32_t n_regions, const int64_t* read_begin_dev,
                         const int32_t* read_len_in_dev, int64_t n_reads, const int64_t* base_off_dev, const int64_t* op_off_dev,
                         int64_t* read_pos_dev, int32_t* read_len_dev, int32_t* read_n_ops_dev, uint8_t* read_flags_dev,
                         uint8_t* read_mapq_dev, uint8_t* bases_dev, uint8_t* quals_dev, uint32_t* cigar_dev,
                         const int64_t* ref_off_dev, uint8_t* ref_dev, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * BAM decoding on the device ("next" row 1 of SURVEY.md 8f; csrc/ingest_gpu.cu, csrc/bam_core.cuh): what
 * BAM_handler::get_reads (pepper_variant/modules/cpp/bam_handler.cpp:115-451) does per byte -- BGZF inflate, record
 * parsing, the cut of every read to its region -- as kernels, so that the packed read batch is written straight into
 * HBM. The host's share (which bytes to read, where the BGZF blocks and the record-chain entry points are) is
 * pv_bam_plan* of pepper_ingest.h. All pointers are device memory; results are bit-identical to pv_ingest_regions.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct PvBgzfBlock {
    int64_t  c_off;      /* offset of the block's DEFLATE payload in the compressed buffer */
    int32_t  c_len;      /* payload bytes */
    int32_t  isize;      /* bytes it inflates to (<= 65536) */
    int64_t  u_off;      /* where they go in the inflated stream (blocks are concatenated) */
    uint32_t crc;        /* CRC-32 of those bytes (BGZF trailer) */
    uint32_t _pad;
} PvBgzfBlock;
typedef struct PvBamPair {   /* one read of the batch: a record cut to one span */
    int64_t rec_off;     /* offset of the record's block_size field in the inflated stream */
    int32_t span;
    int32_t n_ops;       /* kept CIGAR ops */
    int64_t n_bases;     /* kept bases */
    int32_t k_first;     /* first / last kept op of the record's CIGAR: every op between them is kept whole, these two  */
    int32_t k_last;      /* with the lengths below (clipping can only shorten the ends)                                 */
    int32_t first_kept;
    int32_t last_kept;
    int64_t idx0;        /* read index of the first kept base */
    int64_t pos_start;   /* type_read.pos */
    int64_t pos_end;     /* type_read.pos_end */
} PvBamPair;
/* One warp per BGZF block (csrc/inflate_warp.cuh), blocks pulled from a ticket by a grid that fills the device once.
 * n_bad_dev[2]: [0] = blocks that did not decode to isize bytes (or, with verify_crc, whose CRC-32 differs from the
 * trailer), [1] = scratch (the ticket). comp_dev must be 4-byte aligned; it is read in whole 32-bit words, so the
 * allocation must cover comp_bytes rounded up to a multiple of 4. */
int pv_bam_inflate_blocks(const uint8_t* comp_dev, int64_t comp_bytes, const PvBgzfBlock* blocks_dev, int32_t n_blocks,
                          uint8_t* inflated_dev, int64_t inflated_bytes, int32_t verify_crc, int32_t* n_bad_dev, void* stream);
/* Record boundaries. Segment s = [seg_begin[s], seg_end[s]) of the inflated stream starts at a record and ends where the
 * next one starts. rec_off_dev == NULL: count pass -- seg_first_dev[0..n_segments] receives the exclusive scan of the
 * record counts (total in [n_segments]); else fill pass with the same seg_first_dev. *status_dev: bit0 a chain left its
 * segment / ran into a malformed record, bit1 rec_capacity too small. */
int pv_bam_index_records(const uint8_t* inflated_dev, int64_t inflated_bytes, const int64_t* seg_begin_dev,
                         const int64_t* seg_end_dev, int32_t n_segments, int64_t* seg_first_dev, int64_t* rec_off_dev,
                         int64_t rec_capacity, int32_t* status_dev, void* stream);
/* Reads per record: span j = [span_start[j], span_stop[j]] (ascending in both), the get_reads interval of region j.
 * rec_pair_first_dev[0..n_records] receives the exclusive scan of the number of (record, span) pairs that yield a read
 * (flag / mapq filters :137-150, iterator overlap, non-empty cut). status bits: 2 malformed record, 3 internal. */
int pv_bam_clip_count(const uint8_t* inflated_dev, int64_t inflated_bytes, const int64_t* rec_off_dev, int64_t n_records,
                      int32_t tid, const int64_t* span_start_dev, const int64_t* span_stop_dev, int32_t n_spans,
                      int32_t include_supplementary, int32_t min_mapq, int64_t* rec_pair_first_dev, int32_t* status_dev, void* stream);
int64_t pv_bam_clip_workspace_bytes(int64_t n_pairs);
/* The batch's read order (span ascending, file order inside a span) and layout: pairs_sorted_dev[n_pairs],
 * read_base_off_dev / read_cigar_off_dev [n_pairs] (bases padded to 16 per read), region_read_begin_dev[n_spans + 1],
 * totals_dev[2] = {bytes of the base / quality arrays, CIGAR ops}. */
int pv_bam_clip_layout(const uint8_t* inflated_dev, int64_t inflated_bytes, const int64_t* rec_off_dev, int64_t n_records,
                       int32_t tid, const int64_t* span_start_dev, const int64_t* span_stop_dev, int32_t n_spans,
                       int32_t include_supplementary, int32_t min_mapq, const int64_t* rec_pair_first_dev, int64_t n_pairs,
                       void* workspace_dev, int64_t workspace_bytes, PvBamPair* pairs_sorted_dev, int64_t* read_base_off_dev,
                       int64_t* read_cigar_off_dev, int64_t* region_read_begin_dev, int64_t* totals_dev, int32_t* status_dev,
                       void* stream);
/* One warp per read: the PvReadBatch arrays + type_read.pos_end / hp_tag / BAM FLAG, where each query name sits in the
 * inflated stream, and the smallest kept quality (-> PvReadBatch.min_qual). status bit 4: internal. */
int pv_bam_clip_write(const uint8_t* inflated_dev, int64_t inflated_bytes, const PvBamPair* pairs_sorted_dev, int64_t n_pairs,
                      const int64_t* span_start_dev, const int64_t* span_stop_dev, const int64_t* read_base_off_dev,
                      const int64_t* read_cigar_off_dev, int64_t* read_pos_dev, int64_t* read_pos_end_dev, int32_t* read_len_dev,
                      int32_t* read_n_ops_dev, uint8_t* read_flags_dev, uint8_t* read_mapq_dev, int32_t* hp_dev,
                      uint16_t* bam_flag_dev, int64_t* name_off_dev, int32_t* name_len_dev, uint8_t* bases_dev,
                      uint8_t* quals_dev, uint32_t* cigar_dev, int32_t* min_qual_dev, int32_t* status_dev, void* stream);
/* NUL-terminated query names, read i at out_off_dev[i] (the caller's exclusive scan of name_len + 1) */
int pv_bam_gather_names(const uint8_t* inflated_dev, const int64_t* name_off_dev, const int32_t* name_len_dev,
                        const int64_t* out_off_dev, int64_t n, uint8_t* out_dev, void* stream);
/* PvReadBatch.ref from ONE fetch of the contig: region r receives bytes [span_start[r], span_start[r] + region_ref_len[r])
 * of the contig at ref_dev + region_ref_off[r], taken from fetched_dev = contig[fetch_start, fetch_start + fetch_len)
 * ('N' outside it: FASTA_handler clamps at the contig end, the region's reference span can reach one base past it). */
int pv_bam_gather_reference(const uint8_t* fetched_dev, int64_t fetch_len, int64_t fetch_start, const int64_t* span_start_dev,
                            const int64_t* region_ref_off_dev, const int64_t* region_ref_len_dev, int32_t n_spans,
                            int64_t max_ref_len, uint8_t* ref_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PEPPER_B200_H */
