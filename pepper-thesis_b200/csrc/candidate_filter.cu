// Stage 3 of call_variant on the device: the per-candidate decision of small_chunk_stitch
// (/root/reference/pepper_variant/modules/python/CandidateFinder.py:391-529) -- homopolymer annotation of the +-10 bp
// reference context (repeat_annotation, :279-297), genotype = argmax of the three class probabilities, and the
// per-type probability / allele-frequency thresholds -- for every candidate the summary + inference kernels produced,
// without the HDF5 round trip and the process pool. One thread per candidate; everything it reads is already in HBM.
#include "common.cuh"

namespace {

__device__ __forceinline__ uint8_t up(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }
__device__ __forceinline__ bool acgt(uint8_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

struct FilterArgs {
    int64_t n;
    const int64_t* position; const int32_t* region; const int32_t* depth; const int32_t* frequency;
    const uint8_t* allele; const uint8_t* allele_len; const float* probs;
    const int64_t* region_ref_start; const int64_t* region_ref_off; const int64_t* region_ref_len;
    const int64_t* region_contig_len;      // may be null: the region's reference is taken as is
    const uint8_t* ref;
    PvFilterOptions o;
    uint8_t* flags;
};

__global__ void candidate_filter_kernel(const FilterArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const int r = a.region[i];
    const int64_t pos = a.position[i];
    const int64_t rs = a.region_ref_start[r];
    int64_t avail_end = rs + a.region_ref_len[r];                    // reference bytes exist for [rs, avail_end)
    if (a.region_contig_len && a.region_contig_len[r] >= 0 && a.region_contig_len[r] < avail_end) avail_end = a.region_contig_len[r];
    const uint8_t* ref = a.ref + a.region_ref_off[r] - rs;           // ref[p] = base at contig position p
    // full_sequence = ref[max(0, pos-10) : pos] + ref[pos : pos+10], both clipped like faidx does (:393-398)
    int64_t lo = pos - 10; if (lo < 0) lo = 0; if (lo < rs) lo = rs;
    int64_t hi = pos + 10; if (hi > avail_end) hi = avail_end;
    uint8_t s[20];
    const int n_down = pos > lo ? (int)(pos - lo) : 0;
    int len = 0;
    for (int64_t p = lo; p < hi && len < 20; p++) s[len++] = up(ref[p]);
    // repeat_annotation(full_sequence, 1): every position gets the length of the homopolymer run it sits in
    // max over [position_index - 5, position_index + 4) (:403-407)
    int w_lo = n_down - 5; if (w_lo < 0) w_lo = 0;
    int w_hi = n_down + 4; if (w_hi > len) w_hi = len;
    int max_hp = 0;
    for (int k = 0; k < len;) {
        int e = k + 1;
        while (e < len && s[e] == s[k]) e++;
        if (k < w_hi && e > w_lo && e - k > max_hp) max_hp = e - k;
        k = e;
    }
    const bool in_repeat = max_hp >= 5;                              // :412-414
    uint8_t f = in_repeat ? 4 : 0;
    const uint8_t ref_base = (pos >= rs && pos < avail_end) ? up(ref[pos]) : 0;
    if (!acgt(ref_base)) { a.flags[i] = (uint8_t)(f | 64); return; } // :416-417
    const float p0 = a.probs[3 * i], p1 = a.probs[3 * i + 1], p2 = a.probs[3 * i + 2];
    const int g = p1 > p0 ? (p2 > p1 ? 2 : 1) : (p2 > p0 ? 2 : 0);   // np.argmax: first maximum
    f |= (uint8_t)(g << 3);
    const int alen = a.allele_len[i];
    const uint8_t* al = a.allele + i * PV_ALLELE_BYTES;
    bool valid = true;                                               // :436-443
    for (int k = 1; k < alen; k++) valid = valid && acgt(al[k]);
    if (valid && alen > 0) {
        const int type = al[0] - '0';
        if (type == 1 && g != 0) f |= 1;                             // phasing list: SNPs with a non-reference genotype (:444-448)
        // :477 (depth 0 raises ZeroDivisionError in the reference; the host wrapper does the same, here it just must not trap)
        const double vaf = a.depth[i] > 0 ? (double)a.frequency[i] / (double)a.depth[i] : 0.0;
        const double non_alt = (double)(p1 > p2 ? p1 : p2);                              // :478
        const double pv = type == 1 ? (in_repeat ? a.o.snp_p_value_in_lc : a.o.snp_p_value)
                        : type == 2 ? (in_repeat ? a.o.insert_p_value_in_lc : a.o.insert_p_value)
                                    : (in_repeat ? a.o.delete_p_value_in_lc : a.o.delete_p_value);
        const double rep = type == 1 ? a.o.report_snp_above_freq : a.o.report_indel_above_freq;
        if (type >= 1 && type <= 3) {
            if (non_alt >= pv) f |= 2;                               // :480-487, 491-498, 502-512
            else if (0.0 < rep && rep <= vaf) f |= (uint8_t)(2 | (type == 3 ? 32 : 0));   // :488-490, 499-501, 513-515
        }
    }
    a.flags[i] = f;
}

}  // namespace

extern "C" int pv_candidate_filter(int64_t n, const int64_t* position, const int32_t* region, const int32_t* depth,
                                   const int32_t* frequency, const uint8_t* allele, const uint8_t* allele_len,
                                   const float* probs, const int64_t* region_ref_start, const int64_t* region_ref_off,
                                   const int64_t* region_ref_len, const int64_t* region_contig_len, const uint8_t* ref,
                                   const PvFilterOptions* opt, uint8_t* flags, void* stream_) {
    if (n < 0 || !opt) return pv::set_error(PV_EINVAL, "bad argument");
    if (n == 0) return PV_OK;
    if (!position || !region || !depth || !frequency || !allele || !allele_len || !probs || !region_ref_start || !region_ref_off ||
        !region_ref_len || !ref || !flags) return pv::set_error(PV_EINVAL, "null argument");
    if (int rc = pv::require_device()) return rc;
    FilterArgs a;
    a.n = n; a.position = position; a.region = region; a.depth = depth; a.frequency = frequency; a.allele = allele;
    a.allele_len = allele_len; a.probs = probs; a.region_ref_start = region_ref_start; a.region_ref_off = region_ref_off;
    a.region_ref_len = region_ref_len; a.region_contig_len = region_contig_len; a.ref = ref; a.o = *opt; a.flags = flags;
    pv::prof_begin(pv::FAM_FILTER, (cudaStream_t)stream_);
    candidate_filter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(a);
    PV_CUDA_CHECK(cudaGetLastError());
    pv::prof_end(pv::FAM_FILTER, (cudaStream_t)stream_, 1);
    return PV_OK;
}

// host arrays in, host flags out (the records are ~100 bytes per candidate)
extern "C" int pv_candidate_filter_host(int64_t n, const int64_t* position, const int32_t* region, const int32_t* depth,
                                        const int32_t* frequency, const uint8_t* allele, const uint8_t* allele_len,
                                        const float* probs, int32_t n_regions, const int64_t* region_ref_start,
                                        const int64_t* region_ref_off, const int64_t* region_ref_len,
                                        const int64_t* region_contig_len, const uint8_t* ref, int64_t n_ref,
                                        const PvFilterOptions* opt, uint8_t* flags) {
    if (n < 0 || n_regions < 0 || n_ref < 0 || !opt) return pv::set_error(PV_EINVAL, "bad argument");
    if (n == 0) return PV_OK;
    if (int rc = pv::require_device()) return rc;
    for (int64_t i = 0; i < n; i++) if (region[i] < 0 || region[i] >= n_regions) return pv::set_error(PV_EINVAL, "candidate %lld: region out of range", (long long)i);
    const size_t sz[13] = {(size_t)n * 8, (size_t)n * 4, (size_t)n * 4, (size_t)n * 4, (size_t)n * PV_ALLELE_BYTES, (size_t)n, (size_t)n * 12,
                           (size_t)n_regions * 8, (size_t)n_regions * 8, (size_t)n_regions * 8, region_contig_len ? (size_t)n_regions * 8 : 0,
                           (size_t)n_ref, (size_t)n};
    const void* src[12] = {position, region, depth, frequency, allele, allele_len, probs, region_ref_start, region_ref_off,
                           region_ref_len, region_contig_len, ref};
    size_t off[14]; off[0] = 0;
    for (int k = 0; k < 13; k++) off[k + 1] = off[k] + ((sz[k] + 255) & ~(size_t)255);
    uint8_t* d = nullptr;
    PV_CUDA_CHECK(cudaMalloc((void**)&d, off[13] ? off[13] : 256));
    int rc = PV_OK;
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < 12 && e == cudaSuccess; k++) if (sz[k]) e = cudaMemcpy(d + off[k], src[k], sz[k], cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        rc = pv_candidate_filter(n, (const int64_t*)(d + off[0]), (const int32_t*)(d + off[1]), (const int32_t*)(d + off[2]),
                                 (const int32_t*)(d + off[3]), d + off[4], d + off[5], (const float*)(d + off[6]),
                                 (const int64_t*)(d + off[7]), (const int64_t*)(d + off[8]), (const int64_t*)(d + off[9]),
                                 region_contig_len ? (const int64_t*)(d + off[10]) : nullptr, d + off[11], opt, d + off[12], nullptr);
        if (rc == PV_OK) e = cudaMemcpy(flags, d + off[12], sz[12], cudaMemcpyDeviceToHost);
    }
    cudaFree(d);
    if (e != cudaSuccess) return pv::set_error(PV_ECUDA, "candidate filter copy failed: %s", cudaGetErrorString(e));
    return rc;
}
