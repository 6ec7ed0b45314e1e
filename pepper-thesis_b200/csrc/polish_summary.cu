// Polisher pileup summary on sm_100a: SummaryGenerator::generate_summary of the reference's `pepper` module
// (/root/reference/pepper/modules/src/pileup_summary/summary_generator.cpp:47-121, 274-306, 370-392) -- the producer
// of model M-B's input: per reference position 10 normalised uint8 features (A C G T reverse, A C G T forward,
// other reverse, other forward), followed by one row per inserted column up to the longest insert observed at that
// position -- and the 1000/50 chunking of AlignmentSummarizer.chunk_images
// (/root/reference/pepper/modules/python/AlignmentSummarizer.py:19-56).
//
// Kernels (one stream, no host synchronisation except reading back the row count):
//   P0 polish_prefix_kernel   warp per read: CIGAR prefix (here REF_SKIP / PAD advance only the reference, :98-113)
//   P1 polish_count_kernel    warp per read, lane per op: base / deletion counts, coverage (incl. the reference's quirk
//                             of charging a deletion's coverage to the op's first position, :108), longest insert
//   -- cub exclusive scan of (1 + longest insert) -> first output row of every position
//   P2 polish_insert_kernel   warp per read, lane per insert op: counts of the inserted columns
//   P3 polish_emit_kernel     thread per position: normalise (double arithmetic, uint8 conversion as x86 does it) and
//                             write the rows and their (position, insert index) pairs
//   P4 polish_chunk_kernel    gather rows into [n_chunks][chunk][10] windows with zero padding
#include "common.cuh"
#include <cub/device/device_scan.cuh>

namespace {

constexpr int NF = 10;               // ImageSizeOptions.IMAGE_HEIGHT (pepper/modules/python/Options.py:2)

__device__ __forceinline__ int feature_index(uint8_t base, bool rev) {     // get_feature_index, :16-32
    if (base >= 'a' && base <= 'z') base -= 32;
    const int k = base == 'A' ? 0 : base == 'C' ? 1 : base == 'G' ? 2 : base == 'T' ? 3 : -1;
    if (rev) return k >= 0 ? k : 8;
    return k >= 0 ? 4 + k : 9;
}
__device__ __forceinline__ bool is_m(int op) { return op == 0 || op == 7 || op == 8; }

__global__ void polish_prefix_kernel(PvReadBatch b, int32_t* __restrict__ op_ref, int32_t* __restrict__ op_ri) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < b.n_reads; r += n_warps) {
        const int64_t co = b.read_cigar_off[r];
        const int n_ops = b.read_n_ops[r];
        int ref_run = 0, ri_run = 0;
        for (int k0 = 0; k0 < n_ops; k0 += 32) {
            const int k = k0 + lane;
            const uint32_t w = k < n_ops ? b.cigar[co + k] : 0u;
            const int op = (int)(w & 15u), len = k < n_ops ? (int)(w >> 4) : 0;
            const int ra = (is_m(op) || op == 2 || op == 3 || op == 6) ? len : 0;     // :54-113
            const int qa = (is_m(op) || op == 1 || op == 4) ? len : 0;
            int ir = ra, iq = qa;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int tr = __shfl_up_sync(0xffffffffu, ir, d), tq = __shfl_up_sync(0xffffffffu, iq, d);
                if (lane >= d) { ir += tr; iq += tq; }
            }
            if (k < n_ops) { op_ref[co + k] = ref_run + ir - ra; op_ri[co + k] = ri_run + iq - qa; }
            ref_run += __shfl_sync(0xffffffffu, ir, 31);
            ri_run += __shfl_sync(0xffffffffu, iq, 31);
        }
    }
}

struct PolishArgs {
    PvReadBatch b;
    const int64_t* pos_off;        // [n_regions + 1] dense position offset of each region
    const int32_t* read_region;    // [n_reads]
    const int32_t* op_ref; const int32_t* op_ri;
    uint32_t* cnt;                 // [positions][NF]
    uint32_t* cov;                 // [positions]
    uint32_t* longest;             // [positions]
    const int64_t* row_of;         // [positions + 1] first output row of each position (after the scan)
    uint32_t* ins_cnt;             // [rows][NF] (only insert rows are touched)
};

// MODE 0: P1 (bases, deletions, coverage, longest insert); MODE 1: P2 (inserted columns)
template <int MODE>
__global__ void polish_walk_kernel(const PolishArgs a) {
    const PvReadBatch& b = a.b;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < b.n_reads; r += n_warps) {
        if (b.read_mapq[r] == 0) continue;                                          // :376
        const int reg = a.read_region[r];
        const int64_t rs = b.region_ref_start[reg];
        const int64_t L = b.region_ref_end[reg] - rs + 1;
        const int64_t rel = b.read_pos[r] - rs;
        const int64_t g0 = a.pos_off[reg];
        const int64_t co = b.read_cigar_off[r], bo = b.read_base_off[r];
        const int n_ops = b.read_n_ops[r], read_len = b.read_len[r];
        const bool rev = b.read_flags[r] & 1;
        for (int k0 = 0; k0 < n_ops; k0 += 32) {
            const int k = k0 + lane;
            if (k >= n_ops) break;
            const uint32_t w = b.cigar[co + k];
            const int op = (int)(w & 15u);
            const int64_t len = (int64_t)(w >> 4);
            const int64_t p0 = rel + a.op_ref[co + k];                              // region-local reference position
            const int64_t ri = a.op_ri[co + k];
            if (p0 > L - 1) continue;                                               // `if (ref_position > region_end) break`, :55
            if (MODE == 0 && is_m(op)) {                                            // :57-78
                int64_t i = p0 < 0 ? -p0 : 0;
                int64_t i_hi = len; if (i_hi > L - p0) i_hi = L - p0; if (i_hi > read_len - ri) i_hi = read_len - ri;
                for (; i < i_hi; i++) {
                    atomicAdd(&a.cnt[(g0 + p0 + i) * NF + feature_index(b.bases[bo + ri + i], rev)], 1u);
                    atomicAdd(&a.cov[g0 + p0 + i], 1u);
                }
            } else if (op == 1) {                                                   // :79-96
                const int64_t an = p0 - 1;
                if (an < 0 || an > L - 1 || ri > read_len) continue;                // substr(read_index > size) throws in the reference
                int64_t alen = len; if (alen > read_len - ri) alen = read_len - ri; // substr truncation
                if (MODE == 0) { if (alen > 0) atomicMax(&a.longest[g0 + an], (uint32_t)alen); }
                else {
                    const int64_t row = a.row_of[g0 + an] + 1;
                    for (int64_t i = 0; i < alen; i++)
                        atomicAdd(&a.ins_cnt[(row + i) * NF + feature_index(b.bases[bo + ri + i], rev)], 1u);
                }
            } else if (MODE == 0 && (op == 2 || op == 3 || op == 6)) {               // :97-113
                const int64_t i_lo = p0 < 0 ? -p0 : 0;
                int64_t i_hi = len; if (i_hi > L - p0) i_hi = L - p0;
                const int star = feature_index('*', rev);
                for (int64_t i = i_lo; i < i_hi; i++) atomicAdd(&a.cnt[(g0 + p0 + i) * NF + star], 1u);
                // the reference adds this coverage at ref_position (the op's start), once per in-range deleted base
                if (i_hi > i_lo && p0 >= 0) atomicAdd(&a.cov[g0 + p0], (uint32_t)(i_hi - i_lo));
            }
        }
    }
}

__global__ void polish_rows_kernel(const uint32_t* __restrict__ longest, int64_t n, int64_t* __restrict__ rows) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) rows[i] = i < n ? 1 + (int64_t)longest[i] : 0;
}

// uint8_t v = double: x86-64 converts through a truncating 32-bit integer conversion and keeps the low byte
__device__ __forceinline__ uint8_t pixel(uint32_t c, uint32_t cov) {
    const double m = (double)cov > 1.0 ? (double)cov : 1.0;                        // max(1.0, coverage), :283
    const double v = ((double)c / m) * 254.0;                                      // ImageOptions::MAX_COLOR_VALUE
    return (uint8_t)(__double2int_rz(v) & 0xff);
}

__global__ void polish_emit_kernel(const PolishArgs a, int64_t n_pos, const int32_t* __restrict__ pos_region,
                                   uint8_t* __restrict__ image, int64_t* __restrict__ gpos, int32_t* __restrict__ row_region) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pos) return;
    const int reg = pos_region[i];
    const int64_t pos = a.b.region_ref_start[reg] + (i - a.pos_off[reg]);
    const int64_t row = a.row_of[i];
    const uint32_t cov = a.cov[i];
    for (int j = 0; j < NF; j++) image[row * NF + j] = pixel(a.cnt[i * NF + j], cov);
    gpos[2 * row] = pos; gpos[2 * row + 1] = 0; row_region[row] = reg;
    const int64_t n_ins = (int64_t)a.longest[i];
    for (int64_t ii = 0; ii < n_ins; ii++) {                                       // :289-303
        const int64_t rr = row + 1 + ii;
        for (int j = 0; j < NF; j++) image[rr * NF + j] = pixel(a.ins_cnt[rr * NF + j], cov);
        gpos[2 * rr] = pos; gpos[2 * rr + 1] = ii + 1; row_region[rr] = reg;
    }
}

// chunk c of a region: rows [start, start + size) of the region's row range, zero / (-1, -1) padded (chunk_images)
__global__ void polish_chunk_kernel(const uint8_t* __restrict__ image, const int64_t* __restrict__ gpos,
                                    const int64_t* __restrict__ chunk_row, const int64_t* __restrict__ chunk_rows_valid,
                                    int64_t n_chunks, int chunk, uint8_t* __restrict__ out_img, int64_t* __restrict__ out_pos) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;              // over n_chunks * chunk rows
    if (i >= n_chunks * chunk) return;
    const int64_t c = i / chunk, k = i - c * chunk;
    const bool ok = k < chunk_rows_valid[c];
    const int64_t src = chunk_row[c] + k;
    for (int j = 0; j < NF; j++) out_img[i * NF + j] = ok ? image[src * NF + j] : (uint8_t)0;
    out_pos[2 * i] = ok ? gpos[2 * src] : -1; out_pos[2 * i + 1] = ok ? gpos[2 * src + 1] : -1;
}

}  // namespace

// Workspace: op_ref, op_ri (n_ops each), cnt [P][10], cov [P], longest [P], rows [P+1] x2, pos_region [P], read_region [n_reads],
// pos_off [n_regions+1], scan temp. ins_cnt is carved by the caller-visible second call once the row count is known.
extern "C" int64_t pv_polish_workspace_bytes(int64_t n_reads, int64_t n_ops, int32_t n_regions, int64_t total_positions) {
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const int64_t*)nullptr, (int64_t*)nullptr, (int)(total_positions + 1));
    pv::Arena a(nullptr, 0);
    a.take<int32_t>(n_ops); a.take<int32_t>(n_ops); a.take<uint32_t>(total_positions * NF); a.take<uint32_t>(total_positions);
    a.take<uint32_t>(total_positions); a.take<int64_t>(total_positions + 1); a.take<int64_t>(total_positions + 1);
    a.take<int32_t>(total_positions); a.take<int32_t>(n_reads); a.take<int64_t>(n_regions + 1); a.take<uint8_t>((int64_t)tmp);
    return pv::align_up(a.cur, 256);
}

// Step 1: counts + row layout. *n_rows_host receives the total number of output rows (synchronises the stream once).
// region_rows_host (may be NULL): [n_regions + 1] first row of each region.
extern "C" int pv_polish_count(const PvReadBatch* batch, const int64_t* region_len_host, int64_t total_positions,
                               void* workspace, int64_t workspace_bytes, int64_t* n_rows_host, int64_t* region_rows_host,
                               void* stream_) {
    if (!batch || !region_len_host || !workspace || !n_rows_host) return pv::set_error(PV_EINVAL, "null argument");
    if (int rc = pv::require_device()) return rc;
    const PvReadBatch& b = *batch;
    cudaStream_t st = (cudaStream_t)stream_;
    if (pv_polish_workspace_bytes(b.n_reads, b.n_ops, b.n_regions, total_positions) > workspace_bytes)
        return pv::set_error(PV_EINVAL, "polish workspace too small");
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const int64_t*)nullptr, (int64_t*)nullptr, (int)(total_positions + 1));
    pv::Arena ar(workspace, workspace_bytes);
    int32_t* op_ref = ar.take<int32_t>(b.n_ops); int32_t* op_ri = ar.take<int32_t>(b.n_ops);
    uint32_t* cnt = ar.take<uint32_t>(total_positions * NF); uint32_t* cov = ar.take<uint32_t>(total_positions);
    uint32_t* longest = ar.take<uint32_t>(total_positions);
    int64_t* rows_in = ar.take<int64_t>(total_positions + 1); int64_t* row_of = ar.take<int64_t>(total_positions + 1);
    int32_t* pos_region = ar.take<int32_t>(total_positions); int32_t* read_region = ar.take<int32_t>(b.n_reads);
    int64_t* pos_off = ar.take<int64_t>(b.n_regions + 1); void* scan_tmp = ar.take<uint8_t>((int64_t)tmp);

    std::vector<int64_t> po(b.n_regions + 1, 0);
    std::vector<int32_t> pr((size_t)total_positions);
    for (int32_t r = 0; r < b.n_regions; r++) {
        if (region_len_host[r] <= 0) return pv::set_error(PV_EINVAL, "region %d has length %lld", r, (long long)region_len_host[r]);
        po[r + 1] = po[r] + region_len_host[r];
        if (po[r + 1] > total_positions) return pv::set_error(PV_EINVAL, "total_positions does not match region lengths");
        for (int64_t i = po[r]; i < po[r + 1]; i++) pr[(size_t)i] = r;
    }
    if (po[b.n_regions] != total_positions) return pv::set_error(PV_EINVAL, "total_positions does not match region lengths");
    // read -> region (region_read_begin lives on the device: copy it back once; it is tiny)
    std::vector<int64_t> rb(b.n_regions + 1);
    PV_CUDA_CHECK(cudaMemcpyAsync(rb.data(), b.region_read_begin, rb.size() * 8, cudaMemcpyDeviceToHost, st));
    PV_CUDA_CHECK(cudaStreamSynchronize(st));
    std::vector<int32_t> rr((size_t)b.n_reads);
    for (int32_t r = 0; r < b.n_regions; r++) for (int64_t i = rb[r]; i < rb[r + 1]; i++) rr[(size_t)i] = r;
    PV_CUDA_CHECK(cudaMemcpyAsync(pos_off, po.data(), po.size() * 8, cudaMemcpyHostToDevice, st));
    if (total_positions) PV_CUDA_CHECK(cudaMemcpyAsync(pos_region, pr.data(), pr.size() * 4, cudaMemcpyHostToDevice, st));
    if (b.n_reads) PV_CUDA_CHECK(cudaMemcpyAsync(read_region, rr.data(), rr.size() * 4, cudaMemcpyHostToDevice, st));
    PV_CUDA_CHECK(cudaMemsetAsync(cnt, 0, (size_t)total_positions * NF * 4, st));
    PV_CUDA_CHECK(cudaMemsetAsync(cov, 0, (size_t)total_positions * 4, st));
    PV_CUDA_CHECK(cudaMemsetAsync(longest, 0, (size_t)total_positions * 4, st));

    PolishArgs a;
    a.b = b; a.pos_off = pos_off; a.read_region = read_region; a.op_ref = op_ref; a.op_ri = op_ri; a.cnt = cnt; a.cov = cov;
    a.longest = longest; a.row_of = row_of; a.ins_cnt = nullptr;
    const int sms = pv::sm_count();
    if (b.n_reads > 0) {
        int64_t blocks = (b.n_reads + 7) / 8; if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
        pv::prof_begin(pv::FAM_POLISH, st);
        polish_prefix_kernel<<<(unsigned)blocks, 256, 0, st>>>(b, op_ref, op_ri);
        polish_walk_kernel<0><<<(unsigned)blocks, 256, 0, st>>>(a);
        PV_CUDA_CHECK(cudaGetLastError());
        pv::prof_end(pv::FAM_POLISH, st, 2);
    }
    pv::prof_begin(pv::FAM_POLISH, st);
    polish_rows_kernel<<<(unsigned)((total_positions + 256) / 256), 256, 0, st>>>(longest, total_positions, rows_in);
    PV_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(scan_tmp, tmp, rows_in, row_of, (int)(total_positions + 1), st));
    pv::prof_end(pv::FAM_POLISH, st, 2);
    PV_CUDA_CHECK(cudaMemcpyAsync(n_rows_host, row_of + total_positions, 8, cudaMemcpyDeviceToHost, st));
    if (region_rows_host)
        for (int32_t r = 0; r <= b.n_regions; r++)
            PV_CUDA_CHECK(cudaMemcpyAsync(region_rows_host + r, row_of + po[r], 8, cudaMemcpyDeviceToHost, st));
    PV_CUDA_CHECK(cudaStreamSynchronize(st));
    return PV_OK;
}

// Step 2: inserted columns + normalisation. image_dev uint8 [n_rows][10], gpos_dev int64 [n_rows][2], row_region_dev
// int32 [n_rows]; ins_scratch_dev uint32 [n_rows][10].
extern "C" int pv_polish_emit(const PvReadBatch* batch, const int64_t* region_len_host, int64_t total_positions,
                              void* workspace, int64_t workspace_bytes, int64_t n_rows, uint32_t* ins_scratch_dev,
                              uint8_t* image_dev, int64_t* gpos_dev, int32_t* row_region_dev, void* stream_) {
    if (!batch || !workspace || (n_rows > 0 && (!ins_scratch_dev || !image_dev || !gpos_dev || !row_region_dev)))
        return pv::set_error(PV_EINVAL, "null argument");
    (void)region_len_host;
    const PvReadBatch& b = *batch;
    cudaStream_t st = (cudaStream_t)stream_;
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const int64_t*)nullptr, (int64_t*)nullptr, (int)(total_positions + 1));
    pv::Arena ar(workspace, workspace_bytes);
    int32_t* op_ref = ar.take<int32_t>(b.n_ops); int32_t* op_ri = ar.take<int32_t>(b.n_ops);
    uint32_t* cnt = ar.take<uint32_t>(total_positions * NF); uint32_t* cov = ar.take<uint32_t>(total_positions);
    uint32_t* longest = ar.take<uint32_t>(total_positions);
    ar.take<int64_t>(total_positions + 1); int64_t* row_of = ar.take<int64_t>(total_positions + 1);
    int32_t* pos_region = ar.take<int32_t>(total_positions); int32_t* read_region = ar.take<int32_t>(b.n_reads);
    int64_t* pos_off = ar.take<int64_t>(b.n_regions + 1);
    if (n_rows <= 0) return PV_OK;
    PV_CUDA_CHECK(cudaMemsetAsync(ins_scratch_dev, 0, (size_t)n_rows * NF * 4, st));
    PolishArgs a;
    a.b = b; a.pos_off = pos_off; a.read_region = read_region; a.op_ref = op_ref; a.op_ri = op_ri; a.cnt = cnt; a.cov = cov;
    a.longest = longest; a.row_of = row_of; a.ins_cnt = ins_scratch_dev;
    const int sms = pv::sm_count();
    pv::prof_begin(pv::FAM_POLISH, st);
    if (b.n_reads > 0) {
        int64_t blocks = (b.n_reads + 7) / 8; if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
        polish_walk_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(a);
    }
    polish_emit_kernel<<<(unsigned)((total_positions + 255) / 256), 256, 0, st>>>(a, total_positions, pos_region, image_dev, gpos_dev, row_region_dev);
    PV_CUDA_CHECK(cudaGetLastError());
    pv::prof_end(pv::FAM_POLISH, st, 2);
    return PV_OK;
}

// chunk_images (AlignmentSummarizer.py:19-56): chunk_row_dev / chunk_valid_dev int64 [n_chunks] = first source row and
// number of real rows of each chunk (the plan is a few integers per region, computed by the host wrapper)
extern "C" int pv_polish_chunks(const uint8_t* image_dev, const int64_t* gpos_dev, const int64_t* chunk_row_dev,
                                const int64_t* chunk_valid_dev, int64_t n_chunks, int32_t chunk_size, uint8_t* out_images_dev,
                                int64_t* out_positions_dev, void* stream_) {
    if (n_chunks <= 0) return PV_OK;
    if (!image_dev || !gpos_dev || !chunk_row_dev || !chunk_valid_dev || !out_images_dev || !out_positions_dev || chunk_size <= 0)
        return pv::set_error(PV_EINVAL, "bad argument");
    cudaStream_t st = (cudaStream_t)stream_;
    const int64_t n = n_chunks * chunk_size;
    pv::prof_begin(pv::FAM_POLISH, st);
    polish_chunk_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(image_dev, gpos_dev, chunk_row_dev, chunk_valid_dev, n_chunks,
                                                                      chunk_size, out_images_dev, out_positions_dev);
    PV_CUDA_CHECK(cudaGetLastError());
    pv::prof_end(pv::FAM_POLISH, st, 1);
    return PV_OK;
}
