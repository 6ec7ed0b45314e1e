// Polisher pileup summary on sm_100a: SummaryGenerator::generate_summary of the reference's `pepper` module
// (/root/reference/pepper/modules/src/pileup_summary/summary_generator.cpp:47-121, 274-306, 370-392) -- the producer
// of model M-B's input: per reference position 10 normalised uint8 features (A C G T reverse, A C G T forward,
// other reverse, other forward), followed by one row per inserted column up to the longest insert observed at that
// position -- and the 1000/50 chunking of AlignmentSummarizer.chunk_images
// (/root/reference/pepper/modules/python/AlignmentSummarizer.py:19-56).
//
// Kernels (one stream, no host synchronisation except reading back the row count):
//   P0 polish_prefix_kernel   warp per read: CIGAR prefix (here REF_SKIP / PAD advance only the reference, :98-113)
//   P1 polish_tile_kernel     CTA per tile of 1536 positions, counters in shared memory: base / deletion counts, coverage
//                             (incl. the reference's quirk of charging a deletion's coverage to the op's first position,
//                             :108), longest insert
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

// P2: counts of the inserted columns (warp per read, lane per insert op). Inserted bases are ~1.5 % of the bases at ONT
// error rates and their rows are spread over the whole output, so this stays on global atomics.
__global__ void polish_insert_kernel(const PolishArgs a) {
    const PvReadBatch& b = a.b;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < b.n_reads; r += n_warps) {
        if (b.read_mapq[r] == 0) continue;                                          // :376
        int reg = 0;                                                                // last region whose first read is <= r
        { int hi = b.n_regions; while (hi - reg > 1) { const int mid = (reg + hi) >> 1; if (b.region_read_begin[mid] <= r) reg = mid; else hi = mid; } }
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
            if ((w & 15u) != 1u) continue;                                          // :79-96
            const int64_t len = (int64_t)(w >> 4);
            const int64_t p0 = rel + a.op_ref[co + k];                              // region-local reference position
            const int64_t ri = a.op_ri[co + k];
            if (p0 > L - 1) continue;                                               // `if (ref_position > region_end) break`, :55
            const int64_t an = p0 - 1;
            if (an < 0 || an > L - 1 || ri > read_len) continue;                    // substr(read_index > size) throws in the reference
            int64_t alen = len; if (alen > read_len - ri) alen = read_len - ri;     // substr truncation
            const int64_t row = a.row_of[g0 + an] + 1;
            for (int64_t i = 0; i < alen; i++)
                atomicAdd(&a.ins_cnt[(row + i) * NF + feature_index(b.bases[bo + ri + i], rev)], 1u);
        }
    }
}

// ---- P1 as a tile kernel ---------------------------------------------------------------------------------------------
// CTA per tile of PT_P positions of one region; the ten counters of a position live in SHARED memory (position-major,
// so the flush is a straight coalesced copy). A warp takes a read that touches the tile, finds its first op with a
// binary search in the CIGAR prefix, walks the ops lane-per-op (deletions, insert lengths, coverage quirk) and files the
// clipped match runs in a per-warp table; the table is then expanded lane-per-BASE (coalesced base loads, one shared
// atomic per base). Coverage is not counted per base: cov = sum of the ten counters - deleted bases + the reference's
// deletion coverage charged to the op's first position (:108).
constexpr int PT_THREADS = 512;
constexpr int PT_WARPS = PT_THREADS / 32;
constexpr int PT_P = 1536;
constexpr int PT_TBL = 96;
constexpr int PT_LIST = 1024;
struct PtTable { int ri[PT_TBL]; int p0[PT_TBL]; int pre[PT_TBL + 1]; };
constexpr size_t PT_SMEM = (size_t)PT_P * NF * 4 + 3 * (size_t)PT_P * 4 + PT_WARPS * sizeof(PtTable) + PT_LIST * 4 + 64;

__global__ void __launch_bounds__(PT_THREADS, 2) polish_tile_kernel(const PolishArgs a) {
    extern __shared__ __align__(16) uint8_t pt_smem[];
    uint32_t* cnt = (uint32_t*)pt_smem;                       // [PT_P][NF]
    uint32_t* dels = cnt + PT_P * NF;                         // [PT_P] deleted bases ('*' counts that are no coverage)
    uint32_t* delcov = dels + PT_P;                           // [PT_P] coverage charged to a deletion's first position
    uint32_t* longest = delcov + PT_P;                        // [PT_P]
    PtTable* tables = (PtTable*)(longest + PT_P);
    int32_t* s_list = (int32_t*)(tables + PT_WARPS);
    int* s_n = s_list + PT_LIST; int* s_next = s_n + 1;

    const PvReadBatch& b = a.b;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int reg = blockIdx.y;
    const int64_t rs = b.region_ref_start[reg];
    const int64_t L = b.region_ref_end[reg] - rs + 1;
    const int64_t t_lo = (int64_t)blockIdx.x * PT_P;
    if (t_lo >= L) return;
    const int nv = (int)(L - t_lo < PT_P ? L - t_lo : PT_P);
    const int64_t t_hi = t_lo + nv - 1;
    const int64_t g0 = a.pos_off[reg] + t_lo;
    PtTable& tb = tables[warp];

    for (int i = tid; i < PT_P * (NF + 3); i += PT_THREADS) cnt[i] = 0;       // cnt, dels, delcov, longest are contiguous
    const int64_t rb = b.region_read_begin[reg], re = b.region_read_begin[reg + 1];
    for (int64_t base = rb; base < re; base += PT_LIST) {
        __syncthreads();
        if (tid == 0) { *s_n = 0; *s_next = 0; }
        __syncthreads();
        const int64_t end = base + PT_LIST < re ? base + PT_LIST : re;
        for (int64_t r = base + tid; r < end; r += PT_THREADS) {
            if (b.read_mapq[r] == 0) continue;                                      // :376
            const int n_ops = b.read_n_ops[r];
            if (n_ops <= 0) continue;
            const int64_t co = b.read_cigar_off[r];
            const uint32_t wl = b.cigar[co + n_ops - 1];
            const int opl = (int)(wl & 15u);
            const int64_t span = (int64_t)a.op_ref[co + n_ops - 1] + ((is_m(opl) || opl == 2 || opl == 3 || opl == 6) ? (int64_t)(wl >> 4) : 0);
            const int64_t rel = b.read_pos[r] - rs;
            if (rel - 1 > t_hi || rel + span < t_lo) continue;                      // touches [rel - 1, rel + span - 1]
            s_list[atomicAdd(s_n, 1)] = (int32_t)(r - base);
        }
        __syncthreads();
        const int n_list = *s_n;
        while (true) {
            int it = 0;
            if (lane == 0) it = atomicAdd(s_next, 1);
            it = __shfl_sync(0xffffffffu, it, 0);
            if (it >= n_list) break;
            const int64_t r = base + s_list[it];
            const int64_t rel = b.read_pos[r] - rs;
            const int64_t co = b.read_cigar_off[r], bo = b.read_base_off[r];
            const int n_ops = b.read_n_ops[r], read_len = b.read_len[r];
            const bool rev = b.read_flags[r] & 1;
            const int star = feature_index('*', rev);
            // first op that can touch the tile: one before the first op starting at or behind t_lo
            int lo = 0, hi = n_ops;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (rel + (int64_t)a.op_ref[co + mid] >= t_lo) hi = mid; else lo = mid + 1; }
            int k_lo = lo - 1; if (k_lo < 0) k_lo = 0;
            int n_tab = 0, n_base = 0;
            for (int k0 = k_lo; k0 < n_ops; k0 += 32) {
                const int k = k0 + lane;
                const bool have = k < n_ops;
                const uint32_t w = have ? b.cigar[co + k] : 0u;
                const int op = (int)(w & 15u);
                const int64_t len = (int64_t)(w >> 4);
                const int64_t p0 = have ? rel + a.op_ref[co + k] : (int64_t)1 << 40;  // region-local reference position
                const int64_t ri = have ? a.op_ri[co + k] : 0;
                const bool stop = __any_sync(0xffffffffu, have && p0 > t_hi + 1) || k0 + 32 >= n_ops;
                int m_cnt = 0, m_ri = 0, m_p = 0;
                if (have && p0 <= L - 1 && p0 <= t_hi + 1) {                        // `if (ref_position > region_end) break`, :55
                    if (is_m(op)) {                                                 // :57-78
                        int64_t i_lo = p0 < 0 ? -p0 : 0;
                        int64_t i_hi = len; if (i_hi > L - p0) i_hi = L - p0; if (i_hi > read_len - ri) i_hi = read_len - ri;
                        if (i_lo < t_lo - p0) i_lo = t_lo - p0;
                        if (i_hi > t_hi + 1 - p0) i_hi = t_hi + 1 - p0;
                        if (i_hi > i_lo) { m_cnt = (int)(i_hi - i_lo); m_ri = (int)(ri + i_lo); m_p = (int)(p0 + i_lo - t_lo); }
                    } else if (op == 1) {                                           // :79-96
                        const int64_t an = p0 - 1;
                        if (an >= t_lo && an <= t_hi && an <= L - 1 && ri <= read_len) {
                            int64_t alen = len; if (alen > read_len - ri) alen = read_len - ri;
                            if (alen > 0) atomicMax(&longest[an - t_lo], (uint32_t)alen);
                        }
                    } else if (op == 2 || op == 3 || op == 6) {                     // :97-113
                        const int64_t r_lo = p0 < 0 ? -p0 : 0;
                        int64_t r_hi = len; if (r_hi > L - p0) r_hi = L - p0;
                        int64_t i_lo = r_lo < t_lo - p0 ? t_lo - p0 : r_lo;
                        int64_t i_hi = r_hi > t_hi + 1 - p0 ? t_hi + 1 - p0 : r_hi;
                        for (int64_t i = i_lo; i < i_hi; i++) {
                            atomicAdd(&cnt[(p0 + i - t_lo) * NF + star], 1u);
                            atomicAdd(&dels[p0 + i - t_lo], 1u);
                        }
                        // the reference adds this coverage at ref_position (the op's start), once per in-range deleted base
                        if (r_hi > r_lo && p0 >= t_lo && p0 <= t_hi) atomicAdd(&delcov[p0 - t_lo], (uint32_t)(r_hi - r_lo));
                    }
                }
                // file the match runs of these 32 ops with the running number of bases in front of each
                int incl = m_cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                const unsigned nz = __ballot_sync(0xffffffffu, m_cnt > 0);
                if (m_cnt > 0) {
                    const int at = n_tab + __popc(nz & ((1u << lane) - 1u));
                    tb.ri[at] = m_ri; tb.p0[at] = m_p; tb.pre[at] = n_base + incl - m_cnt;
                }
                n_tab += __popc(nz);
                n_base += __shfl_sync(0xffffffffu, incl, 31);
                if (n_tab > PT_TBL - 32 || stop) {
                    if (lane == 0) tb.pre[n_tab] = n_base;
                    __syncwarp();
                    for (int t = lane; t < n_base; t += 32) {                   // lane per base
                        int jl = 0, jh = n_tab;                                  // last piece with pre[j] <= t
                        while (jh - jl > 1) { const int mid = (jl + jh) >> 1; if (tb.pre[mid] <= t) jl = mid; else jh = mid; }
                        const int off = t - tb.pre[jl];
                        atomicAdd(&cnt[(tb.p0[jl] + off) * NF + feature_index(b.bases[bo + tb.ri[jl] + off], rev)], 1u);
                    }
                    __syncwarp();
                    n_tab = 0; n_base = 0;
                }
                if (stop) break;
            }
        }
    }
    __syncthreads();
    // flush: the ten counters (a straight copy), coverage, longest insert
    for (int i = tid; i < nv * NF; i += PT_THREADS) a.cnt[g0 * NF + i] = cnt[i];
    for (int i = tid; i < nv; i += PT_THREADS) {
        uint32_t s = 0;
#pragma unroll
        for (int f = 0; f < NF; f++) s += cnt[i * NF + f];
        a.cov[g0 + i] = s - dels[i] + delcov[i];
        a.longest[g0 + i] = longest[i];
    }
}

__global__ void polish_rows_kernel(const uint32_t* __restrict__ longest, int64_t n, int64_t* __restrict__ rows) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) rows[i] = i < n ? 1 + (int64_t)longest[i] : 0;
}

__global__ void polish_region_rows_kernel(const int64_t* __restrict__ row_of, const int64_t* __restrict__ pos_off, int n,
                                          int64_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = row_of[pos_off[i]];
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
    (void)pos_region;
    int reg = 0;                                                                    // last region whose first position is <= i
    { int hi = a.b.n_regions; while (hi - reg > 1) { const int mid = (reg + hi) >> 1; if (a.pos_off[mid] <= i) reg = mid; else hi = mid; } }
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
    for (int32_t r = 0; r < b.n_regions; r++) {
        if (region_len_host[r] <= 0) return pv::set_error(PV_EINVAL, "region %d has length %lld", r, (long long)region_len_host[r]);
        po[r + 1] = po[r] + region_len_host[r];
    }
    if (po[b.n_regions] != total_positions) return pv::set_error(PV_EINVAL, "total_positions does not match region lengths");
    // the region of a read / of a position is found on the device (binary search in region_read_begin / pos_off): no
    // per-position table is built or uploaded
    (void)pos_region; (void)read_region;
    PV_CUDA_CHECK(cudaMemcpyAsync(pos_off, po.data(), po.size() * 8, cudaMemcpyHostToDevice, st));
    // cnt / cov / longest need no clearing: every position is written by the flush of its tile

    PolishArgs a;
    a.b = b; a.pos_off = pos_off; a.read_region = read_region; a.op_ref = op_ref; a.op_ri = op_ri; a.cnt = cnt; a.cov = cov;
    a.longest = longest; a.row_of = row_of; a.ins_cnt = nullptr;
    const int sms = pv::sm_count();
    if (b.n_reads > 0) {
        int64_t blocks = (b.n_reads + 7) / 8; if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
        pv::prof_begin(pv::FAM_POLISH, st);
        polish_prefix_kernel<<<(unsigned)blocks, 256, 0, st>>>(b, op_ref, op_ri);
        PV_CUDA_CHECK(cudaGetLastError());
        pv::prof_end(pv::FAM_POLISH, st, 1);
    }
    if (total_positions > 0) {
        int64_t max_len = 0;
        for (int32_t r = 0; r < b.n_regions; r++) if (region_len_host[r] > max_len) max_len = region_len_host[r];
        static bool attr_done = false;
        if (!attr_done) { PV_CUDA_CHECK(cudaFuncSetAttribute(polish_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PT_SMEM)); attr_done = true; }
        const dim3 grid((unsigned)((max_len + PT_P - 1) / PT_P), (unsigned)b.n_regions);
        pv::prof_begin(pv::FAM_POLISH, st);
        polish_tile_kernel<<<grid, PT_THREADS, PT_SMEM, st>>>(a);
        PV_CUDA_CHECK(cudaGetLastError());
        pv::prof_end(pv::FAM_POLISH, st, 1);
    }
    pv::prof_begin(pv::FAM_POLISH, st);
    polish_rows_kernel<<<(unsigned)((total_positions + 256) / 256), 256, 0, st>>>(longest, total_positions, rows_in);
    PV_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(scan_tmp, tmp, rows_in, row_of, (int)(total_positions + 1), st));
    pv::prof_end(pv::FAM_POLISH, st, 2);
    // first row of every region (and the total behind the last one): gathered on the device, one copy back
    std::vector<int64_t> rr_host((size_t)b.n_regions + 1, 0);
    polish_region_rows_kernel<<<(unsigned)((b.n_regions + 1 + 255) / 256), 256, 0, st>>>(row_of, pos_off, b.n_regions + 1, rows_in);
    PV_CUDA_CHECK(cudaGetLastError());
    PV_CUDA_CHECK(cudaMemcpyAsync(rr_host.data(), rows_in, rr_host.size() * 8, cudaMemcpyDeviceToHost, st));
    PV_CUDA_CHECK(cudaStreamSynchronize(st));
    *n_rows_host = rr_host[(size_t)b.n_regions];
    if (region_rows_host) memcpy(region_rows_host, rr_host.data(), rr_host.size() * 8);
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
        polish_insert_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
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
