// Reference-predicted bases: the "bases_ref" wire form of a host batch (include/pepper_b200.h).
//
// The host path is bound by the PCIe upload, and most of what it uploads is redundant with data the device already
// holds: an aligned read base nearly always equals the reference base it is aligned to. Every base of a read is
// therefore PREDICTED on the device from the batch's own CIGAR and reference --
//     M / = / X base at an aligned position inside the region's reference_sequence  ->  that reference byte
//     anything else (inserted, soft-clipped, outside the reference, behind the last op)  ->  'A'
// -- and only the bases that differ from their prediction travel, as 16-bit patch entries per read (low byte s:
// 0..254 = the patched base sits s bases behind the cursor, which then moves behind it; 255 = the cursor moves 255 bases,
// no patch; high byte = the base). Any byte value survives (lossless); at ONT error rates the bases cost ~0.5 bits each
// on the wire instead of 2.
//
// The prediction rule is plain BAM semantics (M/=/X consume both, I/S the read, D/N the reference) -- it only has to
// be the SAME in pv_pack_bases_ref (host) and pv_unpack_bases_ref (device); it is not the reference's CIGAR walk.
#include "common.cuh"
#include <atomic>
#include <cmath>
#include <mutex>
#include <thread>
#include <vector>

namespace {

__host__ __device__ __forceinline__ bool wr_match(int op) { return op == 0 || op == 7 || op == 8; }
__host__ __device__ __forceinline__ int wr_read_adv(int op, int len) { return (wr_match(op) || op == 1 || op == 4) ? len : 0; }
__host__ __device__ __forceinline__ int wr_ref_adv(int op, int len) { return (wr_match(op) || op == 2 || op == 3) ? len : 0; }

// ---- host: prediction of one read, calling f(index, predicted byte) for index 0 .. read_len-1 in order ------------------
template <class F>
void predict_read_host(const PvReadBatch& b, int64_t r, int32_t region, F f) {
    const int64_t ref_off = b.region_ref_off[region], ref_len = b.region_ref_len[region];
    const int64_t rel = b.read_pos[r] - b.region_ref_start[region];
    const int64_t co = b.read_cigar_off[r];
    const int n_ops = b.read_n_ops[r], read_len = b.read_len[r];
    int64_t ri = 0, rp = rel;
    for (int k = 0; k < n_ops && ri < read_len; k++) {
        const uint32_t w = b.cigar[co + k];
        const int op = (int)(w & 15u), len = (int)(w >> 4);
        const int qa = wr_read_adv(op, len);
        if (qa) {
            const bool m = wr_match(op);
            for (int j = 0; j < qa && ri + j < read_len; j++) {
                const int64_t p = rp + j;
                f(ri + j, (m && p >= 0 && p < ref_len) ? b.ref[ref_off + p] : (uint8_t)'A');
            }
        }
        ri += qa; rp += wr_ref_adv(op, len);
    }
    for (int64_t i = ri < 0 ? 0 : ri; i < read_len; i++) f(i, (uint8_t)'A');
}

// ---- device ------------------------------------------------------------------------------------------------------
// Warp per read. The ops that consume read bases are filed in a per-warp shared-memory table (first read index, aligned
// reference position or "no reference"); every time the table fills (and at the end of the read) the lanes expand it
// OUTPUT-major: each lane produces one aligned 16-byte block of the read (finding its first op by binary search, then
// walking the table) and stores it with one 16-byte store, so the 1 byte per base this kernel writes leaves coalesced.
constexpr int WR_TBL = 256;
constexpr int WR_WARPS = 8;
constexpr long long WR_NOREF = -(1ll << 62);

struct WrTable { int ri[WR_TBL + 1]; long long rp[WR_TBL]; };

__device__ __forceinline__ uint8_t wr_byte_at(const WrTable& t, int& j, int idx, const uint8_t* __restrict__ ref, int64_t ref_len) {
    while (idx >= t.ri[j + 1]) j++;                       // idx < t.ri[n]: terminates inside the table
    const long long p = t.rp[j] + (idx - t.ri[j]);
    return (p >= 0 && p < ref_len) ? __ldg(ref + p) : (uint8_t)'A';
}
__device__ __forceinline__ int wr_find(const WrTable& t, int n, int idx) {      // last j with t.ri[j] <= idx
    int lo = 0, hi = n;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (t.ri[mid] <= idx) lo = mid; else hi = mid; }
    return lo;
}

__device__ __forceinline__ unsigned long long wr_bytes_below(int x) {      // mask of bytes 0 .. x-1 of a 64-bit word, x in [0, 8]
    return x >= 8 ? ~0ull : ((1ull << (8 * x)) - 1ull);
}

// read indices [a, b) from the table of n entries (t.ri[n] >= b). Aligned 16-byte blocks are produced run by run: a run
// (the part of one op inside the block) is 16 unaligned reference bytes fetched as five aligned words + funnel shifts,
// masked to the run's bytes; a lane takes two neighbouring blocks so it searches its first op once per 32 bytes.
__device__ void wr_expand(const WrTable& t, int n, int a, int b, const uint8_t* __restrict__ ref, int64_t ref_len,
                          uint8_t* __restrict__ out, int lane) {
    if (b <= a) return;
    int blk0 = (a + 15) & ~15;
    if (blk0 > b) blk0 = b;
    if (lane < blk0 - a) { int j = wr_find(t, n, a + lane); out[a + lane] = wr_byte_at(t, j, a + lane, ref, ref_len); }     // head
    const int blk1 = blk0 + ((b - blk0) & ~15);
    for (int s0 = blk0 + lane * 32; s0 < blk1; s0 += 32 * 32) {
        int j = wr_find(t, n, s0);
        int cur = t.ri[j], nxt = t.ri[j + 1];                  // the op that holds the next byte, kept in registers
        long long rp = t.rp[j];
#pragma unroll 1
        for (int blk = s0; blk < s0 + 32 && blk < blk1; blk += 16) {
            unsigned long long lo = 0ull, hi = 0ull;           // bytes 0..7 and 8..15 of the block
            int e = 0;
            while (e < 16) {
                const int idx = blk + e;
                while (idx >= nxt) { j++; cur = nxt; nxt = t.ri[j + 1]; rp = t.rp[j]; }
                int run = nxt - idx; if (run > 16 - e) run = 16 - e;
                const long long q = rp + (idx - cur) - e;      // reference position that lines up with byte 0 of the block
                unsigned long long v_lo, v_hi;
                if (q >= 0 && q + 20 <= ref_len) {
                    const uintptr_t addr = (uintptr_t)(ref + q);
                    const uint32_t* src = (const uint32_t*)(addr & ~(uintptr_t)3);
                    const int sh = (int)(addr & 3) * 8;
                    const uint32_t r0 = __ldg(src), r1 = __ldg(src + 1), r2 = __ldg(src + 2), r3 = __ldg(src + 3), r4 = __ldg(src + 4);
                    v_lo = (unsigned long long)__funnelshift_r(r0, r1, sh) | ((unsigned long long)__funnelshift_r(r1, r2, sh) << 32);
                    v_hi = (unsigned long long)__funnelshift_r(r2, r3, sh) | ((unsigned long long)__funnelshift_r(r3, r4, sh) << 32);
                } else if (rp < -(1ll << 61)) {                // no reference behind this op (insert, soft clip): 'A'
                    v_lo = v_hi = 0x4141414141414141ull;
                } else {                                       // a run at the edge of the region's reference: byte by byte
                    v_lo = v_hi = 0ull;
                    for (int u = e; u < e + run; u++) {
                        const long long pos = q + u;
                        const unsigned long long c = (pos >= 0 && pos < ref_len) ? __ldg(ref + pos) : (uint8_t)'A';
                        if (u < 8) v_lo |= c << (8 * u); else v_hi |= c << (8 * (u - 8));
                    }
                }
                const int e1 = e + run;
                const unsigned long long m_lo = wr_bytes_below(e1 < 8 ? e1 : 8) & ~wr_bytes_below(e < 8 ? e : 8);
                const unsigned long long m_hi = wr_bytes_below(e1 > 8 ? e1 - 8 : 0) & ~wr_bytes_below(e > 8 ? e - 8 : 0);
                lo |= v_lo & m_lo; hi |= v_hi & m_hi;
                e = e1;
            }
            *(uint4*)(out + blk) = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
        }
    }
    if (lane < b - blk1) { int j = wr_find(t, n, blk1 + lane); out[blk1 + lane] = wr_byte_at(t, j, blk1 + lane, ref, ref_len); }   // tail
}

// Read lengths are heavy-tailed (1 .. 100 kbp): warps take reads from a ticket counter instead of a fixed stride, so
// the kernel ends one read after the average warp is done, not after the unluckiest fixed assignment.
__global__ void __launch_bounds__(WR_WARPS * 32) predict_bases_kernel(const PvReadBatch b, uint8_t* __restrict__ bases,
                                                                       unsigned long long* __restrict__ ticket) {
    __shared__ WrTable s_tab[WR_WARPS];
    WrTable& t = s_tab[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    while (true) {
        unsigned long long tk = 0;
        if (lane == 0) tk = atomicAdd(ticket, 1ull);
        const int64_t r = (int64_t)__shfl_sync(0xffffffffu, tk, 0);
        if (r >= b.n_reads) break;
        // region of the read: last region whose first read is <= r
        int lo = 0, hi = b.n_regions;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (b.region_read_begin[mid] <= r) lo = mid; else hi = mid; }
        const uint8_t* ref = b.ref + b.region_ref_off[lo];
        const int64_t ref_len = b.region_ref_len[lo];
        const int64_t rel = b.read_pos[r] - b.region_ref_start[lo];
        const int64_t co = b.read_cigar_off[r];
        const int n_ops = b.read_n_ops[r], read_len = b.read_len[r];
        uint8_t* out = bases + b.read_base_off[r];             // 16-byte aligned
        int ri_run = 0; long long rp_run = rel;
        int n_tab = 0, tab_a = 0;
        for (int k0 = 0; k0 < n_ops && ri_run < read_len; k0 += 32) {
            const int k = k0 + lane;
            const uint32_t w = k < n_ops ? b.cigar[co + k] : 0u;
            const int op = (int)(w & 15u), len = (int)(w >> 4);
            // a (malformed) op longer than the read cannot write more than the read's bases: clamp so the sums stay in range
            int qa = k < n_ops ? wr_read_adv(op, len) : 0;
            if (qa > read_len + 1) qa = read_len + 1;
            const long long ra = k < n_ops ? wr_ref_adv(op, len) : 0;
            int iq = qa; long long ir = ra;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int tq = __shfl_up_sync(0xffffffffu, iq, d); const long long tr = __shfl_up_sync(0xffffffffu, ir, d);
                if (lane >= d) { iq += tq; ir += tr; }
            }
            const unsigned has = __ballot_sync(0xffffffffu, qa > 0);
            if (qa > 0) {
                const int at = n_tab + __popc(has & ((1u << lane) - 1u));
                t.ri[at] = ri_run + iq - qa;
                t.rp[at] = wr_match(op) ? rp_run + ir - ra : WR_NOREF;
            }
            n_tab += __popc(has);
            ri_run += __shfl_sync(0xffffffffu, iq, 31);
            rp_run += __shfl_sync(0xffffffffu, ir, 31);
            const bool last = k0 + 32 >= n_ops || ri_run >= read_len;
            if (n_tab > WR_TBL - 32 || last) {
                if (lane == 0) t.ri[n_tab] = ri_run;
                __syncwarp();
                wr_expand(t, n_tab, tab_a, ri_run < read_len ? ri_run : read_len, ref, ref_len, out, lane);
                __syncwarp();
                n_tab = 0; tab_a = ri_run;
            }
        }
        // behind the last op: 'A' up to the read's end, 0 up to the 16-byte boundary (what pack_regions / synth pad with)
        const int pad_end = ((read_len + 15) & ~15);
        for (int i = (ri_run < read_len ? ri_run : read_len) + lane; i < pad_end; i += 32)
            if (b.read_base_off[r] + i < b.n_bases) out[i] = i < read_len ? (uint8_t)'A' : (uint8_t)0;
    }
}

__global__ void apply_patches_kernel(const PvReadBatch b, const int64_t* __restrict__ patch_off, const uint16_t* __restrict__ patches,
                                     uint8_t* __restrict__ bases) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < b.n_reads; r += n_warps) {
        const int64_t e0 = patch_off[r], e1 = patch_off[r + 1];
        const int read_len = b.read_len[r];
        uint8_t* out = bases + b.read_base_off[r];
        int cursor = 0;
        for (int64_t e = e0; e < e1; e += 32) {
            const bool have = e + lane < e1;
            const uint32_t w = have ? patches[e + lane] : 0u;
            const int s = (int)(w & 0xffu);
            const int adv = have ? (s == 255 ? 255 : s + 1) : 0;
            int inc = adv;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            const int at = cursor + inc - 1;
            if (have && s != 255 && at < read_len) out[at] = (uint8_t)(w >> 8);
            cursor += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
}

template <class F> void wr_parallel(int64_t n, int threads, F f) {
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    std::vector<std::thread> pool;
    const int64_t per = (n + threads - 1) / threads;
    for (int t = 0; t < threads; t++) pool.emplace_back([=]() { const int64_t lo = t * per, hi = lo + per < n ? lo + per : n; if (lo < hi) f(lo, hi); });
    for (auto& th : pool) th.join();
}

}  // namespace

extern "C" int pv_pack_bases_ref(const PvReadBatch* hb, int64_t* read_patch_off, uint16_t* patches, int64_t patch_capacity,
                                 int32_t threads) {
    if (!hb || !read_patch_off) return pv::set_error(PV_EINVAL, "pv_pack_bases_ref: null argument");
    const PvReadBatch& b = *hb;
    if (b.n_reads && (!b.bases || !b.cigar || !b.ref)) return pv::set_error(PV_EINVAL, "pv_pack_bases_ref needs the plain bases, cigar and ref arrays");
    std::vector<int32_t> region_of((size_t)b.n_reads, 0);
    for (int32_t g = 0; g < b.n_regions; g++)
        for (int64_t r = b.region_read_begin[g]; r < b.region_read_begin[g + 1]; r++) region_of[(size_t)r] = g;
    if (!patches) {
        // pass 1: entries per read -> exclusive prefix in read_patch_off[0 .. n_reads]
        wr_parallel(b.n_reads, threads, [&](int64_t lo, int64_t hi) {
            for (int64_t r = lo; r < hi; r++) {
                const uint8_t* seq = b.bases + b.read_base_off[r];
                int64_t n = 0, cursor = 0;
                predict_read_host(b, r, region_of[(size_t)r], [&](int64_t i, uint8_t pred) {
                    if (seq[i] != pred) { int64_t gap = i - cursor; n += gap / 255 + 1; cursor = i + 1; }
                });
                read_patch_off[r + 1] = n;
            }
        });
        read_patch_off[0] = 0;
        for (int64_t r = 0; r < b.n_reads; r++) read_patch_off[r + 1] += read_patch_off[r];
        return PV_OK;
    }
    if (read_patch_off[b.n_reads] > patch_capacity)
        return pv::set_error(PV_EINVAL, "pv_pack_bases_ref: %lld patch entries needed, buffer holds %lld", (long long)read_patch_off[b.n_reads], (long long)patch_capacity);
    wr_parallel(b.n_reads, threads, [&](int64_t lo, int64_t hi) {
        for (int64_t r = lo; r < hi; r++) {
            const uint8_t* seq = b.bases + b.read_base_off[r];
            uint16_t* out = patches + read_patch_off[r];
            int64_t cursor = 0;
            predict_read_host(b, r, region_of[(size_t)r], [&](int64_t i, uint8_t pred) {
                if (seq[i] == pred) return;
                int64_t gap = i - cursor;
                while (gap >= 255) { *out++ = 255; gap -= 255; }
                *out++ = (uint16_t)(gap | ((uint16_t)seq[i] << 8));
                cursor = i + 1;
            });
        }
    });
    return PV_OK;
}

extern "C" int pv_unpack_bases_ref(const PvReadBatch* dev_batch, const int64_t* read_patch_off_dev, const uint16_t* patches_dev,
                                   uint8_t* bases_dev, void* stream) {
    if (!dev_batch || !bases_dev) return pv::set_error(PV_EINVAL, "pv_unpack_bases_ref: null argument");
    const PvReadBatch& b = *dev_batch;
    if (b.n_reads == 0) return PV_OK;
    if (!read_patch_off_dev || !b.cigar || !b.ref) return pv::set_error(PV_EINVAL, "pv_unpack_bases_ref needs cigar, ref and the patch offsets on the device");
    if (int rc = pv::require_device()) return rc;
    int64_t blocks = (b.n_reads + 7) / 8;
    const int64_t cap = (int64_t)pv::sm_count() * 8;
    if (blocks > cap) blocks = cap;
    // one ticket counter per launch, from a small ring (launches on different streams may overlap)
    static unsigned long long* ring = nullptr;
    static std::atomic<unsigned> next{0};
    constexpr unsigned RING = 64;
    if (!ring) {
        static std::mutex mu;
        std::lock_guard<std::mutex> l(mu);
        if (!ring) { unsigned long long* p = nullptr; PV_CUDA_CHECK(cudaMalloc((void**)&p, RING * sizeof(unsigned long long))); ring = p; }
    }
    unsigned long long* ticket = ring + (next.fetch_add(1) % RING);
    PV_CUDA_CHECK(cudaMemsetAsync(ticket, 0, sizeof(unsigned long long), (cudaStream_t)stream));
    predict_bases_kernel<<<(unsigned)blocks, WR_WARPS * 32, 0, (cudaStream_t)stream>>>(b, bases_dev, ticket);
    apply_patches_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(b, read_patch_off_dev, patches_dev, bases_dev);
    PV_CUDA_CHECK(cudaGetLastError());
    return PV_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// 8-bit CIGAR ("cigar8" wire form): one code byte per op -- bit 0 = 0: M of length (c >> 1) + 1 (1..128); bit 0 = 1:
// bits 2:1 = 0 I, 1 D of length (c >> 3) + 1 (1..32), 3 = "escaped": the op's full 32-bit BAM word is the next entry of
// a separate escape stream (any other op type, longer or empty ops). One byte per op makes the code stream op-indexed
// (the read's CIGAR offset addresses it directly) and the decode parallel: an escaped op finds its word by counting
// the escape flags in front of it. At ONT error rates ~1 % of the ops escape: 1.05 bytes per op instead of 2 (cigar16).
// ---------------------------------------------------------------------------------------------------------------------
namespace {

__host__ __device__ __forceinline__ bool c8_plain(uint32_t w) {
    const uint32_t op = w & 15u, len = w >> 4;
    return len >= 1u && ((op == 0u && len <= 128u) || ((op == 1u || op == 2u) && len <= 32u));
}
__host__ __device__ __forceinline__ uint8_t c8_encode(uint32_t w) {          // c8_plain(w) only
    const uint32_t op = w & 15u, len = w >> 4;
    return (uint8_t)(op == 0u ? (len - 1u) << 1 : ((len - 1u) << 3) | ((op - 1u) << 1) | 1u);
}
__host__ __device__ __forceinline__ bool c8_is_escape(uint32_t c) { return (c & 7u) == 7u; }
__host__ __device__ __forceinline__ uint32_t c8_decode(uint32_t c) {         // !c8_is_escape(c)
    return (c & 1u) ? ((((c >> 3) + 1u) << 4) | (((c >> 1) & 3u) + 1u)) : (((c >> 1) + 1u) << 4);
}

__global__ void unpack_cigar8_kernel(const PvReadBatch b, const uint8_t* __restrict__ codes, const int64_t* __restrict__ esc_off,
                                     const uint32_t* __restrict__ esc, uint32_t* __restrict__ cigar) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < b.n_reads; r += n_warps) {
        const int64_t co = b.read_cigar_off[r];
        const int n_ops = b.read_n_ops[r];
        int64_t e = esc_off[r];
        for (int k0 = 0; k0 < n_ops; k0 += 32) {
            const int k = k0 + lane;
            const uint32_t c = k < n_ops ? codes[co + k] : 0u;
            const bool is_esc = k < n_ops && c8_is_escape(c);
            const unsigned m = __ballot_sync(0xffffffffu, is_esc);
            if (k < n_ops) cigar[co + k] = is_esc ? esc[e + __popc(m & ((1u << lane) - 1u))] : c8_decode(c);
            e += __popc(m);
        }
    }
}

}  // namespace

extern "C" int pv_pack_cigar8(const PvReadBatch* hb, uint8_t* codes, int64_t* read_esc_off, uint32_t* escapes, int64_t esc_capacity,
                              int32_t threads) {
    if (!hb || !codes || !read_esc_off) return pv::set_error(PV_EINVAL, "pv_pack_cigar8: null argument");
    const PvReadBatch& b = *hb;
    if (b.n_ops && !b.cigar) return pv::set_error(PV_EINVAL, "pv_pack_cigar8 needs the plain cigar array");
    if (!escapes) {
        // pass 1: the code bytes and the number of escapes per read -> exclusive prefix in read_esc_off[0 .. n_reads]
        wr_parallel(b.n_reads, threads, [&](int64_t lo, int64_t hi) {
            for (int64_t r = lo; r < hi; r++) {
                const int64_t co = b.read_cigar_off[r];
                int64_t n = 0;
                for (int k = 0; k < b.read_n_ops[r]; k++) {
                    const uint32_t w = b.cigar[co + k];
                    if (c8_plain(w)) codes[co + k] = c8_encode(w);
                    else { codes[co + k] = 7; n++; }
                }
                read_esc_off[r + 1] = n;
            }
        });
        read_esc_off[0] = 0;
        for (int64_t r = 0; r < b.n_reads; r++) read_esc_off[r + 1] += read_esc_off[r];
        return PV_OK;
    }
    if (read_esc_off[b.n_reads] > esc_capacity)
        return pv::set_error(PV_EINVAL, "pv_pack_cigar8: %lld escapes, buffer holds %lld", (long long)read_esc_off[b.n_reads], (long long)esc_capacity);
    wr_parallel(b.n_reads, threads, [&](int64_t lo, int64_t hi) {
        for (int64_t r = lo; r < hi; r++) {
            const int64_t co = b.read_cigar_off[r];
            uint32_t* out = escapes + read_esc_off[r];
            for (int k = 0; k < b.read_n_ops[r]; k++) { const uint32_t w = b.cigar[co + k]; if (!c8_plain(w)) *out++ = w; }
        }
    });
    return PV_OK;
}

extern "C" int pv_unpack_cigar8(const PvReadBatch* dev_batch, const uint8_t* codes_dev, const int64_t* read_esc_off_dev,
                                const uint32_t* escapes_dev, uint32_t* cigar_dev, void* stream) {
    if (!dev_batch || !cigar_dev) return pv::set_error(PV_EINVAL, "pv_unpack_cigar8: null argument");
    const PvReadBatch& b = *dev_batch;
    if (b.n_reads == 0 || b.n_ops == 0) return PV_OK;
    if (!codes_dev || !read_esc_off_dev) return pv::set_error(PV_EINVAL, "pv_unpack_cigar8 needs the code bytes and escape offsets on the device");
    if (int rc = pv::require_device()) return rc;
    int64_t blocks = (b.n_reads + 7) / 8;
    const int64_t cap = (int64_t)pv::sm_count() * 16;
    if (blocks > cap) blocks = cap;
    unpack_cigar8_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(b, codes_dev, read_esc_off_dev, escapes_dev, cigar_dev);
    PV_CUDA_CHECK(cudaGetLastError());
    return PV_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Quality predicates ("quals_pred" wire form). The summary reads a base quality in exactly two places
// (/root/reference/pepper_variant/modules/cpp/region_summary.cpp): `q >= min_snp_baseq` for an aligned base (:367,:377,
// :393) and, for an insert, `sum q[anchor .. anchor+len] >= min_indel_baseq * (len + 1)` together with
// `q[anchor] < min_snp_baseq` (:448-463). When the thresholds are known at pack time the qualities themselves need not
// travel: the device array is rebuilt as ONE fill value F = max(ceil(min_snp_baseq), ceil(min_indel_baseq)) plus a
// per-read patch list (the entry format of bases_ref) holding SURROGATE values wherever F would change one of those
// predicates:
//     aligned base below the SNP threshold                          -> 0
//     insert whose sum fails:   inserted bases -> 0, anchor -> ceil(min_snp_baseq) if it passes the SNP test, else 0
//     insert whose sum passes but whose anchor fails the SNP test:  inserted bases -> 255 (anchor already 0)
//     any insert that is not "M-run, insert, then a read-consuming op that is no insert" (chained inserts, insert after
//     a deletion or clip, empty inserts, inserts running over the read's end): the REAL qualities of anchor + insert
// Every predicate the summary evaluates has the same truth value on the surrogate array as on the real qualities (both
// thresholds must be <= 127 so that 255 * len >= thr * (len + 1)); the qualities themselves are NOT recoverable. The read
// index follows the reference's walk, including the N/P fall-through (:556-561).
// ---------------------------------------------------------------------------------------------------------------------
namespace {

struct QpThr { double snp, indel; int fill, amin; };

inline bool qp_read_consuming(int op) { return wr_match(op) || op == 4 || op == 3 || op == 6; }   // M = X S, and N P by fall-through

// surrogate qualities of read r into s[0 .. read_len)
void qp_surrogate_read(const PvReadBatch& b, int64_t r, const QpThr& t, std::vector<uint8_t>& s) {
    const int read_len = b.read_len[r], n_ops = b.read_n_ops[r];
    const uint8_t* q = b.quals + b.read_base_off[r];
    const uint32_t* cg = b.cigar + b.read_cigar_off[r];
    s.assign((size_t)read_len, (uint8_t)t.fill);
    int64_t ri = 0;
    for (int k = 0; k < n_ops; k++) {
        const int op = (int)(cg[k] & 15u);
        const int64_t len = (int64_t)(cg[k] >> 4);
        if (wr_match(op)) {
            for (int64_t i = ri; i < ri + len && i < read_len; i++)
                if (!((double)q[i] >= t.snp)) s[(size_t)i] = 0;
            ri += len;
        } else if (op == 1) {
            if (ri >= 1 && ri - 1 < read_len) {
                bool simple = k > 0 && wr_match((int)(cg[k - 1] & 15u)) && (cg[k - 1] >> 4) != 0u && len >= 1 && ri + len <= read_len;
                for (int k2 = k + 1; simple && k2 < n_ops; k2++) {          // the next op that touches the read index
                    const int op2 = (int)(cg[k2] & 15u);
                    if (op2 == 1) simple = false;
                    else if (qp_read_consuming(op2)) break;
                }
                if (simple) {
                    int64_t bq = 0;
                    for (int64_t i = ri - 1; i < ri + len; i++) bq += q[i];
                    const bool pass = (double)bq >= t.indel * (double)(len + 1);
                    const bool bit = (double)q[ri - 1] >= t.snp;
                    if (pass && !bit) { for (int64_t i = ri; i < ri + len; i++) s[(size_t)i] = 255; }
                    else if (!pass) {
                        for (int64_t i = ri; i < ri + len; i++) s[(size_t)i] = 0;
                        s[(size_t)(ri - 1)] = (uint8_t)(bit ? t.amin : 0);
                    }
                } else {
                    for (int64_t i = ri - 1; i < ri + len && i < read_len; i++) s[(size_t)i] = q[i];
                }
            }
            ri += len;
        } else if (op == 4 || op == 3 || op == 6) {
            ri += len;
        }
        if (ri > read_len) ri = read_len;                                 // malformed CIGAR: nothing behind the read's end is read
    }
}

__global__ void fill_quals_kernel(uint4* __restrict__ out, int64_t n16, uint32_t word) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) out[i] = make_uint4(word, word, word, word);
}

}  // namespace

extern "C" int pv_pack_quals_pred(const PvReadBatch* hb, double min_snp_baseq, double min_indel_baseq, uint8_t* fill_out,
                                  int64_t* read_patch_off, uint16_t* patches, int64_t patch_capacity, int32_t threads) {
    if (!hb || !read_patch_off || !fill_out) return pv::set_error(PV_EINVAL, "pv_pack_quals_pred: null argument");
    const PvReadBatch& b = *hb;
    if (b.n_reads && (!b.quals || !b.cigar)) return pv::set_error(PV_EINVAL, "pv_pack_quals_pred needs the plain quals and cigar arrays");
    if (!(min_snp_baseq <= 127.0) || !(min_indel_baseq <= 127.0))
        return pv::set_error(PV_EINVAL, "pv_pack_quals_pred: thresholds above 127 (%g, %g) cannot be represented by surrogate bytes", min_snp_baseq, min_indel_baseq);
    QpThr t;
    t.snp = min_snp_baseq; t.indel = min_indel_baseq;
    const double cs = std::ceil(min_snp_baseq), ci = std::ceil(min_indel_baseq);
    t.amin = cs < 0.0 ? 0 : (int)cs;
    const int fi = ci < 0.0 ? 0 : (int)ci;
    t.fill = t.amin > fi ? t.amin : fi;
    *fill_out = (uint8_t)t.fill;
    const bool count_only = patches == nullptr;
    if (!count_only && read_patch_off[b.n_reads] > patch_capacity)
        return pv::set_error(PV_EINVAL, "pv_pack_quals_pred: %lld patch entries needed, buffer holds %lld", (long long)read_patch_off[b.n_reads], (long long)patch_capacity);
    wr_parallel(b.n_reads, threads, [&](int64_t lo, int64_t hi) {
        std::vector<uint8_t> s;
        for (int64_t r = lo; r < hi; r++) {
            qp_surrogate_read(b, r, t, s);
            uint16_t* out = count_only ? nullptr : patches + read_patch_off[r];
            int64_t n = 0, cursor = 0;
            const int64_t L = (int64_t)s.size();
            for (int64_t i = 0; i < L; i++) {
                if (s[(size_t)i] == (uint8_t)t.fill) continue;
                int64_t gap = i - cursor;
                n += gap / 255 + 1;
                if (out) {
                    while (gap >= 255) { *out++ = 255; gap -= 255; }
                    *out++ = (uint16_t)(gap | ((uint16_t)s[(size_t)i] << 8));
                }
                cursor = i + 1;
            }
            if (count_only) read_patch_off[r + 1] = n;
        }
    });
    if (count_only) {
        read_patch_off[0] = 0;
        for (int64_t r = 0; r < b.n_reads; r++) read_patch_off[r + 1] += read_patch_off[r];
    }
    return PV_OK;
}

extern "C" int pv_unpack_quals_pred(const PvReadBatch* dev_batch, int32_t fill, const int64_t* read_patch_off_dev, const uint16_t* patches_dev,
                                    uint8_t* quals_dev, void* stream) {
    if (!dev_batch || !quals_dev) return pv::set_error(PV_EINVAL, "pv_unpack_quals_pred: null argument");
    const PvReadBatch& b = *dev_batch;
    if (b.n_reads == 0 || b.n_bases == 0) return PV_OK;
    if (!read_patch_off_dev) return pv::set_error(PV_EINVAL, "pv_unpack_quals_pred needs the patch offsets on the device");
    if (fill < 0 || fill > 255) return pv::set_error(PV_EINVAL, "pv_unpack_quals_pred: fill value %d is no byte", fill);
    if (b.n_bases % 16 || ((uintptr_t)quals_dev & 15)) return pv::set_error(PV_EINVAL, "pv_unpack_quals_pred: the quality array must be 16-byte aligned and a multiple of 16 long");
    if (int rc = pv::require_device()) return rc;
    const int64_t n16 = b.n_bases / 16;
    int64_t fblocks = (n16 + 255) / 256;
    const int64_t fcap = (int64_t)pv::sm_count() * 8;
    if (fblocks > fcap) fblocks = fcap;
    fill_quals_kernel<<<(unsigned)fblocks, 256, 0, (cudaStream_t)stream>>>((uint4*)quals_dev, n16, (uint32_t)fill * 0x01010101u);
    int64_t blocks = (b.n_reads + 7) / 8;
    const int64_t cap = (int64_t)pv::sm_count() * 8;
    if (blocks > cap) blocks = cap;
    apply_patches_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(b, read_patch_off_dev, patches_dev, quals_dev);
    PV_CUDA_CHECK(cudaGetLastError());
    return PV_OK;
}

extern "C" int32_t pv_min_qual(const PvReadBatch* hb, int32_t threads) {
    if (!hb || (hb->n_reads && !hb->quals)) return 0;
    const PvReadBatch& b = *hb;
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    std::vector<int> mins((size_t)threads, 255);
    const int64_t per = (b.n_reads + threads - 1) / threads;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back([&, t]() {
        const int64_t lo = t * per, hi = lo + per < b.n_reads ? lo + per : b.n_reads;
        int m = 255;
        for (int64_t r = lo; r < hi; r++) {
            const uint8_t* q = b.quals + b.read_base_off[r];
            const int n = b.read_len[r];
            for (int i = 0; i < n; i++) m = q[i] < m ? q[i] : m;
        }
        mins[(size_t)t] = m;
    });
    for (auto& th : pool) th.join();
    int m = 255;
    for (int v : mins) m = v < m ? v : m;
    return m;
}
