// BAM decoding on the device ("ingest that can feed the GPU"). The host only reads the compressed
// bytes the BAI query names and walks the BGZF block headers (libpv_ingest.so: pv_bam_plan*); everything per byte happens
// here, and the packed read batch is born in HBM, where the summary kernels take it from:
//
//   inflate_kernel        one WARP per BGZF block (independent <= 64 KiB DEFLATE streams, inflate_warp.cuh): tables in shared
//                         memory built by the warp, matches copied by all lanes, the block's CRC-32 against its trailer
//   chain_kernel          record boundaries: one thread per chain segment (chunk starts + the linear index's record
//                         offsets every 16 kbp), a pointer chase over block_size fields; count pass, scan, fill pass
//   clip_count / _pairs   one thread per record: flag / mapq filters (bam_handler.cpp:137-150), bam_endpos, the spans
//                         (region +- safe bases) the record overlaps, and per (record, span) the cut of
//                         bam_handler.cpp:178-306 (sizes only)
//   radix sort + scans    (span, file order) order of the pairs = read order of the batch; base / CIGAR offsets
//   write_kernel          one warp per read of the batch: lane 0 re-walks the CIGAR and writes the kept ops, all lanes
//                         decode the kept bases (nt16 -> upper-case ASCII) and copy the qualities, 16-byte padded
//
// Results are bit-identical to the host ingest (pv_ingest_regions), which tests pin to the compiled reference.
#include "common.cuh"
#include "bam_core.cuh"
#include "inflate_warp.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace {
using namespace bamcore;

constexpr int INFLATE_WARPS = winf::LIT_ROOT >= 12 ? 2 : 4;     // static shared memory: <= 48 KB per CTA

// one warp per BGZF block (inflate_warp.cuh); with `verify` the block's CRC-32 is checked against its trailer
__global__ void __launch_bounds__(INFLATE_WARPS * 32) inflate_kernel(const uint8_t* __restrict__ comp, int64_t comp_bytes,
                                                                     const PvBgzfBlock* __restrict__ blocks, int n_blocks,
                                                                     uint8_t* __restrict__ out, int64_t out_bytes, int verify, int32_t* n_bad, int32_t* ticket) {
    __shared__ winf::WarpTables T[INFLATE_WARPS];
    __shared__ uint32_t crc_tab[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? winf::CRC_POLY ^ (c >> 1) : c >> 1;
        crc_tab[i] = c;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    winf::WarpTables& Tw = T[threadIdx.x >> 5];
    for (;;) {                                                    // blocks are pulled from a ticket: no wave of warps waits for its slowest
        int i = 0;
        if (lane == 0) i = atomicAdd(ticket, 1);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= n_blocks) break;
        const PvBgzfBlock b = blocks[i];
        bool ok = b.c_off >= 0 && b.c_len >= 0 && b.c_off + b.c_len <= comp_bytes && b.isize >= 0 && b.isize <= 65536 && b.u_off >= 0 &&
                  b.u_off + b.isize <= out_bytes;
        if (ok && b.isize > 0) {
            ok = winf::inflate_warp(comp, comp_bytes, b.c_off, b.c_len, out + b.u_off, b.isize, Tw, lane);
            if (ok && verify) ok = winf::crc32_warp(out + b.u_off, b.isize, crc_tab, lane) == b.crc;
        } else if (ok && verify) {
            ok = b.crc == 0;
        }
        if (!ok && lane == 0) atomicAdd(n_bad, 1);
        __syncwarp();
    }
}

// record chain of one segment; FILL: offsets go to rec_off[first[s] ...]
template <bool FILL>
__global__ void chain_kernel(const uint8_t* __restrict__ U, int64_t u_size, const int64_t* __restrict__ seg_begin,
                             const int64_t* __restrict__ seg_end, int n_seg, int64_t* first, int64_t* rec_off, int64_t cap, int32_t* status) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    int64_t off = seg_begin[s];
    const int64_t end = seg_end[s];
    int64_t n = 0, at = FILL ? first[s] : 0;
    while (off < end) {
        if (off + 4 > u_size) { atomicOr(status, 1); break; }
        const uint32_t bs = ld32(U + off);
        if (bs < 32) { atomicOr(status, 1); break; }
        if (FILL) { if (at + n < cap) rec_off[at + n] = off; else atomicOr(status, 2); }
        n++;
        off += 4 + (int64_t)bs;
    }
    if (off != end) atomicOr(status, 1);                       // the chain must land on the next entry point
    if (!FILL) first[s] = n;
}

// in-place exclusive scan of data[0 .. n), total to data[n]; one CTA
__global__ void scan_kernel(int64_t* data, int64_t n) {
    __shared__ int64_t warp_sum[32];
    __shared__ int64_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += blockDim.x) {
        const int64_t i = base + tid;
        const int64_t v = i < n ? data[i] : 0;
        int64_t x = v;
        for (int d = 1; d < 32; d <<= 1) { const int64_t t = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += t; }
        if (lane == 31) warp_sum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int64_t w = lane < (int)(blockDim.x >> 5) ? warp_sum[lane] : 0;
            for (int d = 1; d < 32; d <<= 1) { const int64_t t = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += t; }
            warp_sum[lane] = w;                                // inclusive over warps
        }
        __syncthreads();
        const int64_t carry = carry_s;
        const int64_t excl = carry + (warp ? warp_sum[warp - 1] : 0) + x - v;
        if (i < n) data[i] = excl;
        __syncthreads();
        if (tid == 0) carry_s = carry + warp_sum[(blockDim.x >> 5) - 1];
        __syncthreads();
    }
    if (tid == 0) data[n] = carry_s;
}

struct ClipArgs {
    const uint8_t* U; int64_t u_size;
    const int64_t* rec_off; int64_t n_rec;
    int32_t tid;
    const int64_t* span_start; const int64_t* span_stop; int32_t n_spans;
    int32_t supp, min_mapq;
};

// spans [j0, ...) a record [pos, endpos) can overlap: spans ascend in start and stop, so the ones whose stop <= pos are a prefix
__device__ __forceinline__ int first_span(const ClipArgs& a, int64_t pos) {
    int lo = 0, hi = a.n_spans;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a.span_stop[mid] <= pos) lo = mid + 1; else hi = mid; }
    return lo;
}

__device__ __forceinline__ int64_t warp_incl_scan(int64_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int64_t t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += t; }
    return v;
}
__device__ __forceinline__ int64_t warp_sum(int64_t v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// bam_endpos by the warp: pos + the reference length of the CIGAR (pos + 1 for unmapped / zero-length alignments)
__device__ int64_t endpos_warp(const uint8_t* U, const RecHdr& h, int lane) {
    int64_t rlen = 0;
    for (int k = lane; k < h.n_ops; k += 32) {
        const uint32_t w = ld32(U + h.ops_off + 4 * (int64_t)k);
        const int op = (int)(w & 15u);
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += w >> 4;
    }
    rlen = warp_sum(rlen);
    return h.pos + (((h.flag & 4) || rlen == 0) ? 1 : rlen);
}

// clip_walk<false> (bam_core.cuh) by the warp, 32 ops per step. Until the walk ends every op advances the reference / read
// position by its whole length, so both are prefix sums; the first match op that keeps a base (k_first) opens the read,
// from there on every op that is reached (its start <= stop) is kept, clipped at stop at most. The result is uniform.
__device__ Clip clip_warp(const uint8_t* U, const RecHdr& h, int64_t start, int64_t stop, int lane) {
    Clip c;
    c.pos_start = -1; c.pos_end = -1; c.idx0 = -1; c.n_bases = 0; c.n_ops = 0; c.bad = false; c.split = false;
    c.k_first = c.k_last = -1; c.first_kept = c.last_kept = 0;
    int64_t cp = h.pos, ci = 0;                                  // positions in front of the step's first op
    for (int k0 = 0; k0 < h.n_ops; k0 += 32) {
        if (cp > stop) break;
        const int k = k0 + lane;
        uint32_t w = 0;
        if (k < h.n_ops) w = ld32(U + h.ops_off + 4 * (int64_t)k);
        const int op = k < h.n_ops ? (int)(w & 15u) : 15;
        const int64_t len = w >> 4;
        const bool is_m = op == 0 || op == 7 || op == 8, is_i = op == 1 || op == 4, is_d = op == 2 || op == 3;
        const int64_t radv = (is_m || is_d) ? len : 0, qadv = (is_m || is_i) ? len : 0;
        const int64_t rincl = warp_incl_scan(radv, lane), qincl = warp_incl_scan(qadv, lane);
        const int64_t p0 = cp + rincl - radv, q0 = ci + qincl - qadv;   // where this op starts
        const bool reached = k < h.n_ops && p0 <= stop;
        // a match op: bases in front of start are skipped, bases behind stop are cut
        const int64_t i0 = (is_m && p0 < start) ? (start - p0 < len ? start - p0 : len) : 0;
        int64_t take = len - i0 < stop - (p0 + i0) + 1 ? len - i0 : stop - (p0 + i0) + 1;
        if (take < 0) take = 0;
        if (c.k_first < 0) {
            const unsigned m = __ballot_sync(0xffffffffu, reached && is_m && take > 0);
            if (m) {
                const int src = __ffs(m) - 1;
                c.k_first = k0 + src;
                c.pos_start = __shfl_sync(0xffffffffu, p0 + i0, src);
                c.idx0 = __shfl_sync(0xffffffffu, q0 + i0, src);
                c.first_kept = (int32_t)__shfl_sync(0xffffffffu, take, src);
            }
        }
        int64_t kept = 0;
        if (c.k_first >= 0 && reached && k >= c.k_first) {
            if (is_m) kept = take;
            else if (is_i) kept = len;
            else if (is_d) kept = len < stop - p0 + 1 ? len : stop - p0 + 1;
        }
        const unsigned km = __ballot_sync(0xffffffffu, kept > 0);
        if (km) {
            const int last = 31 - __clz(km);
            c.k_last = k0 + last;
            c.last_kept = (int32_t)__shfl_sync(0xffffffffu, kept, last);
            c.n_ops += __popc(km);
        }
        c.n_bases += warp_sum((is_m || is_i) ? kept : 0);
        // where the walk stands behind the step's reached ops
        const int64_t rsum = warp_sum(reached ? radv : 0);
        cp += __shfl_sync(0xffffffffu, rincl, 31); ci += __shfl_sync(0xffffffffu, qincl, 31);
        if (c.k_first >= 0) { const int64_t e = (cp - __shfl_sync(0xffffffffu, rincl, 31)) + rsum; c.pos_end = e < stop + 1 ? e : stop + 1; }
        if (__ballot_sync(0xffffffffu, k < h.n_ops && !reached)) break;
    }
    if (c.k_first >= 0 && c.idx0 + c.n_bases > h.l_seq) c.bad = true;
    return c;
}

// PAIRS: writes the record's pairs at first[r] ...; else counts them into first[r]. One warp per record.
template <bool PAIRS>
__global__ void __launch_bounds__(256) clip_kernel(const ClipArgs a, int64_t* first, PvBamPair* pairs, unsigned long long* keys, uint32_t* vals, int32_t* status) {
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= a.n_rec) return;
    int64_t n = 0;
    const int64_t at = PAIRS ? first[r] : 0;
    const RecHdr h = parse_record(a.U, a.rec_off[r], a.u_size);   // every lane: the field offsets
    if (!h.ok) { if (lane == 0) atomicOr(status, 4); }
    else if (h.tid == a.tid && record_passes(h, a.supp, a.min_mapq)) {
        const int64_t endpos = endpos_warp(a.U, h, lane);
        for (int j = first_span(a, h.pos); j < a.n_spans && a.span_start[j] < endpos; j++) {
            const int64_t start = a.span_start[j], stop = a.span_stop[j];
            if (h.pos >= stop || endpos <= start) continue;
            const Clip c = clip_warp(a.U, h, start, stop, lane);
            if (c.bad || c.n_bases == 0) continue;
            if (PAIRS && lane == 0) {
                PvBamPair p; p.rec_off = a.rec_off[r]; p.span = j; p.n_ops = c.n_ops; p.n_bases = c.n_bases;
                p.k_first = c.k_first; p.k_last = c.k_last; p.first_kept = c.first_kept; p.last_kept = c.last_kept;
                p.idx0 = c.idx0; p.pos_start = c.pos_start; p.pos_end = c.pos_end;
                pairs[at + n] = p;
                keys[at + n] = ((unsigned long long)j << 32) | (unsigned long long)(at + n);
                vals[at + n] = (uint32_t)(at + n);
            }
            n++;
        }
    }
    if (!PAIRS && lane == 0) first[r] = n;
}

__global__ void gather_pairs_kernel(const PvBamPair* __restrict__ in, const uint32_t* __restrict__ order, int64_t n, PvBamPair* out,
                                    int64_t* base_size, int64_t* op_size) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const PvBamPair p = in[order[i]];
    out[i] = p;
    base_size[i] = (p.n_bases + 15) & ~(int64_t)15;
    op_size[i] = p.n_ops;
}

__global__ void region_begin_kernel(const unsigned long long* __restrict__ keys, int64_t n, int n_spans, int64_t* begin) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > n_spans) return;
    const unsigned long long k = (unsigned long long)j << 32;
    int64_t lo = 0, hi = n;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (keys[mid] < k) lo = mid + 1; else hi = mid; }
    begin[j] = lo;
}

struct WriteArgs {
    const uint8_t* U; int64_t u_size;
    const PvBamPair* pairs; int64_t n;
    const int64_t* span_start; const int64_t* span_stop;
    const int64_t* base_off; const int64_t* cigar_off;
    int64_t* read_pos; int64_t* read_pos_end; int32_t* read_len; int32_t* read_n_ops; uint8_t* read_flags; uint8_t* read_mapq;
    int32_t* hp; uint16_t* bam_flag; int64_t* name_off; int32_t* name_len;
    uint8_t* bases; uint8_t* quals; uint32_t* cigar; int32_t* min_qual;
};

__global__ void __launch_bounds__(256) write_kernel(const WriteArgs a, int32_t* status) {
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= a.n) return;
    const PvBamPair p = a.pairs[i];
    const RecHdr h = parse_record(a.U, p.rec_off, a.u_size);     // every lane (cheap), so that all have the field offsets
    // the kept ops are ops k_first .. k_last of the record, whole except for the two ends (the pairs pass recorded their kept
    // lengths); ops of other kinds in between (hard clip, pad) drop out: 32 ops per step, compacted by ballot
    {
        uint32_t* out_ops = a.cigar + a.cigar_off[i];
        int written = 0;
        for (int k0 = p.k_first; k0 <= p.k_last && p.k_first >= 0; k0 += 32) {
            const int k = k0 + lane;
            bool keep = false;
            uint32_t word = 0;
            if (k <= p.k_last) {
                const uint32_t w = ld32(a.U + h.ops_off + 4 * (int64_t)k);
                const int op = (int)(w & 15u);
                const uint32_t kept = k == p.k_first ? (uint32_t)p.first_kept : k == p.k_last ? (uint32_t)p.last_kept : (w >> 4);
                keep = kept > 0 && (op == 0 || op == 1 || op == 2 || op == 3 || op == 4 || op == 7 || op == 8);
                word = (kept << 4) | (uint32_t)op;
            }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) out_ops[written + __popc(m & ((1u << lane) - 1u))] = word;
            written += __popc(m);
        }
        if (lane == 0) {
            if (written != p.n_ops || p.k_last >= h.n_ops) atomicOr(status, 16);
            a.read_pos[i] = p.pos_start; a.read_pos_end[i] = p.pos_end;
            a.read_len[i] = (int32_t)p.n_bases; a.read_n_ops[i] = p.n_ops;
            a.read_flags[i] = (h.flag & 0x10) ? 1 : 0;
            a.read_mapq[i] = (uint8_t)h.mapq;
            a.hp[i] = parse_hp(a.U, h.aux_off, h.rec_end);
            a.bam_flag[i] = (uint16_t)h.flag;
            a.name_off[i] = h.name_off;
            int nl = 0;
            while (nl < h.l_name && a.U[h.name_off + nl]) nl++;
            a.name_len[i] = nl;
        }
    }
    const int64_t idx0 = p.idx0;
    const int64_t n = p.n_bases, padded = (n + 15) & ~(int64_t)15, bo = a.base_off[i];
    int mq = 255;
    // four bases per lane and step: nt16 codes -> upper-case ASCII through a 16-byte table held in two registers, qualities
    // copied, both leave as one aligned 32-bit store (the read's slot starts on a 16-byte boundary)
    const unsigned long long NT_LO = 0x565352474d43413dULL, NT_HI = 0x4e42444b48595754ULL;   // "=ACMGRSV", "TWYHKDBN"
    const uint8_t* seq = a.U + h.seq_off;
    const uint8_t* qual = a.U + h.qual_off + idx0;
    uint32_t* ob = (uint32_t*)(a.bases + bo);
    uint32_t* oq = (uint32_t*)(a.quals + bo);
#pragma unroll 2
    for (int64_t k = 4 * (int64_t)lane; k < padded; k += 128) {
        uint32_t bw = 0, qw = 0;
        if (k + 4 <= n) {
            const int64_t s0 = idx0 + k;
            const uint8_t* sp = seq + (s0 >> 1);
            const uint32_t b0 = sp[0], b1 = sp[1], b2 = (s0 & 1) ? sp[2] : 0u;
            const uint32_t nibs = (s0 & 1) ? (((b0 & 15u) << 12) | (b1 << 4) | (b2 >> 4)) : ((b0 << 8) | b1);   // four codes, first in the top nibble
            const uint32_t q0 = qual[k], q1 = qual[k + 1], q2 = qual[k + 2], q3 = qual[k + 3];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t c = (nibs >> (12 - 4 * j)) & 15u;
                bw |= (uint32_t)(((c & 8u) ? NT_HI : NT_LO) >> (8 * (c & 7u)) & 0xffu) << (8 * j);
            }
            qw = q0 | (q1 << 8) | (q2 << 16) | (q3 << 24);
            const int m01 = q0 < q1 ? q0 : q1, m23 = q2 < q3 ? q2 : q3, m = m01 < m23 ? m01 : m23;
            mq = m < mq ? m : mq;
        } else {
            for (int j = 0; j < 4; j++) {
                if (k + j < n) {
                    const uint32_t b = record_base(a.U, h, idx0 + k + j), q = qual[k + j];
                    bw |= b << (8 * j); qw |= q << (8 * j);
                    mq = (int)q < mq ? (int)q : mq;
                }
            }
        }
        ob[k >> 2] = bw; oq[k >> 2] = qw;
    }
    for (int d = 16; d; d >>= 1) { const int o = __shfl_xor_sync(0xffffffffu, mq, d); mq = o < mq ? o : mq; }
    if (lane == 0) atomicMin(a.min_qual, mq);
}

__global__ void names_kernel(const uint8_t* __restrict__ U, const int64_t* __restrict__ name_off, const int32_t* __restrict__ name_len,
                             const int64_t* __restrict__ out_off, int64_t n, uint8_t* out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t s = name_off[i], o = out_off[i];
    const int l = name_len[i];
    for (int k = 0; k < l; k++) out[o + k] = U[s + k];
    out[o + l] = 0;
}

// region r's reference bytes [span_start[r], span_start[r] + ref_len[r]) out of ONE fetch [fetch_start, fetch_start + fetch_len) of the contig;
// 'N' behind the fetch (past the contig end, AlignmentSummarizer's region_end + 1 can reach there)
__global__ void gather_reference_kernel(const uint8_t* __restrict__ fetched, int64_t fetch_len, int64_t fetch_start,
                                        const int64_t* __restrict__ span_start, const int64_t* __restrict__ ref_off,
                                        const int64_t* __restrict__ ref_len, uint8_t* __restrict__ ref) {
    const int r = blockIdx.y;
    const int64_t len = ref_len[r], src = span_start[r] - fetch_start, o = ref_off[r];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = src + i;
        ref[o + i] = (k >= 0 && k < fetch_len) ? fetched[k] : (uint8_t)'N';
    }
}

inline unsigned grid_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

struct ClipWs { unsigned long long* keys_in; unsigned long long* keys_out; uint32_t* vals_in; uint32_t* vals_out; PvBamPair* pairs_in;
                int64_t* base_size; int64_t* op_size; void* sort_tmp; size_t sort_bytes; };
ClipWs clip_ws(pv::Arena& ar, int64_t n_pairs) {
    ClipWs w;
    w.keys_in = ar.take<unsigned long long>(n_pairs + 1); w.keys_out = ar.take<unsigned long long>(n_pairs + 1);
    w.vals_in = ar.take<uint32_t>(n_pairs + 1); w.vals_out = ar.take<uint32_t>(n_pairs + 1);
    w.pairs_in = ar.take<PvBamPair>(n_pairs + 1);
    w.base_size = ar.take<int64_t>(n_pairs + 2); w.op_size = ar.take<int64_t>(n_pairs + 2);
    w.sort_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, w.sort_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)(n_pairs > 0 ? n_pairs : 1), 0, 64);
    w.sort_tmp = ar.take<uint8_t>((int64_t)w.sort_bytes + 256);
    return w;
}

}  // namespace

extern "C" int pv_bam_inflate_blocks(const uint8_t* comp_dev, int64_t comp_bytes, const PvBgzfBlock* blocks_dev, int32_t n_blocks,
                                     uint8_t* inflated_dev, int64_t inflated_bytes, int32_t verify_crc, int32_t* n_bad_dev, void* stream) {
    if (n_blocks < 0 || !n_bad_dev || (n_blocks > 0 && (!comp_dev || !blocks_dev || !inflated_dev)))
        return pv::set_error(PV_EINVAL, "pv_bam_inflate_blocks: bad arguments");
    if ((uintptr_t)comp_dev & 3) return pv::set_error(PV_EINVAL, "pv_bam_inflate_blocks: the compressed buffer must be 4-byte aligned");
    if (int rc = pv::require_device()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    PV_CUDA_CHECK(cudaMemsetAsync(n_bad_dev, 0, 8, st));
    if (n_blocks == 0) return PV_OK;
    static int resident = 0;                                      // CTAs the device holds at a time
    if (!resident) {
        int dev = 0, sms = 0, per_sm = 0;
        PV_CUDA_CHECK(cudaGetDevice(&dev));
        PV_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        PV_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, inflate_kernel, INFLATE_WARPS * 32, 0));
        resident = sms * (per_sm > 0 ? per_sm : 1);
    }
    unsigned grid = grid_for(n_blocks, INFLATE_WARPS);
    if (grid > (unsigned)resident) grid = (unsigned)resident;
    pv::prof_begin(pv::FAM_INGEST, st);
    inflate_kernel<<<grid, INFLATE_WARPS * 32, 0, st>>>(comp_dev, comp_bytes, blocks_dev, n_blocks, inflated_dev, inflated_bytes, verify_crc, n_bad_dev, n_bad_dev + 1);
    pv::prof_end(pv::FAM_INGEST, st, 1);
    PV_CUDA_CHECK(cudaGetLastError());
    return PV_OK;
}

extern "C" int pv_bam_index_records(const uint8_t* inflated_dev, int64_t inflated_bytes, const int64_t* seg_begin_dev,
                                    const int64_t* seg_end_dev, int32_t n_segments, int64_t* seg_first_dev, int64_t* rec_off_dev,
                                    int64_t rec_capacity, int32_t* status_dev, void* stream) {
    if (n_segments < 0 || !seg_first_dev || !status_dev || (n_segments > 0 && (!inflated_dev || !seg_begin_dev || !seg_end_dev)))
        return pv::set_error(PV_EINVAL, "pv_bam_index_records: bad arguments");
    if (int rc = pv::require_device()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    pv::prof_begin(pv::FAM_INGEST, st);
    int launches = 0;
    if (!rec_off_dev) {
        PV_CUDA_CHECK(cudaMemsetAsync(status_dev, 0, 4, st));
        if (n_segments > 0) {
            chain_kernel<false><<<grid_for(n_segments, 64), 64, 0, st>>>(inflated_dev, inflated_bytes, seg_begin_dev, seg_end_dev, n_segments, seg_first_dev, nullptr, 0, status_dev);
            launches++;
        }
        scan_kernel<<<1, 1024, 0, st>>>(seg_first_dev, n_segments);
        launches++;
    } else if (n_segments > 0) {
        chain_kernel<true><<<grid_for(n_segments, 64), 64, 0, st>>>(inflated_dev, inflated_bytes, seg_begin_dev, seg_end_dev, n_segments, seg_first_dev, rec_off_dev, rec_capacity, status_dev);
        launches++;
    }
    pv::prof_end(pv::FAM_INGEST, st, launches);
    PV_CUDA_CHECK(cudaGetLastError());
    return PV_OK;
}

extern "C" int64_t pv_bam_clip_workspace_bytes(int64_t n_pairs) {
    pv::Arena ar(nullptr, 0);
    clip_ws(ar, n_pairs);
    return ar.cur + 256;
}

extern "C" int pv_bam_clip_count(const uint8_t* inflated_dev, int64_t inflated_bytes, const int64_t* rec_off_dev, int64_t n_records,
                                 int32_t tid, const int64_t* span_start_dev, const int64_t* span_stop_dev, int32_t n_spans,
                                 int32_t include_supplementary, int32_t min_mapq, int64_t* rec_pair_first_dev, int32_t* status_dev, void* stream) {
    if (n_records < 0 || n_spans < 0 || !rec_pair_first_dev || !status_dev || (n_records > 0 && (!inflated_dev || !rec_off_dev)) ||
        (n_spans > 0 && (!span_start_dev || !span_stop_dev)))
        return pv::set_error(PV_EINVAL, "pv_bam_clip_count: bad arguments");
    if (int rc = pv::require_device()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const ClipArgs a{inflated_dev, inflated_bytes, rec_off_dev, n_records, tid, span_start_dev, span_stop_dev, n_spans, include_supplementary, min_mapq};
    pv::prof_begin(pv::FAM_INGEST, st);
    int launches = 1;
    if (n_records > 0) { clip_kernel<false><<<grid_for(n_records * 32, 256), 256, 0, st>>>(a, rec_pair_first_dev, nullptr, nullptr, nullptr, status_dev); launches++; }
    scan_kernel<<<1, 1024, 0, st>>>(rec_pair_first_dev, n_records);
    pv::prof_end(pv::FAM_INGEST, st, launches);
    PV_CUDA_CHECK(cudaGetLastError());
    return PV_OK;
}

extern "C" int pv_bam_clip_layout(const uint8_t* inflated_dev, int64_t inflated_bytes, const int64_t* rec_off_dev, int64_t n_records,
                                  int32_t tid, const int64_t* span_start_dev, const int64_t* span_stop_dev, int32_t n_spans,
                                  int32_t include_supplementary, int32_t min_mapq, const int64_t* rec_pair_first_dev, int64_t n_pairs,
                                  void* workspace_dev, int64_t workspace_bytes, PvBamPair* pairs_sorted_dev, int64_t* read_base_off_dev,
                                  int64_t* read_cigar_off_dev, int64_t* region_read_begin_dev, int64_t* totals_dev, int32_t* status_dev,
                                  void* stream) {
    if (n_records < 0 || n_pairs < 0 || n_spans < 0 || !region_read_begin_dev || !totals_dev || !status_dev ||
        (n_pairs > 0 && (!workspace_dev || !pairs_sorted_dev || !read_base_off_dev || !read_cigar_off_dev || !rec_pair_first_dev)))
        return pv::set_error(PV_EINVAL, "pv_bam_clip_layout: bad arguments");
    if (int rc = pv::require_device()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_pairs == 0) {
        PV_CUDA_CHECK(cudaMemsetAsync(region_read_begin_dev, 0, (size_t)(n_spans + 1) * 8, st));
        PV_CUDA_CHECK(cudaMemsetAsync(totals_dev, 0, 16, st));
        return PV_OK;
    }
    if (n_pairs >= (1ll << 32) || n_spans >= (1 << 30)) return pv::set_error(PV_EINVAL, "pv_bam_clip_layout: too many reads for one call");
    pv::Arena ar(workspace_dev, workspace_bytes);
    ClipWs w = clip_ws(ar, n_pairs);
    if (!ar.ok()) return pv::set_error(PV_EINVAL, "pv_bam_clip_layout: workspace of %lld bytes, %lld needed", (long long)workspace_bytes, (long long)ar.cur);
    const ClipArgs a{inflated_dev, inflated_bytes, rec_off_dev, n_records, tid, span_start_dev, span_stop_dev, n_spans, include_supplementary, min_mapq};
    pv::prof_begin(pv::FAM_INGEST, st);
    clip_kernel<true><<<grid_for(n_records * 32, 256), 256, 0, st>>>(a, (int64_t*)rec_pair_first_dev, w.pairs_in, w.keys_in, w.vals_in, status_dev);
    int span_bits = 1;
    while ((1ll << span_bits) < (int64_t)n_spans + 1) span_bits++;
    PV_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(w.sort_tmp, w.sort_bytes, (const unsigned long long*)w.keys_in, w.keys_out,
                                                  (const uint32_t*)w.vals_in, w.vals_out, (int)n_pairs, 0, 32 + span_bits, st));
    gather_pairs_kernel<<<grid_for(n_pairs, 256), 256, 0, st>>>(w.pairs_in, w.vals_out, n_pairs, pairs_sorted_dev, w.base_size, w.op_size);
    scan_kernel<<<1, 1024, 0, st>>>(w.base_size, n_pairs);
    scan_kernel<<<1, 1024, 0, st>>>(w.op_size, n_pairs);
    PV_CUDA_CHECK(cudaMemcpyAsync(read_base_off_dev, w.base_size, (size_t)n_pairs * 8, cudaMemcpyDeviceToDevice, st));
    PV_CUDA_CHECK(cudaMemcpyAsync(read_cigar_off_dev, w.op_size, (size_t)n_pairs * 8, cudaMemcpyDeviceToDevice, st));
    PV_CUDA_CHECK(cudaMemcpyAsync(totals_dev, w.base_size + n_pairs, 8, cudaMemcpyDeviceToDevice, st));
    PV_CUDA_CHECK(cudaMemcpyAsync(totals_dev + 1, w.op_size + n_pairs, 8, cudaMemcpyDeviceToDevice, st));
    region_begin_kernel<<<grid_for(n_spans + 1, 128), 128, 0, st>>>(w.keys_out, n_pairs, n_spans, region_read_begin_dev);
    pv::prof_end(pv::FAM_INGEST, st, 8);
    PV_CUDA_CHECK(cudaGetLastError());
    return PV_OK;
}

extern "C" int pv_bam_clip_write(const uint8_t* inflated_dev, int64_t inflated_bytes, const PvBamPair* pairs_sorted_dev, int64_t n_pairs,
                                 const int64_t* span_start_dev, const int64_t* span_stop_dev, const int64_t* read_base_off_dev,
                                 const int64_t* read_cigar_off_dev, int64_t* read_pos_dev, int64_t* read_pos_end_dev, int32_t* read_len_dev,
                                 int32_t* read_n_ops_dev, uint8_t* read_flags_dev, uint8_t* read_mapq_dev, int32_t* hp_dev,
                                 uint16_t* bam_flag_dev, int64_t* name_off_dev, int32_t* name_len_dev, uint8_t* bases_dev,
                                 uint8_t* quals_dev, uint32_t* cigar_dev, int32_t* min_qual_dev, int32_t* status_dev, void* stream) {
    if (n_pairs < 0 || !min_qual_dev || !status_dev) return pv::set_error(PV_EINVAL, "pv_bam_clip_write: bad arguments");
    if (int rc = pv::require_device()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int32_t init = 255;
    PV_CUDA_CHECK(cudaMemcpyAsync(min_qual_dev, &init, 4, cudaMemcpyHostToDevice, st));
    if (n_pairs == 0) return PV_OK;
    if (!inflated_dev || !pairs_sorted_dev || !span_start_dev || !span_stop_dev || !read_base_off_dev || !read_cigar_off_dev || !read_pos_dev ||
        !read_pos_end_dev || !read_len_dev || !read_n_ops_dev || !read_flags_dev || !read_mapq_dev || !hp_dev || !bam_flag_dev ||
        !name_off_dev || !name_len_dev || !bases_dev || !quals_dev || !cigar_dev)
        return pv::set_error(PV_EINVAL, "pv_bam_clip_write: null output");
    const WriteArgs a{inflated_dev, inflated_bytes, pairs_sorted_dev, n_pairs, span_start_dev, span_stop_dev, read_base_off_dev, read_cigar_off_dev,
                      read_pos_dev, read_pos_end_dev, read_len_dev, read_n_ops_dev, read_flags_dev, read_mapq_dev, hp_dev, bam_flag_dev,
                      name_off_dev, name_len_dev, bases_dev, quals_dev, cigar_dev, min_qual_dev};
    pv::prof_begin(pv::FAM_INGEST, st);
    write_kernel<<<grid_for(n_pairs * 32, 256), 256, 0, st>>>(a, status_dev);
    pv::prof_end(pv::FAM_INGEST, st, 1);
    PV_CUDA_CHECK(cudaGetLastError());
    return PV_OK;
}

extern "C" int pv_bam_gather_names(const uint8_t* inflated_dev, const int64_t* name_off_dev, const int32_t* name_len_dev,
                                   const int64_t* out_off_dev, int64_t n, uint8_t* out_dev, void* stream) {
    if (n < 0 || (n > 0 && (!inflated_dev || !name_off_dev || !name_len_dev || !out_off_dev || !out_dev)))
        return pv::set_error(PV_EINVAL, "pv_bam_gather_names: bad arguments");
    if (int rc = pv::require_device()) return rc;
    if (n == 0) return PV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    pv::prof_begin(pv::FAM_INGEST, st);
    names_kernel<<<grid_for(n, 128), 128, 0, st>>>(inflated_dev, name_off_dev, name_len_dev, out_off_dev, n, out_dev);
    pv::prof_end(pv::FAM_INGEST, st, 1);
    PV_CUDA_CHECK(cudaGetLastError());
    return PV_OK;
}

extern "C" int pv_bam_gather_reference(const uint8_t* fetched_dev, int64_t fetch_len, int64_t fetch_start, const int64_t* span_start_dev,
                                       const int64_t* region_ref_off_dev, const int64_t* region_ref_len_dev, int32_t n_spans,
                                       int64_t max_ref_len, uint8_t* ref_dev, void* stream) {
    if (n_spans < 0 || fetch_len < 0 || max_ref_len < 0 ||
        (n_spans > 0 && max_ref_len > 0 && (!span_start_dev || !region_ref_off_dev || !region_ref_len_dev || !ref_dev || (fetch_len > 0 && !fetched_dev))))
        return pv::set_error(PV_EINVAL, "pv_bam_gather_reference: bad arguments");
    if (n_spans > 65535) return pv::set_error(PV_EINVAL, "pv_bam_gather_reference: at most 65535 regions per call");
    if (int rc = pv::require_device()) return rc;
    if (n_spans == 0 || max_ref_len == 0) return PV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    pv::prof_begin(pv::FAM_INGEST, st);
    unsigned gx = grid_for(max_ref_len, 256 * 8);
    if (gx > 1024) gx = 1024;
    gather_reference_kernel<<<dim3(gx, (unsigned)n_spans), 256, 0, st>>>(fetched_dev, fetch_len, fetch_start, span_start_dev, region_ref_off_dev,
                                                                         region_ref_len_dev, ref_dev);
    pv::prof_end(pv::FAM_INGEST, st, 1);
    PV_CUDA_CHECK(cudaGetLastError());
    return PV_OK;
}
