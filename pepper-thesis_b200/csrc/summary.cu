// Pileup summary on sm_100a: the B200-native replacement of RegionalSummaryGenerator::generate_summary
// (/root/reference/pepper_variant/modules/cpp/region_summary.cpp:337-916, inference mode).
//
// Kernel chain (all on one stream, no host synchronisation):
//   K0 read_entries_kernel / tile_scan_kernel / entry_scatter_kernel
//                            warp per read, ONE pass over the CIGAR: reference span (with the REF_SKIP/PAD -> SOFT_CLIP
//                            fall-through of region_summary.cpp:556-561) and one ENTRY per (tile, read) -- the op the walk
//                            of that tile starts at, with the tile-local position and the read index in front of it --
//                            pooled, counted per tile, scanned, scattered into per-tile lists. K1 never searches a
//                            CIGAR and no per-op prefix array ever goes to HBM.
//   K1 pileup_tile_kernel    CTA per tile of P reference positions of one region, counters in SHARED memory (16 packed
//                            words per position: forward strand in the low half, reverse in the high half). Phase A:
//                            a warp per entry stages up to 128 CIGAR ops at a time in its shared-memory table (coalesced
//                            loads, four ops per lane, one warp scan for positions and read indices); pass 1, lane per
//                            op: run ends into difference arrays, insert/delete anchors; pass 2, lane per 16-byte-aligned
//                            SLICE of the read's bases: each lane finds its first op by bisection in the table and walks
//                            the match runs that cross its slice, 16 bases per load, byte-parallel compares against the
//                            tile's reference -- a matching base costs no atomic, only mismatches and low-quality bases
//                            do. Phase B: prefix sums turn the difference arrays into counts, then the clamped int16
//                            image rows (flushed to HBM), the site thresholds in fp64 exactly like :634-646, and the
//                            candidate sites. Phase C re-walks only the CIGAR ops (no bases) and records the
//                            insert/delete alleles of the registered sites.
//   K2 site_allele_kernel    warp per site: exact de-duplication of the recorded alleles (byte-wise compares, no
//                            hashing), per-allele filters of :682-712, one candidate record per survivor with a
//                            64-bit order key (position, type, allele rank).
//   cub radix sort           of the candidate keys -> the reference's emission order (:669-670).
//   K3 emit_window_kernel    warp per candidate in sorted order: 33x26 int16 window from the dense image with
//                            the centre-row / deletion-span overrides of :848-905, plus position/depth/... .
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cstdlib>
#include <vector>

namespace {

constexpr int NC = 16;           // packed 32-bit counter words per position
#ifndef PV_K1_THREADS
#define PV_K1_THREADS 512
#define PV_K1_PMAX 1216
#define PV_K1_MINB 2
#endif
constexpr int K1_THREADS = PV_K1_THREADS;
constexpr int K1_WARPS = K1_THREADS / 32;
constexpr int MAX_PPT = 4;       // flush: positions per thread  => P <= MAX_PPT * K1_THREADS
constexpr int P_MAX = PV_K1_PMAX;   // 2 CTAs per SM: 16 words x 1216 positions = 76 KB of counters each
constexpr int FREQ_TABLE = 4096; // coverages with a precomputed site-threshold entry (deeper positions divide)
constexpr int K0_CHUNK = 4;      // reads a K0 warp takes per ticket
#ifndef PV_K1_OPL
#define PV_K1_OPL 3
#endif
constexpr int OPL = PV_K1_OPL;   // consecutive CIGAR ops a lane holds in a round
constexpr int ROUND_OPS = 32 * OPL;   // CIGAR ops a warp stages per round
constexpr int QCAP = 63;         // exceptions (mismatches, low-quality bases) a warp queues per round before it counts them
constexpr int WARP_SCRATCH = 4 * ROUND_OPS + 1 + QCAP;   // ints of per-warp scratch: the round's op words, its table of match
                                 // pieces / inserts / deletes (three ints per entry), the exception queue
constexpr int REF_PAD = 16;      // bytes in front of the tile's reference copy (a 16-base chunk may start before the tile)
constexpr int WIN_ELEMS = PV_WINDOW * PV_FEATURES;   // 858

// Counter words (forward strand in the low 16 bits, reverse strand in the high 16 bits unless noted; arithmetic is
// modulo 2^32 on the packed word, so a field may go "negative" as long as its final value is in range). The common
// case -- an aligned base that passes the quality threshold and equals the reference byte -- costs NO atomic at all:
// aligned coverage is kept as a DIFFERENCE array (+1 where a match run enters the tile, -1 behind its last base; bases
// below the quality threshold subtract themselves the same way) that the flush turns into counts with a prefix sum, and
// everything the reference derives from an aligned base (coverage, REFF/REFR, the base's own feature) is
// reconstructed from it:
//   coverage = T_f + T_r + bump      REF = T - SKIP      BASE[class(ref)] = T - DEV      BASE[c != class(ref)] = CLS[c]
enum { C_T = 0,        // difference array -> aligned bases with q >= min_snp_baseq
       C_SKIP = 1,     // ... of which anchor an insert/delete (no REFF/REFR decrement, :381-391)
       C_DEV = 2,      // ... of which have a base class different from the (valid) reference base's class
       C_CLS = 3,      // +0..5 = A C G T I D: explicit class counts (deviating bases, insert/delete anchors)
       C_DELD = 9,     // difference array -> deleted spans ('*', :542-552); sits where class 6 would
       C_COV2 = 10,    // low: insert-anchored coverage bumps (:453-454); high: snp_count of bases that are no dense SNP allele
       C_INSDEL = 11,  // low: insert_count; high: delete_count
       C_SNP = 12 };   // +0..3 = "1A" "1C" "1G" "1T" allele counts
enum { CTR_SITES = 0, CTR_EVENTS = 1, CTR_CANDS = 2, CTR_STATUS = 3, CTR_K0_TICKET = 4, CTR_K2_TICKET = 5, CTR_ENTRIES = 6, CTR_COUNT = 8 };
enum { ST_SITE_OVF = 1, ST_EVENT_OVF = 2, ST_CAND_OVF = 4, ST_INTERNAL = 8,
       ST_NEED_QUALS = 16,     // a quality was needed (insert over the read's end) but the batch came without its quality array
       ST_ENTRY_OVF = 32 };    // more (tile, read) entries than the workspace was sized for (cannot happen with pv_summary_workspace_bytes' bound)
enum { PF_SITE = 1, PF_SNP = 2, PF_INS = 4, PF_DEL = 8, PF_OTHER = 16 };

struct SiteRec {
    int64_t gpos;          // dense position index (region-major)
    int32_t region;
    int32_t local;         // position - region_ref_start
    int32_t cov;
    int32_t flags;         // PF_*
    int32_t ev_off;        // first event slot
    int32_t n_ev;          // events this site will receive
    int32_t fill;          // atomic cursor (phase C)
    uint16_t snp[8];       // "1A","1C","1G","1T" x {fwd, rev}
    int32_t pad;
};

struct Event {             // one recorded allele observation
    int64_t ptr;           // byte offset of the allele's first byte in bases[] (type 1,2) or ref[] (type 3)
    uint32_t info;         // type | rev << 2 | elen << 8
    uint32_t pad;
};

struct CandRec {
    int64_t ptr;           // allele bytes (see Event); for dense SNP alleles: the byte itself
    int32_t site;
    uint32_t info;         // type | dense_snp << 2 | elen << 8
    int32_t nf, nr;
};

// One (tile, read) pair of the work lists K0 builds: where the walk of that tile starts in the read's CIGAR.
struct TileEntry {
    int32_t read;          // read index in the batch
    int32_t k;             // first op of the walk: the op in front of the first op that starts inside the tile (0 if none)
    int32_t a;             // tile-local reference position at which op k starts (negative: the op starts in front of the tile)
    int32_t ri;            // read index in front of op k
};

struct SumParams {
    PvReadBatch b;
    const int64_t* pos_off;       // [n_regions + 1] dense position offset of each region
    const int32_t* tile_region;   // [n_tiles]
    const int32_t* tile_start;    // [n_tiles] first region-relative position of the tile
    const int32_t* tile_base;     // [n_regions + 1] first tile of each region
    int32_t P;                    // tile size
    int32_t n_tiles;
    int32_t* read_region;         // [n_reads] region of each read (K0)
    int32_t* tile_count;          // [n_tiles] entries per tile (K0 count pass)
    int32_t* tile_fill;           // [n_tiles][2] fill cursors from the front / from the back (K0 fill pass)
    int32_t* tile_off;            // [n_tiles + 1] exclusive prefix of tile_count
    TileEntry* entries; int32_t entry_cap;
    void* pool; int32_t* pool_tile;   // [entry_cap] entries in the order K0a found them + the tile each belongs to
    int32_t allq;                 // 1: the batch promises (PvReadBatch.min_qual) that no base quality is below either threshold:
                                  // every quality test passes, qualities are never loaded
    int32_t* read_span;           // [n_reads] total reference advance of the read's CIGAR
    int16_t* img;                 // [total_positions][26]; rows are only valid where a candidate window can read them
    int32_t img_all;              // 1 = write every row (dense-image parity hook)
    SiteRec* sites; int32_t site_cap;
    Event* events; int32_t ev_cap;
    CandRec* cands; unsigned long long* cand_key; int32_t cand_cap;
    int32_t* ctr;                 // CTR_*
    ushort4* freq_min;            // [FREQ_TABLE] smallest snp / insert / delete count that passes its frequency threshold at coverage c
    int32_t qthr;                 // ceil(min_snp_baseq): q >= min_snp_baseq  <=>  q >= qthr for integer q
    PvThresholds t;
};

__device__ __forceinline__ bool valid_ref(uint8_t c) {          // check_ref_base, region_summary.cpp:193-199
    c |= 0x20; return c == 'a' || c == 'c' || c == 'g' || c == 't';
}
__device__ __forceinline__ uint8_t upc(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }
__device__ __forceinline__ int base_class(uint8_t b) {          // get_feature_index, :201-230 (0..6 = A C G T I D *)
    switch (upc(b)) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3;
                      case 'I': return 4; case 'D': return 5; default: return 6; }
}
__device__ __forceinline__ int ref_value(uint8_t b) {           // get_reference_feature_value, :165-172
    switch (upc(b)) { case 'A': return 1; case 'C': return 2; case 'G': return 3; case 'T': return 4; default: return 5; }
}
__device__ __forceinline__ int acgt_code(uint8_t b) {           // upper-case only: dense SNP allele slot
    return b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : b == 'T' ? 3 : -1;
}
__device__ __forceinline__ bool is_match_op(int op) { return (0x181u >> op) & 1u; }     // M, =, X
__device__ __forceinline__ int min125(int v) { return v < 125 ? v : 125; }

// ------------------------------------------------------------------------------------------------------------
// K0: per-tile work lists
// ------------------------------------------------------------------------------------------------------------
// :357-563 -- what an op adds to ref_position / read_index (REF_SKIP and PAD also advance the read: the missing `break`)
__device__ __forceinline__ void op_advance(uint32_t w, int& ra, int& qa) {
    const int op = (int)(w & 15u), len = (int)(w >> 4);
    ra = ((0x1cdu >> op) & 1u) ? len : 0;                     // M D N P = X
    qa = ((0x1dbu >> op) & 1u) ? len : 0;                     // M I N S P = X
}

// four consecutive CIGAR words starting at the 16-byte aligned slot `co_al + k` of a read whose ops occupy slots
// [skip, n_slots): words outside the read come back as 0 (an empty match)
__device__ __forceinline__ void load_ops4(const PvReadBatch& b, int64_t co_al, int k, int skip, int n_slots, uint32_t ws[4]) {
    uint4 w = make_uint4(0u, 0u, 0u, 0u);
    if (k + 3 < n_slots && co_al + k + 3 < b.n_ops) w = __ldg((const uint4*)(b.cigar + co_al + k));
    else {
        if (k + 0 < n_slots) w.x = b.cigar[co_al + k + 0];
        if (k + 1 < n_slots) w.y = b.cigar[co_al + k + 1];
        if (k + 2 < n_slots) w.z = b.cigar[co_al + k + 2];
    }
    ws[0] = w.x; ws[1] = w.y; ws[2] = w.z; ws[3] = w.w;
#pragma unroll
    for (int j = 0; j < 4; j++) if (k + j < skip || k + j >= n_slots) ws[j] = 0u;
}

// K0a, warp per read (K0_CHUNK reads per ticket: op counts are heavy-tailed, a static split leaves the kernel waiting for the
// unluckiest warp): ONE pass over the read's CIGAR -- each lane takes four consecutive ops per step (one 16-byte load),
// two shuffle scans give every op its start -- yields the read's reference span and one entry per tile the read touches:
// the walk of tile t starts at the op in front of the first op that starts at or behind the tile's first position, i.e.
// the one op k with  a_k < t * P <= a_k + (its reference advance)  (a_k = op start, region-relative); tiles that begin
// at or in front of the read's own start begin at op 0. Touched = aligned / deleted positions [rel, rel + span - 1] and
// insert / delete anchors [rel - 1, rel + span - 1], clipped to the region. (A boundary exactly at the read's end gives
// an entry whose walk finds nothing: harmless, the span is not known yet when it is filed.) Entries are collected in a
// per-warp buffer and appended to a global pool 32 at a time (one atomic per 32 entries), tiles counted as they go.
struct PoolEntry { int32_t read, k, a, ri; };
__device__ __forceinline__ void pool_flush(const SumParams& p, const PoolEntry* buf, const int32_t* buf_tile, int n, int lane) {
    int base = 0;
    if (lane == 0) base = atomicAdd(&p.ctr[CTR_ENTRIES], n);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (lane < n) {
        if (base + lane < p.entry_cap) {
            ((int4*)p.pool)[base + lane] = *(const int4*)&buf[lane];
            p.pool_tile[base + lane] = buf_tile[lane];
            atomicAdd(&p.tile_count[buf_tile[lane]], 1);
        } else {
            atomicOr(&p.ctr[CTR_STATUS], ST_ENTRY_OVF);
        }
    }
    __syncwarp();
}

__global__ void read_entries_kernel(const SumParams p) {
    __shared__ __align__(16) PoolEntry s_buf[8][32];
    __shared__ int32_t s_tile[8][32];
    const PvReadBatch& b = p.b;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    PoolEntry* buf = s_buf[warp];
    int32_t* buf_tile = s_tile[warp];
    int n_buf = 0;                                            // warp-uniform
    const int64_t P = p.P;
    int reg = 0;
    for (int64_t r = 0, r_end = 0;; r++) {
        if (r >= r_end) {
            int tk = 0;
            if (lane == 0) tk = atomicAdd(&p.ctr[CTR_K0_TICKET], K0_CHUNK);
            r = __shfl_sync(0xffffffffu, tk, 0);
            if (r >= b.n_reads) break;
            r_end = r + K0_CHUNK < b.n_reads ? r + K0_CHUNK : b.n_reads;
            int lo = 0, hi = b.n_regions - 1;                 // region: first one whose read range ends behind r
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (b.region_read_begin[mid + 1] <= r) lo = mid + 1; else hi = mid; }
            reg = lo;
        }
        while (b.region_read_begin[reg + 1] <= r) reg++;
        const int n_ops = b.read_n_ops[r];
        const int64_t rel = b.read_pos[r] - b.region_ref_start[reg];
        const int64_t L = b.region_ref_end[reg] - b.region_ref_start[reg] + 1;
        const bool live = b.read_mapq[r] != 0 && n_ops > 0 && rel - 1 <= L - 1;      // :619
        const int t_last = (int)((L - 1) / P);                // last tile of the region
        const int tb = p.tile_base[reg];
        int tcur = 0;                                         // next tile boundary behind the read's start
        if (live && rel >= 0) {
            const int t0 = (int)((rel > 0 ? rel - 1 : 0) / P);
            int tf = (int)(rel / P);
            tcur = tf + 1;
            if (tf > t_last) tf = t_last;
            const int n_first = tf - t0 + 1;                  // 1 or 2 tiles begin at or in front of the read's start
            if (lane < n_first) {
                PoolEntry e; e.read = (int32_t)r; e.k = 0; e.a = (int32_t)(rel - (int64_t)(t0 + lane) * P); e.ri = 0;
                buf[n_buf + lane] = e; buf_tile[n_buf + lane] = tb + t0 + lane;
            }
            n_buf += n_first;
        }
        const int64_t co = b.read_cigar_off[r];
        const int64_t co_al = co & ~(int64_t)3;
        const int skip = (int)(co - co_al), n_slots = skip + n_ops;
        int64_t ref_run = 0, ri_run = 0, bpos = (int64_t)tcur * P;
        for (int k0 = 0; k0 < n_slots; k0 += 128) {
            const int k = k0 + lane * 4;
            uint32_t ws[4];
            load_ops4(b, co_al, k, skip, n_slots, ws);
            int ra[4], qa[4], sr = 0, sq = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) { op_advance(ws[j], ra[j], qa[j]); sr += ra[j]; sq += qa[j]; }
            int ir = sr, iq = sq;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int tr = __shfl_up_sync(0xffffffffu, ir, d), tq = __shfl_up_sync(0xffffffffu, iq, d);
                if (lane >= d) { ir += tr; iq += tq; }
            }
            const int64_t a_first = rel + ref_run + (int64_t)(ir - sr), q_first = ri_run + (int64_t)(iq - sq);
            ref_run += __shfl_sync(0xffffffffu, ir, 31);
            ri_run += __shfl_sync(0xffffffffu, iq, 31);
            while (live && tcur <= t_last && bpos <= rel + ref_run) {   // boundaries inside the ops of this step
                if (n_buf > 29) { __syncwarp(); pool_flush(p, buf, buf_tile, n_buf, lane); n_buf = 0; }
                int64_t a_k = a_first, q_k = q_first;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (ra[j] > 0 && a_k < bpos && bpos <= a_k + ra[j]) {
                        PoolEntry e; e.read = (int32_t)r; e.k = k + j - skip; e.a = (int32_t)(a_k - bpos);
                        e.ri = q_k > (1ll << 30) ? (1 << 30) : (int32_t)q_k;
                        buf[n_buf] = e; buf_tile[n_buf] = tb + tcur;
                    }
                    a_k += ra[j]; q_k += qa[j];
                }
                n_buf++; tcur++; bpos += P;
            }
        }
        if (lane == 0) { p.read_span[r] = ref_run > (1ll << 30) ? (1 << 30) : (int32_t)ref_run; p.read_region[r] = reg; }
        if (n_buf > 29) { __syncwarp(); pool_flush(p, buf, buf_tile, n_buf, lane); n_buf = 0; }
    }
    __syncwarp();
    if (n_buf > 0) pool_flush(p, buf, buf_tile, n_buf, lane);
}

// exclusive prefix of the tile counts (one block)
__global__ void tile_scan_kernel(const SumParams p) {
    __shared__ int64_t s_part[1024];
    const int tid = threadIdx.x, n = p.n_tiles;
    const int per = (n + 1023) / 1024;
    const int lo = tid * per, hi = lo + per < n ? lo + per : n;
    int64_t sum = 0;
    for (int i = lo; i < hi; i++) sum += p.tile_count[i];
    s_part[tid] = sum;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const int64_t v = tid >= d ? s_part[tid - d] : 0;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    int64_t run = s_part[tid] - sum;
    for (int i = lo; i < hi; i++) { p.tile_off[i] = (int32_t)run; run += p.tile_count[i]; }
    if (tid == 0) p.tile_off[n] = (int32_t)s_part[1023];
}

// K0b, thread per pool entry: the entry goes into its tile's list. Entries of reads that cross the whole tile are filed
// from the front, reads that start or end inside it (less work) from the back: the warps of a CTA pull the long units
// first and finish closer together.
__global__ void entry_scatter_kernel(const SumParams p) {
    if (p.ctr[CTR_STATUS] & ST_ENTRY_OVF) return;
    const PvReadBatch& b = p.b;
    int n = p.ctr[CTR_ENTRIES];
    if (n > p.entry_cap) n = p.entry_cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 e = ((const int4*)p.pool)[i];
        const int tile = p.pool_tile[i];
        const int reg = p.read_region[e.x];
        const int64_t rel = b.read_pos[e.x] - b.region_ref_start[reg];
        const int64_t t_lo = (int64_t)(tile - p.tile_base[reg]) * p.P;
        const bool whole = rel <= t_lo && rel + (int64_t)p.read_span[e.x] >= t_lo + p.P;
        const int slot = whole ? p.tile_off[tile] + atomicAdd(&p.tile_fill[2 * tile], 1)
                               : p.tile_off[tile + 1] - 1 - atomicAdd(&p.tile_fill[2 * tile + 1], 1);
        if (slot >= p.tile_off[tile] && slot < p.tile_off[tile + 1]) ((int4*)p.entries)[slot] = e;
        else atomicOr(&p.ctr[CTR_STATUS], ST_INTERNAL);
    }
}

// The site test (:634-646) compares count / max(1, coverage) with a frequency threshold in fp64. For a fixed coverage
// that is a threshold on the integer count: the smallest count whose quotient -- the same fp64 division -- reaches the
// threshold. One table entry per coverage spares the tile kernel three divisions per position.
__global__ void freq_table_kernel(const SumParams p) {
    const int cov = blockIdx.x * blockDim.x + threadIdx.x;
    if (cov >= FREQ_TABLE) return;
    const double cv = (double)cov > 1.0 ? (double)cov : 1.0;
    const double thr[3] = {p.t.snp_freq, p.t.insert_freq, p.t.delete_freq};
    unsigned short m[3];
    for (int k = 0; k < 3; k++) {
        int lo = 0, hi = 65535;                              // counts are 16-bit fields; 65535 = no count passes
        if (!((double)hi / cv >= thr[k])) lo = hi;           // also NaN thresholds: nothing passes, as in the reference's >=
        while (lo < hi) { const int mid = (lo + hi) >> 1; if ((double)mid / cv >= thr[k]) hi = mid; else lo = mid + 1; }
        m[k] = (unsigned short)lo;
    }
    p.freq_min[cov] = make_ushort4(m[0], m[1], m[2], 0);
}

// ------------------------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------------------------
struct TileCtx {
    uint32_t* cnt;        // [NC][P] shared
    uint8_t* ref_s;       // [P] shared: reference bytes of the tile
    uint8_t* rcls;        // [P] shared: class 0..3 of the reference base, 0xff when it is not A/C/G/T (any case)
    uint8_t* lut;         // [256] shared: base byte -> class 0..6 | 8 if the byte is an upper-case A/C/G/T
    uint8_t* pflag;       // [P] shared: PF_* of each tile position (phase B onwards)
    uint8_t* near;        // [P] shared: 1 = the position's image row is needed by a window (a site within 16 positions)
    int32_t* site_slot;   // [P] shared (aliases cnt after phase B): site index or -1
    int32_t* scratch;     // this warp's shared op table: [3][TBL_STRIDE] (op word, start position, read index)
    int P;
    int n_valid;          // positions of the tile that exist in the region
    int region;
    int64_t t_lo;         // first tile position, relative to region_ref_start
    int64_t L;            // region length
    int64_t ref_len;      // reference_sequence.length()
    int64_t ref_off;      // offset of the region's reference in ref[]
};

__device__ __forceinline__ void record_event(const SumParams& p, int s, int type, int rev, int elen, int64_t ptr) {
    SiteRec* sr = &p.sites[s];
    const int slot = atomicAdd(&sr->fill, 1);
    if (slot < sr->n_ev) {
        const int idx = sr->ev_off + slot;
        if (idx < p.ev_cap) {
            Event e; e.ptr = ptr; e.info = (uint32_t)type | ((uint32_t)rev << 2) | ((uint32_t)elen << 8); e.pad = 0;
            p.events[idx] = e;
        }
    } else {
        atomicOr(&p.ctr[CTR_STATUS], ST_INTERNAL);
    }
}

// Per-read constants of one work unit, in TILE-LOCAL int32 coordinates (position 0 = first position of the tile).
struct ReadCtx {
    int64_t rel_t;          // read start relative to the tile
    int64_t co, bo;         // first op / first base of the read
    int nv, l_end;          // tile = [0, nv); last region position (an op that starts beyond it is never reached, :355)
    int n_ops, read_len;
    uint32_t rev, inc, dec; // strand; +1 / -1 in the strand's half of a packed counter word
    const uint8_t* quals; const uint8_t* bases;
};

__device__ __forceinline__ ReadCtx make_read_ctx(const SumParams& p, const TileCtx& c, int64_t r) {
    const PvReadBatch& b = p.b;
    ReadCtx x;
    x.rel_t = b.read_pos[r] - b.region_ref_start[c.region] - c.t_lo;
    x.nv = c.n_valid;
    const int64_t l_end64 = c.L - 1 - c.t_lo;
    x.l_end = l_end64 > (1 << 30) ? (1 << 30) : (int)l_end64;
    x.co = b.read_cigar_off[r];
    x.n_ops = b.read_n_ops[r];
    x.read_len = b.read_len[r];
    x.bo = b.read_base_off[r];
    x.rev = b.read_flags[r] & 1u;
    x.inc = x.rev ? 0x10000u : 1u;
    x.dec = 0u - x.inc;
    x.quals = b.quals + x.bo;
    x.bases = b.bases + x.bo;
    return x;
}

// sum of the insert's n = len + 1 qualities starting at the anchor base (:448-450) -> does it pass (:452)?
// qa = quality of the anchor base (already loaded by the caller)
__device__ __forceinline__ bool insert_quality_pass(const SumParams& p, const ReadCtx& x, int ori, int n, int qa) {
    const int i0 = ori - 1;
    if (p.allq && i0 + n <= x.read_len) return true;         // n qualities, each >= min_indel_baseq
    if (p.b.quals == nullptr) {                              // the promise alone cannot decide: the caller re-runs with qualities
        atomicOr(&p.ctr[CTR_STATUS], ST_NEED_QUALS);
        return true;
    }
    const int m = (i0 + n < x.read_len ? i0 + n : x.read_len) - i0;   // qualities that exist (>= 1)
    int64_t bq = qa;
    if (m <= 4) {                                            // the common short insert: independent byte loads
        const int q1 = m > 1 ? (int)__ldg(x.quals + i0 + 1) : 0, q2 = m > 2 ? (int)__ldg(x.quals + i0 + 2) : 0,
                  q3 = m > 3 ? (int)__ldg(x.quals + i0 + 3) : 0;
        bq += q1 + q2 + q3;
    } else {
        for (int i = 1; i < m; i++) bq += x.quals[i0 + i];
    }
    return (double)bq >= p.t.min_indel_baseq * (double)n;
}

// high bit of each byte set where the quality byte is BELOW the threshold (q < qthr, qthr in [0, 256])
__device__ __forceinline__ uint32_t lowq_mask(uint32_t q, int qthr, uint32_t thr4) {
    if (qthr <= 128) {
        const uint32_t s = ((q & 0x7f7f7f7fu) | 0x80808080u) - thr4;   // per byte: 0x80 + (q & 127) - qthr, never borrows
        return ~(s | q) & 0x80808080u;
    }
    if (qthr >= 256) return 0x80808080u;
    return __vcmpltu4(q, thr4) & 0x80808080u;
}
// high bit of each byte set where the byte is non-zero
__device__ __forceinline__ uint32_t nonzero_mask(uint32_t x) {
    return (((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
}
// bits 7, 15, 23, 31 -> bits 0..3
__device__ __forceinline__ uint32_t movemask4(uint32_t m) { return (((m >> 7) * 0x01020408u) >> 24) & 0xfu; }

// A base that passes the quality threshold and differs from the reference byte (raw compare, :394-425):
//   snp_count++;  class(base) != class(valid ref): DEV++, CLS[class]++;  base is an upper-case A/C/G/T: SNP[base]++.
// Nearly always all three happen together, so the dense SNP allele counter stands for all of them (ONE atomic) and the
// flush adds it back: snp_count = X + sum SNP, DEV = DEVX + sum SNP, CLS[c] = CLSX[c] + SNP[c]; the rare other
// combinations correct the explicit counters (which may go "negative" modulo 2^32 in between).
__device__ __forceinline__ void count_mismatch(const TileCtx& c, int pl, uint8_t base, uint32_t inc) {
    const int rc = c.rcls[pl];
    const int lu = c.lut[base];                               // class | dense-allele flag << 3
    const int cb = lu & 7;
    const bool dev = rc != 0xff && cb != rc;
    if (lu & 8) {
        atomicAdd(&c.cnt[(C_SNP + cb) * c.P + pl], inc);
        if (!dev) { atomicAdd(&c.cnt[C_DEV * c.P + pl], 0u - inc); atomicAdd(&c.cnt[(C_CLS + cb) * c.P + pl], 0u - inc); }
    } else {
        atomicAdd(&c.cnt[C_COV2 * c.P + pl], 0x10000u);      // snp_count of the bases that are no dense allele
        if (dev) {
            atomicAdd(&c.cnt[C_DEV * c.P + pl], inc);
            if (cb < 6) atomicAdd(&c.cnt[(C_CLS + cb) * c.P + pl], inc);
            else { atomicAdd(&c.cnt[C_DELD * c.P + pl], inc); if (pl + 1 < c.n_valid) atomicAdd(&c.cnt[C_DELD * c.P + pl + 1], 0u - inc); }
        }
    }
}

// ---- the warp's round table ------------------------------------------------------------------------------------------
// A round stages the (up to) 128 ops [kb, kb + 128) of the read: the words arrive coalesced (lane + 32 u), are transposed
// through shared memory so that every lane holds four CONSECUTIVE ops, and one pair of shuffle scans gives every op its
// tile-local start position and the read index in front of it. The same converged code sorts the ops into ONE table of
// 128 three-word entries -- an op is a match run, an insert, a delete or none of them:
//   match pieces   [0, n_m)               read index of the first base | tile position of it + length << 16 | 16-byte
//                                         sub-pieces in front of the piece
//   inserts        [n_m, n_m + n_ins)     tile-local start | read index in front | length << 1 + "follows a match run"
//   deletes        from the back          the same
// so that pass 1 walks inserts and deletes lane per op without diverging on the op type, and pass 2 walks the sub-pieces
// lane per 16-byte chunk. REF_SKIP / PAD and match bases behind the read's end (both rare) take themselves out of the
// coverage on the spot.
struct OpTable {
    uint32_t* w;          // [ROUND_OPS] op words of the round (0 behind the read's last op)
    int32_t* e0;          // [ROUND_OPS] entry word 0
    int32_t* e1;          // [ROUND_OPS] entry word 1
    int32_t* e2;          // [ROUND_OPS] entry word 2
    int32_t* qn;          // exception queue: fill count ...
    uint32_t* q;          // ... and [QCAP] entries: tile position | base << 16 | low-quality flag << 31
};
__device__ __forceinline__ OpTable op_table(const TileCtx& c) {
    OpTable t; t.w = (uint32_t*)c.scratch; t.e0 = c.scratch + ROUND_OPS; t.e1 = c.scratch + 2 * ROUND_OPS;
    t.e2 = c.scratch + 3 * ROUND_OPS; t.qn = c.scratch + 4 * ROUND_OPS; t.q = (uint32_t*)(t.qn + 1);
    return t;
}
__device__ __forceinline__ void fetch_round(const PvReadBatch& b, const ReadCtx& x, int kb, int lane, uint32_t wv[OPL]) {
#pragma unroll
    for (int u = 0; u < OPL; u++) {
        const int j = kb + lane + 32 * u;
        wv[u] = j < x.n_ops ? __ldg(b.cigar + x.co + j) : 0u;
    }
}
struct Round {
    int cnt;              // staged ops that start at a tile position <= nv (they form a prefix: op starts never decrease; an
                          // insert / delete that starts right behind the tile still anchors on its last position)
    int n_m, n_ins, n_del;   // match pieces / inserts / deletes among them
    int n_sub;            // 16-byte sub-pieces of the match pieces
    int a_next, ri_next;  // state behind op 127
    uint32_t w_last;      // op 127
};
// the part [s, e) of a reference span that lies in the tile leaves the coverage difference array
__device__ __forceinline__ void uncover(const TileCtx& c, const ReadCtx& x, int s, int e) {
    if (s < 0) s = 0;
    if (e > x.nv) e = x.nv;
    if (e <= s) return;
    atomicAdd(&c.cnt[C_T * c.P + s], x.dec);
    if (e < x.nv) atomicAdd(&c.cnt[C_T * c.P + e], x.inc);
}
// ACC: phase A (match pieces filed, coverage corrections applied); else phase C (inserts and deletes only)
template <bool ACC>
__device__ __forceinline__ Round stage_round(const TileCtx& c, const ReadCtx& x, const OpTable& t, const uint32_t wv[OPL], int kb,
                                             int a_run, int ri_run, uint32_t w_prev, int lane) {
#pragma unroll
    for (int u = 0; u < OPL; u++) t.w[lane + 32 * u] = wv[u];
    __syncwarp();
    uint32_t ws[OPL];
#pragma unroll
    for (int j = 0; j < OPL; j++) ws[j] = t.w[OPL * lane + j];
    int ra[OPL], qa[OPL], sr = 0, sq = 0;
#pragma unroll
    for (int j = 0; j < OPL; j++) { op_advance(ws[j], ra[j], qa[j]); sr += ra[j]; sq += qa[j]; }
    int ir = sr, iq = sq;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int tr = __shfl_up_sync(0xffffffffu, ir, d), tq = __shfl_up_sync(0xffffffffu, iq, d);
        if (lane >= d) { ir += tr; iq += tq; }
    }
    uint32_t pw = __shfl_up_sync(0xffffffffu, ws[OPL - 1], 1);   // the op in front of this lane's first one
    if (lane == 0) pw = w_prev;
    int pr = a_run + ir - sr, pq = ri_run + iq - sq;       // state in front of this lane's first op
    int f0[OPL], f1[OPL], f2[OPL];                               // entry words of this lane's ops
    uint32_t kind = 0;                                     // per op: 1 match piece, 2 insert, 3 delete
    int v = 0, n_m = 0, n_i = 0, n_d = 0, nsp = 0;
#pragma unroll
    for (int j = 0; j < OPL; j++) {
        const int op = (int)(ws[j] & 15u), len = (int)(ws[j] >> 4);
        f0[j] = pr; f1[j] = pq; f2[j] = (len << 1) | ((is_match_op((int)(pw & 15u)) && (pw >> 4) != 0u) ? 1 : 0);
        if (kb + OPL * lane + j < x.n_ops && pr <= x.nv) {
            v++;
            if (op == 1) { n_i++; kind |= 2u << (2 * j); }
            else if (op == 2) { n_d++; kind |= 3u << (2 * j); }
            else if (ACC && pr <= x.l_end) {                 // an op that starts beyond ref_end is never reached (:355)
                if (is_match_op(op)) {
                    const int i_lo = pr < 0 ? -pr : 0;
                    int i_hi = x.nv - pr; if (i_hi > len) i_hi = len;
                    if (x.read_len - pq < i_hi) {            // rare: the run continues behind the read's last base; those bases are absent
                        const int have = x.read_len - pq > 0 ? x.read_len - pq : 0;
                        uncover(c, x, pr + have, pr + len);
                        i_hi = have < i_hi ? have : i_hi;
                    }
                    if (i_hi > i_lo) {
                        const int ri0 = pq + i_lo, nb = i_hi - i_lo;
                        f0[j] = ri0; f1[j] = (pr + i_lo) | (nb << 16); f2[j] = nsp;     // f2: sub-pieces of this lane in front (completed below)
                        nsp += ((ri0 + nb - 1) >> 4) - (ri0 >> 4) + 1;
                        n_m++; kind |= 1u << (2 * j);
                    }
                } else if ((op == 3 || op == 6) && len > 0) { // REF_SKIP / PAD advance the reference without aligning a base
                    uncover(c, x, pr, pr + len);
                }
            }
        }
        pw = ws[j];
        pr += ra[j]; pq += qa[j];
    }
    // where this lane's entries go: behind those of the lanes in front of it
    const uint32_t pk = (uint32_t)v | ((uint32_t)n_m << 8) | ((uint32_t)n_i << 16) | ((uint32_t)n_d << 24);
    uint32_t incl = pk; int isub = nsp;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t tv = __shfl_up_sync(0xffffffffu, incl, d); const int ts = __shfl_up_sync(0xffffffffu, isub, d);
        if (lane >= d) { incl += tv; isub += ts; }
    }
    const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
    Round r;
    r.cnt = tot & 0xff; r.n_m = (tot >> 8) & 0xff; r.n_ins = (tot >> 16) & 0xff; r.n_del = (tot >> 24) & 0xff;
    r.n_sub = __shfl_sync(0xffffffffu, isub, 31);
    int at_m = ((incl - pk) >> 8) & 0xff, at_i = r.n_m + (((incl - pk) >> 16) & 0xff), at_d = ROUND_OPS - 1 - (((incl - pk) >> 24) & 0xff);
    const int sub0 = isub - nsp;
#pragma unroll
    for (int j = 0; j < OPL; j++) {
        const uint32_t kd = (kind >> (2 * j)) & 3u;
        if (kd) {
            const int at = kd == 1u ? at_m++ : kd == 2u ? at_i++ : at_d--;
            t.e0[at] = f0[j]; t.e1[at] = f1[j]; t.e2[at] = kd == 1u ? f2[j] + sub0 : f2[j];
        }
    }
    r.a_next = __shfl_sync(0xffffffffu, pr, 31);
    r.ri_next = __shfl_sync(0xffffffffu, pq, 31);
    r.w_last = __shfl_sync(0xffffffffu, ws[OPL - 1], 31);
    __syncwarp();
    return r;
}

// an aligned base that needs counting: queued, so that the warp counts a round's exceptions lane-parallel
__device__ __forceinline__ void count_exception(const TileCtx& c, const ReadCtx& x, uint32_t entry) {
    const int pl = (int)(entry & 0xffffu);
    if (entry >> 31) {                                       // below the quality threshold: the base leaves the coverage
        atomicAdd(&c.cnt[C_T * c.P + pl], x.dec);
        if (pl + 1 < x.nv) atomicAdd(&c.cnt[C_T * c.P + pl + 1], x.inc);
    } else {
        count_mismatch(c, pl, (uint8_t)(entry >> 16), x.inc);
    }
}
__device__ __forceinline__ void push_exception(const TileCtx& c, const ReadCtx& x, const OpTable& t, uint32_t entry) {
    const int slot = atomicAdd(t.qn, 1);
    if (slot < QCAP) t.q[slot] = entry;
    else count_exception(c, x, entry);                       // queue full: count it here
}

// Pass 2 of a round: the match pieces are cut at the read's 16-byte boundaries into SUB-PIECES (a piece of 33 bases has
// 3 of them), one per lane. A lane loads its aligned 16-base chunk of bases and qualities (two 16-byte loads,
// neighbouring lanes mostly neighbouring chunks), lines the tile's reference bytes up with the chunk (five aligned
// shared-memory words + funnel shifts) and tests all 16 bases at once with byte-parallel arithmetic. The code is
// straight-line for every lane; only EXCEPTIONS cost anything further, and they are queued and counted lane-parallel:
//   q <  min_snp_baseq : the base does not count (:378) -> it takes itself out of the coverage difference array
//   base != reference  : snp_count, class deviation, SNP allele count (:394-425)
__device__ void scan_pieces(const SumParams& p, const TileCtx& c, const ReadCtx& x, const OpTable& t, int n_e, int n_sub, int lane) {
    const int32_t* t_beg = t.e0; const uint32_t* t_pc = (const uint32_t*)t.e1; const int32_t* t_sub = t.e2;
    const uint32_t thr4 = (uint32_t)(p.qthr > 255 ? 255 : p.qthr) * 0x01010101u;
    const int range_end = t_beg[n_e - 1] + (int)(t_pc[n_e - 1] >> 16);
    const bool vec_ok = x.bo + (((int64_t)range_end + 15) & ~(int64_t)15) <= p.b.n_bases;
    int e_first = 0;                                         // piece that holds sub-piece jb
    for (int jb = 0; jb < n_sub; jb += 32) {
        // sub-piece v belongs to piece  e_first + #{pieces whose first sub-piece is in (jb, v]}: the (at most 32) pieces
        // that start inside this block mark their start in a bit mask, one warp OR-reduction per block
        const int idx = e_first + 1 + lane;
        const int sv = idx < n_e ? t_sub[idx] - jb - 1 : -1;
        const unsigned mask = __reduce_or_sync(0xffffffffu, (sv >= 0 && sv < 32) ? (1u << sv) : 0u);
        const int e = e_first + __popc(mask & ((1u << lane) - 1u));
        e_first += __popc(mask);
        const int v = jb + lane;
        if (v < n_sub) {
            const int beg = t_beg[e];
            const uint32_t pc = t_pc[e];
            const int c0 = ((beg >> 4) + (v - t_sub[e])) << 4;  // first read index of the chunk
            uint4 ub, uq = make_uint4(0u, 0u, 0u, 0u);
            if (vec_ok) {
                ub = __ldg((const uint4*)(x.bases + c0));
                if (!p.allq) uq = __ldg((const uint4*)(x.quals + c0));
            } else {                                           // last read of a batch whose arrays are not padded
                uint32_t tb[4] = {0u, 0u, 0u, 0u}, tq[4] = {0u, 0u, 0u, 0u};
                for (int j = 0; j < 16 && x.bo + c0 + j < p.b.n_bases; j++) {
                    tb[j >> 2] |= (uint32_t)x.bases[c0 + j] << ((j & 3) * 8);
                    if (!p.allq) tq[j >> 2] |= (uint32_t)x.quals[c0 + j] << ((j & 3) * 8);
                }
                ub = make_uint4(tb[0], tb[1], tb[2], tb[3]); uq = make_uint4(tq[0], tq[1], tq[2], tq[3]);
            }
            int b0 = beg - c0, b1 = b0 + (int)(pc >> 16);      // the piece's bytes of the chunk: [b0, b1) clipped to [0, 16)
            const int rp = (int)(pc & 0xffffu) - b0;           // tile position of the chunk's byte 0 (>= -15)
            b0 = b0 < 0 ? 0 : b0;
            b1 = b1 > 16 ? 16 : b1;
            const uint32_t* sr = (const uint32_t*)(c.ref_s + (rp & ~3));   // ref_s has REF_PAD bytes in front
            const uint32_t r0 = sr[0], r1 = sr[1], r2 = sr[2], r3 = sr[3], r4 = sr[4];
            const int sh = (rp & 3) * 8;
            const uint32_t n0 = nonzero_mask(ub.x ^ __funnelshift_r(r0, r1, sh)), n1 = nonzero_mask(ub.y ^ __funnelshift_r(r1, r2, sh)),
                           n2 = nonzero_mask(ub.z ^ __funnelshift_r(r2, r3, sh)), n3 = nonzero_mask(ub.w ^ __funnelshift_r(r3, r4, sh));
            const uint32_t bm = (0xffffu >> (16 - b1)) & (0xffffu << b0);
            uint32_t lowq16 = 0;
            if (!p.allq) {
                const uint32_t l0 = lowq_mask(uq.x, p.qthr, thr4), l1 = lowq_mask(uq.y, p.qthr, thr4),
                               l2 = lowq_mask(uq.z, p.qthr, thr4), l3 = lowq_mask(uq.w, p.qthr, thr4);
                lowq16 = (movemask4(l0) | (movemask4(l1) << 4) | (movemask4(l2) << 8) | (movemask4(l3) << 12)) & bm;
            }
            uint32_t ex = ((movemask4(n0) | (movemask4(n1) << 4) | (movemask4(n2) << 8) | (movemask4(n3) << 12)) & bm) | lowq16;
            while (ex) {                                        // a few per cent of the bases
                const int j = __ffs(ex) - 1;
                ex &= ex - 1;
                const uint32_t base = __byte_perm(j & 8 ? ub.z : ub.x, j & 8 ? ub.w : ub.y, (uint32_t)(j & 7)) & 0xffu;
                push_exception(c, x, t, (uint32_t)(rp + j) | (base << 16) | (((lowq16 >> j) & 1u) << 31));
            }
        }
        __syncwarp();
    }
    const int nq = *t.qn < QCAP ? *t.qn : QCAP;                 // count the queued exceptions, a lane each
    for (int i = lane; i < nq; i += 32) count_exception(c, x, t.q[i]);
    __syncwarp();
    if (lane == 0) *t.qn = 0;
}

// Phase A work unit: one entry = one read's walk of this tile (whole warp); populate_summary_matrix, :337-566.
// Aligned coverage is a difference array: the read covers [its start, its start + span) of the tile, and everything that
// is not an aligned base with a passing quality takes itself out again -- deleted / skipped spans and bases behind the
// read's end per op, low-quality bases per base. Pass 1, lane per insert and lane per delete: anchor counting (the
// anchor's missing REFF/REFR decrement of :381-391 is charged by the insert/delete op that follows the run) and the
// deleted spans. Pass 2 (scan_pieces) compares the aligned bases.
__device__ void accumulate_entry(const SumParams& p, const TileCtx& c, const TileEntry e, int lane) {
    const PvReadBatch& b = p.b;
    const ReadCtx x = make_read_ctx(p, c, e.read);
    const OpTable t = op_table(c);
    int kb = e.k, a_run = e.a, ri_run = e.ri;
    uint32_t w_prev = kb > 0 ? __ldg(b.cigar + x.co + kb - 1) : 0u;   // the op in front of the round's first one
    uint32_t wv[OPL];
    fetch_round(b, x, kb, lane, wv);
    {   // the bases this walk will compare (about one per tile position from read index e.ri on) start their way from DRAM now
        const int64_t off = (int64_t)e.ri + 128 * lane;
        if (off < x.read_len && 128 * lane < x.nv + 128) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(x.bases + off));
            if (!p.allq) asm volatile("prefetch.global.L2 [%0];" ::"l"(x.quals + off));
        }
    }
    if (lane == 0) {                                          // the read's reference span inside the tile
        int64_t s = x.rel_t, en = x.rel_t + (int64_t)p.read_span[e.read];
        if (s < 0) s = 0;
        if (en > x.nv) en = x.nv;
        if (en > s) {
            atomicAdd(&c.cnt[C_T * c.P + (int)s], x.inc);
            if (en < x.nv) atomicAdd(&c.cnt[C_T * c.P + (int)en], x.dec);
        }
    }
    while (true) {
        const Round rd = stage_round<true>(c, x, t, wv, kb, a_run, ri_run, w_prev, lane);
        const bool more = rd.cnt == ROUND_OPS && kb + ROUND_OPS < x.n_ops;
        if (more) fetch_round(b, x, kb + ROUND_OPS, lane, wv);           // in flight while this round is worked on
        for (int i = lane; i < rd.n_ins; i += 32) {                      // ---- inserts, :431-490
            const int a = t.e0[rd.n_m + i], ori = t.e1[rd.n_m + i], f = t.e2[rd.n_m + i], len = f >> 1;
            const int ol = a - 1;                                        // anchor position
            if (ol < 0 || ol >= x.nv) continue;
            const bool anchor_ok = ori >= 1 && ori - 1 < x.read_len;     // the anchor base exists in the read
            const int qa = !anchor_ok ? 0 : (p.allq ? 255 : (int)x.quals[ori - 1]);
            // anchor rule (:381-391): the base in front of this op is the last base of a match run -> that base keeps
            // its REFF/REFR decrement for itself. Depends only on the op TYPE, not on whether this op is reached.
            if ((f & 1) && anchor_ok && qa >= p.qthr) atomicAdd(&c.cnt[C_SKIP * c.P + ol], x.inc);
            if (a <= x.l_end && anchor_ok) {                             // reached (:355)
                const int n = len + 1;                                   // :442
                int elen = n;                                            // substr truncation, :439
                if (elen > x.read_len - (ori - 1)) elen = x.read_len - (ori - 1);
                if (insert_quality_pass(p, x, ori, n, qa)) {
                    if (qa < p.qthr) atomicAdd(&c.cnt[C_COV2 * c.P + ol], 1u);   // :453-454
                    if (1 + elen <= 61) {                                // :461-464
                        if (c.rcls[ol] != 0xff) atomicAdd(&c.cnt[(C_CLS + 4) * c.P + ol], x.inc);
                        atomicAdd(&c.cnt[C_INSDEL * c.P + ol], 1u);
                    }
                }
            }
        }
        for (int i = lane; i < rd.n_del; i += 32) {                      // ---- deletes, :491-555
            const int a = t.e0[ROUND_OPS - 1 - i], ori = t.e1[ROUND_OPS - 1 - i], f = t.e2[ROUND_OPS - 1 - i], len = f >> 1;
            const bool reached = a <= x.l_end;
            const int ol = a - 1;
            if (ol >= 0 && ol < x.nv) {
                const bool anchor_ok = ori >= 1 && ori - 1 < x.read_len;
                const int qa = !anchor_ok ? 0 : (p.allq ? 255 : (int)x.quals[ori - 1]);
                if ((f & 1) && anchor_ok && qa >= p.qthr) atomicAdd(&c.cnt[C_SKIP * c.P + ol], x.inc);
                if (reached) {
                    const int64_t rem = c.ref_len - c.t_lo - ol;         // reference bytes from the anchor on
                    int elen = len + 1;                                  // substr truncation, :500
                    if ((int64_t)elen > rem) elen = (int)rem;
                    if (c.rcls[ol] != 0xff) atomicAdd(&c.cnt[(C_CLS + 5) * c.P + ol], x.inc);   // :497
                    if (1 + elen <= 61) atomicAdd(&c.cnt[C_INSDEL * c.P + ol], 0x10000u);       // :511-512
                }
            }
            if (reached) {                                               // deleted span, :542-552: '*' instead of an aligned base
                const int s0 = a < 0 ? 0 : a;
                const int e0 = a + len < x.nv ? a + len : x.nv;
                if (e0 > s0) {
                    atomicAdd(&c.cnt[C_DELD * c.P + s0], x.inc);
                    atomicAdd(&c.cnt[C_T * c.P + s0], x.dec);
                    if (e0 < x.nv) { atomicAdd(&c.cnt[C_DELD * c.P + e0], x.dec); atomicAdd(&c.cnt[C_T * c.P + e0], x.inc); }
                }
            }
        }
        if (rd.n_m > 0) scan_pieces(p, c, x, t, rd.n_m, rd.n_sub, lane);
        __syncwarp();
        if (!more) break;
        w_prev = rd.w_last;
        kb += ROUND_OPS; a_run = rd.a_next; ri_run = rd.ri_next;
    }
}

// Phase C work unit: re-walks the ops of one entry (whole warp, lane per op) and records the insert/delete alleles of
// registered sites (the AlleleFrequencyMap updates of :458-487 and :507-535, needed only where the site thresholds passed).
__device__ void record_entry(const SumParams& p, const TileCtx& c, const TileEntry e, int lane) {
    const PvReadBatch& b = p.b;
    const ReadCtx x = make_read_ctx(p, c, e.read);
    const OpTable t = op_table(c);
    int kb = e.k, a_run = e.a, ri_run = e.ri;
    uint32_t wv[OPL];
    fetch_round(b, x, kb, lane, wv);
    while (true) {
        const Round rd = stage_round<false>(c, x, t, wv, kb, a_run, ri_run, 0u, lane);
        const bool more = rd.cnt == ROUND_OPS && kb + ROUND_OPS < x.n_ops;
        if (more) fetch_round(b, x, kb + ROUND_OPS, lane, wv);
        for (int i = lane; i < rd.n_ins + rd.n_del; i += 32) {
            const bool ins = i < rd.n_ins;
            const int at = ins ? rd.n_m + i : ROUND_OPS - 1 - (i - rd.n_ins);
            const int a = t.e0[at], len = t.e2[at] >> 1;
            if (a > x.l_end) continue;
            const int ol = a - 1;
            if (ol < 0 || ol >= x.nv) continue;
            const int s = c.site_slot[ol];
            if (s < 0) continue;
            if (ins) {
                const int ori = t.e1[at];
                if (!(c.pflag[ol] & PF_INS) || ori < 1 || ori - 1 >= x.read_len) continue;
                const int n = len + 1;
                int elen = n;
                if (elen > x.read_len - (ori - 1)) elen = x.read_len - (ori - 1);
                if (1 + elen <= 61 && insert_quality_pass(p, x, ori, n, p.allq ? 255 : (int)x.quals[ori - 1])) record_event(p, s, 2, (int)x.rev, elen, x.bo + ori - 1);
            } else {
                if (!(c.pflag[ol] & PF_DEL)) continue;
                const int64_t rem = c.ref_len - c.t_lo - ol;
                int elen = len + 1;
                if ((int64_t)elen > rem) elen = (int)rem;
                if (1 + elen <= 61) record_event(p, s, 3, (int)x.rev, elen, c.ref_off + c.t_lo + ol);
            }
        }
        __syncwarp();
        if (!more) break;
        kb += ROUND_OPS; a_run = rd.a_next; ri_run = rd.ri_next;
    }
}

// SNP alleles whose byte is not an upper-case A/C/G/T (":398 candidate_string = '1' + alt" keeps the raw byte): rare,
// recorded after the fact for registered sites only. One thread per entry of the tile: walks the read's ops from the
// entry's start until it finds the op that covers tile position ol.
__device__ void record_other_snp(const SumParams& p, const TileCtx& c, int ol, const TileEntry e) {
    const PvReadBatch& b = p.b;
    const int64_t r = e.read;
    const int64_t co = b.read_cigar_off[r];
    const int n_ops = b.read_n_ops[r];
    const int64_t l_end = c.L - 1 - c.t_lo;
    int a = e.a, ri = e.ri;
    for (int k = e.k; k < n_ops; k++) {
        if (a > ol || (int64_t)a > l_end) return;
        const uint32_t w = b.cigar[co + k];
        int ra, qa;
        op_advance(w, ra, qa);
        if (ra > 0 && ol < a + ra) {
            if (!is_match_op((int)(w & 15u))) return;
            const int64_t idx = (int64_t)ri + (ol - a);
            if (idx >= b.read_len[r]) return;
            const int64_t bo = b.read_base_off[r];
            const uint8_t base = b.bases[bo + idx];
            if (!p.allq && (int)b.quals[bo + idx] < p.qthr) return;
            if (base == c.ref_s[ol] || acgt_code(base) >= 0) return;
            record_event(p, c.site_slot[ol], 1, (int)(b.read_flags[r] & 1u), 1, bo + idx);
            return;
        }
        a += ra; ri += qa;
    }
}

// The tile's work list (K0): warps pull entries from a shared ticket counter.
template <int MODE>
__device__ void for_each_entry(const SumParams& p, const TileCtx& c, const TileEntry* ent, int n_ent, int* s_next) {
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) *s_next = 0;
    __syncthreads();
    while (true) {
        int i = 0;
        if (lane == 0) i = atomicAdd(s_next, 1);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= n_ent) break;
        const int4 raw = __ldg((const int4*)(ent + i));
        TileEntry e; e.read = raw.x; e.k = raw.y; e.a = raw.z; e.ri = raw.w;
        if (MODE == 0) accumulate_entry(p, c, e, lane);
        else record_entry(p, c, e, lane);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(K1_THREADS, PV_K1_MINB) pileup_tile_kernel(const SumParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_next, s_any_events, s_any_other;

    const PvReadBatch& b = p.b;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = p.P;

    TileCtx c;
    c.P = P;
    c.cnt = (uint32_t*)smem;
    c.site_slot = (int32_t*)smem;                                     // aliases cnt row 0 after phase B
    c.scratch = (int32_t*)(smem + (size_t)NC * P * 4) + warp * WARP_SCRATCH;
    uint8_t* ref_raw = smem + (size_t)NC * P * 4 + (size_t)K1_WARPS * WARP_SCRATCH * 4;   // REF_PAD + P + 32 bytes
    c.ref_s = ref_raw + REF_PAD;                                       // word loads under- and overrun the tile
    c.pflag = c.ref_s + P + 32;
    c.rcls = c.pflag + P;
    c.lut = c.rcls + P;
    c.near = c.lut + 256;
    if (tid < 256) c.lut[tid] = (uint8_t)(base_class((uint8_t)tid) | (acgt_code((uint8_t)tid) >= 0 ? 8 : 0));
    if (lane == 0) *op_table(c).qn = 0;
    c.region = p.tile_region[blockIdx.x];
    c.t_lo = p.tile_start[blockIdx.x];
    c.L = b.region_ref_end[c.region] - b.region_ref_start[c.region] + 1;
    c.ref_len = b.region_ref_len[c.region];
    c.ref_off = b.region_ref_off[c.region];
    c.n_valid = (int)((c.L - c.t_lo) < P ? (c.L - c.t_lo) : P);
    const int64_t gbase = p.pos_off[c.region] + c.t_lo;

    for (int i = tid; i < NC * P / 4; i += K1_THREADS) ((uint4*)c.cnt)[i] = make_uint4(0u, 0u, 0u, 0u);   // P is a multiple of 4
    if (tid < 32) c.ref_s[P + tid] = 0;
    if (tid < REF_PAD) ref_raw[tid] = 0;
    for (int i = tid; i < P; i += K1_THREADS) {
        const uint8_t rbyte = i < c.n_valid ? b.ref[c.ref_off + c.t_lo + i] : (uint8_t)'N';
        c.ref_s[i] = rbyte;
        c.rcls[i] = valid_ref(rbyte) ? (uint8_t)base_class(rbyte) : (uint8_t)0xff;
        c.pflag[i] = 0;
        c.near[i] = 0;
    }
    if (tid == 0) { s_any_events = 0; s_any_other = 0; }
    __syncthreads();

    // ---- phase A: accumulate -----------------------------------------------------------------------------------
    if (p.ctr[CTR_STATUS] & ST_ENTRY_OVF) return;
    const TileEntry* ent = p.entries + p.tile_off[blockIdx.x];
    const int n_ent = p.tile_off[blockIdx.x + 1] - p.tile_off[blockIdx.x];
    for_each_entry<0>(p, c, ent, n_ent, &s_next);

    // ---- phase B: image rows, site thresholds ---------------------------------------------------------------------
    // difference arrays -> counts: in-place inclusive prefix sums over the tile (packed words, modulo 2^32: the true
    // value of every field of a prefix is a count in [0, 65535]). Thread t owns words 4t .. 4t+3.
    {
        __shared__ uint32_t s_wsum[2][K1_WARPS];
        const bool act = 4 * tid < P;                                   // P is a multiple of 4
        uint4 vt = make_uint4(0u, 0u, 0u, 0u), vd = vt;
        if (act) { vt = *(const uint4*)&c.cnt[C_T * P + 4 * tid]; vd = *(const uint4*)&c.cnt[C_DELD * P + 4 * tid]; }
        vt.y += vt.x; vt.z += vt.y; vt.w += vt.z;
        vd.y += vd.x; vd.z += vd.y; vd.w += vd.z;
        uint32_t it = vt.w, id = vd.w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t a = __shfl_up_sync(0xffffffffu, it, d), e = __shfl_up_sync(0xffffffffu, id, d);
            if (lane >= d) { it += a; id += e; }
        }
        if (lane == 31) { s_wsum[0][warp] = it; s_wsum[1][warp] = id; }
        __syncthreads();
        uint32_t ot = it - vt.w, od = id - vd.w;                        // exclusive prefix inside the warp
        for (int w2 = 0; w2 < warp; w2++) { ot += s_wsum[0][w2]; od += s_wsum[1][w2]; }
        if (act) {
            *(uint4*)&c.cnt[C_T * P + 4 * tid] = make_uint4(vt.x + ot, vt.y + ot, vt.z + ot, vt.w + ot);
            *(uint4*)&c.cnt[C_DELD * P + 4 * tid] = make_uint4(vd.x + od, vd.y + od, vd.z + od, vd.w + od);
        }
        __syncthreads();
    }
    // B1, every position: coverage and the three counts the site thresholds look at (:625-646) -- nothing else of the row
    int my_site[MAX_PPT];
#pragma unroll
    for (int u = 0; u < MAX_PPT; u++) {
        my_site[u] = -1;
        const int i = tid + u * K1_THREADS;
        if (i < c.n_valid) {
            const uint32_t wt = c.cnt[C_T * P + i], wc2 = c.cnt[C_COV2 * P + i], wid = c.cnt[C_INSDEL * P + i];
            uint32_t ws[4];
#pragma unroll
            for (int k = 0; k < 4; k++) ws[k] = c.cnt[(C_SNP + k) * P + i];
            // the dense SNP allele counters stand for snp_count as well (count_mismatch)
            const uint32_t snp_sum = ws[0] + ws[1] + ws[2] + ws[3];
            const int cov = (int)(wt & 0xffffu) + (int)(wt >> 16) + (int)(wc2 & 0xffffu);
            const int dense = (int)(snp_sum & 0xffffu) + (int)(snp_sum >> 16);
            const int snp = (int)(wc2 >> 16) + dense;
            const int ins = (int)(wid & 0xffffu), del = (int)(wid >> 16);
            bool ps, pi, pd;
            if (cov < FREQ_TABLE) {                          // smallest passing count per coverage, computed with the same fp64 division
                const ushort4 m = p.freq_min[cov];
                ps = snp >= (int)m.x; pi = ins >= (int)m.y; pd = del >= (int)m.z;
            } else {
                const double cv = (double)cov > 1.0 ? (double)cov : 1.0;         // :635-637
                ps = (double)snp / cv >= p.t.snp_freq; pi = (double)ins / cv >= p.t.insert_freq; pd = (double)del / cv >= p.t.delete_freq;
            }
            if (ps || pi || pd) {
                const int64_t pos = b.region_ref_start[c.region] + c.t_lo + i;
                if (pos >= b.region_cand_start[c.region] && pos <= b.region_cand_end[c.region] && (double)cov >= p.t.min_coverage) {   // :639-645
                    int flags = PF_SITE | (ps ? PF_SNP : 0) | (pi ? PF_INS : 0) | (pd ? PF_DEL : 0);
                    const int n_other = ps ? snp - dense : 0;
                    if (n_other > 0) flags |= PF_OTHER;
                    const int n_ev = (pi ? ins : 0) + (pd ? del : 0) + n_other;
                    const int s = atomicAdd(&p.ctr[CTR_SITES], 1);
                    if (s < p.site_cap) {
                        int ev_off = 0;
                        if (n_ev > 0) {
                            ev_off = atomicAdd(&p.ctr[CTR_EVENTS], n_ev);
                            if ((int64_t)ev_off + n_ev > p.ev_cap) atomicOr(&p.ctr[CTR_STATUS], ST_EVENT_OVF);
                            s_any_events = 1;
                            if (n_other > 0) s_any_other = 1;
                        }
                        SiteRec sr;
                        sr.gpos = gbase + i; sr.region = c.region; sr.local = (int32_t)(c.t_lo + i); sr.cov = cov;
                        sr.flags = flags; sr.ev_off = ev_off; sr.n_ev = n_ev; sr.fill = 0; sr.pad = 0;
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            sr.snp[2 * k] = (uint16_t)(ws[k] & 0xffffu);
                            sr.snp[2 * k + 1] = (uint16_t)(ws[k] >> 16);
                        }
                        p.sites[s] = sr;
                        my_site[u] = s;
                        c.pflag[i] = (uint8_t)flags;
                        const int j1 = i + (PV_WINDOW / 2) < c.n_valid ? i + (PV_WINDOW / 2) : c.n_valid - 1;
                        for (int j = i - (PV_WINDOW / 2) > 0 ? i - (PV_WINDOW / 2) : 0; j <= j1; j++) c.near[j] = 1;
                    } else {
                        atomicOr(&p.ctr[CTR_STATUS], ST_SITE_OVF);
                    }
                }
            }
        }
    }
    __syncthreads();
    // B2, image rows -> HBM. Only the candidate windows read the image (K3: rows p-16 .. p+16 of a site p), so a row is
    // BUILT and written only when a site of this tile lies within 16 positions, or when it is one of the tile's first /
    // last 16 rows (a site of the neighbouring tile may reach it): ~10 % of the rows.
    for (int i = tid; i < c.n_valid; i += K1_THREADS) {
        if (!(p.img_all || c.near[i] || i < PV_WINDOW / 2 || i >= c.n_valid - PV_WINDOW / 2)) continue;
        uint32_t w[NC];
#pragma unroll
        for (int k = 0; k < NC; k++) w[k] = c.cnt[k * P + i];
        const uint8_t rb = c.ref_s[i];
        int row[PV_FEATURES];
#pragma unroll
        for (int f = 0; f < PV_FEATURES; f++) row[f] = 0;
        row[0] = ref_value(rb);                                             // :174-191
        const int tf = (int)(w[C_T] & 0xffffu), tr = (int)(w[C_T] >> 16);
        row[4] = -(tf - (int)(w[C_SKIP] & 0xffffu));                         // REFF = T - SKIP
        row[15] = -(tr - (int)(w[C_SKIP] >> 16));
        const int rc = c.rcls[i];
        // the dense SNP allele counters stand for DEV and CLS[A..T] as well (count_mismatch)
        w[C_DEV] += w[C_SNP] + w[C_SNP + 1] + w[C_SNP + 2] + w[C_SNP + 3];
#pragma unroll
        for (int k = 0; k < 4; k++) w[C_CLS + k] += w[C_SNP + k];
#pragma unroll
        for (int k = 0; k < 7; k++) {
            int f = (int)(w[C_CLS + k] & 0xffffu), r = (int)(w[C_CLS + k] >> 16);
            if (k == rc) { f += tf - (int)(w[C_DEV] & 0xffffu); r += tr - (int)(w[C_DEV] >> 16); }   // BASE[class(ref)] = T - DEV
            row[8 + k] = rc == 0xff ? 0 : -f;
            row[19 + k] = rc == 0xff ? 0 : -r;
        }
#pragma unroll
        for (int f = 11; f < 25; f++) row[f] = row[f] < -125 ? -125 : row[f];   // :648-653 (all values <= 0 here)
        uint32_t* dst = (uint32_t*)(p.img + (gbase + i) * PV_FEATURES);
#pragma unroll
        for (int f = 0; f < PV_FEATURES; f += 2)
            dst[f >> 1] = ((uint32_t)(uint16_t)(int16_t)row[f]) | ((uint32_t)(uint16_t)(int16_t)row[f + 1] << 16);
    }
    __syncthreads();
    if (!s_any_events) return;

    // ---- phase C: allele events of the registered sites -----------------------------------------------------------
#pragma unroll
    for (int u = 0; u < MAX_PPT; u++) {
        const int i = tid + u * K1_THREADS;
        if (i < P) c.site_slot[i] = my_site[u];
    }
    __threadfence();      // site records (n_ev, ev_off) are read back through global memory by record_event
    __syncthreads();
    for_each_entry<1>(p, c, ent, n_ent, &s_next);
    if (s_any_other) {
        for (int i = 0; i < c.n_valid; i++) {
            if (!(c.pflag[i] & PF_OTHER)) continue;
            for (int j = tid; j < n_ent; j += K1_THREADS) record_other_snp(p, c, i, ent[j]);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// K2
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ const uint8_t* event_bytes(const SumParams& p, const Event& e) {
    return ((e.info & 3u) == 3u ? p.b.ref : p.b.bases) + e.ptr;
}

// std::string order of "<type><bytes>": <0 if x < y, 0 if equal
__device__ int cmp_event(const SumParams& p, const Event& x, const Event& y) {
    const int tx = (int)(x.info & 3u), ty = (int)(y.info & 3u);
    if (tx != ty) return tx < ty ? -1 : 1;
    const int lx = (int)(x.info >> 8), ly = (int)(y.info >> 8);
    const uint8_t* sx = event_bytes(p, x);
    const uint8_t* sy = event_bytes(p, y);
    const int m = lx < ly ? lx : ly;
    for (int i = 0; i < m; i++) {
        const uint8_t a = sx[i], c = sy[i];
        if (a != c) return a < c ? -1 : 1;
    }
    return lx == ly ? 0 : (lx < ly ? -1 : 1);
}

// the per-allele filters of :682-712; fills the candidate record and its order key when the allele survives
__device__ __forceinline__ bool make_candidate(const SumParams& p, const SiteRec& sr, int s, int type, bool dense, int elen, int64_t ptr,
                                               int nf, int nr, uint32_t sub, CandRec& cr, unsigned long long& key) {
    const int depth = min125(sr.cov);                                           // :682
    const int ad = nf + nr;
    if ((double)ad < p.t.candidate_support) return false;                       // :693
    const double cf = (double)ad / ((double)depth > 1.0 ? (double)depth : 1.0); // :689
    if (type != 1 && cf < p.t.indel_candidate_freq) return false;               // :697
    if (type == 1 && cf < p.t.snp_candidate_freq) return false;                 // :700
    if (type != 1 && p.t.skip_indels) return false;                             // :704
    if (!(sr.flags & (1 << type))) return false;                                // :708-712
    cr.ptr = ptr; cr.site = s; cr.info = (uint32_t)type | (dense ? 4u : 0u) | ((uint32_t)elen << 8);
    cr.nf = nf; cr.nr = nr;
    key = ((unsigned long long)sr.gpos << 24) | ((unsigned long long)type << 22) | (unsigned long long)sub;
    return true;
}

// Candidates are collected 32 at a time in the warp's shared-memory buffer and appended to the global list with ONE atomic
// per 32 (one per candidate serialised the whole kernel on a single L2 address). Called by all 32 lanes.
struct CandStage {
    CandRec* rec; unsigned long long* key; int used;
    __device__ __forceinline__ void flush(const SumParams& p, int lane) {
        if (used == 0) return;
        int base = 0;
        if (lane == 0) base = atomicAdd(&p.ctr[CTR_CANDS], used);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane < used) {
            if (base + lane < p.cand_cap) { p.cands[base + lane] = rec[lane]; p.cand_key[base + lane] = key[lane]; }
            else atomicOr(&p.ctr[CTR_STATUS], ST_CAND_OVF);
        }
        __syncwarp();
        used = 0;
    }
    __device__ __forceinline__ void add(const SumParams& p, bool pass, const CandRec& cr, unsigned long long k, int lane) {
        const unsigned m = __ballot_sync(0xffffffffu, pass);
        const int need = __popc(m);
        if (need == 0) return;
        if (used + need > 32) flush(p, lane);
        if (pass) { const int at = used + __popc(m & ((1u << lane) - 1u)); rec[at] = cr; key[at] = k; }
        used += need;
        __syncwarp();
    }
};

// type and the first seven allele bytes of an event as ONE big-endian 64-bit word: comparing two words compares
// "<type><bytes>" up to byte 7; shorter strings are zero-padded, and equal words are told apart by the lengths (both <= 7:
// the shorter string is a prefix of the longer and sorts first) or, rarely, by the remaining bytes
// Delete alleles of one site are all reference substrings that START at the site: one is a prefix of the other, so their
// order and equality are those of their lengths -- the key is exact and needs no byte at all (they are also the long ones:
// the same 10-base deletion seen in 25 reads used to cost 25 x 25 byte-wise compares in global memory).
__device__ __forceinline__ unsigned long long event_key(const SumParams& p, const Event& e) {
    const int len = (int)(e.info >> 8);
    if ((e.info & 3u) == 3u) return (3ull << 56) | (unsigned long long)len;
    const uint8_t* s = event_bytes(p, e);
    unsigned long long k = (unsigned long long)(e.info & 3u) << 56;
    const int m = len < 7 ? len : 7;
    for (int i = 0; i < m; i++) k |= (unsigned long long)s[i] << (48 - 8 * i);
    return k;
}
// bytes 7 .. 14 of the allele (zero-padded): with it, alleles of up to 15 bytes compare without touching memory again
__device__ __forceinline__ unsigned long long event_key2(const SumParams& p, const Event& e) {
    const int len = (int)(e.info >> 8);
    if ((e.info & 3u) == 3u || len <= 7) return 0ull;
    const uint8_t* s = event_bytes(p, e);
    unsigned long long k = 0;
    const int m = len < 15 ? len : 15;
    for (int i = 7; i < m; i++) k |= (unsigned long long)s[i] << (56 - 8 * (i - 7));
    return k;
}

// Warp per site. The alleles a site recorded are de-duplicated EXACTLY (no hashing): every lane holds one event's key in
// registers, the keys of 32 events at a time go round the warp by shuffle, and only equal keys of alleles longer than seven
// bytes fall back to byte-wise compares in global memory. Per distinct allele: forward / reverse counts, its rank in
// std::set<std::string> order, the filters of :682-712, one candidate record.
__global__ void site_allele_kernel(const SumParams p) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    int n_sites = p.ctr[CTR_SITES];
    if (n_sites > p.site_cap) n_sites = p.site_cap;
    __shared__ CandRec s_rec[8][32];
    __shared__ unsigned long long s_key[8][32];
    CandStage st; st.rec = s_rec[threadIdx.x >> 5]; st.key = s_key[threadIdx.x >> 5]; st.used = 0;
    (void)warp; (void)n_warps;
    // 32 sites per ticket. First LANE per site: the dense SNP alleles need nothing but the site record (most sites have no
    // recorded event at all). Then WARP per site for the sites of the block that did record events.
    for (;;) {
        int tk = 0;
        if (lane == 0) tk = atomicAdd(&p.ctr[CTR_K2_TICKET], 32);
        const int base = __shfl_sync(0xffffffffu, tk, 0);
        if (base >= n_sites) break;
        const bool have = base + lane < n_sites;
        SiteRec mine;
        mine.flags = 0; mine.n_ev = 0; mine.ev_off = 0;
        if (have) mine = p.sites[base + lane];
#pragma unroll 1
        for (int a = 0; a < 4; a++) {                                           // dense "1A" "1C" "1G" "1T"
            CandRec cr; unsigned long long key = 0; bool pass = false;
            if (have && (mine.flags & PF_SNP)) {
                const int nf = mine.snp[2 * a], nr = mine.snp[2 * a + 1];
                const uint8_t byte = (uint8_t)("ACGT"[a]);
                pass = nf + nr > 0 && make_candidate(p, mine, base + lane, 1, true, 1, (int64_t)byte, nf, nr, (uint32_t)byte, cr, key);
            }
            st.add(p, pass, cr, key, lane);
        }
        unsigned heavy = __ballot_sync(0xffffffffu, have && mine.n_ev > 0 && (int64_t)mine.ev_off + mine.n_ev <= p.ev_cap);
        while (heavy) {
        const int s = base + __ffs(heavy) - 1;
        heavy &= heavy - 1;
        const SiteRec sr = p.sites[s];
        const int n = sr.n_ev;
        if (sr.fill != n && lane == 0) atomicOr(&p.ctr[CTR_STATUS], ST_INTERNAL);
        const Event* ev = p.events + sr.ev_off;
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            Event ei; ei.ptr = 0; ei.info = 0; ei.pad = 0;
            unsigned long long ki = 0, ki2 = 0;
            if (i < n) { ei = ev[i]; ki = event_key(p, ei); ki2 = event_key2(p, ei); }
            const int li = (int)(ei.info >> 8);
            int less = 0, nf = 0, nr = 0; bool dup = false;
            for (int j0 = 0; j0 < n; j0 += 32) {
                Event ej = ei; unsigned long long kj = ki, kj2 = ki2;        // the block on the diagonal is already in registers
                if (j0 != i0) {
                    ej.ptr = 0; ej.info = 0; kj = 0; kj2 = 0;
                    if (j0 + lane < n) { ej = ev[j0 + lane]; kj = event_key(p, ej); kj2 = event_key2(p, ej); }
                }
                const int cnt = n - j0 < 32 ? n - j0 : 32;
                for (int jj = 0; jj < cnt; jj++) {
                    const unsigned long long k = __shfl_sync(0xffffffffu, kj, jj);
                    const unsigned long long k2 = __shfl_sync(0xffffffffu, kj2, jj);
                    const uint32_t info = __shfl_sync(0xffffffffu, ej.info, jj);
                    if (i >= n) continue;
                    int cc;
                    if (k != ki) cc = k < ki ? -1 : 1;
                    else if (k2 != ki2) cc = k2 < ki2 ? -1 : 1;
                    else {
                        const int lj = (int)(info >> 8);
                        if ((info & 3u) == 3u) cc = 0;                          // deletes: equal keys are equal lengths are equal alleles
                        else if (li <= 15 || lj <= 15) cc = lj == li ? 0 : (lj < li ? -1 : 1);
                        else { const Event e2 = ev[j0 + jj]; cc = cmp_event(p, e2, ei); }     // rare: both longer than the keys
                    }
                    if (cc < 0) less++;
                    else if (cc == 0) { if (j0 + jj < i) dup = true; if ((info >> 2) & 1u) nr++; else nf++; }
                }
            }
            CandRec cr; unsigned long long key = 0; bool pass = false;
            if (i < n && !dup) {
                const int type = (int)(ei.info & 3u);
                const uint32_t sub = type == 1 ? (uint32_t)(*event_bytes(p, ei)) : (uint32_t)less;
                pass = make_candidate(p, sr, s, type, false, li, ei.ptr, nf, nr, sub, cr, key);
            }
            st.add(p, pass, cr, key, lane);
        }
        }
    }
    st.flush(p, lane);
}

// ------------------------------------------------------------------------------------------------------------
// K3
// ------------------------------------------------------------------------------------------------------------
__global__ void iota_kernel(uint32_t* v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (uint32_t)i;
}

__global__ void emit_window_kernel(const SumParams p, const uint32_t* __restrict__ order, PvCandidates out,
                                   int64_t* n_candidates) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const int found = p.ctr[CTR_CANDS];
    int n = found < p.cand_cap ? found : p.cand_cap;
    if (n > out.capacity) n = (int)out.capacity;
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_candidates = found;
    for (int i = warp; i < n; i += n_warps) {
        const CandRec cr = p.cands[order[i]];
        const SiteRec sr = p.sites[cr.site];
        const int type = (int)(cr.info & 3u), elen = (int)(cr.info >> 8);
        const bool dense = (cr.info & 4u) != 0;
        const int64_t L = p.b.region_ref_end[sr.region] - p.b.region_ref_start[sr.region] + 1;
        const int64_t gbase = p.pos_off[sr.region];
        const int64_t o = sr.local;
        const uint8_t rb = p.b.ref[p.b.region_ref_off[sr.region] + o];
        const bool rv = valid_ref(rb);
        const uint8_t* src = dense ? nullptr : ((type == 3 ? p.b.ref : p.b.bases) + cr.ptr);
        const uint8_t a0 = dense ? (uint8_t)cr.ptr : src[0];
        const int nf = min125(cr.nf), nr = min125(cr.nr);
        int end_index = 16 + elen - 1; if (end_index > 31) end_index = 31;      // :885
        const int cls = base_class(a0);
        int16_t* w = out.windows + (int64_t)i * WIN_ELEMS;
        for (int e = lane; e < WIN_ELEMS; e += 32) {
            const int row = e / PV_FEATURES, f = e - row * PV_FEATURES;
            const int64_t q = o - 16 + row;
            int v = (q < 0 || q >= L) ? 0 : (int)p.img[(gbase + q) * PV_FEATURES + f];   // :833-841 (row L is all zero)
            if (row == 16) {
                if (type == 1) {                                                // :853-860
                    if (f == 1) v = ref_value(a0); else if (f == 5) v = nf; else if (f == 16) v = nr;
                    else if (rv && (f == 8 + cls || f == 19 + cls)) v = -v;
                } else if (type == 2) {                                         // :868-875
                    if (f == 2) v = min125(elen); else if (f == 6) v = nf; else if (f == 17) v = nr;
                    else if (rv && (f == 12 || f == 23)) v = -v;
                } else {                                                        // :883-892
                    if (f == 3) v = min125(elen); else if (f == 7) v = nf; else if (f == 18) v = nr;
                    else if (rv && (f == 13 || f == 24)) v = -v;
                }
            } else if (type == 3 && row >= 17 && row <= end_index) {            // :896-904
                if (f == 3) v = min125(elen); else if (f == 7) v = nf; else if (f == 18) v = nr;
                else if (rv && (f == 14 || f == 25)) v = -v;
            }
            w[e] = (int16_t)v;
        }
        for (int e = lane; e < PV_ALLELE_BYTES; e += 32) {
            uint8_t ch = 0;
            if (e == 0) ch = (uint8_t)('0' + type);
            else if (e - 1 < elen) ch = dense ? a0 : src[e - 1];
            out.allele[(int64_t)i * PV_ALLELE_BYTES + e] = ch;
        }
        if (lane == 0) {
            out.position[i] = p.b.region_ref_start[sr.region] + o;
            out.region[i] = sr.region;
            out.depth[i] = min125(sr.cov);
            out.frequency[i] = min125(cr.nf + cr.nr);
            out.allele_len[i] = (uint8_t)(1 + elen);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
struct Plan {
    int P; int64_t n_tiles;
    int64_t site_cap, ev_cap, cand_cap, entry_cap;
    size_t sort_tmp;
};

int choose_tile(int64_t total_positions, int32_t n_regions) {
    // large tiles amortise the per-(tile, read) set-up; small batches need enough CTAs to cover the 148 SMs
    const int sms = 148;
    int P = P_MAX;
    while (P > 320 && (total_positions / P + n_regions) < 4 * sms) P >>= 1;
    return P;
}

size_t k1_smem_bytes(int P) { return (size_t)NC * P * 4 + (size_t)K1_WARPS * WARP_SCRATCH * 4 + REF_PAD + 4 * (size_t)P + 32 + 256 + 16; }

Plan make_plan(int64_t n_reads, int64_t n_ops, int32_t n_regions, int64_t total_positions, int64_t max_region_len, int64_t capacity) {
    Plan pl;
    pl.P = choose_tile(total_positions, n_regions);
    pl.n_tiles = total_positions / pl.P + n_regions + 1;     // upper bound
    // (tile, read) entries: a read touches at most every tile of its own region
    if (max_region_len <= 0 || max_region_len > total_positions) max_region_len = total_positions;
    const int64_t tiles_per_region = max_region_len / pl.P + 1;
    pl.entry_cap = n_reads > 0 ? ((double)n_reads * (double)(tiles_per_region + 1) > 2.0e9 ? 0x7fffff00ll : n_reads * (tiles_per_region + 1)) : 1;
    if (pl.entry_cap > 0x7fffff00ll) pl.entry_cap = 0x7fffff00ll;
    pl.cand_cap = capacity < 1 ? 1 : capacity;
    pl.site_cap = 4 * pl.cand_cap + 4096;
    if (pl.site_cap > total_positions) pl.site_cap = total_positions > 0 ? total_positions : 1;
    pl.ev_cap = n_ops + (1 << 20);                          // grows with the candidate capacity, so the caller's overflow retry
    if (pl.ev_cap < 8 * pl.cand_cap) pl.ev_cap = 8 * pl.cand_cap;   // (larger capacity) also cures an event overflow
    if (pl.ev_cap > 0x7fffff00ll) pl.ev_cap = 0x7fffff00ll;
    if (pl.cand_cap > 0x7fffff00ll) pl.cand_cap = 0x7fffff00ll;
    if (pl.site_cap > 0x7fffff00ll) pl.site_cap = 0x7fffff00ll;
    pl.sort_tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, pl.sort_tmp, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)pl.cand_cap, 0, 64, (cudaStream_t)0);
    return pl;
}

struct WsLayout {
    int64_t* pos_off; int32_t* tile_region; int32_t* tile_start;
    int32_t* tile_base; int32_t* tile_count; int32_t* tile_fill; int32_t* tile_off; int32_t* read_region; TileEntry* entries;
    TileEntry* pool; int32_t* pool_tile;
    int32_t* read_span;
    int16_t* img; SiteRec* sites; Event* events; CandRec* cands;
    unsigned long long* key_in; unsigned long long* key_out; uint32_t* val_in; uint32_t* val_out;
    void* sort_tmp; int32_t* ctr; ushort4* freq_min;
    int64_t bytes;
};

WsLayout carve(void* base, int64_t size, const Plan& pl, int64_t n_reads, int64_t n_ops, int32_t n_regions,
               int64_t total_positions, bool need_img) {
    pv::Arena a(base, size);
    WsLayout w;
    w.ctr = a.take<int32_t>(CTR_COUNT);
    w.freq_min = a.take<ushort4>(FREQ_TABLE);
    w.pos_off = a.take<int64_t>(n_regions + 1);
    w.tile_region = a.take<int32_t>(pl.n_tiles);
    w.tile_start = a.take<int32_t>(pl.n_tiles);
    w.tile_base = a.take<int32_t>(n_regions + 1);
    w.tile_count = a.take<int32_t>(3 * pl.n_tiles);              // counts, then the two fill cursors per tile: zeroed together
    w.tile_fill = w.tile_count ? w.tile_count + pl.n_tiles : nullptr;
    w.tile_off = a.take<int32_t>(pl.n_tiles + 1);
    w.read_region = a.take<int32_t>(n_reads);
    w.entries = a.take<TileEntry>(pl.entry_cap);
    w.pool = a.take<TileEntry>(pl.entry_cap);
    w.pool_tile = a.take<int32_t>(pl.entry_cap);
    w.read_span = a.take<int32_t>(n_reads);
    w.img = need_img ? a.take<int16_t>(total_positions * PV_FEATURES) : nullptr;
    w.sites = a.take<SiteRec>(pl.site_cap);
    w.events = a.take<Event>(pl.ev_cap);
    w.cands = a.take<CandRec>(pl.cand_cap);
    w.key_in = a.take<unsigned long long>(pl.cand_cap);
    w.key_out = a.take<unsigned long long>(pl.cand_cap);
    w.val_in = a.take<uint32_t>(pl.cand_cap);
    w.val_out = a.take<uint32_t>(pl.cand_cap);
    w.sort_tmp = a.take<uint8_t>((int64_t)pl.sort_tmp);
    w.bytes = pv::align_up(a.cur, 256);
    return w;
}

}  // namespace

extern "C" int64_t pv_summary_workspace_bytes(int64_t n_reads, int64_t n_ops, int32_t n_regions,
                                              int64_t total_positions, int64_t max_region_len, int64_t capacity) {
    const Plan pl = make_plan(n_reads, n_ops, n_regions, total_positions, max_region_len, capacity);
    return carve(nullptr, 0, pl, n_reads, n_ops, n_regions, total_positions, true).bytes;
}

extern "C" int pv_summary_regions(const PvReadBatch* batch, const int64_t* region_len_host, int64_t total_positions,
                                  const PvThresholds* thr, int32_t window, int32_t features,
                                  const PvCandidates* out, int64_t* n_candidates_dev, void* workspace_dev,
                                  int64_t workspace_bytes, int16_t* dense_image_dev, void* stream_) {
    if (!batch || !thr || !out || !n_candidates_dev || !region_len_host) return pv::set_error(PV_EINVAL, "null argument");
    if (window != PV_WINDOW - 1 || features != PV_FEATURES)
        return pv::set_error(PV_EINVAL, "only candidate_window_size=32, feature_size=26 are built (got %d, %d)", window, features);
    if (!(thr->min_snp_baseq == thr->min_snp_baseq) || !(thr->min_indel_baseq == thr->min_indel_baseq))
        return pv::set_error(PV_EINVAL, "NaN base-quality threshold");
    if (int rc = pv::require_device()) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    const PvReadBatch& b = *batch;
    if (b.n_regions <= 0) {
        PV_CUDA_CHECK(cudaMemsetAsync(n_candidates_dev, 0, sizeof(int64_t), stream));
        return PV_OK;
    }
    int64_t max_region_len = 0;
    for (int32_t r = 0; r < b.n_regions; r++) if (region_len_host[r] > max_region_len) max_region_len = region_len_host[r];
    const Plan pl = make_plan(b.n_reads, b.n_ops, b.n_regions, total_positions, max_region_len, out->capacity);
    const WsLayout w = carve(workspace_dev, workspace_bytes, pl, b.n_reads, b.n_ops, b.n_regions, total_positions,
                             dense_image_dev == nullptr);
    if (w.bytes > workspace_bytes)
        return pv::set_error(PV_EINVAL, "workspace too small: need %lld bytes, got %lld", (long long)w.bytes, (long long)workspace_bytes);

    // host-side planning: dense position offsets and the tile table
    std::vector<int64_t> pos_off(b.n_regions + 1, 0);
    std::vector<int32_t> t_region, t_start, t_base(b.n_regions + 1, 0);
    for (int32_t r = 0; r < b.n_regions; r++) {
        const int64_t L = region_len_host[r];
        if (L <= 0 || L > 0x7fffffffll) return pv::set_error(PV_EINVAL, "region %d has length %lld", r, (long long)L);
        pos_off[r + 1] = pos_off[r] + L;
        for (int64_t s = 0; s < L; s += pl.P) { t_region.push_back(r); t_start.push_back((int32_t)s); }
        t_base[r + 1] = (int32_t)t_region.size();
    }
    if (pos_off[b.n_regions] != total_positions) return pv::set_error(PV_EINVAL, "total_positions does not match region lengths");
    if (total_positions >= (1ll << 40)) return pv::set_error(PV_EINVAL, "too many positions in one batch");
    const int64_t n_tiles = (int64_t)t_region.size();
    if (n_tiles > pl.n_tiles) return pv::set_error(PV_EINVAL, "internal: tile bound");
    PV_CUDA_CHECK(cudaMemcpyAsync(w.pos_off, pos_off.data(), pos_off.size() * sizeof(int64_t), cudaMemcpyHostToDevice, stream));
    PV_CUDA_CHECK(cudaMemcpyAsync(w.tile_region, t_region.data(), n_tiles * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    PV_CUDA_CHECK(cudaMemcpyAsync(w.tile_start, t_start.data(), n_tiles * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    PV_CUDA_CHECK(cudaMemcpyAsync(w.tile_base, t_base.data(), t_base.size() * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    PV_CUDA_CHECK(cudaMemsetAsync(w.tile_count, 0, 3 * pl.n_tiles * sizeof(int32_t), stream));
    PV_CUDA_CHECK(cudaMemsetAsync(w.ctr, 0, CTR_COUNT * sizeof(int32_t), stream));
    PV_CUDA_CHECK(cudaMemsetAsync(w.key_in, 0xff, pl.cand_cap * sizeof(unsigned long long), stream));

    SumParams p;
    p.b = b; p.pos_off = w.pos_off; p.tile_region = w.tile_region; p.tile_start = w.tile_start; p.P = pl.P;
    p.tile_base = w.tile_base; p.n_tiles = (int32_t)n_tiles; p.read_region = w.read_region; p.tile_count = w.tile_count;
    p.tile_fill = w.tile_fill; p.tile_off = w.tile_off; p.entries = w.entries; p.entry_cap = (int32_t)pl.entry_cap;
    p.pool = w.pool; p.pool_tile = w.pool_tile;
    p.read_span = w.read_span; p.freq_min = w.freq_min;
    p.img = dense_image_dev ? dense_image_dev : w.img;
    p.img_all = dense_image_dev ? 1 : 0;
    p.sites = w.sites; p.site_cap = (int32_t)pl.site_cap; p.events = w.events; p.ev_cap = (int32_t)pl.ev_cap;
    p.cands = w.cands; p.cand_key = w.key_in; p.cand_cap = (int32_t)pl.cand_cap; p.ctr = w.ctr;
    double q = thr->min_snp_baseq; int qi = 0;
    if (q > 256.0) qi = 256; else if (q > 0.0) { qi = (int)q; if ((double)qi < q) qi++; }
    p.qthr = qi;
    // min_qual: "every quality of every read base is >= this" (0 = no promise). When it clears both thresholds (and the
    // insert test's per-base average, min_indel_baseq) K1 never touches the quality array.
    p.allq = (b.min_qual > 0 && b.min_qual >= qi && (double)b.min_qual >= thr->min_indel_baseq) ? 1 : 0;
    { static int off = -1; if (off < 0) { const char* v = getenv("PV_NO_ALLQ"); off = (v && atoi(v)) ? 1 : 0; } if (off && b.quals) p.allq = 0; }
    if (!p.allq && b.quals == nullptr && b.n_bases > 0)
        return pv::set_error(PV_EINVAL, "the batch has no quality array and its min_qual promise (%d) does not clear the thresholds", b.min_qual);
    p.t = *thr;

    const int sms = pv::sm_count();
    {
        int64_t blocks = (b.n_reads + 7) / 8;            // 8 warps per 256-thread block
        if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
        if (blocks < 1) blocks = 1;
        pv::prof_begin(pv::FAM_SUM_PREFIX, stream);
        freq_table_kernel<<<FREQ_TABLE / 256, 256, 0, stream>>>(p);
        PV_CUDA_CHECK(cudaGetLastError());
        read_entries_kernel<<<(unsigned)blocks, 256, 0, stream>>>(p);
        PV_CUDA_CHECK(cudaGetLastError());
        tile_scan_kernel<<<1, 1024, 0, stream>>>(p);
        PV_CUDA_CHECK(cudaGetLastError());
        entry_scatter_kernel<<<(unsigned)(sms * 4), 256, 0, stream>>>(p);
        PV_CUDA_CHECK(cudaGetLastError());
        pv::prof_end(pv::FAM_SUM_PREFIX, stream, 4);
    }
    const size_t smem = k1_smem_bytes(pl.P);
    PV_CUDA_CHECK(cudaFuncSetAttribute(pileup_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pv::prof_begin(pv::FAM_SUM_TILE, stream);
    pileup_tile_kernel<<<(unsigned)n_tiles, K1_THREADS, smem, stream>>>(p);
    PV_CUDA_CHECK(cudaGetLastError());
    pv::prof_end(pv::FAM_SUM_TILE, stream, 1);

    pv::prof_begin(pv::FAM_SUM_ALLELE, stream);
    site_allele_kernel<<<sms * 4, 256, 0, stream>>>(p);
    PV_CUDA_CHECK(cudaGetLastError());
    pv::prof_end(pv::FAM_SUM_ALLELE, stream, 1);

    pv::prof_begin(pv::FAM_SUM_SORT, stream);
    iota_kernel<<<(unsigned)((pl.cand_cap + 255) / 256), 256, 0, stream>>>(w.val_in, (int)pl.cand_cap);
    PV_CUDA_CHECK(cudaGetLastError());
    size_t tmp = pl.sort_tmp;
    int key_bits = 25;                                    // 24 low bits (type, allele rank) + the position bits in use
    while (key_bits < 64 && (total_positions >> (key_bits - 24)) != 0) key_bits++;
    PV_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(w.sort_tmp, tmp, (const unsigned long long*)w.key_in, w.key_out,
                                                  (const uint32_t*)w.val_in, w.val_out, (int)pl.cand_cap, 0, key_bits, stream));
    pv::prof_end(pv::FAM_SUM_SORT, stream, 11);     // iota + histogram + exclusive sum + 8 onesweep passes
    int64_t eblocks = (pl.cand_cap + 7) / 8;
    if (eblocks > (int64_t)sms * 16) eblocks = (int64_t)sms * 16;
    pv::prof_begin(pv::FAM_SUM_EMIT, stream);
    emit_window_kernel<<<(unsigned)eblocks, 256, 0, stream>>>(p, w.val_out, *out, n_candidates_dev);
    PV_CUDA_CHECK(cudaGetLastError());
    pv::prof_end(pv::FAM_SUM_EMIT, stream, 1);
    return PV_OK;
}

// status word of the last pv_summary_regions on this workspace (device int32 at the start of the workspace)
extern "C" int pv_summary_status_offset(void) { return CTR_STATUS * (int)sizeof(int32_t); }
