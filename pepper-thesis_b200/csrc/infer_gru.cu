// Model M-B on sm_100a: the polisher's TransducerGRU.forward
// (/root/reference/pepper/modules/python/models/simple_model.py:27-42): biGRU(10->128, h0 = hidden) ->
// biGRU(256->128, h0 = encoder h_n) -> Linear(256->5), and the sliding-window loop around it
// (/root/reference/pepper/modules/python/models/predict_distributed_gpu.py:63-96).
//
// Two kernels per layer (the recurrence of a GRU is serial in time, its input projection is not):
//   gx GEMM       tc::gemm_kernel<GxEpilogue> (tcgen05/TMA, slot tiles): gx[t] = x_t * W_ih^T + b for ALL time steps and
//                 both directions at once, written as fp16 in the tile order the recurrence streams it back in.
//   recurrence    gru_recur_kernel: one CTA per (128 windows, direction) that stays resident for all T steps. W_hh
//                 (384 x 128 bf16, 96 KB) is TMA-loaded into shared memory ONCE; h_{t-1} lives on chip -- as the bf16
//                 A operand in shared memory (written by the epilogue in the canonical 128-byte-swizzle layout) and as
//                 fp32 state in 128 TMEM columns (tcgen05.st/ld); per step one elected thread issues
//                 tcgen05.mma 128x192x16 into the r/z/n_h accumulators (two column halves, so the epilogue of the
//                 first half overlaps the MMAs of the second), 16 epilogue warps apply the gates and the only
//                 HBM traffic is the gx stream (bulk-copied two half-steps ahead through an mbarrier ring) and the bf16
//                 h_t output. No launch per step, no grid-wide dependency, h never round-trips through L2.
// PyTorch keeps the hidden part of the candidate gate separate (n = tanh(W_in x + b_in + r * (W_hn h + b_hn))): b_hn
// stays in the recurrence, every other bias is folded into gx.
#include "common.cuh"
#include "tc_gemm.cuh"
#include "infer_common.cuh"
#include <cuda_fp16.h>
#include <vector>

namespace {

constexpr int GF = 10;              // input features
constexpr int GH = 128;             // hidden size
constexpr int GC = 2 * GH;          // channels of a layer output
constexpr int GXK = 64;             // encoder input padded to one k block
constexpr int GCLS = 5;
constexpr int GN3 = 3 * GH;         // r, z, n rows of one direction
constexpr int GRU_MAX_CHUNK = 16384;   // windows per pass (gx: 96 KB per 128 windows, step and direction)

// ---- gx layout ---------------------------------------------------------------------------------------------
// [m_blk][dir][t][half][gate r,z,n][quarter 0..3][sub 0..1][row 0..127][8 fp16]: hidden unit j = half*64 + quarter*16 +
// sub*8 + e. One (m_blk, dir, t, half) is a contiguous 48 KB blob = one ring slot of the recurrence; inside it the 16
// bytes of a (row, sub) sit 16 bytes apart from the neighbouring rows', so a warp's reads and writes are contiguous.
constexpr int GX_HALF_BYTES = 3 * 4 * 2 * 128 * 16;      // 49152
__host__ __device__ inline size_t gx_unit_index(int m_blk, int dir, int T, int t, int half, int gate, int quarter, int sub, int r) {
    return ((((((((size_t)m_blk * 2 + dir) * T + t) * 2 + half) * 3 + gate) * 4 + quarter) * 2 + sub) * 128 + r);   // 16-byte units
}

// epilogue of the input-projection GEMM: 768 columns = [dir][gate][128]; `slot` (passed as dir) is the time step
struct GxEpilogue {
    static constexpr int kStages = 4;
    static constexpr int kSmemBytes = 0;
    static constexpr bool kInlinePrefetch = false;
    static constexpr bool kSlotTiles = true;
    __device__ void setup(uint8_t*, int) const {}
    __device__ void prefetch(uint8_t*, int, int, int, int, bool, int, int) const {}

    const float* bias;        // [768]: b_ir+b_hr, b_iz+b_hz, b_in per direction
    uint4* gx;
    int T;
    int rows_alloc;           // rows the gx buffer holds (a multiple of 128)

    __device__ void operator()(uint8_t*, int, int t, int n_blk, int row, bool, uint32_t taddr, int half, int,
                               const tc::NextTile&) const {
        const bool in = row < rows_alloc;
        const int m_blk = row >> 7, r = row & 127;
#pragma unroll 1
        for (int c = 0; c < 8; c += 2) {                       // two 16-column chunks in flight
            float a[2][16], b[2][16];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                tc::tmem_ld16(taddr + (half * 8 + c + u) * 16, a[u]);
                ld16(bias + (n_blk * 16 + half * 8 + c + u) * 16, b[u]);
            }
            tc::tmem_ld_wait();
            if (in) {
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const int col16 = n_blk * 16 + half * 8 + c + u;   // 0..47
                    const int dir = col16 / 24, rem = col16 - dir * 24;
                    const int gate = rem >> 3, c8 = rem & 7;
                    uint32_t hp[8];
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const __half2 h2 = __floats2half2_rn(a[u][i] + b[u][i], a[u][i + 1] + b[u][i + 1]);
                        hp[i >> 1] = *(const uint32_t*)&h2;
                    }
                    const size_t x = gx_unit_index(m_blk, dir, T, t, c8 >> 2, gate, c8 & 3, 0, r);
                    gx[x] = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                    gx[x + 128] = make_uint4(hp[4], hp[5], hp[6], hp[7]);
                }
            }
        }
    }
};

// ---- recurrence ----------------------------------------------------------------------------------------------
constexpr int RC_EPI_WARPS = 16;
constexpr int RC_THREADS = 128 + RC_EPI_WARPS * 32;            // 4 control warps + 16 epilogue warps
constexpr int RC_W_BYTES = GN3 * GH * 2;                       // 98304: two k blocks of [384 rows][128 B]
constexpr int RC_A_BYTES = 128 * GH * 2;                       // 32768: two k blocks of [128 rows][128 B]
constexpr int RC_SMEM = RC_W_BYTES + RC_A_BYTES + 2 * GX_HALF_BYTES + GH * 4 + 128 + 1024;
// FOLD variant (layer 1, 10 input features): no gx stream. The x-part of the r and z gates rides on the MMA (a third k atom:
// A = [h | x x 0], B = [W_hh | W_ir,iz hi lo]), the x-part of the candidate gate (which PyTorch keeps outside r * (...)) is 160
// FMAs per thread and half step on fp32 W_in; x_t arrives as 16 bytes per window through a four-slot ring.
constexpr int RCF_XSLOTS = 4;
constexpr int RCF_X_BYTES = 128 * 16;                          // one step: 128 windows x (10 counts + 6 zero bytes)
constexpr int RCF_W_BYTES = RC_W_BYTES + GN3 * 128;            // + the x atom of the B operand
constexpr int RCF_A_BYTES = RC_A_BYTES + 128 * 128;            // + the x atom of the A operand
constexpr int RCF_SMEM = RCF_W_BYTES + RCF_A_BYTES + RCF_XSLOTS * RCF_X_BYTES + GH * 4 + 3 * GH * 4 + GF * GH * 4 + 128 + 1024;

struct RecurParams {
    const uint8_t* gx;        // gx layout above (not FOLD)
    const uint8_t* xt;        // FOLD: [m_blk][T][128 rows][16] counts of x_t
    const float* bias;        // FOLD: [2][3][128] b_ir + b_hr, b_iz + b_hz, b_in
    const float* w_in;        // FOLD: [2][10][128] fp32 W_in (candidate gate, input part)
    const float* b_hn;        // [2][128]
    float* h_state;           // [M][2][128] fp32: h_0 in, h_T out
    __nv_bfloat16* out;       // [M][S][256]: h_t -> slot t + 1, channels dir*128 ..
    int M, T, S;
    long long* trace;         // optional (-DPV_TRACE): CTA 0 records [step < 16][event] SM-clock stamps
};

#ifdef PV_TRACE
#define GRU_TR(step, ev) do { if (p.trace && blockIdx.x == 0 && (step) < 16) p.trace[(step) * 16 + (ev)] = clock64(); } while (0)
#else
#define GRU_TR(step, ev) do { } while (0)
#endif

template <bool FOLD>
__global__ void __launch_bounds__(RC_THREADS, 1)
gru_recur_kernel(const __grid_constant__ CUtensorMap tmW /* [2*384][128] bf16, box 64 x 128 rows */,
                 const __grid_constant__ CUtensorMap tmWx /* FOLD: [2*384][64] bf16 input weights of r, z (hi | lo), box 64 x 128 rows */,
                 const __grid_constant__ CUtensorMap tmOut /* [rows][S][256] bf16, box 64 x 1 x 128 rows */, const RecurParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_w = smem;
    uint8_t* smem_a = smem + (FOLD ? RCF_W_BYTES : RC_W_BYTES);
    uint8_t* smem_gx = smem_a + (FOLD ? RCF_A_BYTES : RC_A_BYTES);     // FOLD: the x ring
    float* s_bhn = (float*)(smem_gx + (FOLD ? RCF_XSLOTS * RCF_X_BYTES : 2 * GX_HALF_BYTES));
    float* s_bias = s_bhn + GH;                                        // FOLD only: [3][128]
    float* s_win = s_bias + (FOLD ? 3 * GH : 0);                       // FOLD only: [10][128]
    uint64_t* bars = (uint64_t*)(s_win + (FOLD ? GF * GH : 0));
    uint64_t* w_bar = bars;            // W_hh landed
    uint64_t* h_ready = bars + 1;      // h_{t-1} operand written, accumulators drained (16 warp arrivals)
    uint64_t* acc_full = bars + 2;     // [2] accumulators of a column half complete
    uint64_t* gx_full = bars + 4;      // [2]; FOLD: x_full[4]
    uint64_t* gx_empty = bars + 8;     // [2] (16 warp arrivals); FOLD: x_empty[4]
    uint32_t* tmem_slot = (uint32_t*)(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_blk = (int)blockIdx.x >> 1, dir = (int)blockIdx.x & 1;
    const int T = p.T;

    if (warp == 0 && lane == 0) { tc::tma_prefetch_desc(&tmW); tc::tma_prefetch_desc(&tmOut); }
    if (warp == 1 && lane == 0) {
        tc::mbar_init(w_bar, 1);
        tc::mbar_init(h_ready, RC_EPI_WARPS);
        for (int i = 0; i < 2; i++) tc::mbar_init(&acc_full[i], 1);
        for (int i = 0; i < (FOLD ? RCF_XSLOTS : 2); i++) { tc::mbar_init(&gx_full[i], 1); tc::mbar_init(&gx_empty[i], RC_EPI_WARPS); }
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, tc::TMEM_COLS);
    if (warp == 3) {
        for (int i = lane; i < GH; i += 32) s_bhn[i] = p.b_hn[dir * GH + i];
        if constexpr (FOLD) {
            for (int i = lane; i < 3 * GH; i += 32) s_bias[i] = p.bias[dir * 3 * GH + i];
            for (int i = lane; i < GF * GH; i += 32) s_win[i] = p.w_in[dir * GF * GH + i];
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== loads: W_hh once, then the gx ring =====
        if (lane == 0) {
            tc::mbar_expect_tx(w_bar, FOLD ? RCF_W_BYTES : RC_W_BYTES);
            for (int kb = 0; kb < 2; kb++)
                for (int rb = 0; rb < 3; rb++)
                    tc::tma_load_2d(smem_w + kb * (GN3 * 128) + rb * (128 * 128), &tmW, w_bar, kb * 64, dir * GN3 + rb * 128);
            if constexpr (FOLD)
                for (int rb = 0; rb < 3; rb++)
                    tc::tma_load_2d(smem_w + 2 * (GN3 * 128) + rb * (128 * 128), &tmWx, w_bar, 0, dir * GN3 + rb * 128);
            for (int s = 0; s < T; s++) {
                const int t = dir ? T - 1 - s : s;
                if constexpr (FOLD) {
                    const int slot = s & (RCF_XSLOTS - 1);
                    tc::mbar_wait(&gx_empty[slot], (uint32_t)((s / RCF_XSLOTS) & 1) ^ 1u);
                    tc::mbar_expect_tx(&gx_full[slot], RCF_X_BYTES);
                    tc::bulk_load(smem_gx + slot * RCF_X_BYTES, p.xt + ((size_t)m_blk * T + t) * RCF_X_BYTES, RCF_X_BYTES, &gx_full[slot]);
                } else
                for (int hf = 0; hf < 2; hf++) {
                    tc::mbar_wait(&gx_empty[hf], (uint32_t)(s & 1) ^ 1u);
                    tc::mbar_expect_tx(&gx_full[hf], GX_HALF_BYTES);
                    const uint8_t* src = p.gx + gx_unit_index(m_blk, dir, T, t, hf, 0, 0, 0, 0) * 16;
#pragma unroll
                    for (int c = 0; c < 3; c++)
                        tc::bulk_load(smem_gx + hf * GX_HALF_BYTES + c * (GX_HALF_BYTES / 3), src + c * (GX_HALF_BYTES / 3),
                                      GX_HALF_BYTES / 3, &gx_full[hf]);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = tc::make_idesc(128, 192);
            uint64_t adesc[3], bdesc[2][3];
#pragma unroll
            for (int kb = 0; kb < 3; kb++) {
                adesc[kb] = tc::make_smem_desc(tc::smem_u32(smem_a + kb * (128 * 128)));
#pragma unroll
                for (int hf = 0; hf < 2; hf++)
                    bdesc[hf][kb] = tc::make_smem_desc(tc::smem_u32(smem_w + kb * (GN3 * 128) + hf * (192 * 128)));
            }
            tc::mbar_wait(w_bar, 0);
            for (int s = 0; s <= T; s++) {
                GRU_TR(s, 0);
                tc::mbar_wait(h_ready, (uint32_t)s & 1u);
                GRU_TR(s, 1);
                tc::tc_fence_after();
#pragma unroll
                for (int hf = 0; hf < 2; hf++) {
                    if (s < T) {
#pragma unroll
                        for (int kb = 0; kb < 2; kb++)
#pragma unroll
                            for (int k = 0; k < 4; k++)
                                tc::umma_bf16(tmem_base + (uint32_t)(hf * 192), adesc[kb] + (uint64_t)(k * 2), bdesc[hf][kb] + (uint64_t)(k * 2),
                                              idesc, (kb | k) != 0 ? 1u : 0u);
                        if constexpr (FOLD) {                  // x atom: columns 0..31 = [x | x | 0] against [W hi | W lo | 0]
#pragma unroll
                            for (int k = 0; k < 2; k++)
                                tc::umma_bf16(tmem_base + (uint32_t)(hf * 192), adesc[2] + (uint64_t)(k * 2), bdesc[hf][2] + (uint64_t)(k * 2), idesc, 1u);
                        }
                    }
                    if (hf == 0) {
                        if (s < T) tc::umma_commit(&acc_full[0]);
                        GRU_TR(s, 2);
                        if (s > 0) {
                            // the operand tile IS h of the step just finished (bf16, 128-byte swizzle): store it as the
                            // layer output with two TMA tile stores instead of 32-byte stores from every epilogue thread
                            const int tp = dir ? T - s : s - 1;
                            tc::tma_store_3d(&tmOut, smem_a, dir * GH, tp + 1, m_blk * 128);
                            tc::tma_store_3d(&tmOut, smem_a + 128 * 128, dir * GH + 64, tp + 1, m_blk * 128);
                            tc::bulk_commit();
                        }
                    } else if (s < T) {
                        // the epilogue overwrites the operand tile once it has seen acc_full[1]: the tile store must have read it
                        tc::bulk_wait_read0();
                        tc::umma_commit(&acc_full[1]);
                        GRU_TR(s, 3);
                    }
                }
            }
            tc::bulk_wait0();
        }
    } else if (warp >= 4) {
        // ===== epilogue: thread = (row, 16 hidden units of each column half) =====
        const int q = warp & 3;                                // TMEM lane quarter this warp may access
        const int gq = (warp - 4) >> 2;                        // 16-unit group inside a half
        const int r = q * 32 + lane;
        const int row = m_blk * 128 + r;
        const bool ok = row < p.M;
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
        // operand copy of this thread's units in the swizzled A tile: k block = half, 16-byte chunks 2*gq and 2*gq + 1
        const uint32_t a_row = tc::smem_u32(smem_a) + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
        const uint32_t sw0 = (uint32_t)(((2 * gq) ^ (r & 7)) << 4), sw1 = (uint32_t)(((2 * gq + 1) ^ (r & 7)) << 4);
        const uint32_t gx_thr = tc::smem_u32(smem_gx) + (uint32_t)((gq * 256 + r) * 16);   // + half, gate * 16384, sub * 2048
        const uint32_t bhn_thr = tc::smem_u32(s_bhn) + (uint32_t)(gq * 64);                // + half * 256
        // FOLD: this row's 16 count bytes in an x ring slot; its 16-byte chunk gq of the x atom of the A operand
        // (columns 8 gq .. 8 gq + 7 of [x0..x9 | x0..x9 | 0 ...]); biases / W_in of this thread's 16 units (+ half * 256)
        const uint32_t x_thr = tc::smem_u32(smem_gx) + (uint32_t)(r * 16);
        const uint32_t xa_thr = a_row + 2 * (128 * 128) + (uint32_t)((gq ^ (r & 7)) << 4);
        const uint32_t bias_thr = tc::smem_u32(s_bias) + (uint32_t)(gq * 64);
        const uint32_t win_thr = tc::smem_u32(s_win) + (uint32_t)(gq * 64);
        auto write_x_atom = [&](int step) {                      // x of `step` -> bf16 chunk of the A operand
            const int slot = step & (RCF_XSLOTS - 1);
            if (lane == 0) tc::mbar_wait(&gx_full[slot], (uint32_t)((step / RCF_XSLOTS) & 1));
            __syncwarp();
            const uint4 xv = tc::lds128(x_thr + slot * RCF_X_BYTES);
            uint32_t hp[4];
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
                float v[2];
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const int j = 8 * gq + e + u;                // column of the atom
                    const int idx = j < GF ? j : j - GF;         // which count
                    const uint32_t wsel = idx < 4 ? xv.x : idx < 8 ? xv.y : xv.z;
                    v[u] = j < 2 * GF ? (float)((wsel >> (8 * (idx & 3))) & 0xffu) : 0.f;
                }
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[0], v[1]);
                hp[e >> 1] = *(const uint32_t*)&h2;
            }
            tc::sts128(xa_thr, make_uint4(hp[0], hp[1], hp[2], hp[3]));
        };

        // h_0: fp32 state -> TMEM columns 384.., bf16 -> A tile
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
            float h[16];
            if (ok) ld16(p.h_state + ((size_t)row * 2 + dir) * GH + hf * 64 + gq * 16, h);
            else {
#pragma unroll
                for (int i = 0; i < 16; i++) h[i] = 0.f;
            }
            tc::tmem_st16(tlane + 384 + hf * 64 + gq * 16, h);
            uint32_t hp[8];
#pragma unroll
            for (int i = 0; i < 16; i += 2) { const __nv_bfloat162 h2 = __floats2bfloat162_rn(h[i], h[i + 1]); hp[i >> 1] = *(const uint32_t*)&h2; }
            tc::sts128(a_row + hf * (128 * 128) + sw0, make_uint4(hp[0], hp[1], hp[2], hp[3]));
            tc::sts128(a_row + hf * (128 * 128) + sw1, make_uint4(hp[4], hp[5], hp[6], hp[7]));
        }
        if constexpr (FOLD) write_x_atom(0);
        tc::tmem_st_wait();
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(h_ready);

        for (int s = 0; s < T; s++) {
            const uint32_t ph = (uint32_t)s & 1u;
            uint32_t keep[8];                                  // first half's bf16 h_t until the second half's MMAs have read h_{t-1}
#pragma unroll
            for (int hf = 0; hf < 2; hf++) {
                if (lane == 0) {
                    tc::mbar_wait(&acc_full[hf], ph);
                    if (warp == 4) GRU_TR(s, 4 + hf * 4);
                    if constexpr (!FOLD) tc::mbar_wait(&gx_full[hf], ph);   // FOLD: x_t's slot was awaited when its atom was written
                    if (warp == 4) GRU_TR(s, 5 + hf * 4);
                }
                __syncwarp();
                tc::tc_fence_after();
                // two register-light stages: (r, n_h) -> candidate n, then (z, h_{t-1}) -> h_t
                float ar[16], an[16];
                tc::tmem_ld16(tlane + hf * 192 + 0 * 64 + gq * 16, ar);
                tc::tmem_ld16(tlane + hf * 192 + 2 * 64 + gq * 16, an);
                const uint32_t gxa = gx_thr + hf * GX_HALF_BYTES;
                uint4 g4[4];
                float bh[16];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint4 v = tc::lds128(bhn_thr + hf * 256 + i * 16);
                    bh[4 * i] = __uint_as_float(v.x); bh[4 * i + 1] = __uint_as_float(v.y);
                    bh[4 * i + 2] = __uint_as_float(v.z); bh[4 * i + 3] = __uint_as_float(v.w);
                }
                float az[16], h[16];
                uint32_t hp[8];
                if constexpr (FOLD) {
                    // the accumulators already hold W_h h + W_i x for r and z; the candidate's input part is computed here
                    const uint4 xv = tc::lds128(x_thr + (s & (RCF_XSLOTS - 1)) * RCF_X_BYTES);
                    float xf[GF];
#pragma unroll
                    for (int f = 0; f < GF; f++) xf[f] = (float)(((f < 4 ? xv.x : f < 8 ? xv.y : xv.z) >> (8 * (f & 3))) & 0xffu);
                    float nx[16];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const uint4 v = tc::lds128(bias_thr + 2 * GH * 4 + hf * 256 + i * 16);       // b_in
                        nx[4 * i] = __uint_as_float(v.x); nx[4 * i + 1] = __uint_as_float(v.y);
                        nx[4 * i + 2] = __uint_as_float(v.z); nx[4 * i + 3] = __uint_as_float(v.w);
                    }
#pragma unroll
                    for (int f = 0; f < GF; f++) {
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const uint4 v = tc::lds128(win_thr + f * GH * 4 + hf * 256 + i * 16);
                            nx[4 * i] = fmaf(__uint_as_float(v.x), xf[f], nx[4 * i]); nx[4 * i + 1] = fmaf(__uint_as_float(v.y), xf[f], nx[4 * i + 1]);
                            nx[4 * i + 2] = fmaf(__uint_as_float(v.z), xf[f], nx[4 * i + 2]); nx[4 * i + 3] = fmaf(__uint_as_float(v.w), xf[f], nx[4 * i + 3]);
                        }
                    }
                    float br[16];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const uint4 v = tc::lds128(bias_thr + hf * 256 + i * 16);                    // b_ir + b_hr
                        br[4 * i] = __uint_as_float(v.x); br[4 * i + 1] = __uint_as_float(v.y);
                        br[4 * i + 2] = __uint_as_float(v.z); br[4 * i + 3] = __uint_as_float(v.w);
                    }
                    tc::tmem_ld_wait();
                    if (warp == 4 && lane == 0) GRU_TR(s, 6 + hf * 4);
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const float r0 = sigmoid_f(ar[i] + br[i]);
                        an[i] = tanh_f(fmaf(r0, an[i] + bh[i], nx[i]));
                    }
                    tc::tmem_ld16(tlane + hf * 192 + 1 * 64 + gq * 16, az);
                    tc::tmem_ld16(tlane + 384 + hf * 64 + gq * 16, h);
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const uint4 v = tc::lds128(bias_thr + GH * 4 + hf * 256 + i * 16);           // b_iz + b_hz
                        br[4 * i] = __uint_as_float(v.x); br[4 * i + 1] = __uint_as_float(v.y);
                        br[4 * i + 2] = __uint_as_float(v.z); br[4 * i + 3] = __uint_as_float(v.w);
                    }
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const float z0 = sigmoid_f(az[i] + br[i]), z1 = sigmoid_f(az[i + 1] + br[i + 1]);
                        h[i] = fmaf(z0, h[i] - an[i], an[i]);          // (1 - z) n + z h
                        h[i + 1] = fmaf(z1, h[i + 1] - an[i + 1], an[i + 1]);
                        const __nv_bfloat162 h2 = __floats2bfloat162_rn(h[i], h[i + 1]);
                        hp[i >> 1] = *(const uint32_t*)&h2;
                    }
                } else {
                g4[0] = tc::lds128(gxa); g4[1] = tc::lds128(gxa + 2048);
                g4[2] = tc::lds128(gxa + 2 * 16384); g4[3] = tc::lds128(gxa + 2 * 16384 + 2048);
                tc::tmem_ld_wait();
                if (warp == 4 && lane == 0) GRU_TR(s, 6 + hf * 4);
                {
                    const __half2* gr = (const __half2*)&g4[0]; const __half2* gn = (const __half2*)&g4[2];
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const float2 xr = __half22float2(gr[i >> 1]), xn = __half22float2(gn[i >> 1]);
                        const float r0 = sigmoid_f(ar[i] + xr.x), r1 = sigmoid_f(ar[i + 1] + xr.y);
                        an[i] = tanh_f(fmaf(r0, an[i] + bh[i], xn.x));
                        an[i + 1] = tanh_f(fmaf(r1, an[i + 1] + bh[i + 1], xn.y));
                    }
                }
                tc::tmem_ld16(tlane + hf * 192 + 1 * 64 + gq * 16, az);
                tc::tmem_ld16(tlane + 384 + hf * 64 + gq * 16, h);
                g4[0] = tc::lds128(gxa + 16384); g4[1] = tc::lds128(gxa + 16384 + 2048);
                tc::tmem_ld_wait();
                {
                    const __half2* gz = (const __half2*)&g4[0];
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const float2 xz = __half22float2(gz[i >> 1]);
                        const float z0 = sigmoid_f(az[i] + xz.x), z1 = sigmoid_f(az[i + 1] + xz.y);
                        h[i] = fmaf(z0, h[i] - an[i], an[i]);          // (1 - z) n + z h
                        h[i + 1] = fmaf(z1, h[i + 1] - an[i + 1], an[i + 1]);
                        const __nv_bfloat162 h2 = __floats2bfloat162_rn(h[i], h[i + 1]);
                        hp[i >> 1] = *(const uint32_t*)&h2;
                    }
                }
                }
                __syncwarp();
                if (warp == 4 && lane == 0) GRU_TR(s, 7 + hf * 4);
                if constexpr (!FOLD) { if (lane == 0) tc::mbar_arrive(&gx_empty[hf]); }   // the slot may be refilled with the next step's half
                tc::tmem_st16(tlane + 384 + hf * 64 + gq * 16, h);
                if (ok && s == T - 1) {                         // h_T (fp32) back to the caller's hidden buffer
                    float* hg = p.h_state + ((size_t)row * 2 + dir) * GH + hf * 64 + gq * 16;
#pragma unroll
                    for (int i = 0; i < 16; i += 4) *(float4*)(hg + i) = make_float4(h[i], h[i + 1], h[i + 2], h[i + 3]);
                }
                if (hf == 0) {
#pragma unroll
                    for (int i = 0; i < 8; i++) keep[i] = hp[i];
                } else {
                    // acc_full[1] has been observed: every MMA of this step has consumed the old operand tile (and the
                    // tile store of h_{t-1} has read it). The layer output h_t leaves the SM from this tile (TMA store).
                    tc::sts128(a_row + sw0, make_uint4(keep[0], keep[1], keep[2], keep[3]));
                    tc::sts128(a_row + sw1, make_uint4(keep[4], keep[5], keep[6], keep[7]));
                    tc::sts128(a_row + 128 * 128 + sw0, make_uint4(hp[0], hp[1], hp[2], hp[3]));
                    tc::sts128(a_row + 128 * 128 + sw1, make_uint4(hp[4], hp[5], hp[6], hp[7]));
                    if constexpr (FOLD) {
                        if (s + 1 < T) write_x_atom(s + 1);     // the MMAs of this step have read the x atom too
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(&gx_empty[s & (RCF_XSLOTS - 1)]);   // x_t is not needed any more
                    }
                }
            }
            tc::tmem_st_wait();
            tc::fence_proxy_async();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(h_ready);
            if (warp == 4 && lane == 0) GRU_TR(s, 12);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, tc::TMEM_COLS); }
}

// images uint8 [n][row_stride/10 ...]: x[r][t][f] = img[r*row_stride + (t0+t)*10 + f] -> bf16 [n][T][64]; one thread per
// (window, t): columns 0..9 meet the bf16 high part of W_ih, columns 10..19 (the same counts again) its bf16 remainder
__global__ void gru_prep_kernel(const uint8_t* __restrict__ img, int64_t row_stride, int t0, __nv_bfloat16* __restrict__ x,
                                int64_t n, int T) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * T) return;
    const int64_t r = i / T; const int t = (int)(i - r * T);
    const uint8_t* src = img + r * row_stride + (int64_t)(t0 + t) * GF;
    uint32_t w[12];                                            // 24 bf16: [c0..c9 c0..c9 0 0 0 0]
    uint16_t v[24];
#pragma unroll
    for (int f = 0; f < GF; f++) {
        const __nv_bfloat16 b = __float2bfloat16_rn((float)src[f]);
        v[f] = v[GF + f] = *(const uint16_t*)&b;
    }
#pragma unroll
    for (int f = 2 * GF; f < 24; f++) v[f] = 0;
#pragma unroll
    for (int k = 0; k < 12; k++) w[k] = (uint32_t)v[2 * k] | ((uint32_t)v[2 * k + 1] << 16);
    uint4* dst = (uint4*)(x + i * GXK);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    dst[2] = make_uint4(w[8], w[9], w[10], w[11]);
#pragma unroll
    for (int k = 3; k < 8; k++) dst[k] = make_uint4(0u, 0u, 0u, 0u);
}

// FOLD path: x_t as the recurrence streams it: [m_blk][T][128 rows][16 bytes] = the 10 counts of (window, t) + 6 zero bytes;
// rows behind n are zero. One thread per (row of the padded chunk, t).
__global__ void gru_pack_x_kernel(const uint8_t* __restrict__ img, int64_t row_stride, int t0, uint8_t* __restrict__ xt, int64_t n,
                                  int64_t rows_alloc, int T) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows_alloc * T) return;
    const int64_t r = i / T; const int t = (int)(i - r * T);
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (r < n) {
        const uint8_t* src = img + r * row_stride + (int64_t)(t0 + t) * GF;
#pragma unroll
        for (int f = 0; f < GF; f++) w[f >> 2] |= (uint32_t)src[f] << (8 * (f & 3));
    }
    *(uint4*)(xt + (((r >> 7) * T + t) * 128 + (r & 127)) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
}

// dense1 (256 -> 5) per position: one THREAD per (window, t) row -- no cross-lane reduction; the 5 x 256 weights sit in
// shared memory (every lane reads the same address: broadcast), the row streams through 16-byte loads, 8 in flight.
// mode 0: write logits; mode 1: add softmax into acc
__global__ void __launch_bounds__(128) gru_head_kernel(const __nv_bfloat16* __restrict__ dec_out, int S, const float* __restrict__ w,
                                const float* __restrict__ b, float* __restrict__ dst, int64_t dst_row_stride, int t0,
                                int64_t n, int T, int mode) {
    __shared__ float4 w4[GC];                                  // classes 0..3 of channel k
    __shared__ float w1[GC];                                   // class 4
    for (int k = threadIdx.x; k < GC; k += blockDim.x) {
        w4[k] = make_float4(w[0 * GC + k], w[1 * GC + k], w[2 * GC + k], w[3 * GC + k]);
        w1[k] = w[4 * GC + k];
    }
    __syncthreads();
    const int64_t rows = n * T;
    for (int64_t wid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; wid < rows; wid += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = wid / T; const int t = (int)(wid - r * T);
        const uint4* x = (const uint4*)(dec_out + ((size_t)r * S + t + 1) * GC);
        float s0 = b[0], s1 = b[1], s2 = b[2], s3 = b[3], s4 = b[4];
#pragma unroll 1
        for (int k0 = 0; k0 < GC / 8; k0 += 8) {               // 8 loads of 8 channels in flight
            uint4 xv[8];
#pragma unroll
            for (int u = 0; u < 8; u++) xv[u] = __ldg(x + k0 + u);
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const __nv_bfloat162* x2 = (const __nv_bfloat162*)&xv[u];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float2 v = __bfloat1622float2(x2[j]);
                    const int k = (k0 + u) * 8 + 2 * j;
                    const float4 a = w4[k], c = w4[k + 1];
                    s0 = fmaf(v.x, a.x, s0); s1 = fmaf(v.x, a.y, s1); s2 = fmaf(v.x, a.z, s2); s3 = fmaf(v.x, a.w, s3);
                    s4 = fmaf(v.x, w1[k], s4);
                    s0 = fmaf(v.y, c.x, s0); s1 = fmaf(v.y, c.y, s1); s2 = fmaf(v.y, c.z, s2); s3 = fmaf(v.y, c.w, s3);
                    s4 = fmaf(v.y, w1[k + 1], s4);
                }
            }
        }
        float* o = dst + r * dst_row_stride + (int64_t)(t0 + t) * GCLS;
        if (mode == 0) {
            o[0] = s0; o[1] = s1; o[2] = s2; o[3] = s3; o[4] = s4;
        } else {
            const float m = fmaxf(fmaxf(fmaxf(s0, s1), fmaxf(s2, s3)), s4);
            const float e0 = expf(s0 - m), e1 = expf(s1 - m), e2 = expf(s2 - m), e3 = expf(s3 - m), e4 = expf(s4 - m);
            const float inv = 1.f / (e0 + e1 + e2 + e3 + e4);
            o[0] += e0 * inv; o[1] += e1 * inv; o[2] += e2 * inv; o[3] += e3 * inv; o[4] += e4 * inv;   // windows one after another
        }
    }
}

__global__ void gru_argmax_kernel(const float* __restrict__ acc, uint8_t* __restrict__ labels, int64_t n_pos) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pos) return;
    int best = 0; float bv = acc[i * GCLS];
#pragma unroll
    for (int c = 1; c < GCLS; c++) { const float v = acc[i * GCLS + c]; if (v > bv) { bv = v; best = c; } }
    labels[i] = (uint8_t)best;
}

int64_t gru_head_blocks(int64_t rows, int sms) {
    const int64_t want = (rows + 127) / 128;                   // one thread per row
    const int64_t cap = (int64_t)sms * 16;
    return want < cap ? want : cap;
}

struct GruWs {
    __nv_bfloat16 *xin, *enc_out, *dec_out;
    float* h_state;
    uint8_t* gx;
    uint8_t* xt;              // FOLD path of layer 1
    int64_t bytes;
};

GruWs gru_carve(void* base, int64_t size, int64_t chunk, int T) {
    pv::Arena a(base, size);
    GruWs w;
    const int S = T + 2;
    w.xin = a.take<__nv_bfloat16>(chunk * T * GXK);
    w.enc_out = a.take<__nv_bfloat16>(chunk * S * GC);
    w.dec_out = a.take<__nv_bfloat16>(chunk * S * GC);
    w.h_state = a.take<float>(chunk * 2 * GH);
    w.gx = a.take<uint8_t>((chunk / 128) * 2 * (int64_t)T * 2 * GX_HALF_BYTES);
    w.xt = a.take<uint8_t>((chunk / 128) * (int64_t)T * RCF_X_BYTES);
    w.bytes = pv::align_up(a.cur, 256);
    return w;
}

int64_t gru_chunk_for(int64_t n) {
    int64_t c = (n + 127) / 128 * 128;
    if (c < 128) c = 128;
    return c < GRU_MAX_CHUNK ? c : GRU_MAX_CHUNK;
}

// Input projection: rows [dir][gate r,z,n][128] = PyTorch's weight_ih rows, columns = k_total (zero padded).
// hi_lo: the weights are stored as bf16 high part + bf16 remainder in two column groups (the raw 0..254 counts are
// exact in bf16 and appear twice in the operand, so x * W_ih keeps ~16 mantissa bits instead of 8).
// Bias folded into gx: b_ir + b_hr, b_iz + b_hz, b_in.
void pack_gru_ih(const float* const w_ih[2], const float* const b_ih[2], const float* const b_hh[2], int in_dim, int k_total,
                 bool hi_lo, std::vector<uint16_t>& w, std::vector<float>& b) {
    w.assign((size_t)2 * GN3 * k_total, 0);
    b.assign((size_t)2 * GN3, 0.f);
    for (int dir = 0; dir < 2; dir++)
        for (int row = 0; row < GN3; row++) {
            uint16_t* wr = &w[((size_t)dir * GN3 + row) * k_total];
            for (int k = 0; k < in_dim; k++) {
                const float v = w_ih[dir][(size_t)row * in_dim + k];
                wr[k] = f2bf(v);
                if (hi_lo) wr[in_dim + k] = f2bf(v - bf2f(wr[k]));
            }
            b[(size_t)dir * GN3 + row] = b_ih[dir][row] + (row < 2 * GH ? b_hh[dir][row] : 0.f);
        }
}

// Recurrent weights in accumulator-column order: [dir][half][gate r,z,n][64] <- weight_hh row gate*128 + half*64 + j
void pack_gru_hh(const float* const w_hh[2], const float* const b_hh[2], std::vector<uint16_t>& w, std::vector<float>& bhn) {
    w.assign((size_t)2 * GN3 * GH, 0);
    bhn.assign((size_t)2 * GH, 0.f);
    for (int dir = 0; dir < 2; dir++) {
        for (int half = 0; half < 2; half++)
            for (int gate = 0; gate < 3; gate++)
                for (int j = 0; j < 64; j++) {
                    const int src = gate * GH + half * 64 + j;
                    uint16_t* wr = &w[((size_t)dir * GN3 + (half * 3 + gate) * 64 + j) * GH];
                    for (int k = 0; k < GH; k++) wr[k] = f2bf(w_hh[dir][(size_t)src * GH + k]);
                }
        for (int j = 0; j < GH; j++) bhn[(size_t)dir * GH + j] = b_hh[dir][2 * GH + j];
    }
}

}  // namespace

long long* pv_gru_trace_buffer = nullptr;
extern "C" long long* pv_gru_trace_ptr(void) { return pv_gru_trace_buffer; }

struct PvGruModel {
    __nv_bfloat16 *enc_wih, *dec_wih, *enc_whh, *dec_whh;
    float *enc_b, *dec_b, *enc_bhn, *dec_bhn, *dense_w, *dense_b;
    CUtensorMap map_enc_wih, map_dec_wih, map_enc_whh, map_dec_whh;
    // FOLD path of layer 1: input weights of r, z in accumulator-column order (hi | lo | 0), fp32 W_in [dir][10][128]
    __nv_bfloat16* enc_wx; float* enc_win; CUtensorMap map_enc_wx;
    int fold;
    int sms;
};

extern "C" int pv_gru_create(const PvGruWeights* hw, PvGruModel** out) {
    if (!hw || !out) return pv::set_error(PV_EINVAL, "null argument");
    if (int rc = pv::require_device()) return rc;
    PvGruModel* m = new PvGruModel();
    memset(m, 0, sizeof(*m));
    m->sms = pv::sm_count();
    std::vector<uint16_t> w; std::vector<float> b;
    pack_gru_ih(hw->enc_w_ih, hw->enc_b_ih, hw->enc_b_hh, GF, GXK, true, w, b);
    if (int rc = upload(&m->enc_wih, w.data(), w.size() * 2)) return rc;
    if (int rc = upload(&m->enc_b, b.data(), b.size() * 4)) return rc;
    pack_gru_ih(hw->dec_w_ih, hw->dec_b_ih, hw->dec_b_hh, GC, GC, false, w, b);
    if (int rc = upload(&m->dec_wih, w.data(), w.size() * 2)) return rc;
    if (int rc = upload(&m->dec_b, b.data(), b.size() * 4)) return rc;
    pack_gru_hh(hw->enc_w_hh, hw->enc_b_hh, w, b);
    if (int rc = upload(&m->enc_whh, w.data(), w.size() * 2)) return rc;
    if (int rc = upload(&m->enc_bhn, b.data(), b.size() * 4)) return rc;
    pack_gru_hh(hw->dec_w_hh, hw->dec_b_hh, w, b);
    if (int rc = upload(&m->dec_whh, w.data(), w.size() * 2)) return rc;
    if (int rc = upload(&m->dec_bhn, b.data(), b.size() * 4)) return rc;
    if (int rc = upload(&m->dense_w, hw->dense_w, GCLS * GC * 4)) return rc;
    if (int rc = upload(&m->dense_b, hw->dense_b, GCLS * 4)) return rc;
    if (int rc = make_map2(&m->map_enc_wih, m->enc_wih, 2 * GN3, GXK)) return rc;
    if (int rc = make_map2(&m->map_dec_wih, m->dec_wih, 2 * GN3, GC)) return rc;
    if (int rc = make_map2(&m->map_enc_whh, m->enc_whh, 2 * GN3, GH)) return rc;
    if (int rc = make_map2(&m->map_dec_whh, m->dec_whh, 2 * GN3, GH)) return rc;
    {   // layer-1 weights for the FOLD recurrence
        std::vector<uint16_t> wx((size_t)2 * GN3 * GXK, 0);
        std::vector<float> win((size_t)2 * GF * GH, 0.f);
        for (int dir = 0; dir < 2; dir++) {
            for (int half = 0; half < 2; half++)
                for (int gate = 0; gate < 2; gate++)           // r, z only: the candidate's rows stay zero
                    for (int j = 0; j < 64; j++) {
                        const int src = gate * GH + half * 64 + j;
                        uint16_t* wr = &wx[((size_t)dir * GN3 + (half * 3 + gate) * 64 + j) * GXK];
                        for (int k = 0; k < GF; k++) {
                            const float v = hw->enc_w_ih[dir][(size_t)src * GF + k];
                            wr[k] = f2bf(v);
                            wr[GF + k] = f2bf(v - bf2f(wr[k]));
                        }
                    }
            for (int f = 0; f < GF; f++)
                for (int j = 0; j < GH; j++) win[((size_t)dir * GF + f) * GH + j] = hw->enc_w_ih[dir][(size_t)(2 * GH + j) * GF + f];
        }
        if (int rc = upload(&m->enc_wx, wx.data(), wx.size() * 2)) return rc;
        if (int rc = upload(&m->enc_win, win.data(), win.size() * 4)) return rc;
        if (int rc = make_map2(&m->map_enc_wx, m->enc_wx, 2 * GN3, GXK)) return rc;
        // measured (round 2, tools/probe_gru.py): 11.54 vs 11.80 ms at 65536 windows, 2.96 vs 3.01 at 16384, but 1.01 vs 0.78 ms at
        // 256 (the candidate gate's input part on the CUDA cores lengthens a step from 3.2 to 5.7 us): opt-in, PV_GRU_FOLD=1
        const char* e = getenv("PV_GRU_FOLD");
        m->fold = (e && atoi(e)) ? 1 : 0;
    }
    PV_CUDA_CHECK(cudaFuncSetAttribute(gru_recur_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RC_SMEM));
    PV_CUDA_CHECK(cudaFuncSetAttribute(gru_recur_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RCF_SMEM));
    *out = m;
    return PV_OK;
}

extern "C" void pv_gru_destroy(PvGruModel* m) {
    if (!m) return;
    cudaFree(m->enc_wih); cudaFree(m->dec_wih); cudaFree(m->enc_whh); cudaFree(m->dec_whh);
    cudaFree(m->enc_b); cudaFree(m->dec_b); cudaFree(m->enc_bhn); cudaFree(m->dec_bhn); cudaFree(m->dense_w); cudaFree(m->dense_b);
    cudaFree(m->enc_wx); cudaFree(m->enc_win);
    delete m;
}

extern "C" int64_t pv_gru_workspace_bytes(int64_t max_batch, int32_t seq_len) {
    return gru_carve(nullptr, 0, gru_chunk_for(max_batch), seq_len).bytes;
}

namespace {

// one TransducerGRU.forward over rows [0, nb) of a chunk; h_state [nb][2][GH] in/out
int gru_forward_chunk(PvGruModel* m, const GruWs& w, const CUtensorMap& map_x, const CUtensorMap& map_enc,
                      const CUtensorMap& map_dec, const uint8_t* images, int64_t img_row_stride, int t0, int64_t nb, int64_t chunk, int T,
                      float* h_state, cudaStream_t st) {
    const int S = T + 2;
    const int64_t elems = nb * T;
    const int m_blks = (int)((nb + 127) / 128);
    pv::prof_begin(pv::FAM_GRU_MISC, st);
    if (m->fold) {
        const int64_t rows = (int64_t)m_blks * 128;
        gru_pack_x_kernel<<<(unsigned)((rows * T + 255) / 256), 256, 0, st>>>(images, img_row_stride, t0, w.xt, nb, rows, T);
    } else {
        gru_prep_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, st>>>(images, img_row_stride, t0, w.xin, nb, T);
    }
    PV_CUDA_CHECK(cudaGetLastError());
    pv::prof_end(pv::FAM_GRU_MISC, st, 1);
    for (int layer = 0; layer < 2; layer++) {
        const bool fold = layer == 0 && m->fold;
        RecurParams rp;
        memset(&rp, 0, sizeof(rp));
        if (fold) {
            // layer 1 without a gx round trip: the recurrence projects its 10-feature input itself
            rp.xt = w.xt; rp.bias = m->enc_b; rp.w_in = m->enc_win;
            rp.b_hn = m->enc_bhn; rp.h_state = h_state; rp.out = w.enc_out; rp.M = (int)nb; rp.T = T; rp.S = S; rp.trace = nullptr;
            pv::prof_begin(pv::FAM_GRU_STEP, st);
            gru_recur_kernel<true><<<(unsigned)(m_blks * 2), RC_THREADS, RCF_SMEM, st>>>(m->map_enc_whh, m->map_enc_wx, map_enc, rp);
            PV_CUDA_CHECK(cudaGetLastError());
            pv::prof_end(pv::FAM_GRU_STEP, st, 1);
            continue;
        }
        // gx for every time step and both directions: [nb x T, K_in] x [K_in, 768]
        tc::GemmShape g;
        memset(&g, 0, sizeof(g));
        g.M = (int)nb; g.m_blks = m_blks; g.n_blks = 2 * GN3 / tc::BLOCK_N; g.dirs = 1; g.slots = T;
        g.kb0 = 0; g.kb1 = layer == 0 ? GXK / tc::BLOCK_K : GC / tc::BLOCK_K;
        g.a1_slot[0] = layer == 0 ? 0 : 1;                     // x_t sits in slot t of xin, slot t + 1 of enc_out
        GxEpilogue ge;
        ge.bias = layer == 0 ? m->enc_b : m->dec_b; ge.gx = (uint4*)w.gx; ge.T = T; ge.rows_alloc = (int)chunk;
        pv::prof_begin(pv::FAM_GRU_GX, st);
        if (int rc = launch_gemm(map_x, layer == 0 ? map_x : map_enc, layer == 0 ? m->map_enc_wih : m->map_dec_wih, g, ge,
                                 m->sms, st)) return rc;
        pv::prof_end(pv::FAM_GRU_GX, st, 1);
        rp.gx = w.gx; rp.b_hn = layer == 0 ? m->enc_bhn : m->dec_bhn; rp.h_state = h_state;
        rp.out = layer == 0 ? w.enc_out : w.dec_out; rp.M = (int)nb; rp.T = T; rp.S = S; rp.trace = nullptr;
#ifdef PV_TRACE
        { static long long* tr = nullptr; if (!tr) { cudaMalloc((void**)&tr, 16 * 16 * 8); cudaMemset(tr, 0, 16 * 16 * 8); } rp.trace = tr; pv_gru_trace_buffer = tr; }
#endif
        pv::prof_begin(pv::FAM_GRU_STEP, st);
        gru_recur_kernel<false><<<(unsigned)(m_blks * 2), RC_THREADS, RC_SMEM, st>>>(layer == 0 ? m->map_enc_whh : m->map_dec_whh,
                                                                                        layer == 0 ? m->map_enc_whh : m->map_dec_whh,
                                                                                        layer == 0 ? map_enc : map_dec, rp);
        PV_CUDA_CHECK(cudaGetLastError());
        pv::prof_end(pv::FAM_GRU_STEP, st, 1);
    }
    return PV_OK;
}

}  // namespace

extern "C" int pv_gru_forward(PvGruModel* m, const uint8_t* images, int64_t n, int32_t T, float* hidden, float* logits,
                              void* workspace, int64_t workspace_bytes, void* stream_) {
    if (!m || !images || !hidden || !logits || !workspace) return pv::set_error(PV_EINVAL, "null argument");
    if (n <= 0 || T <= 0) return PV_OK;
    cudaStream_t st = (cudaStream_t)stream_;
    int64_t chunk = gru_chunk_for(n);
    while (chunk > 128 && gru_carve(nullptr, 0, chunk, T).bytes > workspace_bytes) chunk -= 128;
    const GruWs w = gru_carve(workspace, workspace_bytes, chunk, T);
    if (w.bytes > workspace_bytes) return pv::set_error(PV_EINVAL, "workspace too small: need at least %lld bytes", (long long)w.bytes);
    const int S = T + 2;
    CUtensorMap map_x, map_enc, map_dec;
    if (int rc = make_map3(&map_x, w.xin, chunk, T, GXK, (int64_t)T * GXK)) return rc;
    if (int rc = make_map3(&map_enc, w.enc_out, chunk, S, GC, (int64_t)S * GC)) return rc;
    if (int rc = make_map3(&map_dec, w.dec_out, chunk, S, GC, (int64_t)S * GC)) return rc;
    for (int64_t off = 0; off < n; off += chunk) {
        const int64_t nb = n - off < chunk ? n - off : chunk;
        if (int rc = gru_forward_chunk(m, w, map_x, map_enc, map_dec, images + off * T * GF, (int64_t)T * GF, 0, nb, chunk, T,
                                       hidden + off * 2 * GH, st)) return rc;
        pv::prof_begin(pv::FAM_GRU_HEAD, st);
        gru_head_kernel<<<(unsigned)gru_head_blocks(nb * T, m->sms), 128, 0, st>>>(w.dec_out, S, m->dense_w, m->dense_b,
                                                                                logits + off * T * GCLS, (int64_t)T * GCLS, 0, nb, T, 0);
        PV_CUDA_CHECK(cudaGetLastError());
        pv::prof_end(pv::FAM_GRU_HEAD, st, 1);
    }
    return PV_OK;
}

extern "C" int pv_gru_predict_chunks(PvGruModel* m, const uint8_t* images, int64_t n, int32_t L, int32_t window, int32_t stride,
                                     float* prob_sum, uint8_t* labels, void* workspace, int64_t workspace_bytes, void* stream_) {
    if (!m || !images || !prob_sum || !labels || !workspace) return pv::set_error(PV_EINVAL, "null argument");
    if (n <= 0 || L <= 0) return PV_OK;
    if (window <= 0 || stride <= 0 || window > L) return pv::set_error(PV_EINVAL, "bad window/stride");
    cudaStream_t st = (cudaStream_t)stream_;
    const int T = window, S = T + 2;
    int64_t chunk = gru_chunk_for(n);
    while (chunk > 128 && gru_carve(nullptr, 0, chunk, T).bytes > workspace_bytes) chunk -= 128;
    const GruWs w = gru_carve(workspace, workspace_bytes, chunk, T);
    if (w.bytes > workspace_bytes) return pv::set_error(PV_EINVAL, "workspace too small: need at least %lld bytes", (long long)w.bytes);
    CUtensorMap map_x, map_enc, map_dec;
    if (int rc = make_map3(&map_x, w.xin, chunk, T, GXK, (int64_t)T * GXK)) return rc;
    if (int rc = make_map3(&map_enc, w.enc_out, chunk, S, GC, (int64_t)S * GC)) return rc;
    if (int rc = make_map3(&map_dec, w.dec_out, chunk, S, GC, (int64_t)S * GC)) return rc;
    PV_CUDA_CHECK(cudaMemsetAsync(prob_sum, 0, (size_t)n * L * GCLS * 4, st));
    for (int64_t off = 0; off < n; off += chunk) {
        const int64_t nb = n - off < chunk ? n - off : chunk;
        PV_CUDA_CHECK(cudaMemsetAsync(w.h_state, 0, (size_t)nb * 2 * GH * 4, st));          // hidden = zeros (:63)
        for (int t0 = 0; t0 + window <= L; t0 += stride) {                                   // :70-73
            if (int rc = gru_forward_chunk(m, w, map_x, map_enc, map_dec, images + off * L * GF, (int64_t)L * GF, t0, nb, chunk, T,
                                           w.h_state, st)) return rc;
            pv::prof_begin(pv::FAM_GRU_HEAD, st);
            gru_head_kernel<<<(unsigned)gru_head_blocks(nb * T, m->sms), 128, 0, st>>>(w.dec_out, S, m->dense_w, m->dense_b,
                                                                                    prob_sum + off * L * GCLS, (int64_t)L * GCLS, t0, nb, T, 1);
            PV_CUDA_CHECK(cudaGetLastError());
            pv::prof_end(pv::FAM_GRU_HEAD, st, 1);
        }
    }
    const int64_t n_pos = n * L;
    gru_argmax_kernel<<<(unsigned)((n_pos + 255) / 256), 256, 0, st>>>(prob_sum, labels, n_pos);
    PV_CUDA_CHECK(cudaGetLastError());
    return PV_OK;
}
