// Model M-A on sm_100a: pepper_variant TransducerGRU.forward
// (/root/reference/pepper_variant/modules/python/models/simple_model.py:48-82): biLSTM(26->256) -> biLSTM(512->256)
// -> flatten(33*512) -> 5 x (Linear + SELU) -> Linear(512->3) -> softmax, batch_first, h0 = c0 = 0, eval mode.
//
// Every dense product runs on the tcgen05/TMA GEMM core (tc_gemm.cuh):
//   * one launch per time step advances BOTH directions of a layer: A = [h_{t-1} | x_t] (two tensor maps),
//     W = [W_hh | W_ih] re-ordered so a 256-column tile holds i,f,g,o of 64 hidden units; the LSTM cell
//     (sigmoid/tanh, c update, h = o*tanh(c)) is the GEMM epilogue, h_t is written as bf16 straight into the
//     layer output [B][T+2][512] where the next step (and the next layer) TMA-loads it from;
//   * the encoder input (raw int16 counts) is split exactly into bf16 hi + lo parts (|x| < 2^16 exact) and W_ih into
//     a bf16 high part + bf16 remainder, so the input projection keeps ~16 mantissa bits even on unnormalised,
//     deep-coverage counts;
//   * the MLP runs the same kernel with a bias+SELU epilogue; the 512->3 head, softmax and argmax are one small kernel.
#include "common.cuh"
#include "tc_gemm.cuh"
#include "infer_common.cuh"
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace {

constexpr int T = PV_WINDOW;        // 33 time steps
constexpr int F = PV_FEATURES;      // 26 input features
constexpr int H = 256;              // hidden size of both LSTMs
constexpr int C = 2 * H;            // channels of a layer output
constexpr int S = T + 2;            // slots per window in a layer output (slot t+1 = time t; 0 and T+1 unused here)
constexpr int XK = 128;             // encoder input, two k blocks: [x_hi(26) x_lo(26) 0(12)] [x_hi(26) 0(38)]
constexpr int ENC_K = H + XK;       // 384
constexpr int DEC_K = H + C;        // 768
constexpr int LIN = 512;
constexpr int FLAT = T * C;         // 16896

// ---- epilogues ---------------------------------------------------------------------------------------------
// DEEP = true  (decoder, 12 k blocks per tile: bound by operand latency -- TMA from L2 ~2000 cycles, one k block of MMA
//               ~550): FOUR 48 KB stages in flight. That leaves 32 KB for the epilogue: ONE c-state buffer (every thread
//               owns its slots, so it re-fills them for the next tile the moment it has read them), biases through L1.
// DEEP = false (encoder, 6 k blocks per tile: bound by the epilogue): three stages, biases in shared memory, the c tile
//               double-buffered and requested by the kernel one tile ahead.
template <bool DEEP>
struct LstmEpilogueT {
    static constexpr int kStages = DEEP ? 4 : 3;
    static constexpr int kBiasBytes = DEEP ? 0 : 2 * 4 * 4 * 64 * 4;      // whole layer: [dir][n_blk][gate][64] fp32
    static constexpr int kStateBytes = tc::EPI_THREADS * 2 * 64;           // one tile: 256 threads x 2 chunks x 16 fp32
    static constexpr int kSmemBytes = kBiasBytes + (DEEP ? 1 : 2) * kStateBytes;
    static constexpr bool kInlinePrefetch = DEEP;

    const float* bias;        // [dirs][n_blks][4 gates][64]  (b_ih + b_hh, tile order)
    float* c_state;           // cell state, fp32, private TILE layout [m_blk][dir][n_blk][half][chunk][j][row 0..127][4]:
                              // every warp-wide 16-byte load/store of the epilogue touches 512 contiguous bytes
    __nv_bfloat16* out;       // [M][S][C]
    int n_blks;
    int out_slot[2];
    int first;                // c_{t-1} == 0
    int debug;                // PV_DEBUG_EPI: 1 = skip the cell math (timing experiments only)

    __device__ void setup(uint8_t* scratch, int te) const {
        if constexpr (!DEEP) {
            float4* sb = (float4*)scratch;
            const float4* gb = (const float4*)bias;
            for (int i = te; i < kBiasBytes / 16; i += tc::EPI_THREADS) sb[i] = __ldg(gb + i);
        }
    }
    // thread te keeps its 2 x 64 bytes of a tile at [(cc*4 + j) * 256 + te] 16-byte slots: conflict-free both ways
    __device__ __forceinline__ float* c_ptr(int dir, int n_blk, int row, int half, int cc, int j) const {
        const size_t tile = ((size_t)(row >> 7) * 2 + dir) * n_blks + n_blk;
        return c_state + ((((tile * 2 + half) * 2 + cc) * 4 + j) * 128 + (row & 127)) * 4;
    }
    __device__ __forceinline__ void fetch_chunk(uint8_t* scratch, int dir, int n_blk, int row, int half, int cc, int te) const {
#pragma unroll
        for (int j = 0; j < 4; j++) tc::cp_async16(scratch + ((cc * 4 + j) * tc::EPI_THREADS + te) * 16, c_ptr(dir, n_blk, row, half, cc, j));
    }
    // DEEP: the kernel's first tile only (later tiles are requested from inside operator()); else: every next tile
    __device__ void prefetch(uint8_t* scratch, int buf, int dir, int n_blk, int row, bool ok, int half, int te) const {
        if (first || !ok) return;
        uint8_t* dst = scratch + kBiasBytes + (DEEP ? 0 : buf) * kStateBytes;
        fetch_chunk(dst, dir, n_blk, row, half, 0, te);
        fetch_chunk(dst, dir, n_blk, row, half, 1, te);
    }
    __device__ void operator()(uint8_t* scratch, int buf, int dir, int n_blk, int row, bool ok, uint32_t taddr, int half, int te,
                               const tc::NextTile& nx) const {
        const float* sb = (DEEP ? bias : (const float*)scratch) + (size_t)((dir * n_blks + n_blk) * 4) * 64;
        uint8_t* cst = scratch + kBiasBytes + (DEEP ? 0 : buf) * kStateBytes;
#pragma unroll 1
        for (int cc = 0; cc < 2; cc++) {
            const int ch = half * 2 + cc;                      // 16 hidden units per chunk
            float ai[16], af[16], ag[16], ao[16], c[16];
            if (debug != 3) {
                tc::tmem_ld16(taddr + 0 * 64 + ch * 16, ai);
                tc::tmem_ld16(taddr + 1 * 64 + ch * 16, af);
                tc::tmem_ld16(taddr + 2 * 64 + ch * 16, ag);
                tc::tmem_ld16(taddr + 3 * 64 + ch * 16, ao);
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++) { ai[i] = 0.1f * i; af[i] = 0.2f; ag[i] = -0.1f * i; ao[i] = 0.3f; }
            }
            if (first || !ok) {
#pragma unroll
                for (int i = 0; i < 16; i++) c[i] = 0.f;
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float4 v = *(const float4*)(cst + ((cc * 4 + j) * tc::EPI_THREADS + te) * 16);
                    c[4 * j] = v.x; c[4 * j + 1] = v.y; c[4 * j + 2] = v.z; c[4 * j + 3] = v.w;
                }
            }
            // these slots are free again: request the same chunk of the NEXT tile's cell state
            if constexpr (DEEP) {
                if (nx.valid && nx.ok && !first) fetch_chunk(cst, nx.dir, nx.n_blk, nx.row, half, cc, te);
                tc::cp_async_commit();
            }
            tc::tmem_ld_wait();
            if (debug == 1) continue;
            if (ok) {
                float bi[16], bf[16], bg[16], bo[16];              // the whole warp reads the same 64 bytes (L1 or smem broadcast)
                if constexpr (DEEP) {
                    ld16(sb + 0 * 64 + ch * 16, bi); ld16(sb + 1 * 64 + ch * 16, bf);
                    ld16(sb + 2 * 64 + ch * 16, bg); ld16(sb + 3 * 64 + ch * 16, bo);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        bi[i] = sb[0 * 64 + ch * 16 + i]; bf[i] = sb[1 * 64 + ch * 16 + i];
                        bg[i] = sb[2 * 64 + ch * 16 + i]; bo[i] = sb[3 * 64 + ch * 16 + i];
                    }
                }
                // stage-wise over the 16 cells (not cell by cell): every stage is 16 independent instructions, so the MUFU
                // and FMA pipes stay full instead of waiting on one cell's dependency chain
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    ai[i] = 0.5f * (ai[i] + bi[i]); af[i] = 0.5f * (af[i] + bf[i]);
                    ag[i] = ag[i] + bg[i];          ao[i] = 0.5f * (ao[i] + bo[i]);
                }
#pragma unroll
                for (int i = 0; i < 16; i++) ai[i] = tanh_f(ai[i]);
#pragma unroll
                for (int i = 0; i < 16; i++) af[i] = tanh_f(af[i]);
#pragma unroll
                for (int i = 0; i < 16; i++) ag[i] = tanh_f(ag[i]);
#pragma unroll
                for (int i = 0; i < 16; i++) ao[i] = tanh_f(ao[i]);
#pragma unroll
                for (int i = 0; i < 16; i++)                    // sigmoid(x) = 0.5 tanh(x/2) + 0.5
                    c[i] = fmaf(fmaf(0.5f, af[i], 0.5f), c[i], fmaf(0.5f, ai[i], 0.5f) * ag[i]);
#pragma unroll
                for (int i = 0; i < 16; i++) ag[i] = tanh_f(c[i]);
                uint32_t hp[8];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(fmaf(0.5f, ao[i], 0.5f) * ag[i], fmaf(0.5f, ao[i + 1], 0.5f) * ag[i + 1]);
                    hp[i >> 1] = *(const uint32_t*)&h2;
                }
                if (debug == 2 && hp[0] != 0x12345678u) continue;
#pragma unroll
                for (int j = 0; j < 4; j++)
                    *(float4*)c_ptr(dir, n_blk, row, half, cc, j) = make_float4(c[4 * j], c[4 * j + 1], c[4 * j + 2], c[4 * j + 3]);
                __nv_bfloat16* op = out + ((size_t)row * S + (dir ? out_slot[1] : out_slot[0])) * C + dir * H + n_blk * 64 + ch * 16;
                *(uint4*)op = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                *(uint4*)(op + 8) = make_uint4(hp[4], hp[5], hp[6], hp[7]);
            }
        }
    }
};

struct SeluEpilogue {
    static constexpr int kStages = 4;
    static constexpr int kSmemBytes = 0;
    static constexpr bool kInlinePrefetch = false;
    __device__ void setup(uint8_t*, int) const {}
    __device__ void prefetch(uint8_t*, int, int, int, int, bool, int, int) const {}

    const float* bias;        // [N]
    __nv_bfloat16* out;       // [M][ldo]
    int ldo;

    __device__ void operator()(uint8_t*, int, int dir, int n_blk, int row, bool ok, uint32_t taddr, int half, int,
                               const tc::NextTile&) const {
        (void)dir;
        const float alpha = 1.6732632423543772f, lambda = 1.0507009873554805f;
#pragma unroll 1
        for (int ch = half * 8; ch < half * 8 + 8; ch++) {
            float a[16], bb[16];
            tc::tmem_ld16(taddr + ch * 16, a);
            const int col = n_blk * tc::BLOCK_N + ch * 16;
            ld16(bias + col, bb);
            tc::tmem_ld_wait();
            if (ok) {
                uint32_t hp[8];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const float x0 = a[i] + bb[i], x1 = a[i + 1] + bb[i + 1];
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(lambda * (x0 > 0.f ? x0 : alpha * (__expf(x0) - 1.f)),
                                                                    lambda * (x1 > 0.f ? x1 : alpha * (__expf(x1) - 1.f)));
                    hp[i >> 1] = *(const uint32_t*)&h2;
                }
                __nv_bfloat16* op = out + (size_t)row * ldo + col;
                *(uint4*)op = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                *(uint4*)(op + 8) = make_uint4(hp[4], hp[5], hp[6], hp[7]);
            }
        }
    }
};

// ---- small kernels -----------------------------------------------------------------------------------------
// windows int16 [n][33][26] -> xhl bf16 [n][33][128]: cols 0..25 = x_hi, 26..51 = x_lo (x == x_hi + x_lo exactly; both
// meet bf16(W_ih)), cols 64..89 = x_hi again (meets the bf16 remainder of W_ih), rest 0
__global__ void prep_input_kernel(const int16_t* __restrict__ win, __nv_bfloat16* __restrict__ xhl, int64_t n, int wrap_int8) {
    // a thread writes 8 consecutive output columns (one 16-byte store); 16 threads share one 52-byte input row. The column
    // blocks change their source at even columns only (26, 52, 64, 90) and a row starts 4-byte aligned (52 bytes per row), so
    // a thread reads its features as pairs: four 32-bit loads instead of eight 16-bit ones
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // over n*33*16
    if (i >= n * T * (XK / 8)) return;
    const int c0 = (int)(i & (XK / 8 - 1)) * 8;
    const int64_t rt = i / (XK / 8);
    const uint32_t* row = (const uint32_t*)(win + rt * F);
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
        const int col = c0 + k;
        const int f = col < 2 * F ? (col < F ? col : col - F) : (col >= 64 && col < 64 + F ? col - 64 : -1);
        uint32_t b2 = 0u;
        if (f >= 0) {
            const uint32_t pair = __ldg(row + (f >> 1));
            int x0 = (int)(int16_t)(pair & 0xffffu), x1 = (int)(int16_t)(pair >> 16);
            if (wrap_int8) { x0 = (int)(int8_t)x0; x1 = (int)(int8_t)x1; }      // DataStore.py:68 int8 round trip
            const float f0 = (float)x0, f1 = (float)x1;
            const float h0 = __bfloat162float(__float2bfloat16_rn(f0)), h1 = __bfloat162float(__float2bfloat16_rn(f1));
            const bool lo = col >= F && col < 2 * F;
            const __nv_bfloat162 v = __floats2bfloat162_rn(lo ? f0 - h0 : h0, lo ? f1 - h1 : h1);
            b2 = *(const uint32_t*)&v;
        }
        o[k >> 1] = b2;
    }
    *(uint4*)(xhl + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
}

// output_layer_type (512 -> 3) + softmax + argmax; one warp per window
__global__ void head_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                            float* __restrict__ probs, uint8_t* __restrict__ argmax, int64_t n) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int k = lane; k < LIN; k += 32) {
        const float v = __bfloat162float(x[row * LIN + k]);
        s0 += v * w[k]; s1 += v * w[LIN + k]; s2 += v * w[2 * LIN + k];
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, d); s1 += __shfl_xor_sync(0xffffffffu, s1, d); s2 += __shfl_xor_sync(0xffffffffu, s2, d);
    }
    if (lane == 0) {
        s0 += b[0]; s1 += b[1]; s2 += b[2];
        const float m = fmaxf(s0, fmaxf(s1, s2));
        const float e0 = expf(s0 - m), e1 = expf(s1 - m), e2 = expf(s2 - m);
        const float inv = 1.f / (e0 + e1 + e2);
        probs[row * 3 + 0] = e0 * inv; probs[row * 3 + 1] = e1 * inv; probs[row * 3 + 2] = e2 * inv;
        if (argmax) argmax[row] = (uint8_t)(s1 > s0 ? (s2 > s1 ? 2 : 1) : (s2 > s0 ? 2 : 0));   // first maximum wins (torch.argmax)
    }
}

}  // namespace

#ifdef PV_TRACE
long long* pv_trace_buffer = nullptr;
extern "C" long long* pv_trace_ptr(void) { return pv_trace_buffer; }
#endif

struct PvLstmModel {
    __nv_bfloat16 *enc_w, *dec_w, *lin_w[5];
    float *enc_b, *dec_b, *lin_b[5], *out_w, *out_b;
    CUtensorMap map_enc_w, map_dec_w, map_lin_w[5];
    int device;
    int sms;
    // host-wrapper buffers
    void* ws; int64_t ws_bytes;
    int16_t* win_dev; float* probs_dev; uint8_t* arg_dev; int64_t io_cap;
    cudaStream_t stream;
    // small batches (the reference calls the model on 512 windows at a time): the 72 launches of a pass replayed as one CUDA graph
    struct GraphEntry { void* ws; int64_t ws_bytes; int64_t n; int32_t wrap; int32_t uses; cudaGraphExec_t exec; };
    GraphEntry graphs[8];
    int n_graphs;
    cudaStream_t capture_stream;
};

namespace {

struct Ws {
    __nv_bfloat16 *xhl, *enc_out, *dec_out, *act0, *act1;
    float* c_state;
    int16_t* win_stage; float* probs_stage; uint8_t* arg_stage;   // fixed addresses for the graph of a small pass
    int64_t bytes;
};
constexpr int64_t GRAPH_MAX_WINDOWS = 4096;

Ws carve_ws(void* base, int64_t size, int64_t chunk) {
    pv::Arena a(base, size);
    Ws w;
    w.xhl = a.take<__nv_bfloat16>(chunk * T * XK);
    w.enc_out = a.take<__nv_bfloat16>(chunk * S * C);
    w.dec_out = a.take<__nv_bfloat16>(chunk * S * C);
    w.act0 = a.take<__nv_bfloat16>(chunk * LIN);
    w.act1 = a.take<__nv_bfloat16>(chunk * LIN);
    w.c_state = a.take<float>(((chunk + 127) / 128 * 128) * 2 * H);
    const int64_t stage = chunk < GRAPH_MAX_WINDOWS ? chunk : GRAPH_MAX_WINDOWS;
    w.win_stage = a.take<int16_t>(stage * T * F);
    w.probs_stage = a.take<float>(stage * 3);
    w.arg_stage = a.take<uint8_t>(stage);
    w.bytes = pv::align_up(a.cur, 256);
    return w;
}

int64_t chunk_for(int64_t n) {
    int64_t c = (n + 127) / 128 * 128;
    if (c < 128) c = 128;
    return c < MAX_CHUNK ? c : MAX_CHUNK;
}

// pack one LSTM layer: rows in tile order (dir, n_blk, gate, j) <- PyTorch row gate*256 + n_blk*64 + j
void pack_lstm(const float* const w_ih[2], const float* const w_hh[2], const float* const b_ih[2], const float* const b_hh[2],
               int in_dim, int k_total, bool hi_lo, std::vector<uint16_t>& w, std::vector<float>& b) {
    w.assign((size_t)2 * 4 * H * k_total, 0);
    b.assign((size_t)2 * 4 * H, 0.f);
    for (int dir = 0; dir < 2; dir++)
        for (int nb = 0; nb < 4; nb++)
            for (int gate = 0; gate < 4; gate++)
                for (int j = 0; j < 64; j++) {
                    const int src = gate * H + nb * 64 + j;
                    const size_t dst = ((size_t)(dir * 4 + nb) * 4 + gate) * 64 + j;
                    uint16_t* wr = &w[dst * k_total];
                    for (int k = 0; k < H; k++) wr[k] = f2bf(w_hh[dir][(size_t)src * H + k]);
                    for (int k = 0; k < in_dim; k++) {
                        const float wv = w_ih[dir][(size_t)src * in_dim + k];
                        const uint16_t v = f2bf(wv);
                        wr[H + k] = v;
                        if (hi_lo) { wr[H + in_dim + k] = v; wr[H + 64 + k] = f2bf(wv - bf2f(v)); }
                    }
                    b[dst] = b_ih[dir][src] + b_hh[dir][src];
                }
}

}  // namespace

extern "C" int pv_lstm_create(const PvLstmWeights* hw, PvLstmModel** out) {
    if (!hw || !out) return pv::set_error(PV_EINVAL, "null argument");
    if (int rc = pv::require_device()) return rc;
    PvLstmModel* m = new PvLstmModel();
    memset(m, 0, sizeof(*m));
    cudaGetDevice(&m->device);
    m->sms = pv::sm_count();
    std::vector<uint16_t> w; std::vector<float> b;
    pack_lstm(hw->enc_w_ih, hw->enc_w_hh, hw->enc_b_ih, hw->enc_b_hh, F, ENC_K, true, w, b);
    if (int rc = upload(&m->enc_w, w.data(), w.size() * 2)) return rc;
    if (int rc = upload(&m->enc_b, b.data(), b.size() * 4)) return rc;
    pack_lstm(hw->dec_w_ih, hw->dec_w_hh, hw->dec_b_ih, hw->dec_b_hh, C, DEC_K, false, w, b);
    if (int rc = upload(&m->dec_w, w.data(), w.size() * 2)) return rc;
    if (int rc = upload(&m->dec_b, b.data(), b.size() * 4)) return rc;
    for (int l = 0; l < 5; l++) {
        const int k = l == 0 ? FLAT : LIN;
        w.resize((size_t)LIN * k);
        for (size_t i = 0; i < w.size(); i++) w[i] = f2bf(hw->lin_w[l][i]);
        if (int rc = upload(&m->lin_w[l], w.data(), w.size() * 2)) return rc;
        if (int rc = upload(&m->lin_b[l], hw->lin_b[l], LIN * 4)) return rc;
        if (int rc = make_map2(&m->map_lin_w[l], m->lin_w[l], LIN, k)) return rc;
    }
    if (int rc = upload(&m->out_w, hw->out_w, 3 * LIN * 4)) return rc;
    if (int rc = upload(&m->out_b, hw->out_b, 3 * 4)) return rc;
    if (int rc = make_map2(&m->map_enc_w, m->enc_w, 2 * 4 * H, ENC_K)) return rc;
    if (int rc = make_map2(&m->map_dec_w, m->dec_w, 2 * 4 * H, DEC_K)) return rc;
    *out = m;
    return PV_OK;
}

extern "C" void pv_lstm_destroy(PvLstmModel* m) {
    if (!m) return;
    cudaFree(m->enc_w); cudaFree(m->dec_w); cudaFree(m->enc_b); cudaFree(m->dec_b);
    for (int l = 0; l < 5; l++) { cudaFree(m->lin_w[l]); cudaFree(m->lin_b[l]); }
    cudaFree(m->out_w); cudaFree(m->out_b);
    if (m->ws) cudaFree(m->ws);
    if (m->win_dev) cudaFree(m->win_dev);
    if (m->probs_dev) cudaFree(m->probs_dev);
    if (m->arg_dev) cudaFree(m->arg_dev);
    if (m->stream) cudaStreamDestroy(m->stream);
    for (int i = 0; i < m->n_graphs; i++) if (m->graphs[i].exec) cudaGraphExecDestroy(m->graphs[i].exec);
    if (m->capture_stream) cudaStreamDestroy(m->capture_stream);
    delete m;
}

extern "C" int64_t pv_lstm_workspace_bytes(int64_t max_windows) {
    return carve_ws(nullptr, 0, chunk_for(max_windows)).bytes;
}

namespace {
int lstm_run(PvLstmModel* m, const int16_t* windows, int64_t n, int32_t wrap_int8, float* probs, uint8_t* argmax, void* workspace,
             int64_t workspace_bytes, cudaStream_t st);
}

extern "C" int pv_lstm_infer(PvLstmModel* m, const int16_t* windows, int64_t n, int32_t wrap_int8, float* probs,
                             uint8_t* argmax, void* workspace, int64_t workspace_bytes, void* stream_) {
    if (!m || !windows || !probs || !workspace) return pv::set_error(PV_EINVAL, "null argument");
    if ((uintptr_t)windows & 3) return pv::set_error(PV_EINVAL, "pv_lstm_infer: windows must be 4-byte aligned (a window is 1716 bytes, so any whole-window offset into an aligned array is)");
    if (n <= 0) return PV_OK;
    cudaStream_t st = (cudaStream_t)stream_;
    static int use_graph = -1;
    if (use_graph < 0) { const char* v = getenv("PV_LSTM_GRAPH"); use_graph = (v && !atoi(v)) ? 0 : 1; }
    if (!use_graph || n > GRAPH_MAX_WINDOWS || pv::prof_enabled())
        return lstm_run(m, windows, n, wrap_int8, probs, argmax, workspace, workspace_bytes, st);
    // A small pass is 72 dependent launches of a few microseconds each. The second time a (workspace, n) pair is seen the
    // launches are captured -- reading and writing fixed staging slots of the workspace -- and from then on one graph launch
    // between two small device copies replaces them.
    const int64_t chunk = chunk_for(n);
    const Ws w = carve_ws(workspace, workspace_bytes, chunk);
    if (w.bytes > workspace_bytes) return lstm_run(m, windows, n, wrap_int8, probs, argmax, workspace, workspace_bytes, st);
    static std::mutex graph_mu;                                  // the cache belongs to the model: one caller at a time edits it
    std::lock_guard<std::mutex> lock(graph_mu);
    PvLstmModel::GraphEntry* e = nullptr;
    for (int i = 0; i < m->n_graphs; i++)
        if (m->graphs[i].ws == workspace && m->graphs[i].ws_bytes == workspace_bytes && m->graphs[i].n == n && m->graphs[i].wrap == wrap_int8) e = &m->graphs[i];
    if (!e) {
        if (m->n_graphs == 8) {                                  // full: forget the oldest
            if (m->graphs[0].exec) cudaGraphExecDestroy(m->graphs[0].exec);
            for (int i = 1; i < 8; i++) m->graphs[i - 1] = m->graphs[i];
            m->n_graphs = 7;
        }
        e = &m->graphs[m->n_graphs++];
        e->ws = workspace; e->ws_bytes = workspace_bytes; e->n = n; e->wrap = wrap_int8; e->uses = 0; e->exec = nullptr;
    }
    if (!e->exec && e->uses++ == 0)                              // first sight: plain launches (also sets the kernels' attributes)
        return lstm_run(m, windows, n, wrap_int8, probs, argmax, workspace, workspace_bytes, st);
    if (!e->exec) {
        if (!m->capture_stream) PV_CUDA_CHECK(cudaStreamCreateWithFlags(&m->capture_stream, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        PV_CUDA_CHECK(cudaStreamBeginCapture(m->capture_stream, cudaStreamCaptureModeThreadLocal));
        const int rc = lstm_run(m, w.win_stage, n, wrap_int8, w.probs_stage, w.arg_stage, workspace, workspace_bytes, m->capture_stream);
        const cudaError_t ce = cudaStreamEndCapture(m->capture_stream, &graph);
        if (rc != PV_OK || ce != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            e->uses = -1000000;                                  // do not try again for this pair
            return lstm_run(m, windows, n, wrap_int8, probs, argmax, workspace, workspace_bytes, st);
        }
        const cudaError_t ie = cudaGraphInstantiate(&e->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) {
            e->exec = nullptr; e->uses = -1000000; cudaGetLastError();
            return lstm_run(m, windows, n, wrap_int8, probs, argmax, workspace, workspace_bytes, st);
        }
    }
    PV_CUDA_CHECK(cudaMemcpyAsync(w.win_stage, windows, (size_t)n * T * F * 2, cudaMemcpyDeviceToDevice, st));
    PV_CUDA_CHECK(cudaGraphLaunch(e->exec, st));
    PV_CUDA_CHECK(cudaMemcpyAsync(probs, w.probs_stage, (size_t)n * 12, cudaMemcpyDeviceToDevice, st));
    if (argmax) PV_CUDA_CHECK(cudaMemcpyAsync(argmax, w.arg_stage, (size_t)n, cudaMemcpyDeviceToDevice, st));
    pv::count_launches(pv::FAM_LSTM_PREP, 1); pv::count_launches(pv::FAM_LSTM_ENC, T); pv::count_launches(pv::FAM_LSTM_DEC, T);
    pv::count_launches(pv::FAM_LSTM_MLP, 6);
    return PV_OK;
}

namespace {
int lstm_run(PvLstmModel* m, const int16_t* windows, int64_t n, int32_t wrap_int8, float* probs, uint8_t* argmax, void* workspace,
             int64_t workspace_bytes, cudaStream_t st) {
    // the largest chunk the workspace can hold
    int64_t chunk = chunk_for(n);
    while (chunk > 128 && carve_ws(nullptr, 0, chunk).bytes > workspace_bytes) chunk -= 128;
    const Ws w = carve_ws(workspace, workspace_bytes, chunk);
    if (w.bytes > workspace_bytes) return pv::set_error(PV_EINVAL, "workspace too small: need at least %lld bytes", (long long)w.bytes);

    CUtensorMap map_x, map_enc, map_dec, map_flat, map_act0, map_act1;
    if (int rc = make_map3(&map_x, w.xhl, chunk, T, XK, (int64_t)T * XK)) return rc;
    if (int rc = make_map3(&map_enc, w.enc_out, chunk, S, C, (int64_t)S * C)) return rc;
    if (int rc = make_map3(&map_dec, w.dec_out, chunk, S, C, (int64_t)S * C)) return rc;
    // linear_1 reads the decoder output of a window as one row of 33*512 channels starting at slot 1
    if (int rc = make_map3(&map_flat, w.dec_out + C, chunk, 1, FLAT, (int64_t)S * C)) return rc;
    if (int rc = make_map3(&map_act0, w.act0, chunk, 1, LIN, LIN)) return rc;
    if (int rc = make_map3(&map_act1, w.act1, chunk, 1, LIN, LIN)) return rc;

    for (int64_t off = 0; off < n; off += chunk) {
        const int64_t nb = n - off < chunk ? n - off : chunk;
        const int64_t elems = nb * T * (XK / 8);
        pv::prof_begin(pv::FAM_LSTM_PREP, st);
        prep_input_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, st>>>(windows + off * T * F, w.xhl, nb, wrap_int8);
        PV_CUDA_CHECK(cudaGetLastError());
        pv::prof_end(pv::FAM_LSTM_PREP, st, 1);

        tc::GemmShape g;
        memset(&g, 0, sizeof(g));
        g.M = (int)nb; g.m_blks = (int)((nb + 127) / 128); g.n_blks = 4; g.dirs = 2;
        g.w_row[0] = 0; g.w_row[1] = 4 * H;
        g.a0_col[0] = 0; g.a0_col[1] = H;
#ifdef PV_TRACE
        { static long long* tr = nullptr; if (!tr) { cudaMalloc((void**)&tr, 3 * 16 * 8 * 8); } g.trace = tr; pv_trace_buffer = tr; }
#endif
        for (int layer = 0; layer < 2; layer++) {
            LstmEpilogueT<true> e;
            e.bias = layer == 0 ? m->enc_b : m->dec_b;
            e.c_state = w.c_state;
            e.out = layer == 0 ? w.enc_out : w.dec_out;
            e.n_blks = 4;
            { static int dbg = -1; if (dbg < 0) { const char* v = getenv("PV_DEBUG_EPI"); dbg = v ? atoi(v) : 0; } e.debug = dbg; }
            g.kb1 = layer == 0 ? XK / tc::BLOCK_K : C / tc::BLOCK_K;
            pv::prof_begin(layer == 0 ? pv::FAM_LSTM_ENC : pv::FAM_LSTM_DEC, st);
            for (int s = 0; s < T; s++) {
                const int tf = s, tb = T - 1 - s;               // time handled by the forward / reverse direction
                e.first = s == 0;
                e.out_slot[0] = tf + 1; e.out_slot[1] = tb + 1;
                g.kb0 = s == 0 ? 0 : H / tc::BLOCK_K;           // h_{-1} == 0: skip the recurrent k blocks
                g.w_kb_off = s == 0 ? H / tc::BLOCK_K : 0;
                g.a0_slot[0] = tf; g.a0_slot[1] = tb + 2;       // slot of h_{t-1} (forward) / h_{t+1} (reverse)
                if (layer == 0) { g.a1_slot[0] = tf; g.a1_slot[1] = tb; g.a1_col[0] = g.a1_col[1] = 0; }
                else { g.a1_slot[0] = tf + 1; g.a1_slot[1] = tb + 1; g.a1_col[0] = g.a1_col[1] = 0; }
#ifdef PV_TRACE
                {   // trace only the launch PV_TRACE_STEP (0..32 encoder, 33..65 decoder) of each pass; default: the last one
                    static int target = -2; if (target == -2) { const char* v = getenv("PV_TRACE_STEP"); target = v ? atoi(v) : 65; }
                    g.trace = (layer * T + s == target) ? pv_trace_buffer : nullptr;
                }
#endif
                if (layer == 0) {
                    LstmEpilogueT<false> e0;                   // same fields, encoder configuration
                    e0.bias = e.bias; e0.c_state = e.c_state; e0.out = e.out; e0.n_blks = e.n_blks; e0.debug = e.debug;
                    e0.first = e.first; e0.out_slot[0] = e.out_slot[0]; e0.out_slot[1] = e.out_slot[1];
                    if (int rc = launch_gemm(map_enc, map_x, m->map_enc_w, g, e0, m->sms, st)) return rc;
                } else {
                    if (int rc = launch_gemm(map_dec, map_enc, m->map_dec_w, g, e, m->sms, st)) return rc;
                }
            }
            pv::prof_end(layer == 0 ? pv::FAM_LSTM_ENC : pv::FAM_LSTM_DEC, st, T);
        }
        // MLP
        pv::prof_begin(pv::FAM_LSTM_MLP, st);
        tc::GemmShape gl;
        memset(&gl, 0, sizeof(gl));
        gl.M = (int)nb; gl.m_blks = g.m_blks; gl.n_blks = LIN / tc::BLOCK_N; gl.dirs = 1;
        for (int l = 0; l < 5; l++) {
            SeluEpilogue e;
            e.bias = m->lin_b[l]; e.out = (l & 1) ? w.act1 : w.act0; e.ldo = LIN;
            gl.kb0 = 0; gl.kb1 = (l == 0 ? FLAT : LIN) / tc::BLOCK_K;
            const CUtensorMap& a = l == 0 ? map_flat : ((l & 1) ? map_act0 : map_act1);
            if (int rc = launch_gemm(a, a, m->map_lin_w[l], gl, e, m->sms, st)) return rc;
        }
        head_kernel<<<(unsigned)((nb * 32 + 255) / 256), 256, 0, st>>>(w.act0, m->out_w, m->out_b, probs + off * 3,
                                                                        argmax ? argmax + off : nullptr, nb);
        PV_CUDA_CHECK(cudaGetLastError());
        pv::prof_end(pv::FAM_LSTM_MLP, st, 6);
    }
    return PV_OK;
}
}  // namespace

extern "C" int pv_lstm_infer_host(PvLstmModel* m, const int16_t* windows_host, int64_t n, int32_t wrap_int8,
                                  float* probs_host, uint8_t* argmax_host) {
    if (!m || !windows_host || !probs_host) return pv::set_error(PV_EINVAL, "null argument");
    if (n <= 0) return PV_OK;
    if (int rc = pv::require_device()) return rc;
    if (!m->stream) PV_CUDA_CHECK(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    const int64_t need = pv_lstm_workspace_bytes(n);
    if (need > m->ws_bytes) {
        if (m->ws) cudaFree(m->ws);
        m->ws = nullptr; m->ws_bytes = 0;
        PV_CUDA_CHECK(cudaMalloc(&m->ws, (size_t)need));
        m->ws_bytes = need;
    }
    if (n > m->io_cap) {
        if (m->win_dev) cudaFree(m->win_dev);
        if (m->probs_dev) cudaFree(m->probs_dev);
        if (m->arg_dev) cudaFree(m->arg_dev);
        m->io_cap = 0;
        PV_CUDA_CHECK(cudaMalloc((void**)&m->win_dev, (size_t)n * T * F * 2));
        PV_CUDA_CHECK(cudaMalloc((void**)&m->probs_dev, (size_t)n * 3 * 4));
        PV_CUDA_CHECK(cudaMalloc((void**)&m->arg_dev, (size_t)n));
        m->io_cap = n;
    }
    PV_CUDA_CHECK(cudaMemcpyAsync(m->win_dev, windows_host, (size_t)n * T * F * 2, cudaMemcpyHostToDevice, m->stream));
    if (int rc = pv_lstm_infer(m, m->win_dev, n, wrap_int8, m->probs_dev, m->arg_dev, m->ws, m->ws_bytes, m->stream)) return rc;
    PV_CUDA_CHECK(cudaMemcpyAsync(probs_host, m->probs_dev, (size_t)n * 12, cudaMemcpyDeviceToHost, m->stream));
    if (argmax_host) PV_CUDA_CHECK(cudaMemcpyAsync(argmax_host, m->arg_dev, (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    PV_CUDA_CHECK(cudaStreamSynchronize(m->stream));
    return PV_OK;
}
