// Model M-B on sm_100a: the polisher's TransducerGRU.forward
// (/root/reference/pepper/modules/python/models/simple_model.py:27-42): biGRU(10->128, h0 = hidden) ->
// biGRU(256->128, h0 = encoder h_n) -> Linear(256->5), and the sliding-window loop around it
// (/root/reference/pepper/modules/python/models/predict_distributed_gpu.py:63-96).
//
// Same machinery as model M-A (tc_gemm.cuh): one tcgen05/TMA launch per time step advances both directions,
// A = [h_{t-1} | x_t]. PyTorch's GRU keeps the hidden part of the candidate gate separate
// (n = tanh(W_in x + b_in + r * (W_hn h + b_hn))), so a 256-column tile holds FOUR accumulators for 64 hidden units:
// r, z (h and x parts summed by the K-concatenation), n_h (x columns of W zeroed) and n_x (h columns zeroed).
// The hidden state used by the element-wise update stays fp32 ([n][2][128], the caller's `hidden` buffer itself);
// only the MMA operand copy of h is bf16.
#include "common.cuh"
#include "tc_gemm.cuh"
#include "infer_common.cuh"
#include <vector>

namespace {

constexpr int GF = 10;              // input features
constexpr int GH = 128;             // hidden size
constexpr int GC = 2 * GH;          // channels of a layer output
constexpr int GXK = 64;             // encoder input padded to one k block
constexpr int GENC_K = GH + GXK;    // 192
constexpr int GDEC_K = GH + GC;     // 384
constexpr int GCLS = 5;

struct GruEpilogue {
    static constexpr int kStages = 3;
    static constexpr int kBiasBytes = 2 * 2 * 4 * 64 * 4;                  // whole layer: [dir][n_blk][gate][64] fp32
    static constexpr int kStateBytes = tc::EPI_THREADS * 2 * 64;
    static constexpr int kSmemBytes = kBiasBytes + 2 * kStateBytes;
    static constexpr bool kInlinePrefetch = false;

    const float* bias;        // [dirs][n_blks][4][64]: b_ir+b_hr, b_iz+b_hz, b_hn, b_in
    float* h_state;           // [M][2][GH] fp32, read as h_{t-1}, overwritten with h_t
    __nv_bfloat16* out;       // [M][S][GC]
    int n_blks, S;
    int out_slot[2];

    __device__ void setup(uint8_t* scratch, int te) const {
        float4* sb = (float4*)scratch;
        const float4* gb = (const float4*)bias;
        for (int i = te; i < kBiasBytes / 16; i += tc::EPI_THREADS) sb[i] = __ldg(gb + i);
    }
    __device__ void prefetch(uint8_t* scratch, int buf, int dir, int n_blk, int row, bool ok, int half, int te) const {
        if (!ok) return;
        uint8_t* dst = scratch + kBiasBytes + buf * kStateBytes;
#pragma unroll
        for (int cc = 0; cc < 2; cc++) {
            const float* hp = h_state + ((size_t)row * 2 + dir) * GH + n_blk * 64 + (half * 2 + cc) * 16;
#pragma unroll
            for (int j = 0; j < 4; j++) tc::cp_async16(dst + ((cc * 4 + j) * tc::EPI_THREADS + te) * 16, hp + j * 4);
        }
    }
    __device__ void operator()(uint8_t* scratch, int buf, int dir, int n_blk, int row, bool ok, uint32_t taddr, int half, int te,
                               const tc::NextTile&) const {
        const float* sb = (const float*)scratch + (size_t)((dir * n_blks + n_blk) * 4) * 64;
        const uint8_t* hst = scratch + kBiasBytes + buf * kStateBytes;
#pragma unroll 1
        for (int cc = 0; cc < 2; cc++) {
            const int ch = half * 2 + cc;
            float ar[16], az[16], ah[16], ax[16], h[16];
            tc::tmem_ld16(taddr + 0 * 64 + ch * 16, ar);
            tc::tmem_ld16(taddr + 1 * 64 + ch * 16, az);
            tc::tmem_ld16(taddr + 2 * 64 + ch * 16, ah);
            tc::tmem_ld16(taddr + 3 * 64 + ch * 16, ax);
            if (ok) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float4 v = *(const float4*)(hst + ((cc * 4 + j) * tc::EPI_THREADS + te) * 16);
                    h[4 * j] = v.x; h[4 * j + 1] = v.y; h[4 * j + 2] = v.z; h[4 * j + 3] = v.w;
                }
            }
            tc::tmem_ld_wait();
            if (ok) {
                const float* br = sb + 0 * 64 + ch * 16; const float* bz = sb + 1 * 64 + ch * 16;
                const float* bh = sb + 2 * 64 + ch * 16; const float* bx = sb + 3 * 64 + ch * 16;
                uint32_t hp[8];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const float r = sigmoid_f(ar[i + e] + br[i + e]);
                        const float z = sigmoid_f(az[i + e] + bz[i + e]);
                        const float n = tanh_f(ax[i + e] + bx[i + e] + r * (ah[i + e] + bh[i + e]));
                        h[i + e] = (1.f - z) * n + z * h[i + e];
                    }
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(h[i], h[i + 1]);
                    hp[i >> 1] = *(const uint32_t*)&h2;
                }
                float* hg = h_state + ((size_t)row * 2 + dir) * GH + n_blk * 64 + ch * 16;
#pragma unroll
                for (int i = 0; i < 16; i += 4) *(float4*)(hg + i) = make_float4(h[i], h[i + 1], h[i + 2], h[i + 3]);
                __nv_bfloat16* op = out + ((size_t)row * S + (dir ? out_slot[1] : out_slot[0])) * GC + dir * GH + n_blk * 64 + ch * 16;
                *(uint4*)op = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                *(uint4*)(op + 8) = make_uint4(hp[4], hp[5], hp[6], hp[7]);
            }
        }
    }
};

// images uint8 [n][row_stride/10 ...]: x[r][t][f] = img[r*row_stride + (t0+t)*10 + f] -> bf16 [n][T][64]
__global__ void gru_prep_kernel(const uint8_t* __restrict__ img, int64_t row_stride, int t0, __nv_bfloat16* __restrict__ x,
                                int64_t n, int T) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * T * GXK) return;
    const int col = (int)(i % GXK);
    const int64_t rt = i / GXK;
    const int64_t r = rt / T; const int t = (int)(rt % T);
    // columns 0..9 meet the bf16 high part of W_ih, columns 10..19 (the same counts again) its bf16 remainder
    x[i] = __float2bfloat16_rn(col < 2 * GF ? (float)img[r * row_stride + (int64_t)(t0 + t) * GF + (col % GF)] : 0.f);
}

// bf16 operand copies of the initial hidden state: forward h0 -> slot 0, reverse h0 -> slot T+1
__global__ void gru_init_slots_kernel(const float* __restrict__ h_state, __nv_bfloat16* __restrict__ out, int64_t n, int S) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // over n*2*GH
    if (i >= n * 2 * GH) return;
    const int j = (int)(i % GH); const int dir = (int)((i / GH) % 2); const int64_t r = i / (2 * GH);
    out[((size_t)r * S + (dir ? S - 1 : 0)) * GC + dir * GH + j] = __float2bfloat16_rn(h_state[i]);
}

// dense1 (256 -> 5) per position; one warp per (window, t). mode 0: write logits; mode 1: add softmax into acc
__global__ void gru_head_kernel(const __nv_bfloat16* __restrict__ dec_out, int S, const float* __restrict__ w,
                                const float* __restrict__ b, float* __restrict__ dst, int64_t dst_row_stride, int t0,
                                int64_t n, int T, int mode) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= n * T) return;
    const int64_t r = wid / T; const int t = (int)(wid % T);
    const __nv_bfloat16* x = dec_out + ((size_t)r * S + t + 1) * GC;
    float s[GCLS] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = lane; k < GC; k += 32) {
        const float v = __bfloat162float(x[k]);
#pragma unroll
        for (int c = 0; c < GCLS; c++) s[c] += v * w[c * GC + k];
    }
#pragma unroll
    for (int c = 0; c < GCLS; c++)
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) s[c] += __shfl_xor_sync(0xffffffffu, s[c], d);
    if (lane == 0) {
        float* o = dst + r * dst_row_stride + (int64_t)(t0 + t) * GCLS;
        float m = -1e30f;
#pragma unroll
        for (int c = 0; c < GCLS; c++) { s[c] += b[c]; m = fmaxf(m, s[c]); }
        if (mode == 0) {
#pragma unroll
            for (int c = 0; c < GCLS; c++) o[c] = s[c];
        } else {
            float e[GCLS], sum = 0.f;
#pragma unroll
            for (int c = 0; c < GCLS; c++) { e[c] = expf(s[c] - m); sum += e[c]; }
            const float inv = 1.f / sum;
#pragma unroll
            for (int c = 0; c < GCLS; c++) o[c] += e[c] * inv;          // windows are processed one after another
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

struct GruWs {
    __nv_bfloat16 *xin, *enc_out, *dec_out;
    float* h_state;
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
    w.bytes = pv::align_up(a.cur, 256);
    return w;
}

int64_t gru_chunk_for(int64_t n) {
    int64_t c = (n + 127) / 128 * 128;
    if (c < 128) c = 128;
    return c < MAX_CHUNK ? c : MAX_CHUNK;
}

// rows in tile order (dir, n_blk, gate{r,z,nh,nx}, j) <- PyTorch rows r: j, z: GH + j, n: 2*GH + j
// hi_lo: the input weights are stored as bf16 high part + bf16 remainder in two column groups (the raw 0..254 counts
// are exact in bf16, so x * W_ih keeps ~16 mantissa bits instead of 8)
void pack_gru(const float* const w_ih[2], const float* const w_hh[2], const float* const b_ih[2], const float* const b_hh[2],
              int in_dim, int k_total, bool hi_lo, std::vector<uint16_t>& w, std::vector<float>& b) {
    w.assign((size_t)2 * 4 * GH * k_total, 0);
    b.assign((size_t)2 * 4 * GH, 0.f);
    for (int dir = 0; dir < 2; dir++)
        for (int nb = 0; nb < 2; nb++)
            for (int gate = 0; gate < 4; gate++)
                for (int j = 0; j < 64; j++) {
                    const int hj = nb * 64 + j;
                    const int src = (gate < 2 ? gate : 2) * GH + hj;
                    const size_t dst = ((size_t)(dir * 2 + nb) * 4 + gate) * 64 + j;
                    uint16_t* wr = &w[dst * k_total];
                    if (gate != 3) for (int k = 0; k < GH; k++) wr[k] = f2bf(w_hh[dir][(size_t)src * GH + k]);
                    if (gate != 2) for (int k = 0; k < in_dim; k++) {
                        const float v = w_ih[dir][(size_t)src * in_dim + k];
                        wr[GH + k] = f2bf(v);
                        if (hi_lo) wr[GH + in_dim + k] = f2bf(v - bf2f(wr[GH + k]));
                    }
                    b[dst] = gate < 2 ? b_ih[dir][src] + b_hh[dir][src] : (gate == 2 ? b_hh[dir][src] : b_ih[dir][src]);
                }
}

}  // namespace

struct PvGruModel {
    __nv_bfloat16 *enc_w, *dec_w;
    float *enc_b, *dec_b, *dense_w, *dense_b;
    CUtensorMap map_enc_w, map_dec_w;
    int sms;
};

extern "C" int pv_gru_create(const PvGruWeights* hw, PvGruModel** out) {
    if (!hw || !out) return pv::set_error(PV_EINVAL, "null argument");
    if (int rc = pv::require_device()) return rc;
    PvGruModel* m = new PvGruModel();
    memset(m, 0, sizeof(*m));
    m->sms = pv::sm_count();
    std::vector<uint16_t> w; std::vector<float> b;
    pack_gru(hw->enc_w_ih, hw->enc_w_hh, hw->enc_b_ih, hw->enc_b_hh, GF, GENC_K, true, w, b);
    if (int rc = upload(&m->enc_w, w.data(), w.size() * 2)) return rc;
    if (int rc = upload(&m->enc_b, b.data(), b.size() * 4)) return rc;
    pack_gru(hw->dec_w_ih, hw->dec_w_hh, hw->dec_b_ih, hw->dec_b_hh, GC, GDEC_K, false, w, b);
    if (int rc = upload(&m->dec_w, w.data(), w.size() * 2)) return rc;
    if (int rc = upload(&m->dec_b, b.data(), b.size() * 4)) return rc;
    if (int rc = upload(&m->dense_w, hw->dense_w, GCLS * GC * 4)) return rc;
    if (int rc = upload(&m->dense_b, hw->dense_b, GCLS * 4)) return rc;
    if (int rc = make_map2(&m->map_enc_w, m->enc_w, 2 * 4 * GH, GENC_K)) return rc;
    if (int rc = make_map2(&m->map_dec_w, m->dec_w, 2 * 4 * GH, GDEC_K)) return rc;
    *out = m;
    return PV_OK;
}

extern "C" void pv_gru_destroy(PvGruModel* m) {
    if (!m) return;
    cudaFree(m->enc_w); cudaFree(m->dec_w); cudaFree(m->enc_b); cudaFree(m->dec_b); cudaFree(m->dense_w); cudaFree(m->dense_b);
    delete m;
}

extern "C" int64_t pv_gru_workspace_bytes(int64_t max_batch, int32_t seq_len) {
    return gru_carve(nullptr, 0, gru_chunk_for(max_batch), seq_len).bytes;
}

namespace {

// one TransducerGRU.forward over rows [0, nb) of a chunk; h_state [nb][2][GH] in/out
int gru_forward_chunk(PvGruModel* m, const GruWs& w, const CUtensorMap& map_x, const CUtensorMap& map_enc,
                      const CUtensorMap& map_dec, const uint8_t* images, int64_t img_row_stride, int t0, int64_t nb, int T,
                      float* h_state, cudaStream_t st) {
    const int S = T + 2;
    const int64_t elems = nb * T * GXK;
    pv::prof_begin(pv::FAM_GRU_MISC, st);
    gru_prep_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, st>>>(images, img_row_stride, t0, w.xin, nb, T);
    PV_CUDA_CHECK(cudaGetLastError());
    pv::prof_end(pv::FAM_GRU_MISC, st, 1);
    tc::GemmShape g;
    memset(&g, 0, sizeof(g));
    g.M = (int)nb; g.m_blks = (int)((nb + 127) / 128); g.n_blks = 2; g.dirs = 2;
    g.w_row[0] = 0; g.w_row[1] = 4 * GH;
    g.a0_col[0] = 0; g.a0_col[1] = GH;
    g.kb0 = GH / tc::BLOCK_K; g.w_kb_off = 0;
    for (int layer = 0; layer < 2; layer++) {
        __nv_bfloat16* out = layer == 0 ? w.enc_out : w.dec_out;
        gru_init_slots_kernel<<<(unsigned)((nb * 2 * GH + 255) / 256), 256, 0, st>>>(h_state, out, nb, S);
        PV_CUDA_CHECK(cudaGetLastError());
        GruEpilogue e;
        e.bias = layer == 0 ? m->enc_b : m->dec_b; e.h_state = h_state; e.out = out; e.n_blks = 2; e.S = S;
        g.kb1 = layer == 0 ? 1 : GC / tc::BLOCK_K;
        pv::prof_begin(pv::FAM_GRU_STEP, st);
        for (int s = 0; s < T; s++) {
            const int tf = s, tb = T - 1 - s;
            e.out_slot[0] = tf + 1; e.out_slot[1] = tb + 1;
            g.a0_slot[0] = tf; g.a0_slot[1] = tb + 2;
            if (layer == 0) { g.a1_slot[0] = tf; g.a1_slot[1] = tb; }
            else { g.a1_slot[0] = tf + 1; g.a1_slot[1] = tb + 1; }
            g.a1_col[0] = g.a1_col[1] = 0;
            if (int rc = launch_gemm(layer == 0 ? map_enc : map_dec, layer == 0 ? map_x : map_enc,
                                     layer == 0 ? m->map_enc_w : m->map_dec_w, g, e, m->sms, st)) return rc;
        }
        pv::prof_end(pv::FAM_GRU_STEP, st, T + 1);
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
        if (int rc = gru_forward_chunk(m, w, map_x, map_enc, map_dec, images + off * T * GF, (int64_t)T * GF, 0, nb, T,
                                       hidden + off * 2 * GH, st)) return rc;
        gru_head_kernel<<<(unsigned)((nb * T * 32 + 255) / 256), 256, 0, st>>>(w.dec_out, S, m->dense_w, m->dense_b,
                                                                                logits + off * T * GCLS, (int64_t)T * GCLS, 0, nb, T, 0);
        PV_CUDA_CHECK(cudaGetLastError());
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
            if (int rc = gru_forward_chunk(m, w, map_x, map_enc, map_dec, images + off * L * GF, (int64_t)L * GF, t0, nb, T,
                                           w.h_state, st)) return rc;
            gru_head_kernel<<<(unsigned)((nb * T * 32 + 255) / 256), 256, 0, st>>>(w.dec_out, S, m->dense_w, m->dense_b,
                                                                                    prob_sum + off * L * GCLS, (int64_t)L * GCLS, t0, nb, T, 1);
            PV_CUDA_CHECK(cudaGetLastError());
        }
    }
    const int64_t n_pos = n * L;
    gru_argmax_kernel<<<(unsigned)((n_pos + 255) / 256), 256, 0, st>>>(prob_sum, labels, n_pos);
    PV_CUDA_CHECK(cudaGetLastError());
    return PV_OK;
}
