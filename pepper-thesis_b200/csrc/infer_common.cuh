// Helpers shared by the two recurrent models (tensor maps, bf16 packing, activations, GEMM launch).
#pragma once
#include "common.cuh"
#include "tc_gemm.cuh"
#include <cstdlib>
#include <vector>

namespace {

constexpr int MAX_CHUNK = 32768;    // windows per pass: many tiles per CTA per launch amortise pipeline fill/drain

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// bf16 tensor [rows][slots][chan] (chan contiguous), box = 64 chan x 1 slot x 128 rows, 128-byte swizzle
int make_map3(CUtensorMap* m, const void* base, int64_t rows, int64_t slots, int64_t chan, int64_t row_stride_elems) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return pv::set_error(PV_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dim[3] = {(cuuint64_t)chan, (cuuint64_t)slots, (cuuint64_t)rows};
    cuuint64_t stride[2] = {(cuuint64_t)chan * 2, (cuuint64_t)row_stride_elems * 2};
    cuuint32_t box[3] = {(cuuint32_t)tc::BLOCK_K, 1, (cuuint32_t)tc::BLOCK_M};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)base, dim, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return pv::set_error(PV_ECUDA, "cuTensorMapEncodeTiled(3d) failed: %d", (int)r);
    return PV_OK;
}
// bf16 weight matrix [rows][k] (k contiguous), box = 64 k x (256 / CLUSTER) rows: one CTA's multicast slice of a W tile
int make_map2(CUtensorMap* m, const void* base, int64_t rows, int64_t k) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return pv::set_error(PV_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dim[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    cuuint64_t stride[1] = {(cuuint64_t)k * 2};
    cuuint32_t box[2] = {(cuuint32_t)tc::BLOCK_K, (cuuint32_t)tc::W_SLICE_ROWS};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dim, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return pv::set_error(PV_ECUDA, "cuTensorMapEncodeTiled(2d) failed: %d", (int)r);
    return PV_OK;
}

uint16_t f2bf(float f) {                 // round-to-nearest-even float -> bf16 bits
    uint32_t u; memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

float bf2f(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }

// Gate non-linearities on the MUFU tanh unit (tanh.approx.f32, max relative error 2^-11 -- an order of magnitude
// below the bf16 rounding of the MMA operands): 1 MUFU op per activation instead of ex2 + rcp.
// -DPV_EXACT_ACT switches to exp/divide based forms.
#ifdef PV_EXACT_ACT
__device__ __forceinline__ float ex2_f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_f(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_f(float x) { return rcp_f(1.f + ex2_f(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_f(float x) { return fmaf(2.f, rcp_f(1.f + ex2_f(-2.8853900817779268f * x)), -1.f); }
#else
__device__ __forceinline__ float tanh_f(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_f(float x) { return fmaf(0.5f, tanh_f(0.5f * x), 0.5f); }
#endif
__device__ __forceinline__ void ld16(const float* p, float* v) {     // 64 contiguous, 16-byte aligned bytes
#pragma unroll
    for (int i = 0; i < 16; i += 4) { const float4 t = __ldg((const float4*)(p + i)); v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w; }
}

template <class Epi>
int launch_gemm(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const tc::GemmShape& g, const Epi& epi,
                int sms, cudaStream_t st) {
    static bool attr_done = false;
    constexpr int smem = tc::smem_bytes<Epi>();
    if (!attr_done) {
        PV_CUDA_CHECK(cudaFuncSetAttribute(tc::gemm_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_done = true;
    }
    const int groups = ((g.m_blks + tc::CLUSTER - 1) / tc::CLUSTER) * g.n_blks * (tc::slot_tiles<Epi>::value ? g.slots : g.dirs);
    int clusters = sms / tc::CLUSTER;
    if (clusters > groups) clusters = groups;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(clusters * tc::CLUSTER));
    cfg.blockDim = dim3(tc::threads_of<Epi>());
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = tc::CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    // programmatic dependent launch (tc_gemm.cuh: launch_dependents / grid_dependency_wait); PV_NO_PDL=1 switches it off
    static int pdl = -1;
    if (pdl < 0) { const char* v = getenv("PV_NO_PDL"); pdl = (v && atoi(v)) ? 0 : 1; }
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 2 : 1;
    PV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, tc::gemm_kernel<Epi>, a0, a1, w, g, epi));
    return PV_OK;
}

template <typename Tv>
int upload(Tv** dst, const void* src, size_t bytes) {
    PV_CUDA_CHECK(cudaMalloc((void**)dst, bytes));
    PV_CUDA_CHECK(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
    return PV_OK;
}


}  // namespace
