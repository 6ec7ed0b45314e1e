// Hand-written tcgen05 + TMA GEMM core for sm_100a with pluggable fused epilogues.
//
//   D[M, N] = A[M, K] * W[N, K]^T        bf16 operands (K-major, 128-byte swizzle), fp32 accumulators in TMEM
//
// One persistent CTA per SM, 256 threads, warp-specialised:
//   warp 0   TMA producer   cp.async.bulk.tensor (UTMALDG) A/W tiles -> 4-stage shared-memory ring (full/empty mbarriers)
//   warp 1   MMA issuer     one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (UTCHMMA), 128x256x16 per
//                           instruction, accumulating into one of two 256-column TMEM stages; tcgen05.commit frees
//                           the smem stage / publishes the accumulator
//   warp 2   TMEM allocator (512 columns)
//   warps 4-11 epilogue     tcgen05.ld (LDTM) 32 lanes x 16 columns per call -> registers -> fused epilogue -> HBM,
//                           overlapping the next tile's MMAs through the second TMEM stage; two warps per TMEM lane
//                           quarter (warp % 4), each taking half of the tile's columns
//
// The A operand is the K-concatenation of up to two tensors (A0 then A1), each addressed through its own 3-D
// tensor map [rows][slot][channels]: this is how a recurrent step computes  [h_{t-1} | x_t] * [W_hh | W_ih]^T
// in ONE pass (input projection fused with the recurrent product, nothing materialised in HBM).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;                       // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KB
constexpr int W_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;   // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + W_STAGE_BYTES;
constexpr int EPI_THREADS = 256;
// Thread-block cluster along M: the CLUSTER CTAs of a cluster work on CLUSTER consecutive row tiles of the SAME
// (direction, column tile), so they need the same W tile: each CTA TMA-loads 1/CLUSTER of it and MULTICASTS the slice
// into the shared memory of all of them. L2->SM operand traffic per CTA drops from 48 KB to 16 + 32/CLUSTER KB per
// k block (the kernel is bound by exactly that traffic).
constexpr int CLUSTER = 2;
constexpr int W_SLICE_ROWS = BLOCK_N / CLUSTER;
// PV_TWO_CTA = 1: the two CTAs of a cluster issue ONE tcgen05.mma.cta_group::2 of M = 256 per k step (the pair's two row tiles
// stacked): every CTA keeps only ITS half of the W tile (128 of the 256 rows: the tensor core reads the other half from the
// peer's shared memory), so a k block brings 16 + 16 KB into an SM instead of 16 + 32 KB. The leader (cluster rank 0) issues
// the MMAs for both; both CTAs run their own TMA producer and their own epilogue on their own 128 accumulator rows.
// Measured (round 2, 32768 windows, `tools/probe_lstm.py`, both builds in one box): decoder steps 3.08 vs 3.10 ms, MLP 0.64 vs
// 0.66 ms, encoder steps 2.16 vs 1.97 ms -- halving the W bytes that enter an SM buys nothing, so the multicast form stays the
// default; build with -DPV_TWO_CTA=1 (PV_NVCC_FLAGS) for the pair form (all model parity tests pass with it).
#ifndef PV_TWO_CTA
#define PV_TWO_CTA 0
#endif
constexpr bool TWO_CTA = PV_TWO_CTA != 0;
constexpr int W_BYTES_PER_CTA = TWO_CTA ? W_STAGE_BYTES / CLUSTER : W_STAGE_BYTES;   // W bytes that land in one CTA per k block
// dynamic shared memory of gemm_kernel<Epilogue>: operand ring + alignment slack + barriers + epilogue scratch
template <class Epilogue> constexpr int smem_bytes() {
    return Epilogue::kStages * STAGE_BYTES + 1024 + 256 + Epilogue::kSmemBytes;
}
constexpr int THREADS = 384;                      // 4 control warps + 8 epilogue warps
constexpr int TMEM_COLS = 512;

// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// multicast variant: the box lands at the same shared-memory offset in every CTA of `mask`, and completes on the
// mbarrier at the same offset in each of them
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(smem_u32(smem_dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
// 2-CTA forms: the data lands in the issuing CTA's shared memory, the bytes complete on the LEADER's mbarrier (same offset,
// peer bit of the shared::cluster address cleared: cute::Sm100MmaPeerBitMask)
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"((uint64_t)map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"((uint64_t)map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
// arrive on the barrier at this offset in the cluster's CTA `rank`
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Programmatic dependent launch: a recurrent step is ~50 us long and is launched 66 times per pass, each launch depending on
// the one before. launch_dependents lets the NEXT launch's CTAs take over an SM the moment this launch's CTA has left it (its
// barrier / TMEM / cluster set-up and the launch latency then hide behind this grid's tail); grid_dependency_wait blocks
// until the PREVIOUS grid has completed and its writes are visible -- every thread that touches global memory calls it first.
// Both are no-ops for a launch without cudaLaunchAttributeProgrammaticStreamSerialization.
__device__ __forceinline__ void launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// the pair's forms: the same warp of BOTH CTAs executes them
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// M = 256 over the pair: issued by the leader only; A rows / W rows of the second half come from the peer's shared memory
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit_mc_2sm(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives (when the MMAs issued so far retire) on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp receives row (lane_base + i), columns col..col+15
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
// the same for 8 columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// the reverse: thread i of the warp writes 16 consecutive fp32 columns of row (lane_base + i)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                   "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                   "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                   "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// shared memory -> global tile store through a tensor map (bulk group of the issuing thread)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"((uint64_t)map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// explicit shared-space 16-byte accesses (a computed generic pointer makes nvcc emit generic LD/ST)
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// contiguous bytes global -> shared (bulk asynchronous copy, completes on an mbarrier); size a multiple of 16
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// 16-byte asynchronous global -> shared copy (LDGSTS) and its group fences; used to prefetch per-tile epilogue state
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
template <int THREADS_> __device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(THREADS_) : "memory"); }   // the epilogue warps only

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, rows of 128 bytes, 8-row groups 1024 bytes apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout=2 [61,64))
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor, kind::f16: D=f32 (1<<4), A=bf16 (1<<7), B=bf16 (1<<10), both K-major, N>>3 at 17, M>>4 at 24
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---- problem description -----------------------------------------------------------------------------------
struct GemmShape {
    int M;            // rows (windows)
    int m_blks;       // ceil(M / 128)
    int n_blks;       // N / 256 per direction
    int dirs;         // 1 or 2 independent directions (own W rows, own slots)
    int kb0, kb1;     // 64-wide k blocks taken from A0 then from A1
    int a0_col[2];    // first channel of A0 per direction
    int a0_slot[2];   // slot (time index) of A0 per direction
    int a1_col[2];
    int a1_slot[2];
    int w_row[2];     // first W row per direction
    int w_kb_off;     // first W k block (non-zero when the h block is skipped because h == 0)
    int slots;        // slot-tile epilogues only (Epilogue::kSlotTiles): a tile is (row group, slot, column tile) with
                      // dirs == 1; A1 is read at slot a1_slot[0] + slot, and the epilogue receives the slot as `dir`
    long long* trace; // optional (debug builds with -DPV_TRACE): CTA 0 records [role][tile][event] SM-clock stamps here
};

#ifdef PV_TRACE
#define PV_TR(role, it, ev) do { if (g.trace && blockIdx.x == 0 && (it) < 16) g.trace[((role) * 16 + (it)) * 8 + (ev)] = clock64(); } while (0)
#else
#define PV_TR(role, it, ev) do { } while (0)
#endif

// the tile an epilogue thread will see next (for epilogues that prefetch their per-row state themselves)
struct NextTile { bool valid; int dir, n_blk, row; bool ok; };

// Epilogue functor interface (called by the 256 epilogue threads, two per accumulator row, half = 0/1, te = 0..255):
//   static constexpr int kStages, kSmemBytes            operand ring depth, bytes of epilogue scratch in shared memory
//   void setup(uint8_t* scratch, int te)                once per kernel (e.g. biases -> shared memory)
//   void prefetch(uint8_t* scratch, int buf, int dir, int n_blk, int row, bool ok, int half, int te)
//                                                       cp.async the NEXT tile's per-row state into buffer `buf`
//   void operator()(uint8_t* scratch, int buf, int dir, int n_blk, int row, bool ok, uint32_t taddr, int half, int te,
//                   const NextTile& nx)                 reads its half of the 256 accumulator columns (tmem_ld16)
//   static constexpr bool kInlinePrefetch               true: one state buffer; the functor itself issues the next tile's
//                                                       cp.async from inside operator() as soon as it has consumed its slots

// epilogue warps of gemm_kernel<E>: 8 (two per TMEM lane quarter, each taking half of the tile's columns) unless the functor
// asks for 16 (E::kEpiWarps: four per lane quarter, a quarter of the columns each -- for epilogue-bound tiles)
template <class E, class = void> struct epi_warps { static constexpr int value = 8; };
template <class E> struct epi_warps<E, decltype((void)E::kEpiWarps)> { static constexpr int value = E::kEpiWarps; };
template <class E> constexpr int threads_of() { return 128 + 32 * epi_warps<E>::value; }

template <class E, class = void> struct slot_tiles { static constexpr bool value = false; };
template <class E> struct slot_tiles<E, decltype((void)E::kSlotTiles)> { static constexpr bool value = E::kSlotTiles; };

struct TileIdx { int dir, n_blk, m_grp, slot; };
// step tiles: dirs and n_blks are powers of two (1/2 and 2/4): tile -> (dir, n_blk, m group) with shifts, no integer
// division. Slot tiles (few, large tiles): n_blk fastest, then the slot, then the row group.
template <bool SLOT>
__device__ __forceinline__ TileIdx decode_tile(int tile, const GemmShape& g, int dir_sh, int nb_sh, int dir_mask, int nb_mask) {
    TileIdx t;
    if constexpr (SLOT) {
        t.dir = 0;
        t.n_blk = tile % g.n_blks;
        const int r = tile / g.n_blks;
        t.slot = r % g.slots;
        t.m_grp = r / g.slots;
    } else {
        t.dir = tile & dir_mask;
        t.n_blk = (tile >> dir_sh) & nb_mask;
        t.m_grp = tile >> (dir_sh + nb_sh);
        t.slot = 0;
    }
    return t;
}

template <class Epilogue>
__global__ void __launch_bounds__(threads_of<Epilogue>(), 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
            const __grid_constant__ CUtensorMap tmW /* box = 64 k x W_SLICE_ROWS rows */, const GemmShape g, const Epilogue epi) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    constexpr int STAGES = Epilogue::kStages;
    uint8_t* smem_w = smem + STAGES * A_STAGE_BYTES;
    uint64_t* bars = (uint64_t*)(smem + STAGES * STAGE_BYTES);
    uint8_t* epi_scratch = smem + STAGES * STAGE_BYTES + 256;
    uint64_t* full_bar = bars;                 // [STAGES]
    uint64_t* empty_bar = bars + STAGES;       // [STAGES]
    uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // work item = a GROUP of CLUSTER consecutive row tiles with the same (column tile, direction); the cluster's CTA r
    // takes row tile  m_grp * CLUSTER + r  (possibly beyond M at the ragged end: loads are zero-filled, stores skipped)
    const int crank = (int)cluster_ctarank();
    const int cluster_id = (int)blockIdx.x / CLUSTER, n_clusters = (int)gridDim.x / CLUSTER;
    const int m_grps = (g.m_blks + CLUSTER - 1) / CLUSTER;
    constexpr bool SLOT = slot_tiles<Epilogue>::value;
    const int n_tiles = m_grps * g.n_blks * (SLOT ? g.slots : g.dirs);   // groups
    const int dir_sh = g.dirs >> 1, nb_sh = 31 - __clz(g.n_blks), dir_mask = g.dirs - 1, nb_mask = g.n_blks - 1;
    const int kbt = g.kb0 + g.kb1;
    constexpr uint16_t kMask = (uint16_t)((1u << CLUSTER) - 1u);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA0); tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmW);
    }
    if (warp == 1 && lane == 0) {
        // a stage may be refilled only when ALL CTAs of the cluster have consumed it (peers multicast into it)
        // TWO_CTA: one commit (the leader's, multicast) frees a stage in each CTA; the leader's accumulator-free barrier
        // collects the epilogue warps of BOTH CTAs
        for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], TWO_CTA ? 1 : CLUSTER); }
        for (int s = 0; s < 2; s++) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], (TWO_CTA ? CLUSTER : 1) * epi_warps<Epilogue>::value); }
        fence_barrier_init();
    }
    if constexpr (TWO_CTA) {
        __syncthreads();
        cluster_sync_all();                                    // both CTAs are resident before the pair allocates
        if (warp == 2) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    } else {
        if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                        // every CTA's barriers are initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    launch_dependents();

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            grid_dependency_wait();                            // A (and the state behind it) is the previous launch's output
            int stage = 0; uint32_t phase = 0; int itp = 0;
            for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, itp++) {
                const TileIdx ti = decode_tile<SLOT>(tile, g, dir_sh, nb_sh, dir_mask, nb_mask);
                const int dir = ti.dir, n_blk = ti.n_blk;
                const int m_blk = ti.m_grp * CLUSTER + crank;
                PV_TR(0, itp, 0);
                for (int kb = 0; kb < kbt; kb++) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    const bool from_a0 = kb < g.kb0;
                    const CUtensorMap* tmA = from_a0 ? &tmA0 : &tmA1;
                    const int a_col = from_a0 ? (dir ? g.a0_col[1] : g.a0_col[0]) + kb * BLOCK_K : (dir ? g.a1_col[1] : g.a1_col[0]) + (kb - g.kb0) * BLOCK_K;
                    const int a_slot = from_a0 ? (dir ? g.a0_slot[1] : g.a0_slot[0]) : (dir ? g.a1_slot[1] : g.a1_slot[0]) + ti.slot;
                    // W columns: A0's k blocks first, then A1's (the host packs [W_hh | W_ih] that way)
                    const int w_col = (g.w_kb_off + kb) * BLOCK_K, w_row = (dir ? g.w_row[1] : g.w_row[0]) + n_blk * BLOCK_N + crank * W_SLICE_ROWS;
                    if constexpr (TWO_CTA) {
                        // both CTAs' bytes complete on the leader's barrier; this CTA's half of the W tile stays here
                        if (crank == 0) mbar_expect_tx(&full_bar[stage], CLUSTER * (A_STAGE_BYTES + W_BYTES_PER_CTA));
                        tma_load_3d_2sm(smem_a + stage * A_STAGE_BYTES, tmA, &full_bar[stage], a_col, a_slot, m_blk * BLOCK_M);
                        tma_load_2d_2sm(smem_w + stage * W_STAGE_BYTES, &tmW, &full_bar[stage], w_col, w_row);
                    } else {
                        mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                        tma_load_3d(smem_a + stage * A_STAGE_BYTES, tmA, &full_bar[stage], a_col, a_slot, m_blk * BLOCK_M);
                        // this CTA's slice of the W tile, multicast to every CTA of the cluster
                        tma_load_2d_mc(smem_w + stage * W_STAGE_BYTES + crank * W_SLICE_ROWS * BLOCK_K * 2, &tmW, &full_bar[stage], w_col, w_row, kMask);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                PV_TR(0, itp, 1);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0 && (!TWO_CTA || crank == 0)) {
            constexpr uint32_t idesc = make_idesc(TWO_CTA ? CLUSTER * BLOCK_M : BLOCK_M, BLOCK_N);
            int stage = 0; uint32_t phase = 0; int it = 0;
            for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, it++) {
                const int acc = it & 1;
                const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
                PV_TR(1, it, 0);
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                PV_TR(1, it, 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int kb = 0; kb < kbt; kb++) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = make_smem_desc(smem_u32(smem_a + stage * A_STAGE_BYTES));
                    const uint64_t bdesc = make_smem_desc(smem_u32(smem_w + stage * W_STAGE_BYTES));
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; k++) {
                        if constexpr (TWO_CTA)
                            umma_bf16_2sm(d_tmem, adesc + (uint64_t)(k * UMMA_K * 2 / 16), bdesc + (uint64_t)(k * UMMA_K * 2 / 16),
                                          idesc, (kb | k) != 0 ? 1u : 0u);
                        else
                            umma_bf16(d_tmem, adesc + (uint64_t)(k * UMMA_K * 2 / 16), bdesc + (uint64_t)(k * UMMA_K * 2 / 16),
                                      idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    // tells every CTA of the cluster this consumer is done with the stage
                    if constexpr (TWO_CTA) umma_commit_mc_2sm(&empty_bar[stage], kMask); else umma_commit_mc(&empty_bar[stage], kMask);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                // accumulator complete (TWO_CTA: in both CTAs' tensor memory, announced to both epilogues)
                if constexpr (TWO_CTA) umma_commit_mc_2sm(&tfull_bar[acc], kMask); else umma_commit(&tfull_bar[acc]);
                PV_TR(1, it, 2);
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue =====
        const int q = warp & 3;                                // TMEM lane quarter this warp may access
        const int half = (warp - 4) >> 2;                       // which part of the columns: 0..1 (8 warps) or 0..3 (16 warps)
        const int te = threadIdx.x - 128;
        epi.setup(epi_scratch, te);                            // constants only (biases): may run ahead of the previous grid's end
        grid_dependency_wait();
        epi_barrier<32 * epi_warps<Epilogue>::value>();
        int it = 0;
        if (cluster_id < n_tiles) {                            // state of the first tile
            const TileIdx t0 = decode_tile<SLOT>(cluster_id, g, dir_sh, nb_sh, dir_mask, nb_mask);
            const int row = (t0.m_grp * CLUSTER + crank) * BLOCK_M + q * 32 + lane;
            epi.prefetch(epi_scratch, 0, SLOT ? t0.slot : t0.dir, t0.n_blk, row, row < g.M, half, te);
        }
        cp_async_commit();
        for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, it++) {
            const TileIdx ti = decode_tile<SLOT>(tile, g, dir_sh, nb_sh, dir_mask, nb_mask);
            const int dir = SLOT ? ti.slot : ti.dir, n_blk = ti.n_blk;
            const int m_blk = ti.m_grp * CLUSTER + crank;
            const int acc = it & 1;
            const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
            const int nxt = tile + n_clusters;
            NextTile nx;
            nx.valid = nxt < n_tiles;
            const TileIdx tn = decode_tile<SLOT>(nxt, g, dir_sh, nb_sh, dir_mask, nb_mask);
            nx.dir = SLOT ? tn.slot : tn.dir; nx.n_blk = tn.n_blk;
            nx.row = (tn.m_grp * CLUSTER + crank) * BLOCK_M + q * 32 + lane;
            nx.ok = nx.row < g.M;
            if constexpr (!Epilogue::kInlinePrefetch) {
                if (nx.valid)                                  // next tile's state: lands while this tile is computed
                    epi.prefetch(epi_scratch, acc ^ 1, nx.dir, nx.n_blk, nx.row, nx.ok, half, te);
                cp_async_commit();
            }
            if (te == 0) PV_TR(2, it, 0);
            if (lane == 0) mbar_wait(&tfull_bar[acc], acc_phase);   // one poller per warp
            __syncwarp();
            if (te == 0) PV_TR(2, it, 1);
            tc_fence_after();
            if constexpr (Epilogue::kInlinePrefetch) cp_async_wait<0>();   // this tile's state was requested a tile ago
            else cp_async_wait<1>();                           // everything but the group just committed has landed
            if (te == 0) PV_TR(2, it, 2);
            const int row = m_blk * BLOCK_M + q * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            epi(epi_scratch, acc, dir, n_blk, row, row < g.M, taddr, half, te, nx);
            if (te == 0) PV_TR(2, it, 3);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (TWO_CTA) { if (crank == 0) mbar_arrive(&tempty_bar[acc]); else mbar_arrive_cluster(&tempty_bar[acc], 0); }
                else mbar_arrive(&tempty_bar[acc]);
            }
        }
        cp_async_wait<0>();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                        // no CTA leaves while a peer may still multicast into it
    if (warp == 2) {
        tc_fence_after();
        if constexpr (TWO_CTA) tmem_dealloc_2sm(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace tc
