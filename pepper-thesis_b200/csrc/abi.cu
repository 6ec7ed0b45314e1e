// C-ABI plumbing of libpepper_b200.so: error state, device checks, batch validation and the host-buffer
// convenience entry point for the summary path (include/pepper_b200.h).
#include "common.cuh"
#include <mutex>
#include <thread>
#include <vector>

namespace pv {

char* err_buf() {
    static thread_local char buf[1024] = {0};
    return buf;
}

int set_error(int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(err_buf(), 1024, fmt, ap);
    va_end(ap);
    return code;
}

int require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return set_error(PV_ENODEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    int dev = 0;
    PV_CUDA_CHECK(cudaGetDevice(&dev));
    int major = 0;
    PV_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) return set_error(PV_ENODEVICE, "device %d has compute capability %d.x; kernels are built for sm_100a only", dev, major);
    return PV_OK;
}

int sm_count() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

// ---- launch counting / profiling -----------------------------------------------------------------------------------
namespace {
struct ProfRec { int fam; cudaEvent_t a, b; };
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<ProfRec> g_prof_open, g_prof_done;
std::vector<cudaEvent_t> g_prof_pool;
int64_t g_launches[FAM_COUNT] = {0};
double g_ms[FAM_COUNT] = {0};
int64_t g_prof_launches[FAM_COUNT] = {0};
cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
}  // namespace

void prof_begin(int fam, cudaStream_t stream) {
    std::lock_guard<std::mutex> l(g_prof_mu);
    if (!g_prof_on) return;
    ProfRec r; r.fam = fam; r.a = prof_event(); r.b = nullptr;
    cudaEventRecord(r.a, stream);
    g_prof_open.push_back(r);
}

bool prof_enabled() {
    std::lock_guard<std::mutex> l(g_prof_mu);
    return g_prof_on;
}

void count_launches(int fam, int launches) {
    std::lock_guard<std::mutex> l(g_prof_mu);
    g_launches[fam] += launches;
}

void prof_end(int fam, cudaStream_t stream, int launches) {
    std::lock_guard<std::mutex> l(g_prof_mu);
    g_launches[fam] += launches;
    if (!g_prof_on) return;
    for (size_t i = g_prof_open.size(); i-- > 0;)
        if (g_prof_open[i].fam == fam) {
            ProfRec r = g_prof_open[i];
            g_prof_open.erase(g_prof_open.begin() + i);
            r.b = prof_event();
            cudaEventRecord(r.b, stream);
            g_prof_done.push_back(r);
            g_prof_launches[fam] += launches;
            break;
        }
}

// grow-only device buffers reused across host-wrapper calls (one set per process; one process per GPU)
struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return PV_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { cudaGetLastError(); return set_error(PV_ENOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e)); }
        cap = want;
        return PV_OK;
    }
};

}  // namespace pv

extern "C" const char* pv_version(void) { return "pepper_b200 0.1.0 (sm_100a)"; }
extern "C" const char* pv_last_error(void) { return pv::err_buf(); }

extern "C" void pv_profile_enable(int on) {
    std::lock_guard<std::mutex> l(pv::g_prof_mu);
    pv::g_prof_on = on != 0;
}

// Waits for the recorded events, adds their elapsed times to the per-family totals and returns the totals since the
// last pv_profile_reset: ms[FAM_COUNT], launches[FAM_COUNT] (profiled launches only). Returns the family count.
extern "C" int pv_profile_collect(double* ms, int64_t* launches) {
    std::lock_guard<std::mutex> l(pv::g_prof_mu);
    for (auto& r : pv::g_prof_done) {
        float t = 0.f;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) pv::g_ms[r.fam] += t;
        pv::g_prof_pool.push_back(r.a); pv::g_prof_pool.push_back(r.b);
    }
    pv::g_prof_done.clear();
    for (int i = 0; i < pv::FAM_COUNT; i++) { if (ms) ms[i] = pv::g_ms[i]; if (launches) launches[i] = pv::g_prof_launches[i]; }
    return pv::FAM_COUNT;
}

extern "C" void pv_profile_reset(void) {
    std::lock_guard<std::mutex> l(pv::g_prof_mu);
    for (int i = 0; i < pv::FAM_COUNT; i++) { pv::g_ms[i] = 0; pv::g_prof_launches[i] = 0; }
}

// kernels launched by this library since load (all families)
extern "C" int64_t pv_launch_count(void) {
    std::lock_guard<std::mutex> l(pv::g_prof_mu);
    int64_t n = 0;
    for (int i = 0; i < pv::FAM_COUNT; i++) n += pv::g_launches[i];
    return n;
}

extern "C" int pv_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; d++) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ok++;
    }
    return ok;
}

extern "C" int pv_batch_validate(const PvReadBatch* b) {
    if (!b) return pv::set_error(PV_EINVAL, "null batch");
    if (b->n_reads < 0 || b->n_bases < 0 || b->n_ops < 0 || b->n_regions < 0 || b->n_ref < 0)
        return pv::set_error(PV_EINVAL, "negative size");
    if (b->bases2 && (b->n_base_exceptions < 0 || (b->n_base_exceptions && !b->base_exceptions))) return pv::set_error(PV_EINVAL, "bases2 needs its exception list");
    if ((b->bases4 || b->bases2 || b->quals_packed) && (b->n_bases & 15)) return pv::set_error(PV_EINVAL, "bases4 / quals_packed need n_bases to be a multiple of 16");
    if (b->quals_packed && (b->qual_bits < 1 || b->qual_bits > 7)) return pv::set_error(PV_EINVAL, "quals_packed needs qual_bits in 1..7");
    if (!b->quals_packed && !b->quals && b->n_bases) return pv::set_error(PV_EINVAL, "neither quals nor quals_packed given");
    if (!b->cigar16 && !b->cigar && b->n_ops) return pv::set_error(PV_EINVAL, "neither cigar nor cigar16 given");
    if (!b->bases4 && !b->bases2 && !b->bases && b->n_bases) return pv::set_error(PV_EINVAL, "neither bases nor a packed form of them given");
    if (b->n_regions == 0) return PV_OK;
    if (!b->region_read_begin || b->region_read_begin[0] != 0 || b->region_read_begin[b->n_regions] != b->n_reads)
        return pv::set_error(PV_EINVAL, "region_read_begin must start at 0 and end at n_reads");
    for (int32_t r = 0; r < b->n_regions; r++) {
        const int64_t L = b->region_ref_end[r] - b->region_ref_start[r] + 1;
        if (L <= 0) return pv::set_error(PV_EINVAL, "region %d: region_end < region_start", r);
        if (L > 0x7fffffffll) return pv::set_error(PV_EINVAL, "region %d longer than 2^31", r);
        if (b->region_ref_len[r] < L) return pv::set_error(PV_EINVAL, "region %d: reference_sequence shorter than the region", r);
        if (b->region_ref_off[r] < 0 || b->region_ref_off[r] + b->region_ref_len[r] > b->n_ref)
            return pv::set_error(PV_EINVAL, "region %d: reference outside ref[]", r);
        const int64_t n = b->region_read_begin[r + 1] - b->region_read_begin[r];
        if (n < 0) return pv::set_error(PV_EINVAL, "region_read_begin not monotone at region %d", r);
        if (n > 32767) return pv::set_error(PV_EINVAL, "region %d has %lld reads; more than 32767 would not fit the int16 windows", r, (long long)n);
    }
    for (int64_t i = 0; i < b->n_reads; i++) {
        if (b->read_base_off[i] < 0 || (b->read_base_off[i] & 15) || b->read_len[i] < 0 ||
            b->read_base_off[i] + b->read_len[i] > b->n_bases)
            return pv::set_error(PV_EINVAL, "read %lld: bases outside bases[] or not 16-byte aligned", (long long)i);
        if (b->read_cigar_off[i] < 0 || b->read_n_ops[i] < 0 || b->read_cigar_off[i] + b->read_n_ops[i] > b->n_ops)
            return pv::set_error(PV_EINVAL, "read %lld: cigar outside cigar[]", (long long)i);
        int64_t tot = 0;
        if (b->cigar) {
            const uint32_t* c = b->cigar + b->read_cigar_off[i];
            for (int32_t k = 0; k < b->read_n_ops[i]; k++) tot += (int64_t)(c[k] >> 4);
        } else {
            const uint16_t* c = b->cigar16 + b->read_cigar_off[i];
            for (int32_t k = 0; k < b->read_n_ops[i]; k++) tot += (int64_t)(c[k] >> 4);
        }
        if (tot > 0x3fffffffll) return pv::set_error(PV_EINVAL, "read %lld: CIGAR longer than 2^30", (long long)i);
    }
    if (b->min_qual < 0 || b->min_qual > 255) return pv::set_error(PV_EINVAL, "min_qual %d is no quality", b->min_qual);
    if (b->min_qual > 0 && b->quals) {                       // the promise is checkable wherever the plain qualities are at hand
        for (int64_t i = 0; i < b->n_reads; i++) {
            const uint8_t* q = b->quals + b->read_base_off[i];
            for (int32_t k = 0; k < b->read_len[i]; k++)
                if ((int32_t)q[k] < b->min_qual)
                    return pv::set_error(PV_EINVAL, "read %lld base %d has quality %d, below the batch's min_qual promise %d",
                                         (long long)i, k, (int)q[k], b->min_qual);
        }
    }
    return PV_OK;
}

namespace {

// 4-bit -> 8-bit bases: a 256-entry shared-memory table maps one packed byte to its two ASCII letters
__global__ void unpack_bases4_kernel(const uint8_t* __restrict__ packed, int64_t n_vec, uint8_t* __restrict__ bases) {
    __shared__ uint16_t lut[256];
    const char* nt16 = "=ACMGRSVTWYHKDBN";
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = (uint16_t)((uint8_t)nt16[i >> 4] | ((uint16_t)(uint8_t)nt16[i & 15] << 8));
    __syncthreads();
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec; v += (int64_t)gridDim.x * blockDim.x) {
        const uint2 in = __ldg((const uint2*)packed + v);                       // 8 packed bytes = 16 bases
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t w = j < 2 ? in.x : in.y;
            const uint32_t b0 = (w >> ((j & 1) * 16)) & 0xffu, b1 = (w >> ((j & 1) * 16 + 8)) & 0xffu;
            o[j] = (uint32_t)lut[b0] | ((uint32_t)lut[b1] << 16);
        }
        ((uint4*)bases)[v] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// dense bit stream -> bytes: a thread expands 32 qualities (`bits` aligned words in, two 16-byte stores out)
template <int BITS>
__global__ void unpack_quals_kernel(const uint32_t* __restrict__ packed, int64_t n_grp, int64_t n_bases, uint8_t* __restrict__ quals) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_grp; g += (int64_t)gridDim.x * blockDim.x) {
        uint32_t w[BITS + 1];
#pragma unroll
        for (int j = 0; j < BITS; j++) w[j] = __ldg(packed + g * BITS + j);
        w[BITS] = 0;
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 32; i++) {
            const int bit = i * BITS, wi = bit >> 5, sh = bit & 31;
            const uint32_t v = __funnelshift_r(w[wi], w[wi + 1], sh) & ((1u << BITS) - 1u);
            if ((i & 3) == 0) o[i >> 2] = v; else o[i >> 2] |= v << ((i & 3) * 8);
        }
        uint4* dst = (uint4*)(quals + g * 32);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        if (g * 32 + 16 < n_bases) dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
}

// 2-bit bases: a thread expands one packed word (16 bases)
__global__ void unpack_bases2_kernel(const uint32_t* __restrict__ packed, int64_t n_vec, uint8_t* __restrict__ bases) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec; v += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t w = __ldg(packed + v);
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t x = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) x |= ((0x54474341u >> (((w >> (8 * j + 2 * k)) & 3u) * 8)) & 0xffu) << (8 * k);   // "ACGT"
            o[j] = x;
        }
        ((uint4*)bases)[v] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}
__global__ void apply_base_exceptions_kernel(const uint64_t* __restrict__ exc, int64_t n, uint8_t* __restrict__ bases) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t e = exc[i];
        bases[e >> 8] = (uint8_t)(e & 0xffu);
    }
}

__global__ void unpack_cigar16_kernel(const uint16_t* __restrict__ packed, int64_t n, uint32_t* __restrict__ cigar) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) cigar[i] = packed[i];
}

template <class F> void parallel_ranges(int64_t n, int threads, F f) {
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    std::vector<std::thread> pool;
    const int64_t per = (n + threads - 1) / threads;
    for (int t = 0; t < threads; t++) pool.emplace_back([=]() { const int64_t lo = t * per, hi = lo + per < n ? lo + per : n; if (lo < hi) f(t, lo, hi); });
    for (auto& th : pool) th.join();
}

struct HostCtx {
    std::mutex mu;
    pv::DevBuf arr[18];      // batch arrays in PvReadBatch order
    pv::DevBuf ws, win, pos, reg, dep, frq, al, aln, cnt, dense, packed, packed_q, packed_c, packed_e;
    cudaStream_t stream = nullptr;
};
HostCtx& host_ctx() { static HostCtx c; return c; }

}  // namespace

extern "C" int pv_summary_status_offset(void);

extern "C" int pv_unpack_bases4(const uint8_t* packed_dev, int64_t n_bases, uint8_t* bases_dev, void* stream) {
    if (!packed_dev || !bases_dev || n_bases < 0 || (n_bases & 15)) return pv::set_error(PV_EINVAL, "pv_unpack_bases4: bad arguments (n_bases must be a multiple of 16)");
    if (n_bases == 0) return PV_OK;
    if (int rc = pv::require_device()) return rc;
    const int64_t n_vec = n_bases / 16;
    int64_t blocks = (n_vec + 255) / 256;
    const int64_t cap = (int64_t)pv::sm_count() * 16;
    if (blocks > cap) blocks = cap;
    unpack_bases4_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(packed_dev, n_vec, bases_dev);
    PV_CUDA_CHECK(cudaGetLastError());
    return PV_OK;
}

extern "C" int pv_unpack_bases2(const uint8_t* packed_dev, int64_t n_bases, const uint64_t* exc_dev, int64_t n_exc, uint8_t* bases_dev,
                                void* stream) {
    if (!packed_dev || !bases_dev || n_bases < 0 || (n_bases & 15) || n_exc < 0 || (n_exc && !exc_dev) || ((uintptr_t)packed_dev & 3))
        return pv::set_error(PV_EINVAL, "pv_unpack_bases2: bad arguments (n_bases %% 16 == 0, 4-byte aligned input)");
    if (n_bases == 0) return PV_OK;
    if (int rc = pv::require_device()) return rc;
    const int64_t n_vec = n_bases / 16, cap = (int64_t)pv::sm_count() * 16;
    int64_t blocks = (n_vec + 255) / 256;
    if (blocks > cap) blocks = cap;
    unpack_bases2_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint32_t*)packed_dev, n_vec, bases_dev);
    if (n_exc) {
        blocks = (n_exc + 255) / 256;
        if (blocks > cap) blocks = cap;
        apply_base_exceptions_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(exc_dev, n_exc, bases_dev);
    }
    PV_CUDA_CHECK(cudaGetLastError());
    return PV_OK;
}

extern "C" int pv_pack_bases2(const uint8_t* bases, int64_t n_bases, uint8_t* packed, uint64_t* exc, int64_t* n_exc, int32_t threads) {
    if (!bases || !packed || !n_exc || n_bases < 0 || (n_bases & 3)) return pv::set_error(PV_EINVAL, "pv_pack_bases2: bad arguments");
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    const int64_t n_out = n_bases / 4;
    std::vector<int64_t> cnt((size_t)threads + 1, 0);
    auto code = [](uint8_t c) -> int { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; };
    parallel_ranges(n_out, threads, [&](int t, int64_t lo, int64_t hi) {
        int64_t c = 0;
        for (int64_t i = lo; i < hi; i++) {
            uint8_t o = 0;
            for (int k = 0; k < 4; k++) {
                const uint8_t b = bases[4 * i + k];
                const int v = code(b);
                if (v >= 0) o |= (uint8_t)(v << (2 * k)); else if (b != 0) c++;
            }
            packed[i] = o;
        }
        cnt[(size_t)t + 1] = c;
    });
    for (int t = 0; t < threads; t++) cnt[(size_t)t + 1] += cnt[(size_t)t];
    const int64_t total = cnt[(size_t)threads];
    if (exc) {
        if (*n_exc != total) return pv::set_error(PV_EINVAL, "pv_pack_bases2: exception buffer holds %lld entries, %lld needed", (long long)*n_exc, (long long)total);
        parallel_ranges(n_out, threads, [&](int t, int64_t lo, int64_t hi) {
            int64_t at = cnt[(size_t)t];
            for (int64_t i = 4 * lo; i < 4 * hi; i++) { const uint8_t b = bases[i]; if (b != 0 && code(b) < 0) exc[at++] = ((uint64_t)i << 8) | b; }
        });
    }
    *n_exc = total;
    return PV_OK;
}

extern "C" int pv_unpack_quals(const uint8_t* packed_dev, int64_t n_bases, int32_t bits, uint8_t* quals_dev, void* stream) {
    if (!packed_dev || !quals_dev || n_bases < 0 || (n_bases & 15) || bits < 1 || bits > 7 || ((uintptr_t)packed_dev & 3))
        return pv::set_error(PV_EINVAL, "pv_unpack_quals: bad arguments (n_bases %% 16 == 0, qual_bits in 1..7, 4-byte aligned input)");
    if (n_bases == 0) return PV_OK;
    if (int rc = pv::require_device()) return rc;
    const int64_t n_grp = (n_bases + 31) / 32;
    int64_t blocks = (n_grp + 255) / 256;
    const int64_t cap = (int64_t)pv::sm_count() * 16;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t* p = (const uint32_t*)packed_dev;
    switch (bits) {
        case 1: unpack_quals_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(p, n_grp, n_bases, quals_dev); break;
        case 2: unpack_quals_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(p, n_grp, n_bases, quals_dev); break;
        case 3: unpack_quals_kernel<3><<<(unsigned)blocks, 256, 0, st>>>(p, n_grp, n_bases, quals_dev); break;
        case 4: unpack_quals_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(p, n_grp, n_bases, quals_dev); break;
        case 5: unpack_quals_kernel<5><<<(unsigned)blocks, 256, 0, st>>>(p, n_grp, n_bases, quals_dev); break;
        case 6: unpack_quals_kernel<6><<<(unsigned)blocks, 256, 0, st>>>(p, n_grp, n_bases, quals_dev); break;
        default: unpack_quals_kernel<7><<<(unsigned)blocks, 256, 0, st>>>(p, n_grp, n_bases, quals_dev); break;
    }
    PV_CUDA_CHECK(cudaGetLastError());
    return PV_OK;
}

extern "C" int32_t pv_qual_bits(const uint8_t* quals, int64_t n, int32_t threads) {
    if (!quals || n <= 0) return 1;
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    std::vector<uint8_t> mx((size_t)threads, 0);
    parallel_ranges(n, threads, [&](int t, int64_t lo, int64_t hi) { uint8_t m = 0; for (int64_t i = lo; i < hi; i++) m |= quals[i]; mx[(size_t)t] = m; });
    uint8_t m = 0;
    for (uint8_t v : mx) m |= v;
    int bits = 1;
    while (bits < 8 && (m >> bits)) bits++;
    return bits;
}

extern "C" int pv_pack_quals(const uint8_t* quals, int64_t n_bases, int32_t bits, uint8_t* packed, int32_t threads) {
    if (!quals || !packed || n_bases < 0 || (n_bases & 15) || bits < 1 || bits > 7) return pv::set_error(PV_EINVAL, "pv_pack_quals: bad arguments");
    const int64_t n_grp = (n_bases + 31) / 32;
    std::vector<int> bad(64, 0);
    parallel_ranges(n_grp, threads, [&](int t, int64_t lo, int64_t hi) {
        for (int64_t g = lo; g < hi; g++) {
            uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int i = 0; i < 32; i++) {
                const int64_t idx = g * 32 + i;
                const uint32_t v = idx < n_bases ? quals[idx] : 0u;
                if (v >> bits) bad[(size_t)t] = 1;
                const int bit = i * bits, wi = bit >> 5, sh = bit & 31;
                w[wi] |= v << sh;
                if (sh + bits > 32) w[wi + 1] |= v >> (32 - sh);
            }
            memcpy(packed + (size_t)g * bits * 4, w, (size_t)bits * 4);
        }
    });
    for (int v : bad) if (v) return pv::set_error(PV_EINVAL, "pv_pack_quals: a quality does not fit %d bits", bits);
    return PV_OK;
}

extern "C" int pv_unpack_cigar16(const uint16_t* packed_dev, int64_t n_ops, uint32_t* cigar_dev, void* stream) {
    if (n_ops < 0 || (n_ops && (!packed_dev || !cigar_dev))) return pv::set_error(PV_EINVAL, "pv_unpack_cigar16: bad arguments");
    if (n_ops == 0) return PV_OK;
    if (int rc = pv::require_device()) return rc;
    int64_t blocks = (n_ops + 255) / 256;
    const int64_t cap = (int64_t)pv::sm_count() * 16;
    if (blocks > cap) blocks = cap;
    unpack_cigar16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(packed_dev, n_ops, cigar_dev);
    PV_CUDA_CHECK(cudaGetLastError());
    return PV_OK;
}

extern "C" int pv_pack_cigar16(const uint32_t* cigar, int64_t n_ops, uint16_t* packed, int32_t threads) {
    if (n_ops < 0 || (n_ops && (!cigar || !packed))) return pv::set_error(PV_EINVAL, "pv_pack_cigar16: bad arguments");
    std::vector<int> bad(64, 0);
    parallel_ranges(n_ops, threads, [&](int t, int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; i++) { if (cigar[i] >> 16) bad[(size_t)t] = 1; packed[i] = (uint16_t)cigar[i]; }
    });
    for (int v : bad) if (v) return pv::set_error(PV_EINVAL, "pv_pack_cigar16: an op length is >= 4096");
    return PV_OK;
}

extern "C" int pv_pack_bases4(const uint8_t* bases, int64_t n_bases, uint8_t* packed, int32_t threads) {
    if (!bases || !packed || n_bases < 0 || (n_bases & 1)) return pv::set_error(PV_EINVAL, "pv_pack_bases4: bad arguments");
    uint8_t code[256];
    memset(code, 0xff, sizeof(code));
    const char* nt16 = "=ACMGRSVTWYHKDBN";
    for (int i = 0; i < 16; i++) code[(uint8_t)nt16[i]] = (uint8_t)i;
    code[0] = 0;                                            // padding
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    std::vector<int> bad((size_t)threads, 0);
    std::vector<std::thread> pool;
    const int64_t n_out = n_bases / 2, per = (n_out + threads - 1) / threads;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&, t]() {
            const int64_t lo = t * per, hi = lo + per < n_out ? lo + per : n_out;
            for (int64_t i = lo; i < hi; i++) {
                const uint8_t a = code[bases[2 * i]], c = code[bases[2 * i + 1]];
                if ((a | c) & 0xf0) bad[(size_t)t] = 1;
                packed[i] = (uint8_t)((a << 4) | (c & 15));
            }
        });
    for (auto& th : pool) th.join();
    for (int v : bad) if (v) return pv::set_error(PV_EINVAL, "pv_pack_bases4: a base is outside the nt16 alphabet");
    return PV_OK;
}

extern "C" int pv_summary_regions_host(const PvReadBatch* hb, const PvThresholds* thr, int32_t window, int32_t features,
                                       const PvCandidates* out, int64_t* n_candidates, int16_t* dense_image_host) {
    if (!hb || !thr || !out || !n_candidates) return pv::set_error(PV_EINVAL, "null argument");
    *n_candidates = 0;
    if (int rc = pv_batch_validate(hb)) return rc;
    if (int rc = pv::require_device()) return rc;
    if (hb->n_regions == 0) return PV_OK;
    HostCtx& h = host_ctx();
    std::lock_guard<std::mutex> lock(h.mu);
    if (!h.stream) PV_CUDA_CHECK(cudaStreamCreateWithFlags(&h.stream, cudaStreamNonBlocking));
    cudaStream_t st = h.stream;

    std::vector<int64_t> rlen(hb->n_regions);
    int64_t total = 0;
    for (int32_t r = 0; r < hb->n_regions; r++) { rlen[r] = hb->region_ref_end[r] - hb->region_ref_start[r] + 1; total += rlen[r]; }

    const void* src[18] = {hb->read_pos, hb->read_base_off, hb->read_len, hb->read_cigar_off, hb->read_n_ops, hb->read_flags,
                           hb->read_mapq, hb->bases, hb->quals, hb->cigar, hb->region_ref_start, hb->region_ref_end,
                           hb->region_cand_start, hb->region_cand_end, hb->region_ref_off, hb->region_ref_len,
                           hb->region_read_begin, hb->ref};
    const size_t bytes[18] = {(size_t)hb->n_reads * 8, (size_t)hb->n_reads * 8, (size_t)hb->n_reads * 4, (size_t)hb->n_reads * 8,
                              (size_t)hb->n_reads * 4, (size_t)hb->n_reads, (size_t)hb->n_reads, (size_t)hb->n_bases,
                              (size_t)hb->n_bases, (size_t)hb->n_ops * 4, (size_t)hb->n_regions * 8, (size_t)hb->n_regions * 8,
                              (size_t)hb->n_regions * 8, (size_t)hb->n_regions * 8, (size_t)hb->n_regions * 8,
                              (size_t)hb->n_regions * 8, (size_t)(hb->n_regions + 1) * 8, (size_t)hb->n_ref};
    // A batch whose min_qual promise clears both quality thresholds passes every quality test by construction (same rule as
    // pv_summary_regions): its quality array -- half of the plain bytes -- is not uploaded. The one case the promise cannot
    // decide (a read whose CIGAR runs over its own end) comes back as status bit 4 and the call is repeated with qualities.
    double qd = thr->min_snp_baseq; int qi = 0;
    if (qd > 256.0) qi = 256; else if (qd > 0.0) { qi = (int)qd; if ((double)qi < qd) qi++; }
    bool skip_q = hb->min_qual > 0 && hb->min_qual >= qi && (double)hb->min_qual >= thr->min_indel_baseq;
    for (int i = 0; i < 18; i++) {
        if (i == 8 && skip_q) continue;
        if (int rc = h.arr[i].reserve(bytes[i] + 16)) return rc;
        if (i == 7 && hb->bases2 && hb->n_bases) {          // bases travel 2-bit packed (+ exception list)
            if (int rc = h.packed.reserve((size_t)hb->n_bases / 4 + 16)) return rc;
            if (int rc = h.packed_e.reserve((size_t)hb->n_base_exceptions * 8 + 16)) return rc;
            PV_CUDA_CHECK(cudaMemcpyAsync(h.packed.p, hb->bases2, (size_t)hb->n_bases / 4, cudaMemcpyHostToDevice, st));
            if (hb->n_base_exceptions)
                PV_CUDA_CHECK(cudaMemcpyAsync(h.packed_e.p, hb->base_exceptions, (size_t)hb->n_base_exceptions * 8, cudaMemcpyHostToDevice, st));
            if (int rc = pv_unpack_bases2((const uint8_t*)h.packed.p, hb->n_bases, (const uint64_t*)h.packed_e.p, hb->n_base_exceptions,
                                          (uint8_t*)h.arr[7].p, st)) return rc;
            continue;
        }
        if (i == 7 && hb->bases4 && hb->n_bases) {          // bases travel 4-bit packed and are expanded on the device
            if (int rc = h.packed.reserve((size_t)hb->n_bases / 2 + 16)) return rc;
            PV_CUDA_CHECK(cudaMemcpyAsync(h.packed.p, hb->bases4, (size_t)hb->n_bases / 2, cudaMemcpyHostToDevice, st));
            if (int rc = pv_unpack_bases4((const uint8_t*)h.packed.p, hb->n_bases, (uint8_t*)h.arr[7].p, st)) return rc;
            continue;
        }
        if (i == 8 && hb->quals_packed && hb->n_bases) {     // qualities travel bit-packed
            const size_t pb = (size_t)((hb->n_bases + 31) / 32) * hb->qual_bits * 4;
            if (int rc = h.packed_q.reserve(pb + 16)) return rc;
            PV_CUDA_CHECK(cudaMemcpyAsync(h.packed_q.p, hb->quals_packed, pb, cudaMemcpyHostToDevice, st));
            if (int rc = pv_unpack_quals((const uint8_t*)h.packed_q.p, hb->n_bases, hb->qual_bits, (uint8_t*)h.arr[8].p, st)) return rc;
            continue;
        }
        if (i == 9 && hb->cigar16 && hb->n_ops) {            // CIGAR words travel as 16 bits
            if (int rc = h.packed_c.reserve((size_t)hb->n_ops * 2 + 16)) return rc;
            PV_CUDA_CHECK(cudaMemcpyAsync(h.packed_c.p, hb->cigar16, (size_t)hb->n_ops * 2, cudaMemcpyHostToDevice, st));
            if (int rc = pv_unpack_cigar16((const uint16_t*)h.packed_c.p, hb->n_ops, (uint32_t*)h.arr[9].p, st)) return rc;
            continue;
        }
        if (bytes[i]) PV_CUDA_CHECK(cudaMemcpyAsync(h.arr[i].p, src[i], bytes[i], cudaMemcpyHostToDevice, st));
    }
    PvReadBatch db = *hb;
    db.bases4 = nullptr; db.quals_packed = nullptr; db.qual_bits = 0; db.cigar16 = nullptr;
    db.bases2 = nullptr; db.base_exceptions = nullptr; db.n_base_exceptions = 0;
    db.read_pos = (const int64_t*)h.arr[0].p; db.read_base_off = (const int64_t*)h.arr[1].p; db.read_len = (const int32_t*)h.arr[2].p;
    db.read_cigar_off = (const int64_t*)h.arr[3].p; db.read_n_ops = (const int32_t*)h.arr[4].p; db.read_flags = (const uint8_t*)h.arr[5].p;
    db.read_mapq = (const uint8_t*)h.arr[6].p; db.bases = (const uint8_t*)h.arr[7].p; db.quals = skip_q ? nullptr : (const uint8_t*)h.arr[8].p;
    db.cigar = (const uint32_t*)h.arr[9].p; db.region_ref_start = (const int64_t*)h.arr[10].p; db.region_ref_end = (const int64_t*)h.arr[11].p;
    db.region_cand_start = (const int64_t*)h.arr[12].p; db.region_cand_end = (const int64_t*)h.arr[13].p;
    db.region_ref_off = (const int64_t*)h.arr[14].p; db.region_ref_len = (const int64_t*)h.arr[15].p;
    db.region_read_begin = (const int64_t*)h.arr[16].p; db.ref = (const uint8_t*)h.arr[17].p;

    const int64_t cap = out->capacity > 0 ? out->capacity : 1;
    int64_t max_len = 0;
    for (int32_t r = 0; r < hb->n_regions; r++) if (rlen[r] > max_len) max_len = rlen[r];
    const int64_t ws_bytes = pv_summary_workspace_bytes(hb->n_reads, hb->n_ops, hb->n_regions, total, max_len, cap);
    if (int rc = h.ws.reserve((size_t)ws_bytes)) return rc;
    if (int rc = h.win.reserve((size_t)cap * PV_WINDOW * PV_FEATURES * 2)) return rc;
    if (int rc = h.pos.reserve((size_t)cap * 8)) return rc;
    if (int rc = h.reg.reserve((size_t)cap * 4)) return rc;
    if (int rc = h.dep.reserve((size_t)cap * 4)) return rc;
    if (int rc = h.frq.reserve((size_t)cap * 4)) return rc;
    if (int rc = h.al.reserve((size_t)cap * PV_ALLELE_BYTES)) return rc;
    if (int rc = h.aln.reserve((size_t)cap)) return rc;
    if (int rc = h.cnt.reserve(64)) return rc;
    int16_t* dense_dev = nullptr;
    if (dense_image_host) {
        if (int rc = h.dense.reserve((size_t)total * PV_FEATURES * 2)) return rc;
        dense_dev = (int16_t*)h.dense.p;
    }
    PvCandidates dout;
    dout.capacity = cap; dout.windows = (int16_t*)h.win.p; dout.position = (int64_t*)h.pos.p; dout.region = (int32_t*)h.reg.p;
    dout.depth = (int32_t*)h.dep.p; dout.frequency = (int32_t*)h.frq.p; dout.allele = (uint8_t*)h.al.p; dout.allele_len = (uint8_t*)h.aln.p;

    int64_t found = 0; int32_t status = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        if (int rc = pv_summary_regions(&db, rlen.data(), total, thr, window, features, &dout, (int64_t*)h.cnt.p, h.ws.p,
                                        (int64_t)h.ws.cap, dense_dev, st)) return rc;
        PV_CUDA_CHECK(cudaMemcpyAsync(&found, h.cnt.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        PV_CUDA_CHECK(cudaMemcpyAsync(&status, (const uint8_t*)h.ws.p + pv_summary_status_offset(), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        PV_CUDA_CHECK(cudaStreamSynchronize(st));
        if (!(status & 16) || !skip_q) break;
        skip_q = false;                                      // rare: the qualities are needed after all
        if (int rc = h.arr[8].reserve(bytes[8] + 16)) return rc;
        if (hb->quals) PV_CUDA_CHECK(cudaMemcpyAsync(h.arr[8].p, hb->quals, bytes[8], cudaMemcpyHostToDevice, st));
        else {
            const size_t pb = (size_t)((hb->n_bases + 31) / 32) * hb->qual_bits * 4;
            if (int rc = h.packed_q.reserve(pb + 16)) return rc;
            PV_CUDA_CHECK(cudaMemcpyAsync(h.packed_q.p, hb->quals_packed, pb, cudaMemcpyHostToDevice, st));
            if (int rc = pv_unpack_quals((const uint8_t*)h.packed_q.p, hb->n_bases, hb->qual_bits, (uint8_t*)h.arr[8].p, st)) return rc;
        }
        db.quals = (const uint8_t*)h.arr[8].p;
    }
    *n_candidates = found;
    if (status & 40) return pv::set_error(PV_ECUDA, "internal inconsistency in the summary kernels (status %d)", status);
    if (status & 3) return pv::set_error(PV_EOVERFLOW, "site/event scratch overflow (status %d): raise the candidate capacity", status);
    const int64_t n = found < out->capacity ? found : out->capacity;
    if (n > 0) {
        PV_CUDA_CHECK(cudaMemcpyAsync(out->windows, dout.windows, (size_t)n * PV_WINDOW * PV_FEATURES * 2, cudaMemcpyDeviceToHost, st));
        PV_CUDA_CHECK(cudaMemcpyAsync(out->position, dout.position, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        PV_CUDA_CHECK(cudaMemcpyAsync(out->region, dout.region, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        PV_CUDA_CHECK(cudaMemcpyAsync(out->depth, dout.depth, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        PV_CUDA_CHECK(cudaMemcpyAsync(out->frequency, dout.frequency, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        PV_CUDA_CHECK(cudaMemcpyAsync(out->allele, dout.allele, (size_t)n * PV_ALLELE_BYTES, cudaMemcpyDeviceToHost, st));
        PV_CUDA_CHECK(cudaMemcpyAsync(out->allele_len, dout.allele_len, (size_t)n, cudaMemcpyDeviceToHost, st));
    }
    if (dense_image_host)
        PV_CUDA_CHECK(cudaMemcpyAsync(dense_image_host, dense_dev, (size_t)total * PV_FEATURES * 2, cudaMemcpyDeviceToHost, st));
    PV_CUDA_CHECK(cudaStreamSynchronize(st));
    if (found > out->capacity)
        return pv::set_error(PV_EOVERFLOW, "%lld candidates found but capacity is %lld", (long long)found, (long long)out->capacity);
    return PV_OK;
}
