// Host-side packing of a group's big arrays into the compact wire forms that travel over PCIe (include/pepper_b200.h:
// bases2 + base_exceptions, cigar16), fast enough to run INSIDE the timed end-to-end path: the plain batch is what the
// caller owns (the analogue of the reference's pre-built type_read lists, read.h:60-108), one base per byte; a B200 behind a
// 50 GB/s PCIe link summarises faster than those bytes arrive, so the host squeezes each group into 2 bits per base and
// 16 bits per CIGAR op while the group before it is on the wire. One streaming pass per array: AVX2 (runtime-dispatched;
// a portable 64-bit path otherwise), non-temporal stores into the page-locked staging buffer, one thread per slice.
//
// Compiled by g++ (not nvcc: target attributes + immintrin) and linked into libpepper_b200.so.
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <unistd.h>

#include "pepper_b200.h"

#if defined(__x86_64__)
#include <immintrin.h>
#define PV_X86 1
#else
#define PV_X86 0
#endif

namespace pv { int set_error(int code, const char* fmt, ...); }

namespace {

// 2-bit code of an upper-case A/C/G/T byte: A 0x41 -> 0, C 0x43 -> 1, G 0x47 -> 2, T 0x54 -> 3
inline uint32_t code2(uint8_t b) { return ((uint32_t)(b >> 1) ^ (uint32_t)(b >> 2)) & 3u; }
inline bool plain_base(uint8_t b) { return b == 'A' || b == 'C' || b == 'G' || b == 'T'; }

// bases [lo, hi) (lo, hi multiples of 4) -> packed[lo/4, hi/4); exceptions appended as (index << 8 | byte)
void pack2_scalar(const uint8_t* bases, int64_t lo, int64_t hi, uint8_t* packed, std::vector<uint64_t>& exc) {
    for (int64_t i = lo; i < hi; i += 4) {
        uint32_t o = 0;
        for (int k = 0; k < 4; k++) {
            const uint8_t b = bases[i + k];
            if (plain_base(b)) o |= code2(b) << (2 * k);
            else if (b != 0) exc.push_back(((uint64_t)(i + k) << 8) | b);
        }
        packed[i >> 2] = (uint8_t)o;
    }
}

#if PV_X86
// 32 bases -> one packed byte in every 32-bit lane; *valid is ANDed with 0xff where the byte is A/C/G/T or 0. Both the byte a
// low nibble has to belong to and its 2-bit code come from 16-entry tables (pshufb): nibble 1 'A', 3 'C', 7 'G', 4 'T', 0 the
// padding byte; every other nibble expects a byte with a different nibble, i.e. never matches.
__attribute__((target("avx2"))) inline __m256i pack2_32(const __m256i v, __m256i* valid) {
    const __m256i expect = _mm256_setr_epi8(0, 'A', 1, 'C', 'T', 1, 1, 'G', 1, 1, 1, 1, 1, 1, 1, 1, 0, 'A', 1, 'C', 'T', 1, 1, 'G', 1, 1, 1, 1, 1, 1, 1, 1);
    const __m256i codes = _mm256_setr_epi8(0, 0, 0, 1, 3, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 3, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i nib = _mm256_and_si256(v, _mm256_set1_epi8(0x0f));
    const __m256i ok = _mm256_cmpeq_epi8(_mm256_shuffle_epi8(expect, nib), v);
    *valid = _mm256_and_si256(*valid, ok);
    const __m256i code = _mm256_and_si256(_mm256_shuffle_epi8(codes, nib), ok);
    // c0 + 4 c1 per 16-bit lane, then + 16 (c2 + 4 c3) per 32-bit lane
    return _mm256_madd_epi16(_mm256_maddubs_epi16(code, _mm256_set1_epi16(0x0401)), _mm256_set1_epi32(0x00100001));
}

__attribute__((target("avx2"))) void pack2_avx2(const uint8_t* bases, int64_t lo, int64_t hi, uint8_t* packed, std::vector<uint64_t>& exc) {
    int64_t i = lo;
    // head up to a 32-byte boundary of the OUTPUT (the streaming stores need packed + i/4 aligned)
    while (i < hi && (((uintptr_t)(packed + (i >> 2))) & 31)) { pack2_scalar(bases, i, i + 4, packed, exc); i += 4; }
    const __m256i order = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
    for (; i + 128 <= hi; i += 128) {
        __m256i valid = _mm256_set1_epi8(-1);
        const __m256i a = pack2_32(_mm256_loadu_si256((const __m256i*)(bases + i)), &valid);
        const __m256i b = pack2_32(_mm256_loadu_si256((const __m256i*)(bases + i + 32)), &valid);
        const __m256i c = pack2_32(_mm256_loadu_si256((const __m256i*)(bases + i + 64)), &valid);
        const __m256i d = pack2_32(_mm256_loadu_si256((const __m256i*)(bases + i + 96)), &valid);
        // dwords -> bytes: per 128-bit lane a0-3 b0-3 c0-3 d0-3 | a4-7 b4-7 c4-7 d4-7, then the lanes interleaved dword-wise
        const __m256i pk = _mm256_packus_epi16(_mm256_packus_epi32(a, b), _mm256_packus_epi32(c, d));
        _mm256_stream_si256((__m256i*)(packed + (i >> 2)), _mm256_permutevar8x32_epi32(pk, order));
        if (_mm256_movemask_epi8(valid) != -1)
            for (int64_t j = i; j < i + 128; j++) { const uint8_t x = bases[j]; if (x != 0 && !plain_base(x)) exc.push_back(((uint64_t)j << 8) | x); }
    }
    _mm_sfence();
    if (i < hi) pack2_scalar(bases, i, hi, packed, exc);
}

// 16 CIGAR words -> 16 half words; returns the OR of the words (bits 16.. set = an op length >= 4096)
__attribute__((target("avx2"))) uint32_t cigar16_avx2(const uint32_t* cigar, int64_t lo, int64_t hi, uint16_t* packed) {
    uint32_t acc = 0;
    int64_t i = lo;
    while (i < hi && (((uintptr_t)(packed + i)) & 31)) { acc |= cigar[i]; packed[i] = (uint16_t)cigar[i]; i++; }
    __m256i vacc = _mm256_setzero_si256();
    for (; i + 16 <= hi; i += 16) {
        const __m256i x = _mm256_loadu_si256((const __m256i*)(cigar + i)), y = _mm256_loadu_si256((const __m256i*)(cigar + i + 8));
        vacc = _mm256_or_si256(vacc, _mm256_or_si256(x, y));
        const __m256i m = _mm256_set1_epi32(0xffff);
        const __m256i pk = _mm256_packus_epi32(_mm256_and_si256(x, m), _mm256_and_si256(y, m));   // lanes interleaved: x0-3 y0-3 | x4-7 y4-7
        _mm256_stream_si256((__m256i*)(packed + i), _mm256_permute4x64_epi64(pk, 0xd8));
    }
    _mm_sfence();
    alignas(32) uint32_t t[8];
    _mm256_store_si256((__m256i*)t, vacc);
    for (int k = 0; k < 8; k++) acc |= t[k];
    for (; i < hi; i++) { acc |= cigar[i]; packed[i] = (uint16_t)cigar[i]; }
    return acc;
}
bool have_avx2() { static const bool v = __builtin_cpu_supports("avx2"); return v; }
#else
bool have_avx2() { return false; }
#endif

void pack2_range(const uint8_t* bases, int64_t lo, int64_t hi, uint8_t* packed, std::vector<uint64_t>& exc) {
#if PV_X86
    if (have_avx2()) { pack2_avx2(bases, lo, hi, packed, exc); return; }
#endif
    pack2_scalar(bases, lo, hi, packed, exc);
}
uint32_t cigar16_range(const uint32_t* cigar, int64_t lo, int64_t hi, uint16_t* packed) {
#if PV_X86
    if (have_avx2()) return cigar16_avx2(cigar, lo, hi, packed);
#endif
    uint32_t acc = 0;
    for (int64_t i = lo; i < hi; i++) { acc |= cigar[i]; packed[i] = (uint16_t)cigar[i]; }
    return acc;
}

// Persistent workers: a group is packed in 2-4 ms, so spawning a dozen threads per call would cost a tenth of it. Worker i
// runs slice i + 1 of a job, the caller slice 0. One job at a time (callers queue on `gate`). The pool is leaked on purpose:
// its detached workers sleep on the condition variable until the process ends.
class Pool {
  public:
    void run(int n, const std::function<void(int)>& f) {
        std::lock_guard<std::mutex> one(gate);
        {
            std::unique_lock<std::mutex> lk(mu);
            while ((int)workers < n - 1) { std::thread(&Pool::loop, this, (int)workers).detach(); workers++; }
            job = &f; n_jobs = n; pending = n - 1; gen++;
        }
        cv_work.notify_all();
        f(0);
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending == 0; });
        job = nullptr;
    }

  private:
    void loop(int idx) {
        uint64_t seen = 0;
        {   // a worker born during a job's set-up starts with that job
            std::unique_lock<std::mutex> lk(mu);
            seen = gen - 1;
        }
        while (true) {
            const std::function<void(int)>* f = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return gen != seen; });
                seen = gen;
                if (idx + 1 < n_jobs) f = job;
            }
            if (f) {
                (*f)(idx + 1);
                std::unique_lock<std::mutex> lk(mu);
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
    std::mutex gate, mu;
    std::condition_variable cv_work, cv_done;
    const std::function<void(int)>* job = nullptr;
    int n_jobs = 0, pending = 0;
    size_t workers = 0;
    uint64_t gen = 1;
};
Pool& pool() {          // a forked child owns no workers: it gets a pool of its own
    static std::mutex m;
    static Pool* p = nullptr;
    static pid_t owner = 0;
    std::lock_guard<std::mutex> lk(m);
    if (!p || owner != getpid()) { p = new Pool(); owner = getpid(); }
    return *p;
}

}  // namespace

// see include/pepper_b200.h
extern "C" int pv_pack_group(const uint8_t* bases, int64_t n_bases, uint8_t* bases2, uint64_t* exceptions, int64_t exception_cap,
                             int64_t* n_exceptions, const uint32_t* cigar, int64_t n_ops, uint16_t* cigar16, int32_t* cigar_fits,
                             int32_t threads) {
    if (n_bases < 0 || (n_bases & 3) || n_ops < 0 || exception_cap < 0 || !n_exceptions || !cigar_fits ||
        (n_bases && (!bases || !bases2)) || (n_ops && (!cigar || !cigar16)) || (exception_cap && !exceptions))
        return pv::set_error(PV_EINVAL, "pv_pack_group: bad arguments (n_bases %% 4 == 0, outputs non-null)");
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    // Chunks of 1 MiB of bases / 256 Ki ops, claimed from a shared counter: a worker that loses its core for a while (the
    // caller's other threads drive the GPU on the same cores) holds up one chunk, not a fixed share of the group.
    constexpr int64_t BASE_CHUNK = 1 << 20, OP_CHUNK = 1 << 18;
    const int64_t nb_chunks = (n_bases + BASE_CHUNK - 1) / BASE_CHUNK, no_chunks = (n_ops + OP_CHUNK - 1) / OP_CHUNK;
    std::vector<std::vector<uint64_t>> exc((size_t)nb_chunks);
    std::vector<uint32_t> hib((size_t)threads, 0u);
    std::atomic<int64_t> next{0};
    auto work = [&](int t) {
        while (true) {
            const int64_t c = next.fetch_add(1, std::memory_order_relaxed);
            if (c >= nb_chunks + no_chunks) break;
            if (c < nb_chunks) {
                pack2_range(bases, c * BASE_CHUNK, std::min(n_bases, (c + 1) * BASE_CHUNK), bases2, exc[(size_t)c]);
            } else {
                const int64_t k = c - nb_chunks;
                hib[(size_t)t] |= cigar16_range(cigar, k * OP_CHUNK, std::min(n_ops, (k + 1) * OP_CHUNK), cigar16);
            }
        }
    };
    if (threads == 1 || nb_chunks + no_chunks <= 1) work(0);
    else pool().run((int)std::min<int64_t>(threads, nb_chunks + no_chunks), work);
    int64_t total = 0;
    for (auto& v : exc) total += (int64_t)v.size();
    *n_exceptions = total;
    if (total <= exception_cap) {
        int64_t at = 0;
        for (auto& v : exc) { if (!v.empty()) memcpy(exceptions + at, v.data(), v.size() * 8); at += (int64_t)v.size(); }
    }
    uint32_t acc = 0;
    for (uint32_t v : hib) acc |= v;
    *cigar_fits = (acc >> 16) ? 0 : 1;
    return PV_OK;
}
