// Shared host/device helpers of libpepper_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include "pepper_b200.h"

namespace pv {

// thread-local last-error message (pv_last_error)
char* err_buf();
int set_error(int code, const char* fmt, ...);

#define PV_CUDA_CHECK(expr)                                                                        \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return pv::set_error(PV_ECUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,     \
                                 cudaGetErrorString(e_));                                          \
    } while (0)

// 0 when a usable sm_100 device is current, else PV_ENODEVICE with the message set.
int require_device();
int sm_count();

// kernel families for launch counting and optional CUDA-event profiling (pv_profile_*)
enum Family { FAM_SUM_PREFIX = 0, FAM_SUM_TILE, FAM_SUM_ALLELE, FAM_SUM_SORT, FAM_SUM_EMIT, FAM_LSTM_PREP, FAM_LSTM_ENC,
              FAM_LSTM_DEC, FAM_LSTM_MLP, FAM_GRU_STEP, FAM_GRU_MISC, FAM_FILTER, FAM_POLISH, FAM_GRU_GX, FAM_GRU_HEAD, FAM_INGEST, FAM_COUNT };
// brackets `launches` kernel launches of one family on `stream`; records events only while profiling is enabled
void prof_begin(int fam, cudaStream_t stream);
void prof_end(int fam, cudaStream_t stream, int launches);
bool prof_enabled();                                  // per-family timing is on (pv_profile_enable)
void count_launches(int fam, int launches);           // launches that did not pass through prof_end (graph replays)

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// bump allocator over a caller-provided workspace
struct Arena {
    uint8_t* base; int64_t size; int64_t cur;
    Arena(void* p, int64_t n) : base((uint8_t*)p), size(n), cur(0) {}
    template <typename T> T* take(int64_t count) {
        cur = align_up(cur, 256);
        T* p = (T*)(base ? base + cur : nullptr);
        cur += count * (int64_t)sizeof(T);
        return p;
    }
    bool ok() const { return cur <= size; }
};

}  // namespace pv
