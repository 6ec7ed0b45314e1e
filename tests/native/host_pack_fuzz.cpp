// ASan + UBSan harness of pv_pack_group (csrc/host_pack.cpp): random sizes, misaligned exact-size heap buffers, 1-9 threads;
// every output is checked against the formats of include/pepper_b200.h (tests/test_host_pack_cpu.py compiles and runs it).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cstdarg>
namespace pv { int set_error(int code, const char* fmt, ...) { return code; } }
extern "C" int pv_pack_group(const uint8_t*, int64_t, uint8_t*, uint64_t*, int64_t, int64_t*, const uint32_t*, int64_t, uint16_t*, int32_t*, int32_t);
int main() {
    srand(7);
    for (int trial = 0; trial < 120; trial++) {
        int64_t n = (rand() % 3 == 0 ? (int64_t)(rand() % 3000000) : (int64_t)(rand() % 5000)) & ~3ll;
        int64_t m = rand() % 2 ? rand() % 700000 : rand() % 100;
        int offb = rand() % 40, offo = rand() % 40, offc = rand() % 9, offk = rand() % 9;
        // exact-size heap buffers so that any overrun is caught
        uint8_t* bb = (uint8_t*)malloc(n + offb + 1); uint8_t* b = bb + offb;
        uint8_t* oo = (uint8_t*)malloc(n / 4 + offo + 1); uint8_t* o = oo + offo;
        uint32_t* cc = (uint32_t*)malloc((m + offc) * 4 + 4); uint32_t* c = cc + offc;
        uint16_t* kk = (uint16_t*)malloc((m + offk) * 2 + 2); uint16_t* k = kk + offk;
        for (int64_t i = 0; i < n; i++) b[i] = (rand() % 37 == 0) ? (uint8_t)rand() : "ACGT"[rand() & 3];
        for (int64_t i = 0; i < m; i++) c[i] = (uint32_t)rand() % (rand() % 50 == 0 ? 1u << 24 : 1u << 14);
        int64_t cap = rand() % 2 ? 0 : rand() % 100000;
        uint64_t* e = cap ? (uint64_t*)malloc(cap * 8) : nullptr;
        int64_t ne = -1; int32_t fits = -1;
        int rc = pv_pack_group(n ? b : nullptr, n, n ? o : nullptr, e, cap, &ne, m ? c : nullptr, m, m ? k : nullptr, &fits, 1 + rand() % 9);
        if (rc != 0) { printf("rc %d\n", rc); return 1; }
        // verify
        int64_t cnt = 0;
        for (int64_t i = 0; i < n; i++) {
            uint8_t x = b[i]; int code = x == 'A' ? 0 : x == 'C' ? 1 : x == 'G' ? 2 : x == 'T' ? 3 : -1;
            int got = (o[i >> 2] >> (2 * (i & 3))) & 3;
            if (code >= 0 ? got != code : got != 0) { printf("mismatch at %lld\n", (long long)i); return 1; }
            if (code < 0 && x != 0) { if (cnt < cap && ne <= cap && e[cnt] != (((uint64_t)i << 8) | x)) { printf("exc mismatch\n"); return 1; } cnt++; }
        }
        if (cnt != ne) { printf("count %lld vs %lld\n", (long long)cnt, (long long)ne); return 1; }
        uint32_t acc = 0;
        for (int64_t i = 0; i < m; i++) { acc |= c[i]; if (k[i] != (uint16_t)c[i]) { printf("cigar mismatch\n"); return 1; } }
        if (fits != ((acc >> 16) ? 0 : 1)) { printf("fits\n"); return 1; }
        free(bb); free(oo); free(cc); free(kk); free(e);
    }
    printf("asan harness ok\n");
}
