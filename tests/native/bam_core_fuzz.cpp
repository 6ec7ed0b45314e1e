// TEST HARNESS ONLY (built by tests/test_bam_core_cpu.py with g++ -fsanitize=address,undefined): corrupts a valid inflated BAM
// record stream a few bytes at a time and runs the record / clip helpers of the GPU ingest (bam_core.cuh) over it exactly as
// the kernels of ingest_gpu.cu do -- chain over block_size fields, parse_record, record_endpos, clip_walk (count, then
// write), parse_hp, the base / quality reads of the write kernel -- inside a heap buffer of exactly the stream's size, so
// any read outside the stream is an AddressSanitizer report. Usage: bam_core_fuzz <stream file> <first record offset> <iters>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "bam_core.cuh"

using namespace bamcore;

static uint64_t rng_state = 0x9e3779b97f4a7c15ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 11); }

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> clean((size_t)n);
    if (fread(clean.data(), 1, (size_t)n, f) != (size_t)n) return 2;
    fclose(f);
    const int64_t first = atoll(argv[2]);
    const int iters = atoi(argv[3]);
    std::vector<int64_t> offs;
    for (int64_t off = first; off + 4 <= n;) {
        const RecHdr h = parse_record(clean.data(), off, n);
        if (!h.ok) { fprintf(stderr, "clean stream does not parse at %lld\n", (long long)off); return 3; }
        offs.push_back(off); off = h.rec_end;
    }
    long long parsed = 0, refused = 0, reads = 0;
    std::vector<uint32_t> ops(1 << 20);
    for (int it = 0; it < iters; it++) {
        uint8_t* U = (uint8_t*)malloc((size_t)n);              // exact size: ASan's red zone starts right behind the stream
        memcpy(U, clean.data(), (size_t)n);
        const int64_t victim = offs[rnd() % offs.size()];
        const int flips = 1 + (int)(rnd() % 6);
        for (int k = 0; k < flips; k++) {
            // mostly the fixed part and the first aux bytes, where the length fields live
            const int64_t span = (rnd() % 4) ? 64 : 4096;
            int64_t at = victim + (int64_t)(rnd() % span);
            if (at >= n) at = n - 1;
            U[at] = (rnd() % 3) ? (uint8_t)rnd() : (uint8_t)(U[at] ^ (1u << (rnd() % 8)));
        }
        // the chain of chain_kernel, then per record what clip_kernel / write_kernel do
        for (int64_t off = victim; off + 4 <= n;) {
            const uint32_t bs = ld32(U + off);
            if (bs < 32) { refused++; break; }
            const RecHdr h = parse_record(U, off, n);
            if (!h.ok) { refused++; break; }
            parsed++;
            const int64_t endpos = record_endpos(U, h);
            const int64_t start = h.pos > 50 ? h.pos - 50 + (int64_t)(rnd() % 200) : 0, stop = start + 1 + (int64_t)(rnd() % 3000);
            if (h.pos < stop && endpos > start && h.n_ops <= (int32_t)ops.size()) {
                const Clip c = clip_walk<false>(U, h, start, stop, nullptr);
                if (!c.bad && c.n_bases > 0 && !c.split) {
                    const Clip w = clip_walk<true>(U, h, start, stop, ops.data());
                    if (w.n_bases != c.n_bases || w.n_ops != c.n_ops) { fprintf(stderr, "count / write disagree\n"); return 4; }
                    volatile uint32_t sink = 0;
                    for (int64_t i = 0; i < c.n_bases; i++) sink += record_base(U, h, c.idx0 + i) + U[h.qual_off + c.idx0 + i];
                    sink += (uint32_t)parse_hp(U, h.aux_off, h.rec_end);
                    for (int k = 0; k < h.l_name && U[h.name_off + k]; k++) sink++;
                    reads++;
                }
            }
            off = h.rec_end;
            if (off - victim > 200000) break;
        }
        free(U);
    }
    printf("iterations %d records parsed %lld refused %lld reads cut %lld\n", iters, parsed, refused, reads);
    return 0;
}
