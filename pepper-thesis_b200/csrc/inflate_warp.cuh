// DEFLATE (RFC 1951) for BGZF blocks, one WARP per block (the host twin is csrc/fast_inflate.h).
//
// A BGZF block is an independent DEFLATE stream of <= 64 KiB, and its symbols can only be decoded one after the other, so
// the unit of parallelism is the block: a warp owns one, keeps the bit reader in registers (every lane holds the same
// state and executes the same instructions: a warp instruction costs the same with one active lane or 32, and this way
// every lane knows every symbol without a shuffle) and uses its 32 lanes where DEFLATE is parallel:
//   * Huffman tables are built by the warp (counts by shared atomics, canonical order by match_any / popc, replicas
//     filled lane per symbol) into shared memory: a 10-bit table for literals / lengths whose entries carry TWO literals
//     when both codes fit the index, an 8-bit table for distances; longer codes take a canonical bit walk;
//   * literals collect in a register per lane and leave as one 32-byte store per 32 bytes of output;
//   * matches are copied by all lanes (32 bytes per step, periodic patterns of short distances expanded directly);
//   * the block's CRC-32 is 32 partial CRCs over 1/32 of the bytes each, combined with x^(8 n) mod P (zlib's
//     crc32_combine identity), and compared with the BGZF trailer.
// The input arrives through a three-word window (two words shifted by a funnel shift, the third prefetched).
#pragma once
#include <stdint.h>

namespace winf {

#ifndef PV_LIT_ROOT
#define PV_LIT_ROOT 10
#endif
constexpr int LIT_ROOT = PV_LIT_ROOT, DIST_ROOT = 8, CL_ROOT = 7;
constexpr unsigned FULL = 0xffffffffu;
constexpr uint32_t CRC_POLY = 0xedb88320u;

struct alignas(16) WarpTables {
    uint32_t lit[1 << LIT_ROOT];       // n1 | type << 4 | ...   (0 = longer code or none)
    uint32_t dist[1 << DIST_ROOT];     // n | extra << 4 | invalid << 8 | base << 16; the code-length code's table while a header is read
    uint32_t count_lit[16], count_dist[16];   // codes per length (canonical walk of long codes)
    uint32_t cur[16], base[16];        // build: cursor into the sorted symbols per length, first code - first index
    uint16_t sym_lit[288], sym_dist[32];
    uint8_t lens[336];
    uint8_t cl[32];
};

// entry of the literal / length table for symbol s with an n-bit code: type 1 literal (2 = two literals, made by the
// pairing pass), 3 length (extra << 8 | base << 16), 4 end of block, 7 invalid symbol
__device__ __forceinline__ uint32_t lit_entry(int s, int n) {
    if (s < 256) return (uint32_t)n | (1u << 4) | ((uint32_t)s << 8);
    if (s == 256) return (uint32_t)n | (4u << 4);
    const int k = s - 257;
    if (k >= 29) return (uint32_t)n | (7u << 4);
    uint32_t extra = 0, base;
    if (k < 8) base = 3 + k;
    else if (k == 28) base = 258;
    else { extra = (k >> 2) - 1; base = ((4u + (k & 3)) << extra) + 3; }
    return (uint32_t)n | (3u << 4) | (extra << 8) | (base << 16);
}
__device__ __forceinline__ uint32_t dist_entry(int d, int n) {
    if (d >= 30) return (uint32_t)n | 0x100u;
    uint32_t extra = 0, base;
    if (d < 4) base = d + 1;
    else { extra = (d >> 1) - 1; base = ((2u + (d & 1)) << extra) + 1; }
    return (uint32_t)n | (extra << 4) | (base << 16);
}

// canonical walk for codes longer than the table's root; -1 = no such code
__device__ __forceinline__ int slow_decode(uint32_t win, const uint32_t* count, const uint16_t* sym, int& nbits) {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; len++) {
        code |= (int)(win & 1u); win >>= 1;
        const int c = (int)count[len];
        if (code - c < first) { nbits = len; return sym[index + (code - first)]; }
        index += c; first += c; first <<= 1; code <<= 1;
    }
    return -1;
}

// KIND 0: code-length code (entry n | symbol << 4), 1: literal / length, 2: distance. false = over-subscribed code.
template <int KIND>
__device__ bool build_table(const uint8_t* lens, int n, uint32_t* count, uint32_t* cur, uint32_t* base, uint16_t* sym, uint32_t* tab,
                            const int root, const int lane) {
    for (int i = lane; i < (1 << root); i += 32) tab[i] = 0;
    if (lane < 16) count[lane] = 0;
    __syncwarp();
    for (int s = lane; s < n; s += 32) atomicAdd(&count[lens[s]], 1u);
    __syncwarp();
    int left = 1;
    uint32_t code = 0, off = 0, prev = 0;
    bool over = false;
    for (int l = 1; l <= 15; l++) {                               // every lane computes the same; lane l files it
        const uint32_t c = count[l];
        code = (code + prev) << 1;
        left = (left << 1) - (int)c;
        over |= left < 0;
        if (lane == l) { cur[l] = off; base[l] = code - off; }
        off += c; prev = c;
    }
    if (over) return false;
    __syncwarp();
    for (int b0 = 0; b0 < n; b0 += 32) {
        const int s = b0 + lane;
        const int l = s < n ? lens[s] : 0;
        const unsigned m = __match_any_sync(FULL, l);
        const int rank = __popc(m & ((1u << lane) - 1u));
        uint32_t pos = 0;
        if (l) pos = cur[l] + rank;
        __syncwarp();
        if (l && rank == __popc(m) - 1) cur[l] = pos + 1;         // the last lane of the group moves the cursor
        __syncwarp();
        if (l) {
            sym[pos] = (uint16_t)s;
            if (l <= root) {
                const uint32_t c = pos + base[l];
                const uint32_t r = __brev(c) >> (32 - l);
                const uint32_t e = KIND == 0 ? ((uint32_t)l | ((uint32_t)s << 4)) : KIND == 1 ? lit_entry(s, l) : dist_entry(s, l);
                for (uint32_t i = r; i < (1u << root); i += 1u << l) tab[i] = e;
            }
        }
    }
    __syncwarp();
    if (KIND == 1) {                                               // pairing pass: a literal whose successor is decided by the index too
        for (int i = lane; i < (1 << root); i += 32) {
            const uint32_t e = tab[i];
            const int n1 = e & 15;
            if (((e >> 4) & 7) == 1 && n1 < root) {
                const uint32_t e2 = tab[i >> n1];                  // may already be paired: its first literal and n1 do not change
                const int n2 = e2 & 15, t2 = (e2 >> 4) & 7;
                if ((t2 == 1 || t2 == 2) && n1 + n2 <= root)
                    tab[i] = (e & 0xff0fu) | (2u << 4) | ((e2 & 0xff00u) << 8) | ((uint32_t)(n1 + n2) << 24);
            }
        }
        __syncwarp();
    }
    return true;
}

struct Bits {
    const uint32_t* w;     // word that holds the stream's first byte
    int n_words;           // words readable from w (the buffer's last word may be a partial one: see pv_bam_inflate_blocks)
    int wp;                // next word to fetch
    uint32_t lo, hi, nxt;
    int bp;                // bit position inside lo, < 32 between symbols
    __device__ __forceinline__ uint32_t ldw(int i) const { return i < n_words ? w[i] : 0u; }
    __device__ __forceinline__ void seek(int byte_off) {
        const int w0 = byte_off >> 2;
        bp = (byte_off & 3) * 8;
        lo = ldw(w0); hi = ldw(w0 + 1); nxt = ldw(w0 + 2); wp = w0 + 3;
    }
    __device__ __forceinline__ uint32_t window() const { return __funnelshift_r(lo, hi, bp); }
    __device__ __forceinline__ void norm() {
        if (bp >= 32) {
            lo = hi; hi = nxt; nxt = ldw(wp);
            if ((wp & 31) == 0 && wp + 64 < n_words) asm volatile("prefetch.global.L1 [%0];" ::"l"(w + wp + 64));
            wp++; bp -= 32;
        }
    }
    __device__ __forceinline__ int64_t bitpos() const { return (int64_t)(wp - 3) * 32 + bp; }
};

__device__ __forceinline__ uint32_t multmodp(uint32_t a, uint32_t b) {
    uint32_t p = 0;
    for (int i = 0; i < 32; i++) {
        if (a & (0x80000000u >> i)) p ^= b;
        b = (b & 1u) ? (b >> 1) ^ CRC_POLY : b >> 1;
    }
    return p;
}

// x^(2^k) mod P, k = 0 .. 31 (reflected), zlib's x2n_table
__device__ const uint32_t X2N[32] = {
    0x40000000u, 0x20000000u, 0x08000000u, 0x00800000u, 0x00008000u, 0xedb88320u, 0xb1e6b092u, 0xa06a2517u, 0xed627daeu, 0x88d14467u, 0xd7bbfe6au,
    0xec447f11u, 0x8e7ea170u, 0x6427800eu, 0x4d47bae0u, 0x09fe548fu, 0x83852d0fu, 0x30362f1au, 0x7b5a9cc3u, 0x31fec169u, 0x9fec022au, 0x6c8dedc4u,
    0x15d6874du, 0x5fde7a4eu, 0xbad90e37u, 0x2e4e5eefu, 0x4eaba214u, 0xa8a472c0u, 0x429a969eu, 0x148d302au, 0xc40ba6d0u, 0xc4e22c3cu};

__device__ const uint8_t CL_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// CRC-32 of out[0 .. n) by the warp; crc_tab = the byte table in shared memory
__device__ uint32_t crc32_warp(const uint8_t* out, int n, const uint32_t* crc_tab, int lane) {
    const int L = (n + 31) >> 5;
    const int a = min(lane * L, n), b = min(a + L, n);
    uint32_t c = 0xffffffffu;
    for (int i = a; i < b; i++) c = crc_tab[(c ^ out[i]) & 0xffu] ^ (c >> 8);
    c = a < b ? c ^ 0xffffffffu : 0u;
    uint32_t after = (uint32_t)(n - b), p = 0x80000000u;          // p = x^(8 * after)
    for (int k = 3; after; after >>= 1, k++)
        if (after & 1u) p = multmodp(X2N[k & 31], p);
    c = multmodp(p, c);
    for (int d = 16; d; d >>= 1) c ^= __shfl_xor_sync(FULL, c, d);
    return c;
}

struct Dec {
    Bits B;
    int op, pend;          // bytes produced; first byte still held in the lanes' registers (bytes [pend, op) of the current 32)
    uint32_t mine;         // this lane's byte of the current 32-byte group
};

// the literals held in registers go to memory (before a copy reads them, and at the end)
__device__ __forceinline__ void flush_pending(Dec& D, uint8_t* out, const int lane) {
    if (D.pend < D.op) {
        const int pos = ((D.op - 1) & ~31) + lane;
        if (pos >= D.pend && pos < D.op) out[pos] = (uint8_t)D.mine;
        D.pend = D.op;
    }
}
// D.op just reached a multiple of 32: the group leaves as one 32-byte store
__device__ __forceinline__ void flush_group(Dec& D, uint8_t* out, const int lane) {
    const int pos = D.op - 32 + lane;
    if (pos >= D.pend) out[pos] = (uint8_t)D.mine;
    D.pend = D.op;
}

constexpr int FAST_MARGIN = 36;   // fast mode holds while op + FAST_MARGIN <= n_out was true at the last group store / match

// Symbols of one block until its end-of-block code. CAREFUL = false checks the output bound only where a 32-byte group is
// stored and behind a match (literals in between cannot pass it: <= 33 bytes follow a check that left FAST_MARGIN) and
// hands over to the CAREFUL instance near the end of the output. Returns 0 end of block, 1 continue CAREFUL, -1 error.
template <bool CAREFUL>
__device__ __forceinline__ int decode_symbols(Dec& D, WarpTables& T, uint8_t* __restrict__ out, const int n_out, const int lane) {
    Bits& B = D.B;
    for (;;) {
        uint32_t w = B.window();
        uint32_t e = T.lit[w & ((1u << LIT_ROOT) - 1u)];
        uint32_t t = e & 0x70u;
        if (t == 0x20u) {                                         // two literals
            if (CAREFUL && D.op + 2 > n_out) return -1;
            const int pos = D.op & 31;
            B.bp += e >> 24; B.norm();
            if (pos == 31) {
                if (lane == 31) D.mine = e >> 8;
                D.op++;
                flush_group(D, out, lane);
                if (lane == 0) D.mine = e >> 16;
                D.op++;
                if (!CAREFUL && D.op + FAST_MARGIN > n_out) return 1;
            } else {
                if (lane == pos) D.mine = e >> 8;
                if (lane == pos + 1) D.mine = e >> 16;
                D.op += 2;
                if ((D.op & 31) == 0) {
                    flush_group(D, out, lane);
                    if (!CAREFUL && D.op + FAST_MARGIN > n_out) return 1;
                }
            }
            continue;
        }
        if (t == 0) {                                             // a code longer than the table's index (or none)
            int nb;
            const int s = slow_decode(w, T.count_lit, T.sym_lit, nb);
            if (s < 0) return -1;
            e = lit_entry(s, nb);
            t = e & 0x70u;
        }
        if (t == 0x10u) {
            if (CAREFUL && D.op >= n_out) return -1;
            if (lane == (D.op & 31)) D.mine = e >> 8;
            D.op++;
            B.bp += e & 15u; B.norm();
            if ((D.op & 31) == 0) {
                flush_group(D, out, lane);
                if (!CAREFUL && D.op + FAST_MARGIN > n_out) return 1;
            }
        } else if (t == 0x30u) {
            const int n1 = e & 15, ex = (e >> 8) & 15;
            const int len = (int)(e >> 16) + (int)((w >> n1) & ((1u << ex) - 1u));
            B.bp += n1 + ex; B.norm();
            w = B.window();
            uint32_t d = T.dist[w & ((1u << DIST_ROOT) - 1u)];
            if (d == 0) {
                int nb;
                const int s = slow_decode(w, T.count_dist, T.sym_dist, nb);
                if (s < 0) return -1;
                d = dist_entry(s, nb);
            }
            if (d & 0x100u) return -1;
            const int nd = d & 15, dx = (d >> 4) & 15;
            const int dist = (int)(d >> 16) + (int)((w >> nd) & ((1u << dx) - 1u));
            B.bp += nd + dx; B.norm();
            if (dist > D.op || len > n_out - D.op) return -1;
            flush_pending(D, out, lane);
            __syncwarp();
            uint8_t* dst = out + D.op;
            if (dist >= 32 || dist >= len) {
                for (int b0 = 0; b0 < len; b0 += 32) {
                    const int k = b0 + lane;
                    if (k < len) dst[k] = dst[k - dist];
                    __syncwarp();
                }
            } else {                                               // the match overlaps itself inside one step: a pattern of period dist
                if (lane < len) dst[lane] = dst[lane % dist - dist];
                const int d2 = (31 / dist + 1) * dist;             // smallest multiple of the period >= 32
                for (int b0 = 32; b0 < len; b0 += 32) {
                    __syncwarp();
                    const int k = b0 + lane;
                    if (k < len) dst[k] = dst[k - d2];
                }
                __syncwarp();
            }
            D.op += len; D.pend = D.op;
            if (!CAREFUL && D.op + FAST_MARGIN > n_out) return 1;
        } else if (t == 0x40u) {
            B.bp += e & 15u; B.norm();
            return 0;
        } else {
            return -1;
        }
    }
}

// one DEFLATE stream of exactly n_out bytes by one warp; the result is uniform over the lanes. The compressed buffer
// is read in whole 32-bit words (its last word may reach up to 3 bytes past comp_bytes).
__device__ bool inflate_warp(const uint8_t* comp, int64_t comp_bytes, int64_t c_off, int c_len, uint8_t* __restrict__ out, const int n_out,
                             WarpTables& T, const int lane) {
    Dec D;
    Bits& B = D.B;
    const uint8_t* bytes;
    {
        const uintptr_t addr = (uintptr_t)(comp + c_off);
        const int head = (int)(addr & 3);
        bytes = (const uint8_t*)(addr - head);
        B.w = (const uint32_t*)bytes;
        const int64_t avail = (comp_bytes - c_off + head + 3) >> 2;
        B.n_words = (int)(avail < (1ll << 28) ? avail : (1ll << 28));
        B.seek(head);
        c_len += head;                                            // stream end in bytes from B.w
    }
    const int64_t end_bit = 8ll * c_len;
    D.op = 0; D.pend = 0; D.mine = 0;
    for (;;) {
        if (B.bitpos() > end_bit) return false;
        uint32_t w = B.window();
        const uint32_t fin = w & 1u, type = (w >> 1) & 3u;
        B.bp += 3; B.norm();
        if (type == 0) {                                          // stored
            B.bp = (B.bp + 7) & ~7; B.norm();
            w = B.window();
            const uint32_t len = w & 0xffffu, nlen = w >> 16;
            if ((len ^ 0xffffu) != nlen) return false;
            B.bp += 32; B.norm();
            const int p = (int)(B.bitpos() >> 3);
            if ((int64_t)p + len > c_len || (int)len > n_out - D.op) return false;
            flush_pending(D, out, lane);
            for (uint32_t k = lane; k < len; k += 32) out[D.op + k] = bytes[p + k];
            D.op += (int)len; D.pend = D.op;
            B.seek(p + (int)len);
        } else if (type == 3) {
            return false;
        } else {
            __syncwarp();
            if (type == 1) {
                for (int i = lane; i < 288; i += 32) T.lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
                T.lens[288 + lane] = 5;
                __syncwarp();
                if (!build_table<1>(T.lens, 288, T.count_lit, T.cur, T.base, T.sym_lit, T.lit, LIT_ROOT, lane)) return false;
                if (!build_table<2>(T.lens + 288, 32, T.count_dist, T.cur, T.base, T.sym_dist, T.dist, DIST_ROOT, lane)) return false;
            } else {
                w = B.window();
                const int hlit = (int)(w & 31u) + 257, hdist = (int)((w >> 5) & 31u) + 1, hclen = (int)((w >> 10) & 15u) + 4;
                B.bp += 14; B.norm();
                if (hlit > 286 || hdist > 30) return false;
                T.cl[lane] = 0;
                __syncwarp();
                for (int i = 0; i < hclen; i++) {
                    T.cl[CL_ORDER[i]] = (uint8_t)(B.window() & 7u);   // every lane stores the same byte
                    B.bp += 3; B.norm();
                }
                __syncwarp();
                if (!build_table<0>(T.cl, 19, T.count_dist, T.cur, T.base, T.sym_dist, T.dist, CL_ROOT, lane)) return false;
                const int total = hlit + hdist;
                int n = 0;
                while (n < total) {
                    w = B.window();
                    const uint32_t e = T.dist[w & ((1u << CL_ROOT) - 1u)];
                    if (!e) return false;
                    const int nb = e & 15, s = (int)(e >> 4);
                    if (s < 16) { T.lens[n++] = (uint8_t)s; B.bp += nb; B.norm(); continue; }
                    int rep; uint8_t v = 0;
                    if (s == 16) { if (!n) return false; v = T.lens[n - 1]; rep = 3 + (int)((w >> nb) & 3u); B.bp += nb + 2; }
                    else if (s == 17) { rep = 3 + (int)((w >> nb) & 7u); B.bp += nb + 3; }
                    else { rep = 11 + (int)((w >> nb) & 127u); B.bp += nb + 7; }
                    B.norm();
                    if (n + rep > total) return false;
                    for (int k = lane; k < rep; k += 32) T.lens[n + k] = v;
                    n += rep;
                    __syncwarp();
                }
                __syncwarp();
                if (!T.lens[256]) return false;
                if (!build_table<1>(T.lens, hlit, T.count_lit, T.cur, T.base, T.sym_lit, T.lit, LIT_ROOT, lane)) return false;
                if (!build_table<2>(T.lens + hlit, hdist, T.count_dist, T.cur, T.base, T.sym_dist, T.dist, DIST_ROOT, lane)) return false;
            }
            int r = 1;
            if (D.op + FAST_MARGIN <= n_out) r = decode_symbols<false>(D, T, out, n_out, lane);
            if (r == 1) r = decode_symbols<true>(D, T, out, n_out, lane);
            if (r < 0) return false;
        }
        if (fin) break;
    }
    flush_pending(D, out, lane);
    __syncwarp();
    return D.op == n_out && B.bitpos() <= end_bit;
}

}  // namespace winf
