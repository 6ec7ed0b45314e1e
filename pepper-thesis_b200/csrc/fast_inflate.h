// Raw DEFLATE (RFC 1951) decoder for BGZF blocks, written for what BAM payloads are made of: long runs of literals
// (4-bit packed bases and base qualities barely compress), which zlib's inflate decodes at ~110 MB/s on this class of
// host. A BGZF block is at most 64 KiB with its size and CRC-32 known up front, so the decoder works on whole buffers:
//   * 64-bit bit buffer refilled with one unaligned 8-byte load (the caller pads the input with 16 readable bytes);
//   * canonical Huffman codes in two-level tables (11-bit primary table for literals/lengths, 8-bit for distances);
//   * a literal costs one table lookup, one shift and one byte store; up to three literals go out per refill.
// Every write is bounds-checked. The caller verifies the block's CRC-32 and exact size afterwards and falls back to zlib
// when this decoder reports an error or the check fails, so a bug here costs time, never correctness.
#pragma once
#include <cstdint>
#include <cstring>

namespace fastinf {

constexpr int LIT_PB = 11, DIST_PB = 8;
constexpr uint32_t F_LIT = 1u << 15, F_EOB = 1u << 14, F_SUB = 1u << 13, F_BAD = 1u << 12, F_LIT2 = 1u << 9;
// entry: bits 0..3 = code length to consume (sub-table pointer: primary bits), 4..8 = extra bits (length/distance symbols),
// flags, bits 16..31 = literal byte / base value / sub-table offset. F_LIT2: TWO literals whose codes fit the primary index
// together (bits 16..23 the first, 24..31 the second, length = both codes): quality-like data has ~5-bit codes, so most
// lookups emit two bytes and the load -> shift -> load dependency chain is paid once per pair.

struct Tables {
    uint32_t lit[(1 << LIT_PB) + 320 * 16];        // primary + sub-tables (sum of sub-table sizes is bounded by the codes above PB)
    uint32_t dist[(1 << DIST_PB) + 32 * 128];
};

static const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

inline uint32_t rev_bits(uint32_t v, int n) {
    uint32_t r = 0;
    for (int i = 0; i < n; i++) { r = (r << 1) | (v & 1u); v >>= 1; }
    return r;
}

inline uint32_t lit_entry(int sym, int len) {
    if (sym < 256) return ((uint32_t)sym << 16) | F_LIT | (uint32_t)len;
    if (sym == 256) return F_EOB | (uint32_t)len;
    if (sym > 285) return F_BAD | (uint32_t)len;
    return ((uint32_t)LEN_BASE[sym - 257] << 16) | ((uint32_t)LEN_EXTRA[sym - 257] << 4) | (uint32_t)len;
}
inline uint32_t dist_entry(int sym, int len) {
    if (sym > 29) return F_BAD | (uint32_t)len;
    return ((uint32_t)DIST_BASE[sym] << 16) | ((uint32_t)DIST_EXTRA[sym] << 4) | (uint32_t)len;
}

// canonical code of `n` symbols with lengths lens[] -> two-level table; false when the code is over-subscribed or the
// sub-tables would not fit. An incomplete code is accepted (unused entries decode as F_BAD), as zlib accepts the single
// distance code some encoders emit.
template <bool LIT>
inline bool build(const uint8_t* lens, int n, uint32_t* tab, int pb, int cap) {
    int count[16] = {0};
    for (int i = 0; i < n; i++) count[lens[i]]++;
    count[0] = 0;
    uint32_t next[16]; uint32_t code = 0; int64_t left = 1;
    for (int l = 1; l <= 15; l++) {
        left = (left << 1) - count[l];
        if (left < 0) return false;
        code = (code + (uint32_t)count[l - 1]) << 1;
        next[l] = code;
    }
    const int psize = 1 << pb;
    for (int i = 0; i < psize; i++) tab[i] = F_BAD | 1u;
    // sub-table sizes: longest code behind every primary prefix
    uint8_t sub_bits[1 << 11];
    memset(sub_bits, 0, (size_t)psize);
    {
        uint32_t nx[16];
        memcpy(nx, next, sizeof(nx));
        for (int s = 0; s < n; s++) {
            const int l = lens[s];
            if (l <= pb) { if (l) nx[l]++; continue; }
            const uint32_t r = rev_bits(nx[l]++, l);
            uint8_t& sb = sub_bits[r & (uint32_t)(psize - 1)];
            if (l - pb > sb) sb = (uint8_t)(l - pb);
        }
    }
    int used = psize;
    for (int i = 0; i < psize; i++) {
        if (!sub_bits[i]) continue;
        const int sz = 1 << sub_bits[i];
        if (used + sz > cap) return false;
        tab[i] = ((uint32_t)used << 16) | F_SUB | ((uint32_t)sub_bits[i] << 4) | (uint32_t)pb;
        for (int k = 0; k < sz; k++) tab[used + k] = F_BAD | 1u;
        used += sz;
    }
    for (int s = 0; s < n; s++) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t r = rev_bits(next[l]++, l);
        if (l <= pb) {
            const uint32_t e = LIT ? lit_entry(s, l) : dist_entry(s, l);
            for (uint32_t i = r; i < (uint32_t)psize; i += 1u << l) tab[i] = e;
        } else {
            const uint32_t p = tab[r & (uint32_t)(psize - 1)];
            const int sb = (int)((p >> 4) & 31u);
            uint32_t* sub = tab + (p >> 16);
            const uint32_t e = LIT ? lit_entry(s, l - pb) : dist_entry(s, l - pb);
            for (uint32_t i = r >> pb; i < (1u << sb); i += 1u << (l - pb)) sub[i] = e;
        }
    }
    return true;
}

// second pass over the primary literal table: entry i whose first symbol is a literal of l1 bits and whose remaining
// LIT_PB - l1 index bits already hold a complete second literal code becomes a pair entry
inline void pair_literals(uint32_t* tab) {
    uint32_t single[1 << LIT_PB];
    memcpy(single, tab, sizeof(single));
    for (uint32_t i = 0; i < (1u << LIT_PB); i++) {
        const uint32_t e = single[i];
        if (!(e & F_LIT)) continue;
        const uint32_t l1 = e & 15u;
        const uint32_t e2 = single[i >> l1];                  // the unknown high bits read as zero: valid iff the code is short enough
        if ((e2 & F_LIT) && (e2 & 15u) + l1 <= (uint32_t)LIT_PB)
            tab[i] = (e & 0x00ff0000u) | ((e2 & 0x00ff0000u) << 8) | F_LIT | F_LIT2 | (l1 + (e2 & 15u));
    }
}

inline uint64_t load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

// in: n_in bytes of raw DEFLATE, with at least 16 more readable bytes behind them. out: exactly n_out bytes expected.
// Returns true when the stream ended cleanly with exactly n_out bytes written.
inline bool inflate_raw(const uint8_t* in, size_t n_in, uint8_t* out, size_t n_out, Tables& T) {
    const uint8_t* ip = in;
    const uint8_t* const in_end = in + n_in;
    uint8_t* op = out;
    uint8_t* const out_end = out + n_out;
    uint64_t bb = 0; int bc = 0;
#define FI_REFILL() do { bb |= load64(ip) << bc; ip += (63 - bc) >> 3; bc |= 56; } while (0)
#define FI_BITS(n) ((uint32_t)(bb & ((1ull << (n)) - 1ull)))
#define FI_DROP(n) do { bb >>= (n); bc -= (n); } while (0)
    // ip runs up to 7 bytes ahead of the bits consumed so far (ip - (bc >> 3)); that position must stay inside the input, which
    // also keeps every 8-byte load inside the 16 bytes of padding
    for (;;) {
        if (ip - (bc >> 3) > in_end) return false;
        FI_REFILL();
        const uint32_t final = FI_BITS(1), type = (uint32_t)((bb >> 1) & 3u);
        FI_DROP(3);
        if (type == 0) {                                                    // stored
            FI_DROP(bc & 7);
            // the bytes still in the bit buffer belong to the stream: step the input pointer back over them
            ip -= bc >> 3; bb = 0; bc = 0;
            if (in_end - ip < 4) return false;
            const uint32_t len = ip[0] | ((uint32_t)ip[1] << 8), nlen = ip[2] | ((uint32_t)ip[3] << 8);
            ip += 4;
            if ((len ^ 0xffffu) != nlen || (size_t)(in_end - ip) < len || (size_t)(out_end - op) < len) return false;
            memcpy(op, ip, len); op += len; ip += len;
        } else {
            if (type == 3) return false;
            uint8_t lens[320];
            if (type == 1) {                                                // fixed code
                int i = 0;
                for (; i < 144; i++) lens[i] = 8;
                for (; i < 256; i++) lens[i] = 9;
                for (; i < 280; i++) lens[i] = 7;
                for (; i < 288; i++) lens[i] = 8;
                if (!build<true>(lens, 288, T.lit, LIT_PB, (int)(sizeof(T.lit) / 4))) return false;
                pair_literals(T.lit);
                for (i = 0; i < 32; i++) lens[i] = 5;
                if (!build<false>(lens, 32, T.dist, DIST_PB, (int)(sizeof(T.dist) / 4))) return false;
            } else {                                                        // dynamic code
                const int hlit = (int)FI_BITS(5) + 257; FI_DROP(5);
                const int hdist = (int)FI_BITS(5) + 1; FI_DROP(5);
                const int hclen = (int)FI_BITS(4) + 4; FI_DROP(4);
                if (hlit > 286 || hdist > 30) return false;
                static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                uint8_t cl[19] = {0};
                if (ip - (bc >> 3) > in_end) return false;
                FI_REFILL();
                for (int i = 0; i < hclen; i++) { cl[order[i]] = (uint8_t)FI_BITS(3); FI_DROP(3); if (bc < 8) { if (ip - (bc >> 3) > in_end) return false; FI_REFILL(); } }
                // the code-length code: at most 7 bits, one flat table
                uint16_t clt[128];
                {
                    int count[8] = {0}; uint32_t next[8]; uint32_t code = 0; int left = 1;
                    for (int i = 0; i < 19; i++) count[cl[i]]++;
                    count[0] = 0;
                    for (int l = 1; l <= 7; l++) { left = (left << 1) - count[l]; if (left < 0) return false; code = (code + (uint32_t)count[l - 1]) << 1; next[l] = code; }
                    for (int i = 0; i < 128; i++) clt[i] = 0;
                    for (int s = 0; s < 19; s++) {
                        const int l = cl[s];
                        if (!l) continue;
                        const uint32_t r = rev_bits(next[l]++, l);
                        for (uint32_t i = r; i < 128; i += 1u << l) clt[i] = (uint16_t)((s << 4) | l);
                    }
                }
                int n = 0;
                while (n < hlit + hdist) {
                    if (ip - (bc >> 3) > in_end) return false;
                    FI_REFILL();
                    const uint16_t e = clt[FI_BITS(7)];
                    if (!e) return false;
                    FI_DROP(e & 15);
                    const int s = e >> 4;
                    if (s < 16) { lens[n++] = (uint8_t)s; continue; }
                    int rep; uint8_t v = 0;
                    if (s == 16) { if (!n) return false; v = lens[n - 1]; rep = 3 + (int)FI_BITS(2); FI_DROP(2); }
                    else if (s == 17) { rep = 3 + (int)FI_BITS(3); FI_DROP(3); }
                    else { rep = 11 + (int)FI_BITS(7); FI_DROP(7); }
                    if (n + rep > hlit + hdist) return false;
                    while (rep--) lens[n++] = v;
                }
                if (!lens[256]) return false;
                uint8_t dl[32];
                memcpy(dl, lens + hlit, (size_t)hdist);
                if (!build<true>(lens, hlit, T.lit, LIT_PB, (int)(sizeof(T.lit) / 4))) return false;
                pair_literals(T.lit);
                if (!build<false>(dl, hdist, T.dist, DIST_PB, (int)(sizeof(T.dist) / 4))) return false;
            }
            // ---- the symbols of the block ----
            for (;;) {
                if (ip - (bc >> 3) > in_end) return false;
                FI_REFILL();                                                // >= 56 bits: a literal/length code (<= 15) + its
                uint32_t e = T.lit[FI_BITS(LIT_PB)];                        // extra bits (<= 5) + a distance code (<= 15) + 13
#define FI_EMIT(e) do { \
                    if (e & F_LIT2) { if (out_end - op < 2) return false; op[0] = (uint8_t)(e >> 16); op[1] = (uint8_t)(e >> 24); op += 2; } \
                    else { if (op >= out_end) return false; *op++ = (uint8_t)(e >> 16); } \
                    FI_DROP(e & 15u); } while (0)
                if (e & F_LIT) {                                            // fast path: up to three lookups (six literals) per refill
                    FI_EMIT(e);
                    e = T.lit[FI_BITS(LIT_PB)];
                    if (e & F_LIT) {
                        FI_EMIT(e);
                        e = T.lit[FI_BITS(LIT_PB)];
                        if (e & F_LIT) {
                            FI_EMIT(e);
                            continue;
                        }
                    }
                    if (bc < 48) continue;                                  // not enough bits left for a whole match: start over
                }
                if (e & F_SUB) {
                    FI_DROP(e & 15u);
                    e = T.lit[(e >> 16) + FI_BITS((e >> 4) & 31u)];
                    if (e & F_LIT) {
                        if (op >= out_end) return false;
                        FI_DROP(e & 15u); *op++ = (uint8_t)(e >> 16);
                        continue;
                    }
                }
                if (e & (F_EOB | F_BAD)) {
                    if (e & F_BAD) return false;
                    FI_DROP(e & 15u);
                    break;
                }
                FI_DROP(e & 15u);
                const int xb = (int)((e >> 4) & 31u);
                const uint32_t length = (e >> 16) + FI_BITS(xb);
                FI_DROP(xb);
                if (bc < 32) { if (ip - (bc >> 3) > in_end) return false; FI_REFILL(); }
                uint32_t d = T.dist[FI_BITS(DIST_PB)];
                if (d & F_SUB) { FI_DROP(d & 15u); d = T.dist[(d >> 16) + FI_BITS((d >> 4) & 31u)]; }
                if (d & F_BAD) return false;
                FI_DROP(d & 15u);
                const int dxb = (int)((d >> 4) & 31u);
                const uint32_t distance = (d >> 16) + FI_BITS(dxb);
                FI_DROP(dxb);
                if (distance > (size_t)(op - out) || length > (size_t)(out_end - op)) return false;
                const uint8_t* src = op - distance;
                if (distance >= 8 && (size_t)(out_end - op) >= (size_t)length + 8) {
                    uint8_t* const stop = op + length;
                    do { memcpy(op, src, 8); op += 8; src += 8; } while (op < stop);
                    op = stop;
                } else {
                    for (uint32_t i = 0; i < length; i++) op[i] = src[i];
                    op += length;
                }
            }
        }
        if (final) break;
    }
#undef FI_EMIT
#undef FI_REFILL
#undef FI_BITS
#undef FI_DROP
    return op == out_end && ip - (bc >> 3) <= in_end;
}

}  // namespace fastinf
