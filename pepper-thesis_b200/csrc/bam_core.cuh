// Device-side BAM decoding ("ingest that can feed the GPU"): the pieces a thread needs to turn one
// BGZF block into bytes and one BAM record into the clipped read BAM_handler::get_reads returns
// (/root/reference/pepper_variant/modules/cpp/bam_handler.cpp:115-451). Everything here is `__host__ __device__` so the
// same code is exercised on the CPU by tests/test_bam_core_cpu.py (through csrc/bam_core_host.cpp) before it ever runs
// on a GPU; the product path only calls it from the kernels of ingest_gpu.cu.
//
//   inflate_block   RFC 1951 (stored / fixed / dynamic) for ONE BGZF payload of <= 64 KiB, canonical Huffman decoding
//                   with a 9-bit first-level table per code and a bit-serial walk for longer codes; ~2 KB of state.
//                   The single-thread form: the first device decoder, now the CPU-testable reference of the
//                   warp-cooperative inflate_warp.cuh the kernels use
//   crc32_update    the block's CRC-32 (BGZF trailer), table-driven
//   parse_record    fixed part + field offsets of a BAM record, long-CIGAR (CG:B,I) convention of SAMv1 4.2.2
//   clip_walk       the cut of one record to [start, stop] of bam_handler.cpp:178-306 in closed form per op
//   parse_hp        the aux walk for the HP tag, bam_handler.cpp:313-421
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PV_HD __host__ __device__ __forceinline__
#define PV_HDN __host__ __device__
#else
#define PV_HD inline
#define PV_HDN inline
#endif

namespace bamcore {

PV_HD uint32_t ld16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
// 32-bit little-endian load at any alignment. On the device: the aligned word(s) that hold the four bytes and a funnel
// shift (one or two transactions instead of four); the second word is only touched when the value really reaches into it,
// so nothing behind p + 3 is read.
PV_HD uint32_t ld32(const uint8_t* p) {
#if defined(__CUDA_ARCH__)
    const uintptr_t a = (uintptr_t)p;
    const uint32_t* q = (const uint32_t*)(a & ~(uintptr_t)3);
    const int sh = (int)(a & 3) * 8;
    const uint32_t lo = q[0], hi = sh ? q[1] : 0u;
    return __funnelshift_r(lo, hi, sh);
#else
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
#endif
}

// ---- DEFLATE ---------------------------------------------------------------------------------------------------
constexpr int FAST_BITS = 9;
struct Huff {
    uint16_t count[16];        // codes per length
    uint16_t symbol[288];      // symbols in canonical order
    uint16_t fast[1 << FAST_BITS];   // (symbol << 4) | length for codes of <= FAST_BITS bits, 0 = longer code (or none)
};

// canonical code from lens[0..n): returns false when the code is over-subscribed (an incomplete code is accepted, as
// zlib accepts the single distance code some encoders emit)
PV_HDN bool huff_build(Huff& h, const uint8_t* lens, int n) {
    for (int l = 0; l < 16; l++) h.count[l] = 0;
    for (int s = 0; s < n; s++) h.count[lens[s]]++;
    int left = 1;
    for (int l = 1; l <= 15; l++) { left = (left << 1) - (int)h.count[l]; if (left < 0) return false; }
    uint16_t offs[16];
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = (uint16_t)(offs[l] + h.count[l]);
    for (int s = 0; s < n; s++) if (lens[s]) h.symbol[offs[lens[s]]++] = (uint16_t)s;
    for (int i = 0; i < (1 << FAST_BITS); i++) h.fast[i] = 0;
    // first-level table: canonical codes in order, bit-reversed (DEFLATE packs codes MSB first into an LSB-first stream)
    uint32_t code = 0; int idx = 0;
    for (int l = 1; l <= FAST_BITS; l++) {
        for (int k = 0; k < (int)h.count[l]; k++, idx++, code++) {
            uint32_t r = 0;
            for (int b = 0; b < l; b++) r |= ((code >> b) & 1u) << (l - 1 - b);
            const uint16_t e = (uint16_t)((h.symbol[idx] << 4) | l);
            for (uint32_t i = r; i < (1u << FAST_BITS); i += 1u << l) h.fast[i] = e;
        }
        code <<= 1;
    }
    h.count[0] = 0;
    return true;
}

struct BitReader {
    const uint8_t* in; int64_t n_in; int64_t ip; uint64_t bb; int bc; bool overrun;
    PV_HD void init(const uint8_t* p, int64_t n) { in = p; n_in = n; ip = 0; bb = 0; bc = 0; overrun = false; }
    PV_HD void refill() {                                   // at least 32 valid bits behind this (zeros past the end)
        while (bc <= 56) {
            const uint64_t b = ip < n_in ? in[ip] : 0;
            ip++;
            bb |= b << bc; bc += 8;
        }
    }
    PV_HD uint32_t peek(int n) const { return (uint32_t)(bb & ((1ull << n) - 1ull)); }
    PV_HD void drop(int n) { bb >>= n; bc -= n; }
    PV_HD uint32_t take(int n) { const uint32_t v = peek(n); drop(n); return v; }
    // bytes consumed so far (whole bytes still in the buffer are not consumed)
    PV_HD int64_t consumed() const { return ip - (bc >> 3); }
};

// one symbol; -1 on an invalid code. Needs >= 15 bits in the reader.
PV_HD int huff_decode(BitReader& br, const Huff& h) {
    const uint16_t e = h.fast[br.peek(FAST_BITS)];
    if (e) { br.drop(e & 15); return e >> 4; }
    int code = 0, first = 0, index = 0;
    uint32_t bits = br.peek(15);
    for (int len = 1; len <= 15; len++) {
        code |= (int)(bits & 1u); bits >>= 1;
        const int count = h.count[len];
        if (code - count < first) { br.drop(len); return h.symbol[index + (code - first)]; }
        index += count; first += count; first <<= 1; code <<= 1;
    }
    return -1;
}

struct InflateState { Huff lit, dist; uint8_t lens[320]; };

// raw DEFLATE stream of exactly n_out bytes; false on any malformed input / size mismatch
PV_HDN bool inflate_block(const uint8_t* in, int64_t n_in, uint8_t* out, int64_t n_out, InflateState& S) {
    const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    const uint8_t ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    BitReader br; br.init(in, n_in);
    int64_t op = 0;
    for (;;) {
        br.refill();
        const uint32_t fin = br.take(1), type = br.take(2);
        if (type == 0) {                                        // stored
            br.drop(br.bc & 7);
            int64_t p = br.consumed();
            if (n_in - p < 4) return false;
            const uint32_t len = ld16(in + p), nlen = ld16(in + p + 2);
            p += 4;
            if ((len ^ 0xffffu) != nlen || n_in - p < (int64_t)len || n_out - op < (int64_t)len) return false;
            for (uint32_t i = 0; i < len; i++) out[op + i] = in[p + i];
            op += len; p += len;
            br.ip = p; br.bb = 0; br.bc = 0;
        } else if (type == 3) {
            return false;
        } else {
            if (type == 1) {
                int i = 0;
                for (; i < 144; i++) S.lens[i] = 8;
                for (; i < 256; i++) S.lens[i] = 9;
                for (; i < 280; i++) S.lens[i] = 7;
                for (; i < 288; i++) S.lens[i] = 8;
                if (!huff_build(S.lit, S.lens, 288)) return false;
                for (i = 0; i < 30; i++) S.lens[i] = 5;
                if (!huff_build(S.dist, S.lens, 30)) return false;
            } else {
                const int hlit = (int)br.take(5) + 257, hdist = (int)br.take(5) + 1, hclen = (int)br.take(4) + 4;
                if (hlit > 286 || hdist > 30) return false;
                uint8_t cl[19];
                for (int i = 0; i < 19; i++) cl[i] = 0;
                for (int i = 0; i < hclen; i++) { if (br.bc < 3) br.refill(); cl[ORDER[i]] = (uint8_t)br.take(3); }
                if (!huff_build(S.dist, cl, 19)) return false;  // the code-length code borrows the distance table
                int n = 0;
                while (n < hlit + hdist) {
                    br.refill();
                    const int s = huff_decode(br, S.dist);
                    if (s < 0 || s > 18) return false;
                    if (s < 16) { S.lens[n++] = (uint8_t)s; continue; }
                    int rep; uint8_t v = 0;
                    if (s == 16) { if (!n) return false; v = S.lens[n - 1]; rep = 3 + (int)br.take(2); }
                    else if (s == 17) rep = 3 + (int)br.take(3);
                    else rep = 11 + (int)br.take(7);
                    if (n + rep > hlit + hdist) return false;
                    while (rep--) S.lens[n++] = v;
                }
                if (!S.lens[256]) return false;
                if (!huff_build(S.lit, S.lens, hlit)) return false;
                if (!huff_build(S.dist, S.lens + hlit, hdist)) return false;
            }
            for (;;) {
                br.refill();                                     // >= 57 bits: code 15 + extra 5 + code 15 + extra 13 = 48
                if (br.consumed() > n_in) return false;
                int s = huff_decode(br, S.lit);
                if (s < 0) return false;
                if (s < 256) {
                    if (op >= n_out) return false;
                    out[op++] = (uint8_t)s;
                    continue;
                }
                if (s == 256) break;
                s -= 257;
                if (s >= 29) return false;
                const uint32_t length = LEN_BASE[s] + br.take(LEN_EXTRA[s]);
                const int d = huff_decode(br, S.dist);
                if (d < 0 || d >= 30) return false;
                const uint32_t distance = DIST_BASE[d] + br.take(DIST_EXTRA[d]);
                if ((int64_t)distance > op || (int64_t)length > n_out - op) return false;
                for (uint32_t i = 0; i < length; i++) out[op + i] = out[op + i - distance];
                op += length;
            }
        }
        if (fin) break;
    }
    return op == n_out && br.consumed() <= n_in;
}

// CRC-32 (IEEE, reflected) as in the gzip trailer; table[256] is built by crc32_table
PV_HDN void crc32_table(uint32_t* t) {
    for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1; t[i] = c; }
}
PV_HD uint32_t crc32_bytes(const uint32_t* t, const uint8_t* p, int64_t n) {
    uint32_t c = 0xffffffffu;
    for (int64_t i = 0; i < n; i++) c = t[(c ^ p[i]) & 0xffu] ^ (c >> 8);
    return c ^ 0xffffffffu;
}

// ---- BAM records -----------------------------------------------------------------------------------------------
struct RecHdr {
    int32_t tid; int64_t pos; int l_name, mapq, flag; int64_t l_seq; int64_t rec_end;   // rec_end: offset behind the record
    int64_t name_off, seq_off, qual_off, aux_off;      // offsets in the inflated stream
    int64_t ops_off; int32_t n_ops;                    // the record's CIGAR (inside the CG tag for long CIGARs)
    bool ok;
};

PV_HD int aux_size(uint8_t t) {                        // HtslibAuxSize, bam_handler.cpp:58-70
    switch (t) { case 'A': case 'c': case 'C': return 1; case 's': case 'S': return 2; case 'f': case 'i': case 'I': return 4; default: return -1; }
}

// offset of the CG:B,I payload (first op word) and its op count; false when the aux block has none / is malformed
PV_HDN bool find_long_cigar(const uint8_t* U, int64_t s, int64_t end, int64_t& ops_off, int32_t& n_ops) {
    while (end - s >= 4) {
        const bool is_cg = U[s] == 'C' && U[s + 1] == 'G';
        const uint8_t t = U[s + 2];
        s += 3;
        switch (t) {
            case 'A': case 'c': case 'C': if (end - s < 1) return false; s += 1; break;
            case 's': case 'S': if (end - s < 2) return false; s += 2; break;
            case 'i': case 'I': case 'f': if (end - s < 4) return false; s += 4; break;
            case 'Z': case 'H':
                while (s < end && U[s]) ++s;
                if (s >= end) return false;
                ++s;
                break;
            case 'B': {
                if (end - s < 5) return false;
                const uint8_t st = U[s];
                const int es = aux_size(st);
                if (es < 0) return false;
                const uint32_t n = ld32(U + s + 1);
                if ((uint64_t)n * (uint64_t)es > (uint64_t)(end - s - 5)) return false;
                if (is_cg && st == 'I') { ops_off = s + 5; n_ops = (int32_t)n; return true; }
                s += 5 + (int64_t)n * es;
            } break;
            default: return false;
        }
    }
    return false;
}

// HP tag of the aux block [s, end), walking it like bam_handler.cpp:313-421 (stops at the first malformed tag)
PV_HDN int parse_hp(const uint8_t* U, int64_t s, int64_t end) {
    int hp = 0;
    while (end - s >= 4) {
        const bool is_hp = U[s] == 'H' && U[s + 1] == 'P';
        const uint8_t t = U[s + 2];
        s += 3;
        switch (t) {
            case 'A': s += 1; break;
            case 'c': case 'C': case 's': case 'S': case 'i': case 'I': {
                const int sz = aux_size(t);
                if (end - s < sz) return hp;
                int64_t v = 0;
                if (t == 'c') v = (int8_t)U[s]; else if (t == 'C') v = U[s];
                else if (t == 's') v = (int16_t)ld16(U + s); else if (t == 'S') v = ld16(U + s);
                else if (t == 'i') v = (int32_t)ld32(U + s); else v = ld32(U + s);
                if (is_hp) hp = (int)v;
                s += sz;
            } break;
            case 'f': if (end - s < 4) return hp; s += 4; break;
            case 'Z': case 'H': { while (s < end && U[s]) ++s; if (s >= end) return hp; ++s; } break;
            case 'B': {
                if (end - s < 5) return hp;
                const int es = aux_size(U[s]);
                if (es < 0) return hp;
                const uint32_t n = ld32(U + s + 1);
                s += 5 + (int64_t)n * es;
                if (s > end) return hp;
            } break;
            default: return hp;
        }
    }
    return hp;
}

// record whose 4-byte block_size sits at `off`; u_size = bytes of the inflated stream
PV_HDN RecHdr parse_record(const uint8_t* U, int64_t off, int64_t u_size) {
    RecHdr h;
    h.ok = false; h.n_ops = 0; h.ops_off = 0;
    if (off + 36 > u_size) return h;
    const uint32_t bs = ld32(U + off);
    const uint8_t* r = U + off + 4;
    h.rec_end = off + 4 + (int64_t)bs;
    if (bs < 32 || h.rec_end > u_size) return h;
    h.tid = (int32_t)ld32(r);
    h.pos = (int32_t)ld32(r + 4);
    h.l_name = r[8]; h.mapq = r[9];
    const int n_cig = (int)ld16(r + 12);
    h.flag = (int)ld16(r + 14);
    h.l_seq = (int32_t)ld32(r + 16);
    h.name_off = off + 36;
    const int64_t cig_off = h.name_off + h.l_name;
    h.seq_off = cig_off + 4 * (int64_t)n_cig;
    if (h.l_seq < 0) return h;
    h.qual_off = h.seq_off + (h.l_seq + 1) / 2;
    h.aux_off = h.qual_off + h.l_seq;
    if (h.aux_off > h.rec_end) return h;
    h.ops_off = cig_off; h.n_ops = n_cig;
    if (n_cig == 2) {
        const uint32_t c0 = ld32(U + cig_off), c1 = ld32(U + cig_off + 4);
        if ((c0 & 15u) == 4 && (int64_t)(c0 >> 4) == h.l_seq && (c1 & 15u) == 3) {
            int64_t oo; int32_t nn;
            if (find_long_cigar(U, h.aux_off, h.rec_end, oo, nn)) { h.ops_off = oo; h.n_ops = nn; }
        }
    }
    h.ok = true;
    return h;
}

// bam_endpos as the iterator uses it: pos + reference length, pos + 1 for unmapped / zero-length alignments
PV_HDN int64_t record_endpos(const uint8_t* U, const RecHdr& h) {
    int64_t rlen = 0;
    for (int k = 0; k < h.n_ops; k++) {
        const uint32_t w = ld32(U + h.ops_off + 4 * (int64_t)k);
        const int op = (int)(w & 15u);
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += w >> 4;
    }
    return h.pos + (((h.flag & 4) || rlen == 0) ? 1 : rlen);
}

// flag / mapq filters of bam_handler.cpp:137-150
PV_HD bool record_passes(const RecHdr& h, int include_supplementary, int min_mapq) {
    if ((h.flag & 0x200) || (h.flag & 0x400) || (h.flag & 0x100) || (h.flag & 0x4)) return false;
    if (!include_supplementary && (h.flag & 0x800)) return false;
    return h.mapq >= min_mapq;
}

struct Clip {
    int64_t pos_start, pos_end;    // type_read.pos / pos_end
    int64_t idx0;                  // first kept read index (the kept bases are ONE run of read indices)
    int64_t n_bases; int32_t n_ops;
    int32_t k_first, k_last;       // first / last kept op of the record's CIGAR (every kept op lies between them, whole)
    int32_t first_kept, last_kept; // kept lengths of those two ops (the only ones a cut can shorten)
    bool bad;                      // the CIGAR walks past SEQ (the reference reads out of bounds there): read dropped
    bool split;                    // the kept bases were not one contiguous run (cannot happen for a well-formed record)
};

// BAM_handler::get_reads' cut of one record to [start, stop] (clipping inclusive of stop), bam_handler.cpp:178-306.
// WRITE: kept ops go to ops_out[0 .. n_ops).
template <bool WRITE>
PV_HDN Clip clip_walk(const uint8_t* U, const RecHdr& h, int64_t start, int64_t stop, uint32_t* ops_out) {
    Clip c;
    c.pos_start = -1; c.pos_end = -1; c.idx0 = -1; c.n_bases = 0; c.n_ops = 0; c.bad = false; c.split = false;
    c.k_first = c.k_last = -1; c.first_kept = c.last_kept = 0;
    int64_t cur_pos = h.pos, cur_idx = 0;
    for (int k = 0; k < h.n_ops; k++) {
        const uint32_t w = ld32(U + h.ops_off + 4 * (int64_t)k);
        const int op = (int)(w & 15u);
        const int64_t len = w >> 4;
        if (cur_pos > stop) break;                               // :186-188
        int64_t kept = 0, kept_idx = -1;
        switch (op) {
            case 0: case 7: case 8: {                            // :190-229
                int64_t i0 = 0;
                if (cur_pos < start) { i0 = start - cur_pos < len ? start - cur_pos : len; cur_idx += i0; cur_pos += i0; }
                int64_t take = len - i0 < stop - cur_pos + 1 ? len - i0 : stop - cur_pos + 1;
                if (take < 0) take = 0;
                if (take > 0) {
                    if (c.pos_start == -1) { c.pos_start = cur_pos; c.pos_end = cur_pos; }
                    if (cur_idx + take > h.l_seq) { c.bad = true; break; }
                    kept_idx = cur_idx;
                    cur_idx += take; cur_pos += take; c.pos_end += take; kept = take;
                }
            } break;
            case 4: case 1:                                      // :230-262
                if (cur_pos >= start && cur_pos <= stop && c.pos_start != -1) {
                    if (cur_idx + len > h.l_seq) { c.bad = true; break; }
                    kept_idx = cur_idx; kept = len;
                }
                cur_idx += len;
                break;
            case 3: case 2:                                      // :263-291
                if (cur_pos >= start && cur_pos <= stop && c.pos_start != -1) {
                    const int64_t take = len < stop - cur_pos + 1 ? len : stop - cur_pos + 1;
                    kept = take; c.pos_end += take; cur_pos += take;
                } else {
                    cur_pos += len;
                }
                break;
            default: break;                                      // hard clip, pad, back: ignored (:300-303)
        }
        if (c.bad) break;
        if (kept > 0) {
            if (kept_idx >= 0) {
                if (c.idx0 < 0) c.idx0 = kept_idx;
                else if (c.idx0 + c.n_bases != kept_idx) c.split = true;
                c.n_bases += kept;
            }
            if (WRITE) ops_out[c.n_ops] = (uint32_t)(kept << 4) | (uint32_t)op;
            if (c.n_ops == 0) { c.k_first = k; c.first_kept = (int32_t)kept; }
            c.k_last = k; c.last_kept = (int32_t)kept;
            c.n_ops++;
        }
    }
    return c;
}

// base `i` of the record as the upper-case ASCII byte get_reads stores (seq_nt16_str, :213)
PV_HD uint8_t record_base(const uint8_t* U, const RecHdr& h, int64_t i) {
    const char NT16[17] = "=ACMGRSVTWYHKDBN";
    const uint8_t b = U[h.seq_off + (i >> 1)];
    return (uint8_t)NT16[(i & 1) ? (b & 15) : (b >> 4)];
}

}  // namespace bamcore
