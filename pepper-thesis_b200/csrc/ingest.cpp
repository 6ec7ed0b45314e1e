// libpv_ingest.so: BAM / FASTA ingest without htslib -- BGZF blocks inflated with zlib, BAI bin + linear index query,
// .fai random access -- emitting the packed SoA read batch (include/pepper_b200.h) directly instead of the
// reference's AoS type_read strings. Semantics follow BAM_handler::get_reads
// (/root/reference/pepper_variant/modules/cpp/bam_handler.cpp:115-444) and FASTA_handler
// (/root/reference/pepper_variant/modules/cpp/fasta_handler.cpp:31-51); see include/pepper_ingest.h.
//
// BAM / BGZF / BAI layouts are those of the SAM specification (SAMv1 sections 4.1, 4.2, 5.2); the region iterator
// reproduces htslib's sam_itr_queryi contract: records of the contig with pos < end and end-position > beg.
#include "pepper_ingest.h"
#include "fast_inflate.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <string>
#include <thread>
#include <memory>
#include <vector>

namespace {

thread_local char g_err[512];
int fail(int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return code;
}

inline uint16_t le16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint64_t le64(const uint8_t* p) { return (uint64_t)le32(p) | ((uint64_t)le32(p + 4) << 32); }

bool read_file(const std::string& path, std::vector<uint8_t>& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    const bool ok = n <= 0 || fread(out.data(), 1, (size_t)n, f) == (size_t)n;
    fclose(f);
    return ok;
}

// ---- BGZF ----------------------------------------------------------------------------------------------------
// One reader per thread over a shared file descriptor (pread): holds the current inflated block.
std::atomic<uint64_t> g_fast_blocks{0};        // diagnostics: blocks decoded by fast_inflate.h (the rest went through zlib)
const bool g_fast_inflate = []() { const char* v = getenv("PV_INGEST_ZLIB_ONLY"); return !(v && atoi(v)); }();
std::atomic<uint64_t> g_inflated_bytes{0};     // diagnostics: BGZF payload bytes inflated since the library was loaded

struct BgzfReader {
    int fd = -1;
    int64_t file_size = 0;
    int64_t block_coff = -1;         // compressed offset of the block in `data`
    int64_t next_coff = 0;           // compressed offset of the following block
    std::vector<uint8_t> data;       // inflated block
    std::vector<uint8_t> raw;
    size_t upos = 0;
    z_stream zs;
    bool zs_init = false;
    std::vector<fastinf::Tables> tables;   // one element, on the heap (46 KB)

    ~BgzfReader() { if (zs_init) inflateEnd(&zs); }

    // 0 ok, 1 end of file, <0 error
    int load_block(int64_t coff) {
        if (coff >= file_size) { block_coff = coff; next_coff = coff; data.clear(); upos = 0; return 1; }
        uint8_t hdr[18];
        if (pread(fd, hdr, 18, coff) != 18) return fail(PV_EINVAL, "BGZF: short read of a block header at %lld", (long long)coff);
        if (hdr[0] != 31 || hdr[1] != 139 || hdr[2] != 8 || !(hdr[3] & 4)) return fail(PV_EINVAL, "BGZF: bad block magic at %lld", (long long)coff);
        const int xlen = le16(hdr + 10);
        // the BC subfield is normally the first (and only) one; walk the extra field when it is not
        int bsize = -1;
        if (xlen == 6 && hdr[12] == 'B' && hdr[13] == 'C') bsize = le16(hdr + 16);
        else {
            std::vector<uint8_t> extra(xlen);
            if (pread(fd, extra.data(), xlen, coff + 12) != xlen) return fail(PV_EINVAL, "BGZF: short extra field");
            for (int o = 0; o + 4 <= xlen;) {
                const int sl = le16(&extra[o + 2]);
                if (extra[o] == 'B' && extra[o + 1] == 'C' && sl == 2 && o + 6 <= xlen) { bsize = le16(&extra[o + 4]); break; }
                o += 4 + sl;
            }
        }
        if (bsize < 0) return fail(PV_EINVAL, "BGZF: block without BC subfield at %lld", (long long)coff);
        const int total = bsize + 1;
        const int cdata_off = 12 + xlen, cdata_len = total - cdata_off - 8;
        if (cdata_len < 0) return fail(PV_EINVAL, "BGZF: bad block size");
        raw.resize((size_t)total + 16);                       // fast_inflate.h reads up to 16 bytes behind the stream
        if (pread(fd, raw.data(), total, coff) != total) return fail(PV_EINVAL, "BGZF: truncated block at %lld", (long long)coff);
        const uint32_t isize = le32(&raw[total - 4]);
        const uint32_t want_crc = le32(&raw[total - 8]);
        // SAMv1 4.1: a BGZF block holds at most 64 KiB of data; a larger ISIZE is a corrupt (or hostile) trailer, not a
        // reason to allocate up to 4 GiB per reader thread
        if (isize > 65536u) return fail(PV_EINVAL, "BGZF: block at %lld claims %u bytes of data (limit 65536)", (long long)coff, isize);
        data.resize(isize);
        bool done = false;
        if (isize && g_fast_inflate) {
            // own decoder first (written for literal-heavy BAM payloads); zlib below takes over whenever it gives up or the
            // block's CRC-32 does not confirm its output
            if (tables.empty()) tables.resize(1);
            uint8_t tail[16];
            memcpy(tail, raw.data() + total - 8, 8);         // the trailer sits right behind the stream: keep it, zero the padding
            memset(raw.data() + total - 8, 0, 24);
            done = fastinf::inflate_raw(raw.data() + cdata_off, (size_t)cdata_len, data.data(), isize, tables[0]) &&
                   (uint32_t)crc32(crc32(0L, Z_NULL, 0), data.data(), isize) == want_crc;
            memcpy(raw.data() + total - 8, tail, 8);
            if (done) g_fast_blocks.fetch_add(1, std::memory_order_relaxed);
        }
        if (isize && !done) {
            if (!zs_init) { memset(&zs, 0, sizeof(zs)); if (inflateInit2(&zs, -15) != Z_OK) return fail(PV_ENOMEM, "inflateInit2"); zs_init = true; }
            else inflateReset(&zs);
            zs.next_in = raw.data() + cdata_off; zs.avail_in = (uInt)cdata_len;
            zs.next_out = data.data(); zs.avail_out = isize;
            const int rc = inflate(&zs, Z_FINISH);
            if (rc != Z_STREAM_END || zs.avail_out != 0) return fail(PV_EINVAL, "BGZF: inflate failed (%d) at %lld", rc, (long long)coff);
            if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), data.data(), isize) != want_crc)
                return fail(PV_EINVAL, "BGZF: CRC mismatch at %lld", (long long)coff);
        }
        block_coff = coff; next_coff = coff + total; upos = 0;
        g_inflated_bytes.fetch_add(isize, std::memory_order_relaxed);
        return 0;
    }
    int seek(uint64_t voff) {
        const int64_t coff = (int64_t)(voff >> 16);
        if (coff != block_coff) { const int rc = load_block(coff); if (rc < 0) return rc; }
        upos = (size_t)(voff & 0xffff);
        return 0;
    }
    uint64_t tell() const {
        if (upos >= data.size() && block_coff >= 0) return (uint64_t)next_coff << 16;   // htslib reports the next block's start
        return ((uint64_t)block_coff << 16) | (uint64_t)upos;
    }
    // reads n bytes; returns n, 0 at clean EOF (nothing read), <0 on error / truncated data
    int64_t read(uint8_t* dst, int64_t n) {
        int64_t got = 0;
        while (got < n) {
            if (upos >= data.size()) {
                const int rc = load_block(next_coff);
                if (rc < 0) return rc;
                if (rc == 1) break;
                continue;                                  // empty blocks (EOF marker) are skipped
            }
            const size_t take = std::min<size_t>((size_t)(n - got), data.size() - upos);
            memcpy(dst + got, data.data() + upos, take);
            upos += take; got += (int64_t)take;
        }
        if (got != 0 && got != n) return fail(PV_EINVAL, "BGZF: truncated record");
        return got;
    }
};

// ---- BAI -----------------------------------------------------------------------------------------------------
struct Chunk { uint64_t beg, end; };
struct BaiRef {
    std::vector<std::pair<uint32_t, std::vector<Chunk>>> bins;   // sorted by bin id
    std::vector<uint64_t> linear;
};

inline void reg2bins(int64_t beg, int64_t end, std::vector<uint32_t>& out) {   // SAMv1 5.3, 14-bit linear / 5 levels
    out.clear();
    if (end <= beg) return;
    --end;
    if (end >= (1ll << 29)) end = (1ll << 29) - 1;
    if (beg >= (1ll << 29)) return;
    out.push_back(0);
    for (int64_t k = 1 + (beg >> 26); k <= 1 + (end >> 26); ++k) out.push_back((uint32_t)k);
    for (int64_t k = 9 + (beg >> 23); k <= 9 + (end >> 23); ++k) out.push_back((uint32_t)k);
    for (int64_t k = 73 + (beg >> 20); k <= 73 + (end >> 20); ++k) out.push_back((uint32_t)k);
    for (int64_t k = 585 + (beg >> 17); k <= 585 + (end >> 17); ++k) out.push_back((uint32_t)k);
    for (int64_t k = 4681 + (beg >> 14); k <= 4681 + (end >> 14); ++k) out.push_back((uint32_t)k);
}

}  // namespace

struct PvBamFile {
    int fd = -1;
    int64_t file_size = 0;
    std::string text;
    std::vector<std::string> names;
    std::vector<int64_t> lens;
    std::vector<BaiRef> index;
};

struct PvFastaFile {
    int fd = -1;
    struct Seq { std::string name; int64_t len, offset, linebases, linewidth; };
    std::vector<Seq> seqs;
    const Seq* find(const char* n) const { for (const Seq& s : seqs) if (s.name == n) return &s; return nullptr; }
};

struct PvIngestBatch {
    std::vector<int64_t> read_pos, read_base_off, read_cigar_off, read_pos_end;
    std::vector<int32_t> read_len, read_n_ops, hp;
    std::vector<uint16_t> bam_flag;
    std::vector<uint8_t> read_flags, read_mapq, bases, quals, ref;
    std::vector<uint32_t> cigar;
    std::vector<int64_t> region_ref_start, region_ref_end, region_cand_start, region_cand_end, region_ref_off, region_ref_len,
        region_read_begin;
    std::string names;
    int min_qual = 255;            // smallest quality of any read base seen (-> PvReadBatch.min_qual)
    int64_t n_reads() const { return (int64_t)read_pos.size(); }
};

namespace {

static const char NT16[] = "=ACMGRSVTWYHKDBN";

// packed 4-bit bases [start, start + count) of a BAM record -> ASCII, two bases per table lookup
struct Nt16Pairs {
    uint16_t t[256];
    Nt16Pairs() { for (int i = 0; i < 256; i++) t[i] = (uint16_t)((uint8_t)NT16[i >> 4] | ((uint16_t)(uint8_t)NT16[i & 15] << 8)); }
};
static const Nt16Pairs NT16_PAIRS;
inline void decode_bases(const uint8_t* packed, int64_t start, int64_t count, std::vector<uint8_t>& out) {
    if (count <= 0) return;
    const size_t at = out.size();
    out.resize(at + (size_t)count);
    uint8_t* dst = out.data() + at;
    int64_t i = start, end = start + count;
    if (i & 1) { *dst++ = (uint8_t)NT16[packed[i >> 1] & 15]; i++; }
    for (; i + 1 < end; i += 2) { const uint16_t v = NT16_PAIRS.t[packed[i >> 1]]; memcpy(dst, &v, 2); dst += 2; }
    if (i < end) *dst = (uint8_t)NT16[packed[i >> 1] >> 4];
}

// the reads of one span, appended in the packed layout (bases / quals padded to 16 bytes per read)
struct ReadSink {
    PvIngestBatch* b;
    void add(int64_t pos_start, int64_t pos_end, const std::vector<uint8_t>& seq, const std::vector<uint8_t>& q,
             const std::vector<uint32_t>& ops, int flag, int mapq, int hp, const char* qname) {
        const bool reverse = (flag & 0x10) != 0;
        b->bam_flag.push_back((uint16_t)flag);
        b->read_pos.push_back(pos_start);
        b->read_pos_end.push_back(pos_end);
        b->read_base_off.push_back((int64_t)b->bases.size());
        b->read_len.push_back((int32_t)seq.size());
        b->read_cigar_off.push_back((int64_t)b->cigar.size());
        b->read_n_ops.push_back((int32_t)ops.size());
        b->read_flags.push_back(reverse ? 1 : 0);
        b->read_mapq.push_back((uint8_t)(mapq < 0 ? 0 : mapq > 255 ? 255 : mapq));
        b->hp.push_back(hp);
        b->bases.insert(b->bases.end(), seq.begin(), seq.end());
        b->quals.insert(b->quals.end(), q.begin(), q.end());
        int mq = b->min_qual;
        for (uint8_t v : q) mq = v < mq ? v : mq;
        b->min_qual = mq;
        const size_t pad = (16 - (seq.size() & 15)) & 15;
        b->bases.insert(b->bases.end(), pad, 0);
        b->quals.insert(b->quals.end(), pad, 0);
        b->cigar.insert(b->cigar.end(), ops.begin(), ops.end());
        b->names.append(qname); b->names.push_back('\0');
    }
};

int aux_size(uint8_t t) {                 // HtslibAuxSize, bam_handler.cpp:58-70
    switch (t) { case 'A': case 'c': case 'C': return 1; case 's': case 'S': return 2; case 'f': case 'i': case 'I': return 4; default: return -1; }
}

// HP tag of the aux block, walking it like bam_handler.cpp:313-421 (stops at the first malformed tag)
int parse_hp(const uint8_t* s, const uint8_t* end) {
    int hp = 0;
    while (end - s >= 4) {
        const bool is_hp = s[0] == 'H' && s[1] == 'P';
        const uint8_t t = s[2];
        s += 3;
        switch (t) {
            case 'A': s += 1; break;
            case 'c': case 'C': case 's': case 'S': case 'i': case 'I': {
                const int sz = aux_size(t);
                if (end - s < sz) return hp;
                int64_t v = 0;
                if (t == 'c') v = (int8_t)s[0]; else if (t == 'C') v = s[0];
                else if (t == 's') v = (int16_t)le16(s); else if (t == 'S') v = le16(s);
                else if (t == 'i') v = (int32_t)le32(s); else v = le32(s);
                if (is_hp) hp = (int)v;
                s += sz;
            } break;
            case 'f': if (end - s < 4) return hp; s += 4; break;
            case 'Z': case 'H': { while (s < end && *s) ++s; if (s >= end) return hp; ++s; } break;
            case 'B': {
                if (end - s < 5) return hp;
                const int es = aux_size(s[0]);
                if (es < 0) return hp;
                const uint32_t n = le32(s + 1);
                s += 5 + (size_t)n * es;
                if (s > end) return hp;
            } break;
            default: return hp;
        }
    }
    return hp;
}

// real CIGAR of a record whose op count did not fit 16 bits: "<l_seq>S<ref_len>N" placeholder + CG:B,I tag (SAMv1 4.2.2)
bool long_cigar(const uint8_t* aux, const uint8_t* end, std::vector<uint32_t>& out) {
    const uint8_t* s = aux;
    while (end - s >= 4) {
        const bool is_cg = s[0] == 'C' && s[1] == 'G';
        const uint8_t t = s[2];
        s += 3;
        switch (t) {                                          // every advance is checked against the end of the aux block first
            case 'A': case 'c': case 'C': if (end - s < 1) return false; s += 1; break;
            case 's': case 'S': if (end - s < 2) return false; s += 2; break;
            case 'i': case 'I': case 'f': if (end - s < 4) return false; s += 4; break;
            case 'Z': case 'H':
                while (s < end && *s) ++s;
                if (s >= end) return false;                   // unterminated string
                ++s;
                break;
            case 'B': {
                if (end - s < 5) return false;
                const uint8_t st = s[0];
                const int es = aux_size(st);
                if (es < 0) return false;
                const uint32_t n = le32(s + 1);
                if ((uint64_t)n * (uint64_t)es > (uint64_t)(end - s - 5)) return false;
                if (is_cg && st == 'I') { out.resize(n); for (uint32_t i = 0; i < n; i++) out[i] = le32(s + 5 + 4 * (size_t)i); return true; }
                s += 5 + (size_t)n * es;
            } break;
            default: return false;
        }
    }
    return false;
}

// merged BAI chunks that can hold a record overlapping [start, stop): bins overlapping the span, not before the
// linear-index lower bound (what htslib's iterator reads)
std::vector<Chunk> query_chunks(const BaiRef& ref, int64_t start, int64_t stop) {
    std::vector<uint32_t> bins;
    reg2bins(start, stop, bins);
    uint64_t min_off = 0;
    if (!ref.linear.empty()) {
        const size_t w = (size_t)(start >> 14);
        min_off = w < ref.linear.size() ? ref.linear[w] : ref.linear.back();
    }
    std::vector<Chunk> chunks;
    for (uint32_t bin : bins) {
        auto it = std::lower_bound(ref.bins.begin(), ref.bins.end(), bin,
                                   [](const std::pair<uint32_t, std::vector<Chunk>>& a, uint32_t v) { return a.first < v; });
        if (it == ref.bins.end() || it->first != bin) continue;
        // no record that overlaps the window of `start` begins before min_off (that is what the linear index stores), so a
        // chunk of a coarse bin that straddles it is entered there instead of at its own beginning
        for (const Chunk& c : it->second) if (c.end > min_off) chunks.push_back(Chunk{std::max(c.beg, min_off), c.end});
    }
    std::sort(chunks.begin(), chunks.end(), [](const Chunk& a, const Chunk& b) { return a.beg < b.beg; });
    std::vector<Chunk> merged;
    for (const Chunk& c : chunks) {
        if (!merged.empty() && c.beg <= merged.back().end) merged.back().end = std::max(merged.back().end, c.end);
        else merged.push_back(c);
    }
    return merged;
}

// BAM_handler::get_reads for spans [start_j, stop_j] (clipping inclusive of stop, iterator half-open like htslib), the
// spans ascending by start. ONE pass over the records of their union: every BGZF block is inflated and every record
// parsed once, then cut to each span it overlaps (a per-span index query re-inflates the blocks around every shared
// boundary and everything the coarse bins of long reads drag in: 1.4x the file for 100 kbp spans of 12 kbp reads).
// Each span receives exactly the records its own htslib iterator would return, in file order.
struct Span { int64_t start, stop; };
int collect_reads_multi(const PvBamFile& f, int tid, const std::vector<Span>& spans, const PvIngestOptions& o, std::vector<ReadSink>& sinks) {
    if (tid < 0 || tid >= (int)f.index.size() || spans.empty()) return PV_OK;
    const BaiRef& ref = f.index[tid];
    int64_t start = spans[0].start, stop = spans[0].stop;      // the union (spans are ascending by start)
    for (const Span& sp : spans) stop = std::max(stop, sp.stop);
    size_t j_lo = 0;                                           // spans before j_lo end at or before every later record
    const std::vector<Chunk> merged = query_chunks(ref, start, stop);
    if (merged.empty()) return PV_OK;

    BgzfReader rd;
    rd.fd = f.fd; rd.file_size = f.file_size;
    std::vector<uint8_t> rec;
    std::vector<uint8_t> seq, quals; std::vector<uint32_t> ops, cg, kept;
    std::vector<std::pair<int64_t, int64_t>> runs;           // (first read index, length) of the kept bases
    for (const Chunk& ch : merged) {
        if (int rc = rd.seek(ch.beg)) return rc;
        while (rd.tell() < ch.end) {
            uint8_t szb[4];
            const int64_t g = rd.read(szb, 4);
            if (g < 0) return (int)g;
            if (g == 0) break;
            const uint32_t bs = le32(szb);
            if (bs < 32) return fail(PV_EINVAL, "BAM: record smaller than its fixed part");
            rec.resize(bs);
            const int64_t g2 = rd.read(rec.data(), bs);
            if (g2 != (int64_t)bs) return g2 < 0 ? (int)g2 : fail(PV_EINVAL, "BAM: truncated record");
            const uint8_t* r = rec.data();
            const int32_t rtid = (int32_t)le32(r);
            const int64_t pos = (int32_t)le32(r + 4);
            if (rtid != tid || pos >= stop) { if (rtid > tid || (rtid == tid && pos >= stop) || rtid < 0) goto done; continue; }
            {
                const int l_name = r[8];
                const int mapq = r[9];
                int n_cig = le16(r + 12);
                const int flag = le16(r + 14);
                const int64_t l_seq = (int32_t)le32(r + 16);
                const uint8_t* p = r + 32;
                const char* qname = (const char*)p;
                const uint8_t* cig_p = p + l_name;
                const uint8_t* seq_p = cig_p + 4 * (size_t)n_cig;
                const uint8_t* qual_p = seq_p + (l_seq + 1) / 2;
                const uint8_t* aux_p = qual_p + l_seq;
                const uint8_t* end_p = r + bs;
                if (aux_p > end_p || l_seq < 0) return fail(PV_EINVAL, "BAM: record fields overrun the record");
                const uint32_t* cig = nullptr;
                ops.resize(n_cig);
                for (int k = 0; k < n_cig; k++) ops[k] = le32(cig_p + 4 * k);
                if (n_cig == 2 && (ops[0] & 15u) == 4 && (int64_t)(ops[0] >> 4) == l_seq && (ops[1] & 15u) == 3 && long_cigar(aux_p, end_p, cg)) {
                    ops = cg; n_cig = (int)ops.size();
                }
                cig = ops.data();
                // end position on the reference (bam_endpos): iterator keeps records with endpos > start
                int64_t rlen = 0;
                for (int k = 0; k < n_cig; k++) { const int op = cig[k] & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += cig[k] >> 4; }
                const int64_t endpos = pos + ((flag & 4) || rlen == 0 ? 1 : rlen);
                if (endpos <= start) continue;
                while (j_lo < spans.size() && spans[j_lo].stop <= pos) j_lo++;
                // bam_handler.cpp:133-147
                if ((flag & 0x200) || (flag & 0x400) || (flag & 0x100) || (flag & 0x4)) continue;
                if (!o.include_supplementary && (flag & 0x800)) continue;
                if (mapq < o.min_mapq) continue;

                for (size_t j = j_lo; j < spans.size() && spans[j].start < endpos; j++) {
                const int64_t start = spans[j].start, stop = spans[j].stop;      // this span (shadows the union)
                if (pos >= stop || endpos <= start) continue;
                ReadSink& sink = sinks[j];
                // bam_handler.cpp:163-304: cut the read to [start, stop]
                // The kept bases form contiguous runs of read indices (one run for every well-formed record: once the first
                // aligned base inside the span is kept, every later read-consuming op up to `stop` is kept too), so the ops
                // only file index ranges here and bases / qualities are decoded once per run, not once per CIGAR op.
                seq.clear(); quals.clear(); kept.clear(); runs.clear();
                int64_t pos_start = -1, pos_end = -1, cur_pos = pos, cur_idx = 0;
                bool bad = false;                            // CIGAR walks past SEQ (e.g. SEQ '*'): the reference reads out of bounds there
                for (int k = 0; k < n_cig; k++) {
                    const int op = cig[k] & 15;
                    const int64_t len = cig[k] >> 4;
                    if (cur_pos > stop) break;
                    int64_t kept_len = 0;
                    switch (op) {
                        case 0: case 7: case 8: {
                            int64_t i0 = 0;
                            if (cur_pos < start) { i0 = std::min(start - cur_pos, len); cur_idx += i0; cur_pos += i0; }
                            const int64_t take = std::max<int64_t>(0, std::min(len - i0, stop - cur_pos + 1));
                            if (take > 0) {
                                if (pos_start == -1) { pos_start = cur_pos; pos_end = pos_start; }
                                if (cur_idx + take > l_seq) { bad = true; break; }
                                if (!runs.empty() && runs.back().first + runs.back().second == cur_idx) runs.back().second += take;
                                else runs.emplace_back(cur_idx, take);
                                cur_idx += take; cur_pos += take; pos_end += take; kept_len = take;
                            }
                            // the reference leaves the remaining bases of the op unconsumed (its loop breaks at the first
                            // position beyond stop); the next op is then rejected by `cur_pos > stop`
                        } break;
                        case 4: case 1:
                            if (cur_pos >= start && cur_pos <= stop && pos_start != -1) {
                                if (cur_idx + len > l_seq) { bad = true; break; }
                                if (len > 0) {
                                    if (!runs.empty() && runs.back().first + runs.back().second == cur_idx) runs.back().second += len;
                                    else runs.emplace_back(cur_idx, len);
                                }
                                kept_len = len;
                            }
                            cur_idx += len;
                            break;
                        case 3: case 2:
                            if (cur_pos >= start && cur_pos <= stop && pos_start != -1) {
                                const int64_t take = std::min(len, stop - cur_pos + 1);
                                kept_len = take; pos_end += take; cur_pos += take;
                            } else {
                                cur_pos += len;
                            }
                            break;
                        default: break;                       // hard clip, pad, back: ignored (bam_handler.cpp:300-303)
                    }
                    if (bad) break;
                    if (kept_len > 0) kept.push_back((uint32_t)(kept_len << 4) | (uint32_t)op);
                }
                if (!bad)
                    for (const auto& run : runs) {
                        decode_bases(seq_p, run.first, run.second, seq);
                        quals.insert(quals.end(), qual_p + run.first, qual_p + run.first + run.second);
                    }
                if (!bad && !seq.empty())
                    sink.add(pos_start, pos_end, seq, quals, kept, flag, mapq, parse_hp(aux_p, end_p), qname);
                }
            }
        }
    }
done:
    return PV_OK;
}

int collect_reads(const PvBamFile& f, int tid, int64_t start, int64_t stop, const PvIngestOptions& o, ReadSink& sink) {
    std::vector<Span> spans{Span{start, stop}};
    std::vector<ReadSink> sinks{sink};
    return collect_reads_multi(f, tid, spans, o, sinks);
}

int tid_of(const PvBamFile& f, const char* contig) {
    for (size_t i = 0; i < f.names.size(); i++) if (f.names[i] == contig) return (int)i;
    return -1;
}

void append_batch(PvIngestBatch& dst, const PvIngestBatch& src) {
    const int64_t b0 = (int64_t)dst.bases.size(), c0 = (int64_t)dst.cigar.size();
    dst.read_pos.insert(dst.read_pos.end(), src.read_pos.begin(), src.read_pos.end());
    dst.read_pos_end.insert(dst.read_pos_end.end(), src.read_pos_end.begin(), src.read_pos_end.end());
    for (int64_t v : src.read_base_off) dst.read_base_off.push_back(v + b0);
    for (int64_t v : src.read_cigar_off) dst.read_cigar_off.push_back(v + c0);
    dst.read_len.insert(dst.read_len.end(), src.read_len.begin(), src.read_len.end());
    dst.read_n_ops.insert(dst.read_n_ops.end(), src.read_n_ops.begin(), src.read_n_ops.end());
    dst.hp.insert(dst.hp.end(), src.hp.begin(), src.hp.end());
    dst.bam_flag.insert(dst.bam_flag.end(), src.bam_flag.begin(), src.bam_flag.end());
    dst.read_flags.insert(dst.read_flags.end(), src.read_flags.begin(), src.read_flags.end());
    dst.read_mapq.insert(dst.read_mapq.end(), src.read_mapq.begin(), src.read_mapq.end());
    dst.bases.insert(dst.bases.end(), src.bases.begin(), src.bases.end());
    dst.quals.insert(dst.quals.end(), src.quals.begin(), src.quals.end());
    if (src.min_qual < dst.min_qual) dst.min_qual = src.min_qual;
    dst.cigar.insert(dst.cigar.end(), src.cigar.begin(), src.cigar.end());
    dst.names.append(src.names);
}

}  // namespace

extern "C" const char* pv_ingest_last_error(void) { return g_err; }

extern "C" int pv_bam_open(const char* bam_path, const char* bai_path, PvBamFile** out) {
    if (!bam_path || !out) return fail(PV_EINVAL, "null argument");
    PvBamFile* f = new PvBamFile();
    f->fd = open(bam_path, O_RDONLY);
    if (f->fd < 0) { delete f; return fail(PV_EINVAL, "INVALID BAM FILE. PLEASE CHECK IF PATH IS CORRECT: %s", bam_path); }
    struct stat st;
    fstat(f->fd, &st);
    f->file_size = st.st_size;
    BgzfReader rd;
    rd.fd = f->fd; rd.file_size = f->file_size;
    auto bail = [&](int rc) { close(f->fd); delete f; return rc; };
    if (rd.load_block(0) != 0) return bail(fail(PV_EINVAL, "HEADER ERROR: INVALID BAM FILE. PLEASE CHECK IF PATH IS CORRECT: %s", bam_path));
    uint8_t h[8];
    if (rd.read(h, 8) != 8 || memcmp(h, "BAM\1", 4) != 0) return bail(fail(PV_EINVAL, "HEADER ERROR: not a BAM file: %s", bam_path));
    const uint32_t l_text = le32(h + 4);
    f->text.resize(l_text);
    if (l_text && rd.read((uint8_t*)&f->text[0], l_text) != (int64_t)l_text) return bail(fail(PV_EINVAL, "HEADER ERROR: truncated header text"));
    if (rd.read(h, 4) != 4) return bail(fail(PV_EINVAL, "HEADER ERROR: truncated header"));
    const uint32_t n_ref = le32(h);
    for (uint32_t i = 0; i < n_ref; i++) {
        if (rd.read(h, 4) != 4) return bail(fail(PV_EINVAL, "HEADER ERROR: truncated reference list"));
        const uint32_t l = le32(h);
        std::string name(l, '\0');
        if (l && rd.read((uint8_t*)&name[0], l) != (int64_t)l) return bail(fail(PV_EINVAL, "HEADER ERROR: truncated reference name"));
        if (!name.empty() && name.back() == '\0') name.pop_back();
        if (rd.read(h, 4) != 4) return bail(fail(PV_EINVAL, "HEADER ERROR: truncated reference length"));
        f->names.push_back(name);
        f->lens.push_back((int64_t)le32(h));
    }
    // index
    std::vector<uint8_t> bai;
    std::string p1 = bai_path ? std::string(bai_path) : std::string(bam_path) + ".bai";
    bool ok = read_file(p1, bai);
    if (!ok && !bai_path) {
        std::string p2(bam_path);
        if (p2.size() > 4 && p2.substr(p2.size() - 4) == ".bam") { p2 = p2.substr(0, p2.size() - 4) + ".bai"; ok = read_file(p2, bai); }
    }
    if (!ok || bai.size() < 8 || memcmp(bai.data(), "BAI\1", 4) != 0)
        return bail(fail(PV_EINVAL, "INVALID BAM INDEX FILE. PLEASE CHECK IF FILE IS INDEXED: %s", bam_path));
    const uint8_t* q = bai.data() + 4;
    const uint8_t* qe = bai.data() + bai.size();
    const uint32_t n_idx = le32(q); q += 4;
    f->index.resize(n_idx);
    for (uint32_t i = 0; i < n_idx; i++) {
        if (qe - q < 4) return bail(fail(PV_EINVAL, "BAI: truncated"));
        const uint32_t n_bin = le32(q); q += 4;
        BaiRef& br = f->index[i];
        for (uint32_t k = 0; k < n_bin; k++) {
            if (qe - q < 8) return bail(fail(PV_EINVAL, "BAI: truncated"));
            const uint32_t bin = le32(q); const uint32_t n_chunk = le32(q + 4); q += 8;
            if ((uint64_t)(qe - q) < (uint64_t)n_chunk * 16) return bail(fail(PV_EINVAL, "BAI: truncated"));
            std::vector<Chunk> cs(n_chunk);
            for (uint32_t c = 0; c < n_chunk; c++) { cs[c].beg = le64(q); cs[c].end = le64(q + 8); q += 16; }
            if (bin != 37450) br.bins.emplace_back(bin, std::move(cs));   // 37450 = metadata pseudo-bin
        }
        std::sort(br.bins.begin(), br.bins.end(), [](const std::pair<uint32_t, std::vector<Chunk>>& a, const std::pair<uint32_t, std::vector<Chunk>>& b) { return a.first < b.first; });
        if (qe - q < 4) return bail(fail(PV_EINVAL, "BAI: truncated"));
        const uint32_t n_intv = le32(q); q += 4;
        if ((uint64_t)(qe - q) < (uint64_t)n_intv * 8) return bail(fail(PV_EINVAL, "BAI: truncated"));
        br.linear.resize(n_intv);
        for (uint32_t k = 0; k < n_intv; k++) { br.linear[k] = le64(q); q += 8; }
    }
    *out = f;
    return PV_OK;
}

extern "C" void pv_bam_close(PvBamFile* f) { if (!f) return; if (f->fd >= 0) close(f->fd); delete f; }
extern "C" int32_t pv_bam_n_targets(const PvBamFile* f) { return f ? (int32_t)f->names.size() : 0; }
extern "C" const char* pv_bam_target_name(const PvBamFile* f, int32_t i) { return (f && i >= 0 && i < (int)f->names.size()) ? f->names[i].c_str() : nullptr; }
extern "C" int64_t pv_bam_target_len(const PvBamFile* f, int32_t i) { return (f && i >= 0 && i < (int)f->lens.size()) ? f->lens[i] : -1; }

extern "C" int64_t pv_bam_sample_names(const PvBamFile* f, char* out, int64_t out_cap) {
    if (!f) return 0;
    std::set<std::string> samples;                      // bam_handler.cpp:31-56
    size_t p = 0;
    while (p < f->text.size()) {
        size_t e = f->text.find('\n', p);
        if (e == std::string::npos) e = f->text.size();
        const std::string line = f->text.substr(p, e - p);
        p = e + 1;
        if (line.compare(0, 3, "@RG") != 0 || (line.size() > 3 && line[3] != '\t')) continue;
        size_t t = 0;
        while (t < line.size()) {
            size_t te = line.find('\t', t);
            if (te == std::string::npos) te = line.size();
            const std::string tok = line.substr(t, te - t);
            t = te + 1;
            if (tok.compare(0, 3, "SM:") == 0) {
                std::string v = tok.substr(3);
                const size_t c = v.find(':');
                if (c != std::string::npos) v = v.substr(0, c);     // getline(tag_tokenizer, sample_name, ':')
                samples.insert(v);
            }
        }
    }
    std::string joined;
    for (const std::string& s : samples) { if (!joined.empty()) joined.push_back('\n'); joined += s; }
    if (out && out_cap > 0) { const size_t n = std::min<size_t>(joined.size(), (size_t)out_cap - 1); memcpy(out, joined.data(), n); out[n] = 0; }
    return (int64_t)joined.size();
}

extern "C" int pv_fasta_open(const char* path, PvFastaFile** out) {
    if (!path || !out) return fail(PV_EINVAL, "null argument");
    std::vector<uint8_t> fai;
    if (!read_file(std::string(path) + ".fai", fai))
        return fail(PV_EINVAL, "INVALID FASTA FILE. PLEASE CHECK IF PATH IS CORRECT AND FILE IS INDEXED: %s", path);
    PvFastaFile* f = new PvFastaFile();
    f->fd = open(path, O_RDONLY);
    if (f->fd < 0) { delete f; return fail(PV_EINVAL, "INVALID FASTA FILE. PLEASE CHECK IF PATH IS CORRECT AND FILE IS INDEXED: %s", path); }
    std::string text(fai.begin(), fai.end());
    size_t p = 0;
    while (p < text.size()) {
        size_t e = text.find('\n', p);
        if (e == std::string::npos) e = text.size();
        const std::string line = text.substr(p, e - p);
        p = e + 1;
        if (line.empty()) continue;
        PvFastaFile::Seq s;
        char name[1024];
        long long len, off, lb, lw;
        if (sscanf(line.c_str(), "%1023[^\t]\t%lld\t%lld\t%lld\t%lld", name, &len, &off, &lb, &lw) != 5 || lb <= 0 || lw < lb) {
            close(f->fd); delete f; return fail(PV_EINVAL, "malformed .fai line: %s", line.c_str());
        }
        s.name = name; s.len = len; s.offset = off; s.linebases = lb; s.linewidth = lw;
        f->seqs.push_back(s);
    }
    *out = f;
    return PV_OK;
}
extern "C" void pv_fasta_close(PvFastaFile* f) { if (!f) return; if (f->fd >= 0) close(f->fd); delete f; }
extern "C" int32_t pv_fasta_n_seq(const PvFastaFile* f) { return f ? (int32_t)f->seqs.size() : 0; }
extern "C" const char* pv_fasta_seq_name(const PvFastaFile* f, int32_t i) { return (f && i >= 0 && i < (int)f->seqs.size()) ? f->seqs[i].name.c_str() : nullptr; }
extern "C" int64_t pv_fasta_seq_len(const PvFastaFile* f, const char* name) {
    if (!f || !name) return -1;
    const PvFastaFile::Seq* s = f->find(name);
    return s ? s->len : -1;
}

extern "C" int pv_fasta_fetch(const PvFastaFile* f, const char* name, int64_t start, int64_t stop, char* out, int64_t* out_len) {
    if (!f || !name || !out_len) return fail(PV_EINVAL, "null argument");
    *out_len = 0;
    const PvFastaFile::Seq* s = f->find(name);
    if (!s) return fail(PV_EINVAL, "CHROMOSOME NAME NOT PRESENT IN REFERENCE FASTA FILE: %s %lld %lld", name, (long long)start, (long long)stop);
    if (start < 0) start = 0;
    if (stop > s->len) stop = s->len;                    // faidx_fetch_seq clips the end to the sequence
    if (stop <= start) return PV_OK;
    if (!out) return fail(PV_EINVAL, "null output");
    const int64_t first = s->offset + (start / s->linebases) * s->linewidth + start % s->linebases;
    const int64_t last = s->offset + ((stop - 1) / s->linebases) * s->linewidth + (stop - 1) % s->linebases;
    const size_t raw_n = (size_t)(last - first + 1);
    const int64_t want = stop - start;
    int64_t n = 0;
    auto read_all = [&](uint8_t* dst, size_t bytes, int64_t at) -> bool {
        for (size_t got = 0; got < bytes;) {
            const ssize_t k = pread(f->fd, dst + got, bytes - got, at + (int64_t)got);
            if (k <= 0) return false;
            got += (size_t)k;
        }
        return true;
    };
    // line by line (the .fai promises linebases bases + a terminator of linewidth - linebases bytes per line), through a
    // buffer of whole lines that stays in cache: memcpy per line, then ONE long loop over the result for the upper-casing.
    // A terminator that is not made of line breaks, or a line break inside a line, sends the fetch down the generic loop.
    bool regular = s->linebases > 0 && s->linewidth > s->linebases;
    if (regular) {
        const int64_t term = s->linewidth - s->linebases;
        const int64_t lines = std::max<int64_t>(1, (1 << 20) / s->linewidth);
        std::unique_ptr<uint8_t[]> buf(new uint8_t[(size_t)(lines * s->linewidth)]);
        int64_t col = start % s->linebases, fpos = first;
        while (regular && n < want) {
            const int64_t bytes = std::min<int64_t>((s->linebases - col) + term + (lines - 1) * s->linewidth, last + 1 - fpos);
            if (!read_all(buf.get(), (size_t)bytes, fpos))
                return fail(PV_EINVAL, "ENCOUNTERED ERROR IN FETCHING REFERENCE FASTA FILE: %s %lld %lld", name, (long long)start, (long long)stop);
            int64_t j = 0;
            while (n < want && j < bytes) {
                const int64_t take = std::min<int64_t>(s->linebases - col, want - n);
                if (j + take > bytes) { regular = false; break; }
                memcpy(out + n, buf.get() + j, (size_t)take);
                n += take; j += take; col = 0;
                if (n < want) {
                    if (j + term > bytes) { regular = false; break; }
                    for (int64_t t = 0; t < term; t++, j++)
                        if (buf[(size_t)j] != '\n' && buf[(size_t)j] != '\r') regular = false;
                    if (!regular) break;
                }
            }
            fpos += bytes;
        }
        if (regular) {
            int any = 0;
            uint8_t* o = (uint8_t*)out;
            for (int64_t i = 0; i < n; i++) {                     // upper-casing (fasta_handler.cpp:49) and the line-break check
                const uint8_t ch = o[i];
                any |= (ch == '\n') | (ch == '\r');
                o[i] = (uint8_t)(ch - (((ch >= 'a') & (ch <= 'z')) << 5));
            }
            regular = !any;
        }
    }
    if (!regular) {
        std::unique_ptr<uint8_t[]> raw(new uint8_t[raw_n]);
        if (!read_all(raw.get(), raw_n, first))
            return fail(PV_EINVAL, "ENCOUNTERED ERROR IN FETCHING REFERENCE FASTA FILE: %s %lld %lld", name, (long long)start, (long long)stop);
        n = 0;
        for (size_t j = 0; j < raw_n; j++) {
            const uint8_t ch = raw[j];
            if (ch == '\n' || ch == '\r') continue;
            out[n++] = (char)((ch >= 'a' && ch <= 'z') ? ch - 32 : ch);   // fasta_handler.cpp:49
            if (n == want) break;
        }
    }
    *out_len = n;
    return PV_OK;
}

extern "C" int pv_bam_get_reads(PvBamFile* bam, const char* contig, int64_t start, int64_t stop, const PvIngestOptions* opt,
                                PvIngestBatch** out) {
    if (!bam || !contig || !opt || !out) return fail(PV_EINVAL, "null argument");
    PvIngestBatch* b = new PvIngestBatch();
    ReadSink sink{b};
    if (int rc = collect_reads(*bam, tid_of(*bam, contig), start, stop, *opt, sink)) { delete b; return rc; }
    b->region_read_begin = {0, b->n_reads()};
    *out = b;
    return PV_OK;
}

extern "C" int pv_ingest_regions(PvBamFile* bam, const PvFastaFile* fasta, const char* contig, int32_t n_regions,
                                 const int64_t* region_start, const int64_t* region_end, const PvIngestOptions* opt,
                                 PvIngestBatch** out) {
    if (!bam || !fasta || !contig || !opt || !out || (n_regions > 0 && (!region_start || !region_end))) return fail(PV_EINVAL, "null argument");
    const int tid = tid_of(*bam, contig);
    if (tid < 0) return fail(PV_EINVAL, "contig %s is not in the BAM header", contig);
    if (pv_fasta_seq_len(fasta, contig) < 0) return fail(PV_EINVAL, "CHROMOSOME NAME NOT PRESENT IN REFERENCE FASTA FILE: %s", contig);
    std::vector<PvIngestBatch> parts(n_regions);
    std::vector<std::string> refs(n_regions);
    std::vector<int> rcs(n_regions, PV_OK);
    std::vector<std::string> errs(n_regions);
    std::atomic<int> next(0);
    int nt = opt->threads > 0 ? opt->threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > n_regions) nt = n_regions > 0 ? n_regions : 1;
    // Work items = runs of consecutive intervals read in ONE pass over the BAM (collect_reads_multi): intervals must come
    // ascending and close to each other for that (a gap of more than 200 kbp starts a new item, so does an interval out of
    // order).
    std::vector<std::pair<int, int>> items;                    // [first, last) interval
    {
        // few intervals per thread: one item each (ceil(n / threads) intervals); many: about four items per thread
        const int per = n_regions < 8 * nt ? std::max(1, (n_regions + nt - 1) / nt) : n_regions / (4 * nt);
        int first = 0;
        for (int i = 1; i <= n_regions; i++) {
            const bool cut = i == n_regions || i - first >= per || region_start[i] < region_start[i - 1] ||
                             region_start[i] - region_end[i - 1] > 200000;
            if (cut) { items.emplace_back(first, i); first = i; }
        }
    }
    auto work = [&]() {
        for (;;) {
            const int it = next.fetch_add(1);
            if (it >= (int)items.size()) break;
            const int i0 = items[it].first, i1 = items[it].second;
            std::vector<Span> spans;
            std::vector<ReadSink> sinks;
            for (int i = i0; i < i1; i++) {
                spans.push_back(Span{std::max<int64_t>(0, region_start[i] - opt->safe_bases),   // AlignmentSummarizer.py:181-182
                                     region_end[i] + opt->safe_bases});
                sinks.push_back(ReadSink{&parts[i]});
            }
            const int rc = collect_reads_multi(*bam, tid, spans, *opt, sinks);
            if (rc) { rcs[i0] = rc; errs[i0] = g_err; continue; }
            for (int i = i0; i < i1; i++) {
                const int64_t rs = spans[i - i0].start, re = spans[i - i0].stop;
                std::string& ref = refs[i];
                ref.assign((size_t)(re + 1 - rs), 'N');       // region_end + 1 exclusive (:214-216); 'N' past the contig end
                int64_t got = 0;
                rcs[i] = pv_fasta_fetch(fasta, contig, rs, re + 1, &ref[0], &got);
                if (rcs[i]) errs[i] = g_err;
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; t++) pool.emplace_back(work);
    work();
    for (std::thread& t : pool) t.join();
    for (int i = 0; i < n_regions; i++) if (rcs[i]) return fail(rcs[i], "%s", errs[i].c_str());

    PvIngestBatch* b = new PvIngestBatch();
    b->region_read_begin.push_back(0);
    {   // one allocation per big array instead of a doubling chain (each step of which copies what is already there)
        size_t nb = 0, nc = 0, nr = 0, nf = 0, nn = 0;
        for (int i = 0; i < n_regions; i++) {
            nb += parts[i].bases.size(); nc += parts[i].cigar.size(); nr += parts[i].read_pos.size(); nf += refs[i].size(); nn += parts[i].names.size();
        }
        b->bases.reserve(nb); b->quals.reserve(nb); b->cigar.reserve(nc); b->ref.reserve(nf); b->names.reserve(nn);
        b->read_pos.reserve(nr); b->read_pos_end.reserve(nr); b->read_base_off.reserve(nr); b->read_cigar_off.reserve(nr);
        b->read_len.reserve(nr); b->read_n_ops.reserve(nr); b->hp.reserve(nr); b->bam_flag.reserve(nr);
        b->read_flags.reserve(nr); b->read_mapq.reserve(nr);
    }
    for (int i = 0; i < n_regions; i++) {
        const int64_t rs = std::max<int64_t>(0, region_start[i] - opt->safe_bases), re = region_end[i] + opt->safe_bases;
        append_batch(*b, parts[i]);
        parts[i] = PvIngestBatch();
        b->region_read_begin.push_back(b->n_reads());
        b->region_ref_start.push_back(rs); b->region_ref_end.push_back(re);
        b->region_cand_start.push_back(region_start[i]); b->region_cand_end.push_back(region_end[i]);
        b->region_ref_off.push_back((int64_t)b->ref.size());
        b->region_ref_len.push_back((int64_t)refs[i].size());
        b->ref.insert(b->ref.end(), refs[i].begin(), refs[i].end());
    }
    *out = b;
    return PV_OK;
}

// ---- plan for the device-side decode (pv_bam_inflate_blocks / pv_bam_index_records / pv_bam_clip_* of pepper_b200.h) --------
// The host's share of the GPU ingest: which compressed bytes to read (BAI query, as collect_reads_multi), where every BGZF
// block's DEFLATE payload sits in them, where its bytes land in the concatenated inflated stream, and where record chains
// can be entered (chunk starts + the linear index's per-16-kbp record offsets), so that the device can find all record
// boundaries with many short pointer chases instead of one long one.
struct PvBamPlan {
    const PvBamFile* bam = nullptr;
    int tid = -1;
    struct Range { int64_t file_off, bytes, buf_off; uint64_t vbeg, vend; };
    std::vector<Range> ranges;                    // one per merged chunk
    int64_t comp_bytes = 0;
    std::vector<PvBgzfBlock> blocks;
    std::vector<int64_t> block_file_off;          // file offset of every block (ascending)
    std::vector<int64_t> seg_begin, seg_end;
    int64_t inflated_bytes = 0;
};

extern "C" int pv_bam_plan(const PvBamFile* bam, const char* contig, int64_t start, int64_t stop, PvBamPlan** out) {
    if (!bam || !contig || !out) return fail(PV_EINVAL, "null argument");
    PvBamPlan* p = new PvBamPlan();
    p->bam = bam;
    p->tid = tid_of(*bam, contig);
    *out = p;
    if (p->tid < 0 || p->tid >= (int)bam->index.size() || stop <= start) return PV_OK;      // empty plan
    for (const Chunk& c : query_chunks(bam->index[p->tid], start, stop)) {
        PvBamPlan::Range r;
        r.vbeg = c.beg; r.vend = c.end;
        r.file_off = (int64_t)(c.beg >> 16);
        int64_t last = (int64_t)(c.end >> 16) + ((c.end & 0xffff) ? 65536 + 26 : 0);   // the end block is needed unless the chunk ends at its first byte
        if (last > bam->file_size) last = bam->file_size;
        r.bytes = last > r.file_off ? last - r.file_off : 0;
        r.buf_off = p->comp_bytes;
        p->comp_bytes += (r.bytes + 255) & ~(int64_t)255;
        p->ranges.push_back(r);
    }
    p->comp_bytes += 256;                         // the decoder may look a few bytes past a payload
    return PV_OK;
}
extern "C" int64_t pv_bam_plan_comp_bytes(const PvBamPlan* p) { return p ? p->comp_bytes : 0; }
extern "C" int32_t pv_bam_plan_tid(const PvBamPlan* p) { return p ? p->tid : -1; }

// reads the planned byte ranges into comp_out (pv_bam_plan_comp_bytes bytes; page-locked memory lets the upload run at
// PCIe speed) with `threads` readers, then walks the BGZF headers there
extern "C" int pv_bam_plan_load(PvBamPlan* p, uint8_t* comp_out, int32_t threads) {
    if (!p || (!comp_out && p->comp_bytes > 256)) return fail(PV_EINVAL, "null argument");
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    struct Piece { int64_t file_off, bytes, buf_off; };
    std::vector<Piece> pieces;
    for (const auto& r : p->ranges)
        for (int64_t o = 0; o < r.bytes; o += (8 << 20)) pieces.push_back(Piece{r.file_off + o, std::min<int64_t>(8 << 20, r.bytes - o), r.buf_off + o});
    std::atomic<size_t> next(0);
    std::atomic<int> bad(0);
    auto work = [&]() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= pieces.size()) break;
            int64_t got = 0;
            while (got < pieces[i].bytes) {
                const ssize_t n = pread(p->bam->fd, comp_out + pieces[i].buf_off + got, (size_t)(pieces[i].bytes - got), pieces[i].file_off + got);
                if (n <= 0) { bad = 1; return; }
                got += n;
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads && t < (int)pieces.size(); t++) pool.emplace_back(work);
    work();
    for (std::thread& t : pool) t.join();
    if (bad) return fail(PV_EINVAL, "BAM: short read of the planned byte ranges");

    p->blocks.clear(); p->block_file_off.clear(); p->seg_begin.clear(); p->seg_end.clear();
    int64_t u = 0;
    const std::vector<uint64_t>& linear = p->bam->index[p->tid].linear;
    for (const auto& r : p->ranges) {
        const size_t first_block = p->blocks.size();
        const int64_t end_coff = (int64_t)(r.vend >> 16);
        const bool end_block_needed = (r.vend & 0xffff) != 0;
        int64_t o = 0;
        while (o < r.bytes) {
            const int64_t coff = r.file_off + o;
            if (coff > end_coff || (coff == end_coff && !end_block_needed)) break;
            const uint8_t* h = comp_out + r.buf_off + o;
            if (r.bytes - o < 18 || h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) return fail(PV_EINVAL, "BGZF: bad block magic at %lld", (long long)coff);
            const int xlen = le16(h + 10);
            int bsize = -1;
            if (r.bytes - o < 12 + xlen) return fail(PV_EINVAL, "BGZF: truncated extra field at %lld", (long long)coff);
            for (int x = 0; x + 4 <= xlen;) {
                const int sl = le16(h + 12 + x + 2);
                if (h[12 + x] == 'B' && h[12 + x + 1] == 'C' && sl == 2 && x + 6 <= xlen) { bsize = le16(h + 12 + x + 4); break; }
                x += 4 + sl;
            }
            if (bsize < 0) return fail(PV_EINVAL, "BGZF: block without BC subfield at %lld", (long long)coff);
            const int total = bsize + 1, cdata_off = 12 + xlen, cdata_len = total - cdata_off - 8;
            if (cdata_len < 0 || r.bytes - o < total) return fail(PV_EINVAL, "BGZF: truncated block at %lld", (long long)coff);
            const uint32_t isize = le32(h + total - 4);
            if (isize > 65536u) return fail(PV_EINVAL, "BGZF: block at %lld claims %u bytes of data (limit 65536)", (long long)coff, isize);
            PvBgzfBlock b;
            b.c_off = r.buf_off + o + cdata_off; b.c_len = cdata_len; b.isize = (int32_t)isize; b.u_off = u; b.crc = le32(h + total - 8); b._pad = 0;
            p->blocks.push_back(b);
            p->block_file_off.push_back(coff);
            u += isize;
            o += total;
        }
        // virtual offset -> offset in the inflated stream (only offsets inside this range's blocks)
        auto to_u = [&](uint64_t v, int64_t& out_u) -> bool {
            const int64_t coff = (int64_t)(v >> 16);
            auto it = std::lower_bound(p->block_file_off.begin() + first_block, p->block_file_off.end(), coff);
            if (it == p->block_file_off.end() || *it != coff) {
                if (it == p->block_file_off.end() && (v & 0xffff) == 0 && !p->blocks.empty()) { out_u = u; return true; }   // right behind the last block
                return false;
            }
            const PvBgzfBlock& b = p->blocks[(size_t)(it - p->block_file_off.begin())];
            if ((int64_t)(v & 0xffff) > b.isize) return false;
            out_u = b.u_off + (int64_t)(v & 0xffff);
            return true;
        };
        int64_t ub, ue;
        if (!to_u(r.vbeg, ub) || !to_u(r.vend, ue)) return fail(PV_EINVAL, "BAI: a chunk boundary does not fall on a BGZF block of the file");
        std::vector<int64_t> starts{ub};
        for (uint64_t v : linear) {
            int64_t x;
            if (v > r.vbeg && v < r.vend && to_u(v, x) && x > starts.back() && x < ue) starts.push_back(x);
        }
        for (size_t i = 0; i < starts.size(); i++) {
            p->seg_begin.push_back(starts[i]);
            p->seg_end.push_back(i + 1 < starts.size() ? starts[i + 1] : ue);
        }
    }
    p->inflated_bytes = u;
    g_inflated_bytes.fetch_add((uint64_t)u, std::memory_order_relaxed);
    return PV_OK;
}
extern "C" int32_t pv_bam_plan_n_blocks(const PvBamPlan* p) { return p ? (int32_t)p->blocks.size() : 0; }
extern "C" int32_t pv_bam_plan_n_segments(const PvBamPlan* p) { return p ? (int32_t)p->seg_begin.size() : 0; }
extern "C" int64_t pv_bam_plan_inflated_bytes(const PvBamPlan* p) { return p ? p->inflated_bytes : 0; }
extern "C" int pv_bam_plan_tables(const PvBamPlan* p, PvBgzfBlock* blocks_out, int64_t* seg_begin_out, int64_t* seg_end_out) {
    if (!p) return fail(PV_EINVAL, "null argument");
    if (blocks_out && !p->blocks.empty()) memcpy(blocks_out, p->blocks.data(), p->blocks.size() * sizeof(PvBgzfBlock));
    if (seg_begin_out && !p->seg_begin.empty()) memcpy(seg_begin_out, p->seg_begin.data(), p->seg_begin.size() * 8);
    if (seg_end_out && !p->seg_end.empty()) memcpy(seg_end_out, p->seg_end.data(), p->seg_end.size() * 8);
    return PV_OK;
}
extern "C" void pv_bam_plan_free(PvBamPlan* p) { delete p; }

extern "C" int pv_ingest_view(const PvIngestBatch* b, PvReadBatch* v) {
    if (!b || !v) return fail(PV_EINVAL, "null argument");
    memset(v, 0, sizeof(*v));
    v->n_reads = b->n_reads(); v->n_bases = (int64_t)b->bases.size(); v->n_ops = (int64_t)b->cigar.size(); v->n_ref = (int64_t)b->ref.size();
    v->n_regions = (int32_t)b->region_ref_start.size();
    v->read_pos = b->read_pos.data(); v->read_base_off = b->read_base_off.data(); v->read_len = b->read_len.data();
    v->read_cigar_off = b->read_cigar_off.data(); v->read_n_ops = b->read_n_ops.data(); v->read_flags = b->read_flags.data();
    v->read_mapq = b->read_mapq.data(); v->bases = b->bases.data(); v->quals = b->quals.data(); v->cigar = b->cigar.data();
    v->region_ref_start = b->region_ref_start.data(); v->region_ref_end = b->region_ref_end.data();
    v->region_cand_start = b->region_cand_start.data(); v->region_cand_end = b->region_cand_end.data();
    v->region_ref_off = b->region_ref_off.data(); v->region_ref_len = b->region_ref_len.data();
    v->region_read_begin = b->region_read_begin.data(); v->ref = b->ref.data();
    v->min_qual = (b->n_reads() && b->min_qual > 0 && b->min_qual <= 255) ? b->min_qual : 0;   // 0 = no promise
    return PV_OK;
}
extern "C" const int32_t* pv_ingest_hp_tags(const PvIngestBatch* b) { return b ? b->hp.data() : nullptr; }
extern "C" const uint16_t* pv_ingest_bam_flags(const PvIngestBatch* b) { return b ? b->bam_flag.data() : nullptr; }
extern "C" const int64_t* pv_ingest_pos_end(const PvIngestBatch* b) { return b ? b->read_pos_end.data() : nullptr; }
extern "C" const char* pv_ingest_query_names(const PvIngestBatch* b, int64_t* total) {
    if (!b) return nullptr;
    if (total) *total = (int64_t)b->names.size();
    return b->names.data();
}

extern "C" int pv_ingest_select(const PvIngestBatch* b, const int64_t* keep, int64_t n_keep, PvIngestBatch** out) {
    if (!b || !out || (n_keep > 0 && !keep)) return fail(PV_EINVAL, "null argument");
    PvIngestBatch* o = new PvIngestBatch();
    o->region_ref_start = b->region_ref_start; o->region_ref_end = b->region_ref_end;
    o->region_cand_start = b->region_cand_start; o->region_cand_end = b->region_cand_end;
    o->region_ref_off = b->region_ref_off; o->region_ref_len = b->region_ref_len; o->ref = b->ref;
    const int nreg = (int)b->region_ref_start.size();
    // offsets of the query names
    std::vector<int64_t> name_off;
    for (int64_t p = 0; p < (int64_t)b->names.size(); p += (int64_t)strlen(b->names.data() + p) + 1) name_off.push_back(p);
    o->region_read_begin.assign(nreg + 1, 0);
    int reg = 0;
    for (int64_t j = 0; j < n_keep; j++) {
        const int64_t i = keep[j];
        if (i < 0 || i >= b->n_reads()) { delete o; return fail(PV_EINVAL, "read index %lld out of range", (long long)i); }
        int r = 0;
        if (nreg > 0) {
            r = (int)(std::upper_bound(b->region_read_begin.begin(), b->region_read_begin.end(), i) - b->region_read_begin.begin()) - 1;
            if (r < reg) { delete o; return fail(PV_EINVAL, "keep_idx must be grouped by region in region order"); }
            reg = r;
            o->region_read_begin[r + 1]++;
        }
        o->read_pos.push_back(b->read_pos[i]); o->read_pos_end.push_back(b->read_pos_end[i]);
        o->read_base_off.push_back((int64_t)o->bases.size()); o->read_len.push_back(b->read_len[i]);
        o->read_cigar_off.push_back((int64_t)o->cigar.size()); o->read_n_ops.push_back(b->read_n_ops[i]);
        o->read_flags.push_back(b->read_flags[i]); o->read_mapq.push_back(b->read_mapq[i]); o->hp.push_back(b->hp[i]); o->bam_flag.push_back(b->bam_flag[i]);
        const int64_t padded = ((int64_t)b->read_len[i] + 15) & ~15ll;
        o->bases.insert(o->bases.end(), b->bases.begin() + b->read_base_off[i], b->bases.begin() + b->read_base_off[i] + padded);
        o->quals.insert(o->quals.end(), b->quals.begin() + b->read_base_off[i], b->quals.begin() + b->read_base_off[i] + padded);
        o->cigar.insert(o->cigar.end(), b->cigar.begin() + b->read_cigar_off[i], b->cigar.begin() + b->read_cigar_off[i] + b->read_n_ops[i]);
        o->names.append(b->names.data() + name_off[i]); o->names.push_back('\0');
    }
    for (int r = 0; r < nreg; r++) o->region_read_begin[r + 1] += o->region_read_begin[r];
    if (nreg == 0) o->region_read_begin = {0, o->n_reads()};
    o->min_qual = b->min_qual;                                  // a subset keeps the lower bound
    *out = o;
    return PV_OK;
}

extern "C" void pv_ingest_free(PvIngestBatch* b) { delete b; }

extern "C" uint64_t pv_ingest_inflated_bytes(void) { return g_inflated_bytes.load(); }
extern "C" uint64_t pv_ingest_fast_blocks(void) { return g_fast_blocks.load(); }

extern "C" int pv_inflate_raw(const uint8_t* in, int64_t n_in, uint8_t* out, int64_t n_out) {
    if (!in || !out || n_in < 0 || n_out < 0) return fail(PV_EINVAL, "null argument");
    std::vector<uint8_t> padded((size_t)n_in + 16, 0);
    memcpy(padded.data(), in, (size_t)n_in);
    std::vector<fastinf::Tables> t(1);
    return fastinf::inflate_raw(padded.data(), (size_t)n_in, out, (size_t)n_out, t[0]) ? PV_OK : fail(PV_EINVAL, "pv_inflate_raw: not a DEFLATE stream of %lld bytes", (long long)n_out);
}
