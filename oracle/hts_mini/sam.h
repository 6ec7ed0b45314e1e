/* TEST INFRASTRUCTURE ONLY (oracle). A minimal stand-in for the part of htslib 1.9 (the version the reference pins:
 * pepper/modules/htslib.cmake:9; htslib itself is absent from this image and from /root/reference) that the reference's
 * pepper_variant/modules/cpp/bam_handler.cpp calls, so that the UNMODIFIED bam_handler.cpp -- get_reads' clipping,
 * flag / mapq filters, aux walk (:115-451) -- can be compiled and run here as the oracle for the ingest row.
 *
 * What is restated from htslib's published behaviour (SAM/BAM specification + htslib 1.9 sam.c / hts.c), NOT taken from
 * the product's ingest.cpp: BGZF = concatenated gzip members (inflated with zlib), the BAM header and record layout,
 * bam1_t's data block (qname, cigar, seq nt16, qual, aux), bam_tag2cigar (the CG:B,I long-CIGAR convention, applied by
 * bam_read1), bam_endpos, and the iterator contract of sam_itr_queryi / sam_itr_next: records of `tid` whose
 * [pos, bam_endpos) overlaps [beg, end), in file order. The iterator does NOT use the .bai (a linear scan returns what
 * the index would narrow down to), so it also checks the product's index query independently. */
#ifndef PV_HTS_MINI_SAM_H
#define PV_HTS_MINI_SAM_H
#include <zlib.h>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define BAM_CMATCH 0
#define BAM_CINS 1
#define BAM_CDEL 2
#define BAM_CREF_SKIP 3
#define BAM_CSOFT_CLIP 4
#define BAM_CHARD_CLIP 5
#define BAM_CPAD 6
#define BAM_CEQUAL 7
#define BAM_CDIFF 8
#define BAM_CBACK 9

#define BAM_FPAIRED 1
#define BAM_FPROPER_PAIR 2
#define BAM_FUNMAP 4
#define BAM_FMUNMAP 8
#define BAM_FREVERSE 16
#define BAM_FMREVERSE 32
#define BAM_FREAD1 64
#define BAM_FREAD2 128
#define BAM_FSECONDARY 256
#define BAM_FQCFAIL 512
#define BAM_FDUP 1024
#define BAM_FSUPPLEMENTARY 2048

typedef struct { int32_t n_targets; uint32_t l_text; uint32_t* target_len; char** target_name; char* text; } bam_hdr_t;
typedef struct {
    int32_t tid, pos; uint16_t bin; uint8_t qual, l_qname; uint16_t flag; uint8_t unused1, l_extranul; uint32_t n_cigar;
    int32_t l_qseq, mtid, mpos, isize;
} bam1_core_t;
typedef struct { bam1_core_t core; int l_data; uint32_t m_data; uint8_t* data; } bam1_t;

struct htsFile { std::vector<uint8_t> bytes; size_t first_record = 0; };
struct hts_idx_t { int unused; };
struct hts_itr_t { int tid; long long beg, end; size_t cursor; };

static const char seq_nt16_str[] = "=ACMGRSVTWYHKDBN";
#define bam_cigar_op(c) ((c) & 0xf)
#define bam_cigar_oplen(c) ((c) >> 4)
#define bam_get_qname(b) ((char*)(b)->data)
#define bam_get_cigar(b) ((uint32_t*)((b)->data + (b)->core.l_qname))
#define bam_get_seq(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname)
#define bam_get_qual(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1))
#define bam_get_aux(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1) + (b)->core.l_qseq)
#define bam_seqi(s, i) ((s)[(i) >> 1] >> ((~(i) & 1) << 2) & 0xf)

static inline uint32_t hm_u32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline int32_t hm_i32(const uint8_t* p) { int32_t v; memcpy(&v, p, 4); return v; }

/* BGZF: a series of gzip members; inflate them all (file sizes here are test-sized). */
static inline htsFile* sam_open(const char* path, const char*) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    std::vector<uint8_t> raw; uint8_t buf[1 << 16]; size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) raw.insert(raw.end(), buf, buf + n);
    fclose(f);
    htsFile* h = new htsFile;
    z_stream z; memset(&z, 0, sizeof z);
    if (inflateInit2(&z, 15 + 16) != Z_OK) { delete h; return NULL; }
    z.next_in = raw.data(); z.avail_in = (uInt)raw.size();
    std::vector<uint8_t> out(1 << 20);
    while (z.avail_in > 0) {
        z.next_out = out.data(); z.avail_out = (uInt)out.size();
        int r = inflate(&z, Z_NO_FLUSH);
        h->bytes.insert(h->bytes.end(), out.data(), out.data() + (out.size() - z.avail_out));
        if (r == Z_STREAM_END) { if (z.avail_in == 0) break; inflateReset(&z); }
        else if (r != Z_OK) { inflateEnd(&z); delete h; return NULL; }
    }
    inflateEnd(&z);
    if (h->bytes.size() < 12 || memcmp(h->bytes.data(), "BAM\1", 4) != 0) { delete h; return NULL; }
    return h;
}
static inline hts_idx_t* sam_index_load(htsFile*, const char* path) {
    std::string p = std::string(path) + ".bai";
    FILE* f = fopen(p.c_str(), "rb");
    if (!f) return NULL;
    fclose(f);
    return new hts_idx_t;
}
static inline bam_hdr_t* sam_hdr_read(htsFile* h) {
    const uint8_t* p = h->bytes.data();
    bam_hdr_t* hd = new bam_hdr_t;
    size_t o = 4;
    hd->l_text = hm_u32(p + o); o += 4;
    hd->text = (char*)calloc(hd->l_text + 1, 1); memcpy(hd->text, p + o, hd->l_text); o += hd->l_text;
    hd->n_targets = hm_i32(p + o); o += 4;
    hd->target_name = (char**)calloc(hd->n_targets + 1, sizeof(char*));
    hd->target_len = (uint32_t*)calloc(hd->n_targets + 1, sizeof(uint32_t));
    for (int i = 0; i < hd->n_targets; i++) {
        uint32_t l = hm_u32(p + o); o += 4;
        hd->target_name[i] = (char*)calloc(l + 1, 1); memcpy(hd->target_name[i], p + o, l); o += l;
        hd->target_len[i] = hm_u32(p + o); o += 4;
    }
    h->first_record = o;
    return hd;
}
/* the reference destroys the header in two getters AND in its destructor (bam_handler.cpp:53,98,457): a no-op here */
static inline void bam_hdr_destroy(bam_hdr_t*) {}
static inline int bam_name2id(bam_hdr_t* h, const char* name) {
    for (int i = 0; i < h->n_targets; i++) if (strcmp(h->target_name[i], name) == 0) return i;
    return -1;
}
static inline bam1_t* bam_init1() { return (bam1_t*)calloc(1, sizeof(bam1_t)); }
static inline void bam_destroy1(bam1_t* b) { if (b) { free(b->data); free(b); } }
static inline hts_itr_t* sam_itr_queryi(hts_idx_t*, int tid, long long beg, long long end) {
    hts_itr_t* it = new hts_itr_t; it->tid = tid; it->beg = beg; it->end = end; it->cursor = 0; return it;
}
static inline void hts_itr_destroy(hts_itr_t* it) { delete it; }
static inline void hts_idx_destroy(hts_idx_t* i) { delete i; }
static inline void sam_close(htsFile* f) { delete f; }

static inline int hm_cigar2rlen(int n, const uint32_t* c) {
    int l = 0;
    for (int k = 0; k < n; k++) { const int op = bam_cigar_op(c[k]); if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) l += bam_cigar_oplen(c[k]); }
    return l;
}
static inline int32_t bam_endpos(const bam1_t* b) {
    int rlen = 1;
    if (!(b->core.flag & BAM_FUNMAP) && b->core.n_cigar > 0) rlen = hm_cigar2rlen(b->core.n_cigar, bam_get_cigar(b));
    return b->core.pos + (rlen ? rlen : 1);
}
/* long CIGARs (SAM spec 4.2.2, htslib bam_tag2cigar): a record whose CIGAR is `<l_seq>S<rlen>N` carries the real one
 * in a CG:B,I tag; the reader moves it into place and drops the tag. */
static inline void hm_tag2cigar(bam1_t* b) {
    bam1_core_t* c = &b->core;
    if (c->n_cigar == 0 || c->tid < 0 || c->pos < 0) return;
    const uint32_t c0 = bam_get_cigar(b)[0];
    if (bam_cigar_op(c0) != BAM_CSOFT_CLIP || (int32_t)bam_cigar_oplen(c0) != c->l_qseq) return;
    uint8_t* aux = bam_get_aux(b); uint8_t* end = b->data + b->l_data; uint8_t* s = aux; uint8_t* cg = NULL;
    while (end - s >= 4) {                                   /* find CG (a plain walk over well-formed tags) */
        uint8_t* tag = s; const uint8_t t = s[2]; s += 3;
        size_t sz;
        if (t == 'A' || t == 'c' || t == 'C') sz = 1; else if (t == 's' || t == 'S') sz = 2; else if (t == 'i' || t == 'I' || t == 'f') sz = 4;
        else if (t == 'Z' || t == 'H') { sz = strnlen((char*)s, end - s) + 1; }
        else if (t == 'B') { const uint8_t st = s[0]; const size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4; sz = 5 + es * hm_u32(s + 1); }
        else return;
        if (tag[0] == 'C' && tag[1] == 'G') { cg = tag; break; }
        s += sz;
    }
    if (!cg || cg[2] != 'B' || cg[3] != 'I') return;
    const uint32_t n = hm_u32(cg + 4);
    if (n == 0) return;
    const size_t cg_bytes = 8 + (size_t)n * 4;
    const size_t fake = (size_t)c->n_cigar * 4;
    std::vector<uint8_t> nd;
    uint8_t* cig = (uint8_t*)bam_get_cigar(b);
    nd.insert(nd.end(), b->data, cig);                       /* qname */
    nd.insert(nd.end(), cg + 8, cg + 8 + (size_t)n * 4);     /* real cigar */
    nd.insert(nd.end(), cig + fake, cg);                     /* seq, qual, aux in front of CG */
    nd.insert(nd.end(), cg + cg_bytes, end);                 /* aux behind CG */
    free(b->data);
    b->data = (uint8_t*)calloc(nd.size() + 65536, 1);        /* slack: the reference indexes qual/seq unchecked */
    memcpy(b->data, nd.data(), nd.size());
    b->l_data = (int)nd.size(); b->m_data = (uint32_t)nd.size() + 65536;
    c->n_cigar = n;
}
static inline int sam_itr_next(htsFile* f, hts_itr_t* it, bam1_t* b) {
    const std::vector<uint8_t>& v = f->bytes;
    if (it->cursor == 0) it->cursor = f->first_record;
    while (it->cursor + 4 <= v.size()) {
        const uint8_t* p = v.data() + it->cursor;
        const uint32_t bs = hm_u32(p);
        if (bs < 32 || it->cursor + 4 + bs > v.size()) return -2;
        it->cursor += 4 + bs;
        bam1_core_t* c = &b->core;
        c->tid = hm_i32(p + 4); c->pos = hm_i32(p + 8);
        c->l_qname = p[12]; c->qual = p[13]; c->bin = (uint16_t)(p[14] | p[15] << 8);
        c->n_cigar = (uint32_t)(p[16] | p[17] << 8); c->flag = (uint16_t)(p[18] | p[19] << 8);
        c->l_qseq = hm_i32(p + 20); c->mtid = hm_i32(p + 24); c->mpos = hm_i32(p + 28); c->isize = hm_i32(p + 32);
        free(b->data);
        b->l_data = (int)bs - 32; b->m_data = (uint32_t)b->l_data + 65536;
        b->data = (uint8_t*)calloc(b->m_data, 1);
        memcpy(b->data, p + 36, b->l_data);
        hm_tag2cigar(b);
        if (it->tid < 0 || c->tid != it->tid) continue;
        const long long beg = c->pos, end = bam_endpos(b);
        if (end > it->beg && it->end > beg) return 0;
    }
    return -1;
}
static inline int64_t bam_aux2i(const uint8_t* s) {
    const int type = *s++;
    if (type == 'c') return (int8_t)*s;
    if (type == 'C') return *s;
    if (type == 's') { int16_t v; memcpy(&v, s, 2); return v; }
    if (type == 'S') { uint16_t v; memcpy(&v, s, 2); return v; }
    if (type == 'i') { int32_t v; memcpy(&v, s, 4); return v; }
    if (type == 'I') { uint32_t v; memcpy(&v, s, 4); return v; }
    errno = EINVAL;
    return 0;
}
#endif
