/* TEST INFRASTRUCTURE ONLY: stand-in for htslib 1.9 faidx.h as used by the reference's fasta_handler.cpp:8-55 -- .fai
 * lookup (name, length, offset, line bases, line width), faidx_fetch_seq with htslib's clamping of [beg, end] to the
 * contig and the newline-skipping read. Restated from the published faidx format / htslib faidx.c, not from ingest.cpp. */
#ifndef PV_HTS_MINI_FAIDX_H
#define PV_HTS_MINI_FAIDX_H
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
struct faidx_entry { std::string name; long long len, offset; int line_blen, line_len; };
struct faidx_t { std::string path; std::vector<faidx_entry> seqs; };
static inline faidx_t* fai_load(const char* path) {
    FILE* f = fopen((std::string(path) + ".fai").c_str(), "r");
    if (!f) return NULL;
    faidx_t* fa = new faidx_t; fa->path = path;
    char name[1024]; long long len, off; int bl, ll;
    while (fscanf(f, "%1023s %lld %lld %d %d", name, &len, &off, &bl, &ll) == 5) fa->seqs.push_back({name, len, off, bl, ll});
    fclose(f);
    FILE* g = fopen(path, "rb");
    if (!g) { delete fa; return NULL; }
    fclose(g);
    return fa;
}
static inline void fai_destroy(faidx_t* f) { delete f; }
static inline int faidx_nseq(const faidx_t* f) { return (int)f->seqs.size(); }
static inline const char* faidx_iseq(const faidx_t* f, int i) { return f->seqs[i].name.c_str(); }
static inline int faidx_seq_len(const faidx_t* f, const char* name) {
    for (auto& e : f->seqs) if (e.name == name) return (int)e.len;
    return -1;
}
static inline char* faidx_fetch_seq(const faidx_t* f, const char* name, int beg, int end, int* len) {
    const faidx_entry* e = NULL;
    for (auto& s : f->seqs) if (s.name == name) e = &s;
    if (!e) { *len = -2; return NULL; }
    long long b = beg, t = end;
    if (t < b) b = t;
    if (b < 0) b = 0; else if (e->len <= b) b = e->len - 1;
    if (t < 0) t = 0; else if (e->len <= t) t = e->len - 1;
    FILE* g = fopen(f->path.c_str(), "rb");
    if (!g) { *len = -1; return NULL; }
    fseek(g, e->offset + b / e->line_blen * e->line_len + b % e->line_blen, SEEK_SET);
    const long long want = t + 1 - b;
    char* s = (char*)malloc(want + 1); long long l = 0; int c;
    while (l < want && (c = fgetc(g)) != EOF) if (c > ' ' && c <= '~') s[l++] = (char)c;   /* isgraph, like faidx.c */
    s[l] = 0; fclose(g);
    *len = (int)l;
    return s;
}
#endif
