/*
 * Seeded synthetic pileup generator emitting the packed read-batch layout of
 * include/pepper_b200.h directly. Used by tests and bench.py to make inputs for BOTH the CUDA path and the
 * CPU oracle; it is not part of the hot path.
 *
 * Everything is a pure function of (seed, absolute position) or (seed, region, read index): the contig
 * (i.i.d. uniform ACGT), the truth variants (het/hom SNP about every `snp_every` bp, 1-10 bp indel about every
 * `indel_every` bp) and every read (start, length, strand, haplotype, errors, qualities), so any region can be
 * generated independently and in parallel, and regenerated bit-identically anywhere.
 *
 * Reads are produced the way BAM_handler::get_reads hands them to the summary generator
 * (/root/reference/pepper_variant/modules/cpp/bam_handler.cpp:178-306): clipped to [region_start, region_end],
 * upper-case bases, first CIGAR op always a match, ops in {M, I, D}.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>

typedef struct {
    uint64_t seed;
    int64_t contig_len;        /* positions >= contig_len do not exist */
    int64_t region_size;       /* 100000 (CallVariantsArguments.py:65-70) */
    int64_t margin;            /* 100  (Options.py:2 REGION_SAFE_BASES) */
    double coverage;
    double len_median, len_sigma;   /* log-normal read length; sigma==0 -> normal(len_median, len_sd) */
    double len_sd;
    int64_t len_min, len_max;
    double sub_rate, ins_rate, del_rate, indel_geom_p;
    int32_t qual_lo, qual_hi;       /* uniform inclusive */
    int32_t snp_every, indel_every; /* truth variant spacing (0 = none) */
} PvSynthConfig;

typedef struct {
    int64_t n_reads, n_bases /* padded to 16 per read */, n_ops, n_ref;
} PvSynthSizes;

static inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
typedef struct { uint64_t s; } Rng;
static inline uint64_t rng_next(Rng* r) { r->s += 0x9E3779B97F4A7C15ull; uint64_t z = r->s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
static inline double rng_unif(Rng* r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }

static const char ACGT[4] = {'A', 'C', 'G', 'T'};
static inline int ref_code(uint64_t seed, int64_t x) { return (int)(mix64(seed ^ (0xA5A5ull << 32) ^ (uint64_t)x) & 3); }

/* truth variant at absolute position x: type 0 none, 1 SNP, 2 INS, 3 DEL; zyg 1 het (hap 1 only), 2 hom */
typedef struct { int type, zyg, len; int alt[10]; } Variant;
static inline void variant_at(const PvSynthConfig* c, int64_t x, Variant* v) {
    v->type = 0;
    uint64_t h = mix64(c->seed ^ (0x5EEDull << 40) ^ (uint64_t)x * 0x9E37ull);
    if (c->snp_every > 0 && h % (uint64_t)c->snp_every == 0) {
        v->type = 1; v->zyg = ((h >> 32) & 1) ? 2 : 1; v->len = 1;
        v->alt[0] = (ref_code(c->seed, x) + 1 + (int)((h >> 40) % 3)) & 3;
        return;
    }
    uint64_t g = mix64(h ^ 0x1D31ull);
    if (c->indel_every > 0 && g % (uint64_t)c->indel_every == 0) {
        v->type = ((g >> 32) & 1) ? 2 : 3; v->zyg = ((g >> 33) & 1) ? 2 : 1; v->len = 1 + (int)((g >> 40) % 10);
        uint64_t b = mix64(g);
        for (int i = 0; i < v->len; i++) v->alt[i] = (int)((b >> (2 * i)) & 3);
    }
}

static inline int geom(Rng* r, double p) { int n = 1; while (rng_unif(r) > p && n < 50) n++; return n; }

typedef struct {
    /* outputs (NULL in the counting pass) */
    int64_t* read_pos; int64_t* read_base_off; int32_t* read_len; int64_t* read_cigar_off; int32_t* read_n_ops;
    uint8_t* read_flags; uint8_t* read_mapq; uint8_t* bases; uint8_t* quals; uint32_t* cigar; uint8_t* ref;
    int64_t r_cur, b_cur, o_cur;     /* cursors (absolute indices) */
} Out;

static inline void emit_op(Out* o, int write, int op, int64_t len, int* cur_op, int64_t* cur_len, int32_t* n_ops) {
    if (len <= 0) return;
    if (*cur_op == op) { *cur_len += len; return; }
    if (*cur_op >= 0) { if (write) o->cigar[o->o_cur] = (uint32_t)((*cur_len << 4) | (uint32_t)*cur_op); o->o_cur++; (*n_ops)++; }
    *cur_op = op; *cur_len = len;
}

static void region_bounds(const PvSynthConfig* c, int64_t region, int64_t* cs, int64_t* ce, int64_t* rs, int64_t* re) {
    /* ImageGenerationUI.py:292-316 intervals; AlignmentSummarizer.py:181-182 margins */
    *cs = region * c->region_size;
    *ce = *cs + c->region_size; if (*ce > c->contig_len - 1) *ce = c->contig_len - 1;
    *rs = *cs - c->margin; if (*rs < 0) *rs = 0;
    *re = *ce + c->margin; if (*re > c->contig_len - 1) *re = c->contig_len - 1;
}

static void gen_region(const PvSynthConfig* c, int64_t region, Out* o, int write) {
    int64_t cs, ce, rs, re; region_bounds(c, region, &cs, &ce, &rs, &re);
    const int64_t L = re - rs + 1;
    if (write) for (int64_t i = 0; i < L; i++) o->ref[i] = (uint8_t)ACGT[ref_code(c->seed, rs + i)];
    double mean_len = c->len_sigma > 0 ? c->len_median * exp(0.5 * c->len_sigma * c->len_sigma) : c->len_median;
    if (mean_len > (double)c->len_max) mean_len = (double)c->len_max;
    const int64_t n_reads = (int64_t)(c->coverage * ((double)L + mean_len) / mean_len + 0.5);
    for (int64_t r = 0; r < n_reads; r++) {
        Rng g; g.s = mix64(c->seed ^ ((uint64_t)region << 24) ^ (uint64_t)r * 0xD1B54A32D192ED03ull);
        double len_d;
        if (c->len_sigma > 0) {
            double u1 = rng_unif(&g), u2 = rng_unif(&g); if (u1 < 1e-12) u1 = 1e-12;
            len_d = c->len_median * exp(c->len_sigma * sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2));
        } else {
            double u1 = rng_unif(&g), u2 = rng_unif(&g); if (u1 < 1e-12) u1 = 1e-12;
            len_d = c->len_median + c->len_sd * sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
        }
        int64_t len = (int64_t)len_d; if (len < c->len_min) len = c->len_min; if (len > c->len_max) len = c->len_max;
        /* start uniform so that the read overlaps [rs, re] */
        int64_t start = rs - len + 1 + (int64_t)(rng_unif(&g) * (double)(L + len - 1));
        int64_t end = start + len - 1;
        if (start < rs) start = rs; if (end > re) end = re;   /* get_reads clipping */
        if (end < start) continue;
        const int rev = (int)(rng_next(&g) & 1), hap = (int)(rng_next(&g) & 1);
        const int64_t b0 = o->b_cur, o0 = o->o_cur;
        int64_t nb = 0; int32_t n_ops = 0; int cur_op = -1; int64_t cur_len = 0;
#define PUTB(code) do { const int code_ = (code); if (write) { o->bases[b0 + nb] = (uint8_t)ACGT[code_ & 3]; \
        o->quals[b0 + nb] = (uint8_t)(c->qual_lo + (int)(rng_next(&g) % (uint64_t)(c->qual_hi - c->qual_lo + 1))); } \
        else (void)rng_next(&g); nb++; } while (0)
        int64_t x = start;
        while (x <= end) {
            Variant v; variant_at(c, x, &v);
            const int carries = v.type && (v.zyg == 2 || hap == 1);
            int code = ref_code(c->seed, x);
            if (carries && v.type == 1) code = v.alt[0];
            const int first = (x == start), last = (x == end);
            if (!first && !last && rng_unif(&g) < c->sub_rate) code = (code + 1 + (int)(rng_next(&g) % 3)) & 3;
            emit_op(o, write, 0, 1, &cur_op, &cur_len, &n_ops); PUTB(code);
            x++;
            if (last) break;
            if (carries && v.type == 2) {               /* truth insertion after x */
                emit_op(o, write, 1, v.len, &cur_op, &cur_len, &n_ops);
                for (int i = 0; i < v.len; i++) PUTB(v.alt[i]);
            } else if (carries && v.type == 3) {        /* truth deletion of the next v.len bases */
                int64_t dl = v.len; if (x + dl > end) dl = end - x;   /* keep a final match */
                if (dl > 0) { emit_op(o, write, 2, dl, &cur_op, &cur_len, &n_ops); x += dl; }
            } else {
                const double u = rng_unif(&g);
                if (u < c->ins_rate) {
                    const int il = geom(&g, c->indel_geom_p);
                    emit_op(o, write, 1, il, &cur_op, &cur_len, &n_ops);
                    for (int i = 0; i < il; i++) PUTB((int)(rng_next(&g) & 3));
                } else if (u < c->ins_rate + c->del_rate) {
                    int64_t dl = geom(&g, c->indel_geom_p); if (x + dl > end) dl = end - x;
                    if (dl > 0) { emit_op(o, write, 2, dl, &cur_op, &cur_len, &n_ops); x += dl; }
                }
            }
        }
        emit_op(o, write, 15, 1, &cur_op, &cur_len, &n_ops);   /* flush the pending op (15 is never stored) */
#undef PUTB
        if (write) {
            o->read_pos[o->r_cur] = start; o->read_base_off[o->r_cur] = b0; o->read_len[o->r_cur] = (int32_t)nb;
            o->read_cigar_off[o->r_cur] = o0; o->read_n_ops[o->r_cur] = n_ops;
            o->read_flags[o->r_cur] = (uint8_t)rev; o->read_mapq[o->r_cur] = 60;
            const int64_t pad = ((nb + 15) & ~15ll);
            for (int64_t i = nb; i < pad; i++) { o->bases[b0 + i] = 0; o->quals[b0 + i] = 0; }
        }
        o->r_cur++; o->b_cur = b0 + ((nb + 15) & ~15ll);
    }
}

void pv_synth_region_bounds(const PvSynthConfig* c, int64_t region, int64_t* out4) {
    region_bounds(c, region, &out4[0], &out4[1], &out4[2], &out4[3]);
}

typedef struct {
    const PvSynthConfig* c; int64_t r0, r1, stride; PvSynthSizes* sizes; Out base; int write;
    const int64_t* read_begin; const int64_t* base_begin; const int64_t* op_begin; const int64_t* ref_begin;
} Job;

static void* worker(void* p) {
    Job* j = (Job*)p;
    for (int64_t r = j->r0; r < j->r1; r += j->stride) {
        Out o = j->base;
        if (j->write) {
            o.r_cur = j->read_begin[r]; o.b_cur = j->base_begin[r]; o.o_cur = j->op_begin[r];
            o.ref = j->base.ref + j->ref_begin[r];
        } else { o.r_cur = 0; o.b_cur = 0; o.o_cur = 0; }
        gen_region(j->c, r, &o, j->write);
        if (!j->write) {
            int64_t cs, ce, rs, re; region_bounds(j->c, r, &cs, &ce, &rs, &re);
            j->sizes[r].n_reads = o.r_cur; j->sizes[r].n_bases = o.b_cur; j->sizes[r].n_ops = o.o_cur;
            j->sizes[r].n_ref = re - rs + 1;
        }
    }
    return NULL;
}

static void run_jobs(Job* proto, int64_t first, int64_t n_regions, int threads) {
    if (threads < 1) threads = 1; if (threads > 256) threads = 256;
    pthread_t th[256]; Job jobs[256];
    for (int t = 0; t < threads; t++) { jobs[t] = *proto; jobs[t].r0 = first + t; jobs[t].r1 = first + n_regions; jobs[t].stride = threads;
        pthread_create(&th[t], NULL, worker, &jobs[t]); }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
}

/* pass 1: sizes[r - first] for regions [first, first+n) */
void pv_synth_count(const PvSynthConfig* c, int64_t first, int64_t n, int threads, PvSynthSizes* sizes) {
    Job j; memset(&j, 0, sizeof(j)); j.c = c; j.sizes = sizes - first; j.write = 0;
    run_jobs(&j, first, n, threads);
}

/* pass 2: *_begin[r - first] are the exclusive prefix sums of the pass-1 sizes */
void pv_synth_fill(const PvSynthConfig* c, int64_t first, int64_t n, int threads,
                   const int64_t* read_begin, const int64_t* base_begin, const int64_t* op_begin, const int64_t* ref_begin,
                   int64_t* read_pos, int64_t* read_base_off, int32_t* read_len, int64_t* read_cigar_off,
                   int32_t* read_n_ops, uint8_t* read_flags, uint8_t* read_mapq,
                   uint8_t* bases, uint8_t* quals, uint32_t* cigar, uint8_t* ref) {
    Job j; memset(&j, 0, sizeof(j)); j.c = c; j.write = 1;
    j.read_begin = read_begin - first; j.base_begin = base_begin - first; j.op_begin = op_begin - first; j.ref_begin = ref_begin - first;
    j.base.read_pos = read_pos; j.base.read_base_off = read_base_off; j.base.read_len = read_len;
    j.base.read_cigar_off = read_cigar_off; j.base.read_n_ops = read_n_ops; j.base.read_flags = read_flags;
    j.base.read_mapq = read_mapq; j.base.bases = bases; j.base.quals = quals; j.base.cigar = cigar; j.base.ref = ref;
    run_jobs(&j, first, n, threads);
}

/* ---- host half of the DEVICE generator (csrc/synth_device.cu) -------------------------------------------------------
 * reads per region and the clamped read lengths: the only part of a read that goes through libm (exp / log / cos of the
 * length sample), which the GPU's math library does not reproduce bit for bit. 4 bytes per read. */
static int64_t region_reads(const PvSynthConfig* c, int64_t region) {
    int64_t cs, ce, rs, re; region_bounds(c, region, &cs, &ce, &rs, &re);
    const int64_t L = re - rs + 1;
    double mean_len = c->len_sigma > 0 ? c->len_median * exp(0.5 * c->len_sigma * c->len_sigma) : c->len_median;
    if (mean_len > (double)c->len_max) mean_len = (double)c->len_max;
    return (int64_t)(c->coverage * ((double)L + mean_len) / mean_len + 0.5);
}

/* counts[i] = reads of region first + i (before any are dropped: none is, a read always overlaps its region) */
void pv_synth_region_reads(const PvSynthConfig* c, int64_t first, int64_t n, int64_t* counts) {
    for (int64_t i = 0; i < n; i++) counts[i] = region_reads(c, first + i);
}

typedef struct { const PvSynthConfig* c; int64_t first, r0, r1, stride; const int64_t* read_begin; int32_t* len_out; } LenJob;
static void* len_worker(void* p) {
    LenJob* j = (LenJob*)p;
    for (int64_t reg = j->r0; reg < j->r1; reg += j->stride) {
        const PvSynthConfig* c = j->c;
        const int64_t region = j->first + reg, n = j->read_begin[reg + 1] - j->read_begin[reg];
        for (int64_t r = 0; r < n; r++) {
            Rng g; g.s = mix64(c->seed ^ ((uint64_t)region << 24) ^ (uint64_t)r * 0xD1B54A32D192ED03ull);
            double u1 = rng_unif(&g), u2 = rng_unif(&g); if (u1 < 1e-12) u1 = 1e-12;
            double len_d;
            if (c->len_sigma > 0) len_d = c->len_median * exp(c->len_sigma * sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2));
            else len_d = c->len_median + c->len_sd * sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
            int64_t len = (int64_t)len_d; if (len < c->len_min) len = c->len_min; if (len > c->len_max) len = c->len_max;
            j->len_out[j->read_begin[reg] + r] = (int32_t)len;
        }
    }
    return NULL;
}
void pv_synth_read_lengths(const PvSynthConfig* c, int64_t first, int64_t n, int threads, const int64_t* read_begin, int32_t* len_out) {
    if (threads < 1) threads = 1; if (threads > 256) threads = 256;
    pthread_t th[256]; LenJob jobs[256];
    for (int t = 0; t < threads; t++) {
        jobs[t].c = c; jobs[t].first = first; jobs[t].r0 = t; jobs[t].r1 = n; jobs[t].stride = threads;
        jobs[t].read_begin = read_begin; jobs[t].len_out = len_out;
        pthread_create(&th[t], NULL, len_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
}
