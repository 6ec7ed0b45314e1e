/*
 * TEST INFRASTRUCTURE ONLY -- a CPU restatement ("port") of the reference's pileup-summary algorithm.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the
 * product path (pepper-thesis_b200/) never does.
 *
 * Restates, for ONE region, RegionalSummaryGenerator::generate_summary in inference mode
 * (train_mode=false, GENERATE_INDELS=false => base_index == position - ref_start):
 *     /root/reference/pepper_variant/modules/cpp/region_summary.cpp
 *       :174-191  encode_reference_bases        -> feature 0
 *       :201-230  get_feature_index             -> feat_index()
 *       :337-566  populate_summary_matrix       -> walk_read()
 *       :625-654  site thresholds + clamp       -> step 2
 *       :667-915  per-allele filters + windows  -> step 3
 * It is written in the SAME decomposition the CUDA path uses (dense per-position counters with the two
 * strands kept apart, an indel/other-base event list, exact allele de-duplication by sorting) so a
 * disagreement between the CUDA path and the reference can be bisected here.
 *
 * Parity status: pinned against the unmodified reference compiled in oracle/_ref (tests/test_oracle.py)
 * on the SURVEY section-8 known-answer cases and on seeded random fuzz; the reference repo itself holds no
 * golden vectors for this path.
 *
 * Where the reference has undefined behaviour this port (and the CUDA path) define it:
 *   - read index beyond the sequence (possible through the REF_SKIP/PAD fall-through, :556-561): the base is
 *     treated as absent;
 *   - get_feature_index()==-1 used as a vector index in the centre-row overrides (:859-860 etc.): skipped.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define F 26
#define MAXC 125

typedef struct {
    int64_t pos;        /* region-relative anchor position */
    const uint8_t* s;   /* allele bytes (after the type digit) */
    int32_t len;        /* number of allele bytes */
    uint8_t type;       /* 1 SNP, 2 INS, 3 DEL */
    uint8_t rev;
} Event;

static int is_valid_ref(uint8_t c) {                 /* check_ref_base :193-199 */
    return c == 'A' || c == 'a' || c == 'C' || c == 'c' || c == 'G' || c == 'g' || c == 'T' || c == 't';
}
static uint8_t up(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }

static int feat_index(uint8_t ref_base, uint8_t base, int rev) {   /* :201-230 */
    if (!is_valid_ref(ref_base)) return -1;
    int start = rev ? 18 : 7;
    switch (up(base)) {
        case 'A': return start + 1;
        case 'C': return start + 2;
        case 'G': return start + 3;
        case 'T': return start + 4;
        case 'I': return start + 5;
        case 'D': return start + 6;
        default:  return start + 7;
    }
}
static int ref_value(uint8_t base) {                 /* get_reference_feature_value :165-172 */
    switch (up(base)) { case 'A': return 1; case 'C': return 2; case 'G': return 3; case 'T': return 4; default: return 5; }
}

typedef struct {
    int64_t L, ref_len, ref_start, ref_end;
    const uint8_t* ref;
    int32_t* img;       /* [L+1][F] */
    int32_t *cov, *snp, *ins, *del;
    Event* ev; int64_t n_ev, cap_ev;
    double min_snp_baseq, min_indel_baseq;
} Ctx;

static void push_event(Ctx* c, int64_t pos, const uint8_t* s, int32_t len, int type, int rev) {
    if (c->n_ev == c->cap_ev) {
        c->cap_ev = c->cap_ev ? c->cap_ev * 2 : 1024;
        c->ev = (Event*)realloc(c->ev, (size_t)c->cap_ev * sizeof(Event));
    }
    Event* e = &c->ev[c->n_ev++];
    e->pos = pos; e->s = s; e->len = len; e->type = (uint8_t)type; e->rev = (uint8_t)rev;
}

/* populate_summary_matrix :337-566 for one read */
static void walk_read(Ctx* c, int64_t read_pos, const uint8_t* seq, const uint8_t* qual, int32_t read_len,
                      const uint32_t* cig, int32_t n_ops, int rev) {
    int64_t ref_position = read_pos;
    int64_t ri = 0;
    for (int32_t k = 0; k < n_ops; k++) {
        if (ref_position > c->ref_end) break;                               /* :355 */
        const int op = (int)(cig[k] & 15u);
        const int64_t len = (int64_t)(cig[k] >> 4);
        switch (op) {
        case 0: case 7: case 8: {                                           /* MATCH/EQUAL/DIFF :357-430 */
            int next_is_indel = 0;
            if (k != n_ops - 1) { int nop = (int)(cig[k + 1] & 15u); next_is_indel = (nop == 1 || nop == 2); }
            for (int64_t i = 0; i < len; i++) {
                const int64_t p = ref_position + i, idx = ri + i;
                if (p < c->ref_start || p > c->ref_end || idx >= read_len) continue;
                const int64_t o = p - c->ref_start;
                const uint8_t base = seq[idx], rb = c->ref[o];
                if (!((double)qual[idx] >= c->min_snp_baseq)) continue;      /* :378,394,422 */
                c->cov[o] += 1;
                if (!(i == len - 1 && k != n_ops - 1 && next_is_indel))      /* anchor rule :381-391 */
                    c->img[o * F + (rev ? 15 : 4)] -= 1;
                const int fi = feat_index(rb, base, rev);
                if (fi >= 0) c->img[o * F + fi] -= 1;                        /* :396,423 */
                if (rb != base) {                                            /* raw compare :394 */
                    c->snp[o] += 1;
                    push_event(c, o, seq + idx, 1, 1, rev);
                }
            }
            ref_position += len; ri += len;
            break;
        }
        case 1: {                                                           /* IN :431-490 */
            const int64_t a = ref_position - 1;
            if (a >= c->ref_start && a <= c->ref_end && ri - 1 >= 0 && ri - 1 < read_len) {
                const int64_t o = a - c->ref_start;
                const int64_t n = len + 1;                                   /* :442 */
                int64_t elen = n;                                            /* substr truncation :439 */
                if (ri - 1 + elen > read_len) elen = read_len - (ri - 1);
                double bq = 0;
                for (int64_t i = ri - 1; i < ri - 1 + n && i < read_len; i++) bq += qual[i];   /* :448-450 */
                const int pass = bq >= c->min_indel_baseq * (double)n;
                if (pass && (double)qual[ri - 1] < c->min_snp_baseq) c->cov[o] += 1;          /* :453-454 */
                if (1 + elen <= 61 && pass) {                                /* :461 */
                    const int fi = feat_index(c->ref[o], 'I', rev);
                    if (fi >= 0) c->img[o * F + fi] -= 1;
                    c->ins[o] += 1;
                    push_event(c, o, seq + (ri - 1), (int32_t)elen, 2, rev);
                }
            }
            ri += len;
            break;
        }
        case 2: {                                                           /* DEL :491-555 */
            const int64_t a = ref_position - 1;
            if (a >= c->ref_start && a <= c->ref_end) {
                const int64_t o = a - c->ref_start;
                const int fi = feat_index(c->ref[o], 'D', rev);
                if (fi >= 0) c->img[o * F + fi] -= 1;                        /* :497 (unconditional) */
                int64_t elen = len + 1;                                      /* substr truncation :500 */
                if (o + elen > c->ref_len) elen = c->ref_len - o;
                if (1 + elen <= 61) {                                        /* :511 */
                    c->del[o] += 1;
                    push_event(c, o, c->ref + o, (int32_t)elen, 3, rev);
                }
            }
            for (int64_t i = 0; i < len; i++) {                              /* :542-552 */
                const int64_t p = ref_position + i;
                if (p >= c->ref_start && p <= c->ref_end) {
                    const int64_t o = p - c->ref_start;
                    const int fi = feat_index(c->ref[o], '*', rev);
                    if (fi >= 0) c->img[o * F + fi] -= 1;
                }
            }
            ref_position += len;
            break;
        }
        case 3: case 6:                                                     /* REF_SKIP/PAD :556-558 */
            ref_position += len;
            /* FALLTHROUGH -- reproduces the reference's missing break */
        case 4:                                                             /* SOFT_CLIP :559-561 */
            ri += len;
            break;
        default:                                                            /* HARD_CLIP / BACK / unknown: no-op */
            break;
        }
    }
}

static int cmp_event(const void* pa, const void* pb) {
    const Event* a = (const Event*)pa; const Event* b = (const Event*)pb;
    if (a->pos != b->pos) return a->pos < b->pos ? -1 : 1;
    if (a->type != b->type) return a->type < b->type ? -1 : 1;       /* '1' < '2' < '3' */
    const int32_t m = a->len < b->len ? a->len : b->len;
    const int r = memcmp(a->s, b->s, (size_t)m);                      /* std::string order = unsigned bytes */
    if (r) return r;
    if (a->len != b->len) return a->len < b->len ? -1 : 1;
    return 0;
}

static int32_t clampc(int32_t v) { return v > MAXC ? MAXC : (v < -MAXC ? -MAXC : v); }
static int32_t minc(int32_t v) { return v < MAXC ? v : MAXC; }

/*
 * thr[9] = min_snp_baseq, min_indel_baseq, snp_freq, insert_freq, delete_freq, min_coverage,
 *          snp_candidate_freq, indel_candidate_freq, candidate_support     (region_summary.h:191-206)
 * Outputs (caller-allocated, `capacity` candidates): windows int16 [K][33][26], position, depth, frequency,
 * allele [K][64] ("<type digit><bases>", NUL padded), allele_len [K] (bytes incl. the digit).
 * dense_out (optional) receives the clamped image_matrix int32 [L][26]; counts_out (optional) int32 [L][4]
 * = coverage, snp, insert, delete counts.
 * Returns the number of candidates the reference would emit (may exceed capacity; only the first
 * `capacity` are stored), or -1 on bad arguments.
 */
int64_t pv_port_summary_region(const int64_t* read_pos, const int64_t* base_off, const int32_t* read_len,
                               const int64_t* cigar_off, const int32_t* n_ops, const uint8_t* flags,
                               const uint8_t* mapq, const uint8_t* bases, const uint8_t* quals,
                               const uint32_t* cigar, int64_t rbegin, int64_t rend,
                               const uint8_t* ref, int64_t ref_len, int64_t ref_start, int64_t ref_end,
                               const double* thr, int32_t skip_indels, int64_t cand_start, int64_t cand_end,
                               int64_t capacity, int16_t* windows, int64_t* position, int32_t* depth,
                               int32_t* frequency, uint8_t* allele, uint8_t* allele_len,
                               int32_t* dense_out, int32_t* counts_out) {
    const int64_t L = ref_end - ref_start + 1;
    if (L <= 0 || ref_len < L) return -1;
    Ctx c; memset(&c, 0, sizeof(c));
    c.L = L; c.ref_len = ref_len; c.ref_start = ref_start; c.ref_end = ref_end; c.ref = ref;
    c.min_snp_baseq = thr[0]; c.min_indel_baseq = thr[1];
    c.img = (int32_t*)calloc((size_t)(L + 1) * F, sizeof(int32_t));
    c.cov = (int32_t*)calloc((size_t)L, sizeof(int32_t)); c.snp = (int32_t*)calloc((size_t)L, sizeof(int32_t));
    c.ins = (int32_t*)calloc((size_t)L, sizeof(int32_t)); c.del = (int32_t*)calloc((size_t)L, sizeof(int32_t));
    uint8_t* pass = (uint8_t*)calloc((size_t)L, 1);      /* bit0 site, bit1 snp, bit2 ins, bit3 del */

    for (int64_t o = 0; o < L; o++) c.img[o * F] = ref_value(ref[o]);           /* :174-191 */

    /* step 1: accumulate (:617-623) */
    for (int64_t r = rbegin; r < rend; r++) {
        if (mapq[r] == 0) continue;                                               /* :619 */
        walk_read(&c, read_pos[r], bases + base_off[r], quals + base_off[r], read_len[r],
                  cigar + cigar_off[r], n_ops[r], flags[r] & 1);
    }

    /* step 2: site thresholds + clamp (:634-654) */
    for (int64_t o = 0; o < L; o++) {
        const double cv = (double)c.cov[o] > 1.0 ? (double)c.cov[o] : 1.0;
        const double sf = c.snp[o] / cv, inf = c.ins[o] / cv, df = c.del[o] / cv;
        if (sf >= thr[2] || inf >= thr[3] || df >= thr[4]) {
            const int64_t p = ref_start + o;
            if (p >= cand_start && p <= cand_end && (double)c.cov[o] >= thr[5]) {
                pass[o] = 1;
                if (sf >= thr[2]) pass[o] |= 2;
                if (inf >= thr[3]) pass[o] |= 4;
                if (df >= thr[4]) pass[o] |= 8;
            }
        }
        for (int j = 11; j < 25; j++) c.img[o * F + j] = clampc(c.img[o * F + j]);   /* features 11..24 only */
    }
    if (dense_out) memcpy(dense_out, c.img, (size_t)L * F * sizeof(int32_t));
    if (counts_out)
        for (int64_t o = 0; o < L; o++) {
            counts_out[o * 4 + 0] = c.cov[o]; counts_out[o * 4 + 1] = c.snp[o];
            counts_out[o * 4 + 2] = c.ins[o]; counts_out[o * 4 + 3] = c.del[o];
        }

    /* step 3: alleles in std::set<string> order, filters, windows (:669-912) */
    qsort(c.ev, (size_t)c.n_ev, sizeof(Event), cmp_event);
    int64_t K = 0;
    for (int64_t i = 0; i < c.n_ev;) {
        int64_t j = i; int32_t nf = 0, nr = 0;
        while (j < c.n_ev && cmp_event(&c.ev[i], &c.ev[j]) == 0) { if (c.ev[j].rev) nr++; else nf++; j++; }
        const Event* e = &c.ev[i];
        i = j;
        const int64_t o = e->pos;
        if (!(pass[o] & 1)) continue;
        const int32_t dp = minc(c.cov[o]);                                     /* :682 */
        const int32_t ad = nf + nr;
        const double cf = (double)ad / ((double)dp > 1.0 ? (double)dp : 1.0);   /* :689 */
        if ((double)ad < thr[8]) continue;                                      /* :693 */
        if (e->type != 1 && cf < thr[7]) continue;                              /* :697 */
        if (e->type == 1 && cf < thr[6]) continue;                              /* :700 */
        if (e->type != 1 && skip_indels) continue;                              /* :704 */
        if (!(pass[o] & (1u << e->type))) continue;                             /* :708-712 */
        if (K < capacity) {
            int16_t* w = windows + K * 33 * F;
            for (int row = 0; row < 33; row++) {                                /* :833-841 */
                const int64_t q = o - 16 + row;
                for (int f = 0; f < F; f++)
                    w[row * F + f] = (q < 0 || q > L) ? 0 : (int16_t)c.img[q * F + f];
            }
            int16_t* mid = w + 16 * F;
            const uint8_t rb = ref[o];
            if (e->type == 1) {                                                 /* :848-862 */
                const int ff = feat_index(rb, e->s[0], 0), fr = feat_index(rb, e->s[0], 1);
                mid[1] = (int16_t)ref_value(e->s[0]);
                mid[5] = (int16_t)minc(nf); mid[16] = (int16_t)minc(nr);
                if (ff >= 0) { mid[ff] = (int16_t)-mid[ff]; mid[fr] = (int16_t)-mid[fr]; }
            } else if (e->type == 2) {                                          /* :863-877 */
                const int ff = feat_index(rb, 'I', 0), fr = feat_index(rb, 'I', 1);
                mid[2] = (int16_t)minc(e->len);
                mid[6] = (int16_t)minc(nf); mid[17] = (int16_t)minc(nr);
                if (ff >= 0) { mid[ff] = (int16_t)-mid[ff]; mid[fr] = (int16_t)-mid[fr]; }
            } else {                                                            /* :878-905 */
                const int del_len = e->len;
                int end_index = 16 + del_len - 1; if (end_index > 31) end_index = 31;
                int ff = feat_index(rb, 'D', 0), fr = feat_index(rb, 'D', 1);
                mid[3] = (int16_t)minc(del_len);
                mid[7] = (int16_t)minc(nf); mid[18] = (int16_t)minc(nr);
                if (ff >= 0) { mid[ff] = (int16_t)-mid[ff]; mid[fr] = (int16_t)-mid[fr]; }
                ff = feat_index(rb, '*', 0); fr = feat_index(rb, '*', 1);
                for (int idx = 17; idx <= end_index; idx++) {
                    int16_t* row = w + idx * F;
                    row[3] = (int16_t)minc(del_len); row[7] = (int16_t)minc(nf); row[18] = (int16_t)minc(nr);
                    if (ff >= 0) { row[ff] = (int16_t)-row[ff]; row[fr] = (int16_t)-row[fr]; }
                }
            }
            position[K] = ref_start + o; depth[K] = dp; frequency[K] = minc(ad);
            memset(allele + K * 64, 0, 64);
            allele[K * 64] = (uint8_t)('0' + e->type);
            memcpy(allele + K * 64 + 1, e->s, (size_t)e->len);
            allele_len[K] = (uint8_t)(1 + e->len);
        }
        K++;
    }
    free(c.img); free(c.cov); free(c.snp); free(c.ins); free(c.del); free(c.ev); free(pass);
    return K;
}
