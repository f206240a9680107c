/*
 * ipcr_oracle.c -- layer 1 ("truth") of the CPU oracle.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of what the reference's own correctness oracle
 * SimulateBatchBruteForce computes (core/engine/bruteforce.go:11-38):
 *   primer.BaseMatch   core/primer/iupac.go:6-67
 *   primer.RevComp     core/primer/rc.go:8-56
 *   primer.FindMatches core/primer/match.go:30-90
 *   filterLeftTW       core/engine/engine.go:53-68
 *   join               core/engine/engine.go:108-404
 * Go `int` is 64-bit; everything positional is int64_t here.
 */
#include "ipcr_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ---- iupac.go:6-58 ---------------------------------------------------- */
static uint8_t g_mask[256];
static uint8_t g_comp[256];
static int g_tables_ready = 0;

static void init_tables(void) {
    if (g_tables_ready) return;
    memset(g_mask, 0, sizeof g_mask);
    memset(g_comp, 0, sizeof g_comp);
    const uint8_t A = 1, C = 2, G = 4, T = 8;
    struct { char c; uint8_t m; } tab[] = {
        {'A', A}, {'C', C}, {'G', G}, {'T', T}, {'U', T},
        {'R', A | G}, {'Y', C | T}, {'S', C | G}, {'W', A | T}, {'K', G | T}, {'M', A | C},
        {'B', C | G | T}, {'D', A | G | T}, {'H', A | C | T}, {'V', A | C | G},
        {'N', A | C | G | T},
    };
    for (size_t i = 0; i < sizeof tab / sizeof tab[0]; i++) {
        g_mask[(uint8_t)tab[i].c] = tab[i].m;
        g_mask[(uint8_t)(tab[i].c + ('a' - 'A'))] = tab[i].m; /* lowercase mirrors */
    }
    /* rc.go:8-24 -- uppercase only, no U */
    const char *from = "ACGTRYSWKMBVDHN";
    const char *to = "TGCAYRSWMKVBHDN";
    for (size_t i = 0; from[i]; i++) g_comp[(uint8_t)from[i]] = (uint8_t)to[i];
    g_tables_ready = 1;
}

uint8_t orc_iupac_mask(uint8_t c) {
    init_tables();
    return g_mask[c];
}

/* iupac.go:62-67 */
int orc_base_match(uint8_t g, uint8_t p) {
    init_tables();
    if (g != 'A' && g != 'C' && g != 'G' && g != 'T') return 0;
    return (g_mask[p] & g_mask[g]) != 0;
}

/* mismatch.go:3-14 (caller guarantees equal lengths) */
int orc_mismatch_count(const uint8_t *g, const uint8_t *p, int64_t n) {
    int mm = 0;
    for (int64_t i = 0; i < n; i++)
        if (!orc_base_match(g[i], p[i])) mm++;
    return mm;
}

/* rc.go:41-56 */
int64_t orc_revcomp(const uint8_t *in, int64_t n, uint8_t *out) {
    init_tables();
    for (int64_t i = 0; i < n; i++) {
        int64_t src = n - 1 - i;
        uint8_t c = g_comp[in[src]];
        if (c == 0) return src + 1;
        out[i] = c;
    }
    return 0;
}

/* ---- containers -------------------------------------------------------- */
void orc_matchlist_init(orc_matchlist *m) { memset(m, 0, sizeof *m); }
void orc_matchlist_free(orc_matchlist *m) {
    free(m->v);
    free(m->pool);
    memset(m, 0, sizeof *m);
}
static void ml_clear(orc_matchlist *m) { m->n = 0; m->pn = 0; }

static void ml_push(orc_matchlist *m, int64_t pos, int32_t mm, int32_t len, const int32_t *idx,
                    int32_t idx_n) {
    if (m->n == m->cap) {
        m->cap = m->cap ? m->cap * 2 : 16;
        m->v = (orc_match *)realloc(m->v, m->cap * sizeof *m->v);
    }
    if (m->pn + (size_t)idx_n > m->pcap) {
        while (m->pn + (size_t)idx_n > m->pcap) m->pcap = m->pcap ? m->pcap * 2 : 64;
        m->pool = (int32_t *)realloc(m->pool, m->pcap * sizeof *m->pool);
    }
    orc_match *x = &m->v[m->n++];
    x->pos = pos;
    x->mm = mm;
    x->len = len;
    x->idx_off = (int64_t)m->pn;
    x->idx_n = idx_n;
    x->pad = 0;
    if (idx_n > 0) memcpy(m->pool + m->pn, idx, (size_t)idx_n * sizeof *idx);
    m->pn += (size_t)idx_n;
}

void orc_products_init(orc_products *p) { memset(p, 0, sizeof *p); }
void orc_products_free(orc_products *p) {
    free(p->v);
    free(p->pool);
    memset(p, 0, sizeof *p);
}
void orc_products_clear(orc_products *p) { p->n = 0; p->pn = 0; }

static void pr_push(orc_products *o, int32_t pair_idx, int32_t type, int64_t start, int64_t end,
                    int64_t length, const orc_matchlist *fl, const orc_match *f,
                    const orc_matchlist *rl, const orc_match *r, int64_t rev_primer_len) {
    if (o->n == o->cap) {
        o->cap = o->cap ? o->cap * 2 : 16;
        o->v = (orc_product *)realloc(o->v, o->cap * sizeof *o->v);
    }
    size_t need = (size_t)f->idx_n + (size_t)r->idx_n;
    if (o->pn + need > o->pcap) {
        while (o->pn + need > o->pcap) o->pcap = o->pcap ? o->pcap * 2 : 64;
        o->pool = (int32_t *)realloc(o->pool, o->pcap * sizeof *o->pool);
    }
    orc_product *x = &o->v[o->n++];
    x->pair_idx = pair_idx;
    x->type = type;
    x->start = start;
    x->end = end;
    x->length = length;
    x->fwd_mm = f->mm;
    x->rev_mm = r->mm;
    x->fwd_idx_off = (int64_t)o->pn;
    x->fwd_idx_n = f->idx_n;
    for (int32_t i = 0; i < f->idx_n; i++) o->pool[o->pn++] = fl->pool[f->idx_off + i];
    x->rev_idx_off = (int64_t)o->pn;
    x->rev_idx_n = r->idx_n;
    /* flip (engine.go:127-136): out[i] = n-1-idx[i], order preserved */
    for (int32_t i = 0; i < r->idx_n; i++)
        o->pool[o->pn++] = (int32_t)(rev_primer_len - 1 - rl->pool[r->idx_off + i]);
}

/* ---- match.go:30-90 ---------------------------------------------------- */
static int is_unambiguous(const uint8_t *p, int64_t n) { /* match.go:17-24 */
    for (int64_t i = 0; i < n; i++)
        if (p[i] != 'A' && p[i] != 'C' && p[i] != 'G' && p[i] != 'T') return 0;
    return 1;
}

void orc_find_matches(const uint8_t *seq, int64_t n, const uint8_t *primer, int64_t pl,
                      int64_t max_mm, int64_t cap_hits, int64_t tw, orc_matchlist *out) {
    init_tables();
    ml_clear(out);
    if (pl == 0 || n < pl) return;

    /* exact fast path (match.go:38-53): bytes.Index jump scan, overlapping hits */
    if (max_mm == 0 && is_unambiguous(primer, pl)) {
        for (int64_t pos = 0; pos + pl <= n; pos++) {
            if (memcmp(seq + pos, primer, (size_t)pl) != 0) continue;
            ml_push(out, pos, 0, (int32_t)pl, NULL, 0);
            if (cap_hits > 0 && (int64_t)out->n >= cap_hits) break;
        }
        return;
    }

    int64_t end = n - pl;
    int64_t cutoff = pl - tw; /* match.go:59-66 */
    if (tw <= 0) cutoff = pl + 1;
    if (cutoff < 0) cutoff = 0;

    int32_t *idx = (int32_t *)malloc((size_t)(pl > 0 ? pl : 1) * sizeof *idx);
    for (int64_t pos = 0; pos <= end; pos++) {
        int64_t mm = 0;
        int ok = 1;
        for (int64_t j = 0; j < pl; j++) {
            if (!orc_base_match(seq[pos + j], primer[j])) {
                if (j >= cutoff) { ok = 0; break; }
                idx[mm] = (int32_t)j;
                mm++;
                if (mm > max_mm) { ok = 0; break; }
            }
        }
        if (!ok) continue;
        ml_push(out, pos, (int32_t)mm, (int32_t)pl, idx, (int32_t)mm);
        if (cap_hits > 0 && (int64_t)out->n >= cap_hits) break;
    }
    free(idx);
}

/* engine.go:53-68 */
void orc_filter_left_tw(orc_matchlist *ms, int64_t tw) {
    if (tw <= 0) return;
    size_t w = 0;
    for (size_t i = 0; i < ms->n; i++) {
        const orc_match *m = &ms->v[i];
        int drop = 0;
        for (int32_t k = 0; k < m->idx_n; k++)
            if (ms->pool[m->idx_off + k] < tw) { drop = 1; break; }
        if (!drop) ms->v[w++] = *m;
    }
    ms->n = w;
}

/* ---- engine.go:70-93 ---------------------------------------------------- */
typedef struct { int64_t pos; size_t i; } pos_key;
static int cmp_pos_key(const void *a, const void *b) {
    const pos_key *x = (const pos_key *)a, *y = (const pos_key *)b;
    if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
    if (x->i != y->i) return x->i < y->i ? -1 : 1; /* SliceStable: keep input order on ties */
    return 0;
}

static void sort_matches_by_pos(orc_matchlist *ms) { /* engine.go:70-85, stable */
    int sorted = 1;
    for (size_t i = 1; i < ms->n; i++)
        if (ms->v[i].pos < ms->v[i - 1].pos) { sorted = 0; break; }
    if (sorted) return;
    size_t n = ms->n;
    pos_key *k = (pos_key *)malloc(n * sizeof *k);
    orc_match *tmp = (orc_match *)malloc(n * sizeof *tmp);
    for (size_t i = 0; i < n; i++) { k[i].pos = ms->v[i].pos; k[i].i = i; }
    qsort(k, n, sizeof *k, cmp_pos_key);
    for (size_t i = 0; i < n; i++) tmp[i] = ms->v[k[i].i];
    memcpy(ms->v, tmp, n * sizeof *tmp);
    free(tmp);
    free(k);
}

static int64_t lower_bound_pos(const orc_matchlist *ms, int64_t pos) { /* engine.go:87-89 */
    int64_t lo = 0, hi = (int64_t)ms->n;
    while (lo < hi) {
        int64_t mid = lo + (hi - lo) / 2;
        if (ms->v[mid].pos >= pos) hi = mid; else lo = mid + 1;
    }
    return lo;
}
static int64_t upper_bound_pos(const orc_matchlist *ms, int64_t pos) { /* engine.go:91-93 */
    int64_t lo = 0, hi = (int64_t)ms->n;
    while (lo < hi) {
        int64_t mid = lo + (hi - lo) / 2;
        if (ms->v[mid].pos > pos) hi = mid; else lo = mid + 1;
    }
    return lo;
}

/* One direction of forEachJoinedProduct.  `fw` are the forward-orientation
 * sites (A for "forward", B for "revcomp"), `rv` the partner rc sites, `rlen`
 * the partner primer length.  engine.go:142-272 and :274-402 are the same code
 * with the roles of A and B exchanged. */
static void join_dir(const orc_cfg *cfg, int64_t seq_len, int64_t minL, int64_t maxL,
                     int32_t pair_idx, int32_t type, const orc_matchlist *fw,
                     orc_matchlist *rv, int64_t rlen, orc_products *out) {
    sort_matches_by_pos(rv); /* engine.go:143 / :275 */
    for (size_t a = 0; a < fw->n; a++) {
        const orc_match *ma = &fw->v[a];
        int64_t last = seq_len - rlen;
        int64_t lo = ma->pos + 1;
        if (minL > 0) {
            lo = ma->pos + minL - rlen;
            if (lo <= ma->pos) lo = ma->pos + 1;
        }
        if (lo < 0) lo = 0;
        int64_t hi = last;
        if (maxL > 0) {
            hi = ma->pos + maxL - rlen;
            if (hi > last) hi = last;
        }
        if (hi >= lo) {
            int64_t iMin = lower_bound_pos(rv, lo);
            int64_t iMax = upper_bound_pos(rv, hi) - 1;
            if (iMin <= iMax) {
                for (int64_t j = iMax; j >= iMin; j--) { /* descending: engine.go:168 */
                    const orc_match *mb = &rv->v[j];
                    int64_t end = mb->pos + rlen;
                    int64_t length = end - ma->pos;
                    if ((minL != 0 && length < minL) || (maxL != 0 && length > maxL)) continue;
                    pr_push(out, pair_idx, type, ma->pos, end, length, fw, ma, rv, mb, rlen);
                }
            }
        }
        if (cfg->circular) { /* engine.go:209-271 */
            int64_t X = seq_len - ma->pos;
            int64_t loWrap = 0;
            if (minL > 0) {
                int64_t needed = minL - X - rlen;
                if (needed < 0) needed = 0;
                loWrap = needed;
            }
            int64_t hiWrap = ma->pos - 1;
            if (maxL > 0) {
                int64_t allowed = maxL - X - rlen;
                if (allowed < hiWrap) hiWrap = allowed;
            }
            if (hiWrap >= loWrap) {
                int64_t iMinW = lower_bound_pos(rv, loWrap);
                int64_t iMaxW = upper_bound_pos(rv, hiWrap) - 1;
                for (int64_t j = iMaxW; j >= iMinW; j--) {
                    const orc_match *mb = &rv->v[j];
                    if (mb->pos >= ma->pos) continue;
                    int64_t end = mb->pos + rlen;
                    int64_t length = (seq_len - ma->pos) + end;
                    if ((minL != 0 && length < minL) || (maxL != 0 && length > maxL)) continue;
                    pr_push(out, pair_idx, type, ma->pos, end, length, fw, ma, rv, mb, rlen);
                }
            }
        }
    }
}

/* engine.go:108-404 */
void orc_join(const orc_cfg *cfg, int64_t seq_len, const orc_pair *p, int32_t pair_idx,
              const orc_matchlist *fwdA, const orc_matchlist *fwdB, orc_matchlist *revA,
              orc_matchlist *revB, orc_products *out) {
    int64_t minL = p->min_product, maxL = p->max_product; /* engine.go:113-120 */
    if (minL == 0) minL = cfg->min_len;
    if (maxL == 0) maxL = cfg->max_len;
    /* A x rc(B) => "forward"; partner primer is B */
    join_dir(cfg, seq_len, minL, maxL, pair_idx, 0, fwdA, revB, p->rev_len, out);
    /* B x rc(A) => "revcomp"; partner primer is A */
    join_dir(cfg, seq_len, minL, maxL, pair_idx, 1, fwdB, revA, p->fwd_len, out);
}

/* bruteforce.go:11-38 */
int orc_simulate_bruteforce(const orc_cfg *cfg, const orc_pair *pairs, int64_t npairs,
                            const uint8_t *seq, int64_t n, orc_products *out) {
    orc_products_clear(out);
    if (npairs == 0) return 0;
    orc_matchlist fA, fB, rA, rB;
    orc_matchlist_init(&fA);
    orc_matchlist_init(&fB);
    orc_matchlist_init(&rA);
    orc_matchlist_init(&rB);
    int rc = 0;
    int64_t hc = cfg->hit_cap, tw = cfg->terminal_window;
    for (int64_t i = 0; i < npairs && rc == 0; i++) {
        const orc_pair *p = &pairs[i];
        uint8_t *rca = (uint8_t *)malloc((size_t)(p->fwd_len > 0 ? p->fwd_len : 1));
        uint8_t *rcb = (uint8_t *)malloc((size_t)(p->rev_len > 0 ? p->rev_len : 1));
        if (orc_revcomp(p->fwd, p->fwd_len, rca) != 0 || orc_revcomp(p->rev, p->rev_len, rcb) != 0) {
            rc = -1; /* RevComp panics in the reference (rc.go:28-34) */
        } else {
            orc_find_matches(seq, n, p->fwd, p->fwd_len, cfg->max_mm, hc, tw, &fA);
            orc_find_matches(seq, n, p->rev, p->rev_len, cfg->max_mm, hc, tw, &fB);
            orc_find_matches(seq, n, rca, p->fwd_len, cfg->max_mm, hc, 0, &rA);
            orc_filter_left_tw(&rA, tw);
            orc_find_matches(seq, n, rcb, p->rev_len, cfg->max_mm, hc, 0, &rB);
            orc_filter_left_tw(&rB, tw);
            orc_join(cfg, n, p, (int32_t)i, &fA, &fB, &rA, &rB, out);
        }
        free(rca);
        free(rcb);
    }
    orc_matchlist_free(&fA);
    orc_matchlist_free(&fB);
    orc_matchlist_free(&rA);
    orc_matchlist_free(&rB);
    return rc;
}
