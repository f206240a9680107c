/*
 * ipcr_seeded.c -- layer 2 of the CPU oracle: the reference's PRODUCTION scan
 * path restated in C.  TEST INFRASTRUCTURE ONLY (checker + timed CPU baseline).
 *
 *   seed choice / variant enumeration / 2-bit dedupe  core/engine/seed.go:104-367
 *   Aho-Corasick build + scan, verifyAt               core/engine/ac.go:16-213
 *   non-ACGT halos                                    core/engine/halo.go:9-108
 *   matchCollector (hit cap + attempted-start dedupe) core/engine/hit_collect.go:73-148
 *   CompilePanel / ForEachCompiledProduct             core/engine/compiled.go:96-267
 *   worker pool over records / rolling chunks         internal/pipeline/pipeline.go:60-125,
 *                                                     core/fasta/path_ctx.go:83-179
 * The join is layer 1's orc_join (engine.go:108-404).
 */
#include "ipcr_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---- small dynamic helpers --------------------------------------------- */
#define VEC_PUSH(ptr, n, cap, val)                                              \
    do {                                                                        \
        if ((n) == (cap)) {                                                     \
            (cap) = (cap) ? (cap) * 2 : 16;                                     \
            (ptr) = realloc((ptr), (size_t)(cap) * sizeof *(ptr));              \
        }                                                                       \
        (ptr)[(n)++] = (val);                                                   \
    } while (0)

/* ac.go:16-30 */
static int8_t g_base_code[256];
static int g_bc_ready = 0;
static void init_base_code(void) {
    if (g_bc_ready) return;
    memset(g_base_code, -1, sizeof g_base_code);
    g_base_code['A'] = 0; g_base_code['C'] = 1; g_base_code['G'] = 2; g_base_code['T'] = 3;
    g_base_code['a'] = 0; g_base_code['c'] = 1; g_base_code['g'] = 2; g_base_code['t'] = 3;
    g_bc_ready = 1;
}

/* seed.go:11-16 */
typedef struct {
    int32_t pair_idx;
    uint8_t which; /* 'A','B','a','b' */
    int32_t primer_len;
    int32_t seed_offset;
} seed_payload;

/* seed.go:21-24 (Pat is kept 2-bit encoded; seed.go:29-57) */
typedef struct {
    uint64_t code;
    uint8_t len;
    seed_payload *payloads;
    int64_t np, pcap;
} seed_pattern;

/* ac.go:34-39 */
typedef struct {
    uint32_t next[4];
    uint32_t fail;
    uint32_t out_start;
    uint32_t out_len;
} ac_node;

typedef struct {
    uint8_t *bytes;
    int64_t len;
} oseq;

struct orc_panel {
    orc_cfg cfg;
    int64_t npairs;
    orc_pair *pairs;      /* deep copies */
    oseq *fwdA, *fwdB, *rcA, *rcB;
    uint8_t *have;        /* orientation bitmask per pair (compiled.go:8-13) */
    int64_t *seed_off[4]; /* per pair, per orientation (A,B,a,b); -1 = unseeded */
    int64_t *seed_len[4];
    seed_pattern *patterns;
    int64_t npat, patcap;
    ac_node *nodes;
    int64_t nnodes;
    uint32_t *out;
    int64_t nout;
};

static int which_slot(int which) {
    switch (which) {
    case 'A': return 0;
    case 'B': return 1;
    case 'a': return 2;
    case 'b': return 3;
    }
    return -1;
}

/* ---- seed.go:104-115 ---------------------------------------------------- */
static int64_t configured_seed_len(int64_t primer_len, int64_t cfg) {
    if (cfg <= 0) return primer_len < 12 ? primer_len : 12;
    if (cfg > primer_len) return primer_len;
    return cfg;
}

/* seed.go:285-302 */
static uint64_t exact_variant_estimate(const uint8_t *seed, int64_t n) {
    static const uint8_t alpha[4] = {'A', 'C', 'G', 'T'};
    uint64_t score = 1;
    for (int64_t i = 0; i < n; i++) {
        uint64_t k = 0;
        for (int b = 0; b < 4; b++)
            if (orc_base_match(alpha[b], seed[i])) k++;
        if (k == 0) k = 5;
        score *= k; /* wraps like Go uint64 */
    }
    return score;
}

/* seed.go:260-283 */
static int choose_seed_span(const uint8_t *pat, int64_t plen, int64_t seed_len_cfg,
                            int prefer_right, int64_t *off_out, int64_t *len_out) {
    int64_t length = configured_seed_len(plen, seed_len_cfg);
    if (length <= 0 || length > plen) return 0;
    int64_t best_off = 0;
    uint64_t best_score = 0;
    for (int64_t off = 0; off <= plen - length; off++) {
        uint64_t score = exact_variant_estimate(pat + off, length);
        int tie = prefer_right ? (off > best_off) : (off < best_off);
        if (off == 0 || score < best_score || (score == best_score && tie)) {
            best_off = off;
            best_score = score;
        }
    }
    *off_out = best_off;
    *len_out = length;
    return 1;
}

/* seed.go:304-367 : DFS over A,C,G,T; emits 2-bit codes when len<=32, else flags */
typedef struct {
    const uint8_t *seed;
    int64_t seed_n, seed_offset, primer_len, max_mm, left_tw, right_tw, max_variants;
    int64_t emitted;
    int ok;
    uint64_t *codes;
    int64_t ncodes, ccap;
    uint8_t *buf;
} enum_ctx;

static void enum_walk(enum_ctx *c, int64_t pos, int64_t mm) {
    static const uint8_t alpha[4] = {'A', 'C', 'G', 'T'};
    if (!c->ok || mm > c->max_mm) return;
    if (pos == c->seed_n) {
        if (c->emitted >= c->max_variants) { c->ok = 0; return; }
        uint64_t code = 0;
        if (c->seed_n <= 32)
            for (int64_t i = 0; i < c->seed_n; i++) code = (code << 2) | (uint64_t)g_base_code[c->buf[i]];
        VEC_PUSH(c->codes, c->ncodes, c->ccap, code);
        c->emitted++;
        return;
    }
    int64_t full = c->seed_offset + pos;
    int prot = (c->left_tw > 0 && full < c->left_tw) ||
               (c->right_tw > 0 && full >= c->primer_len - c->right_tw);
    for (int b = 0; b < 4; b++) {
        int64_t cost = orc_base_match(alpha[b], c->seed[pos]) ? 0 : 1;
        if (prot && cost != 0) continue;
        if (mm + cost > c->max_mm) continue;
        c->buf[pos] = alpha[b];
        enum_walk(c, pos + 1, mm + cost);
    }
}

/* ---- seedPatternBuilder (seed.go:233-258): open-addressing (code,len)->idx */
typedef struct {
    int64_t *slots; /* pattern index + 1, 0 = empty */
    int64_t nslots, used;
} pat_map;

static uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

static void pat_map_grow(pat_map *m, const seed_pattern *pats) {
    int64_t nn = m->nslots ? m->nslots * 2 : 1024;
    int64_t *ns = (int64_t *)calloc((size_t)nn, sizeof *ns);
    for (int64_t i = 0; i < m->nslots; i++) {
        int64_t v = m->slots[i];
        if (!v) continue;
        const seed_pattern *p = &pats[v - 1];
        uint64_t h = mix64(p->code ^ ((uint64_t)p->len << 58)) & (uint64_t)(nn - 1);
        while (ns[h]) h = (h + 1) & (uint64_t)(nn - 1);
        ns[h] = v;
    }
    free(m->slots);
    m->slots = ns;
    m->nslots = nn;
}

static int64_t builder_add(orc_panel *cp, pat_map *m, uint64_t code, uint8_t len, seed_payload pl) {
    if ((m->used + 1) * 2 > m->nslots) pat_map_grow(m, cp->patterns);
    uint64_t h = mix64(code ^ ((uint64_t)len << 58)) & (uint64_t)(m->nslots - 1);
    while (m->slots[h]) {
        seed_pattern *p = &cp->patterns[m->slots[h] - 1];
        if (p->code == code && p->len == len) {
            VEC_PUSH(p->payloads, p->np, p->pcap, pl);
            return m->slots[h] - 1;
        }
        h = (h + 1) & (uint64_t)(m->nslots - 1);
    }
    seed_pattern np;
    memset(&np, 0, sizeof np);
    np.code = code;
    np.len = len;
    VEC_PUSH(np.payloads, np.np, np.pcap, pl);
    VEC_PUSH(cp->patterns, cp->npat, cp->patcap, np);
    m->slots[h] = cp->npat;
    m->used++;
    return cp->npat - 1;
}

/* seed.go:170-213 addOrientation */
static void add_orientation(orc_panel *cp, pat_map *m, int64_t pair_idx, int which, const oseq *pat,
                            int prefer_right, int64_t left_tw, int64_t right_tw) {
    if (pat->len == 0) return;
    int64_t off, sl;
    if (!choose_seed_span(pat->bytes, pat->len, cp->cfg.seed_len, prefer_right, &off, &sl)) return;
    enum_ctx c;
    memset(&c, 0, sizeof c);
    c.seed = pat->bytes + off;
    c.seed_n = sl;
    c.seed_offset = off;
    c.primer_len = pat->len;
    c.max_mm = cp->cfg.max_mm < 0 ? 0 : cp->cfg.max_mm;
    c.left_tw = left_tw < 0 ? 0 : left_tw;
    c.right_tw = right_tw < 0 ? 0 : right_tw;
    c.max_variants = 50000; /* seed.go:97 */
    c.ok = 1;
    c.buf = (uint8_t *)malloc((size_t)sl);
    enum_walk(&c, 0, 0);
    free(c.buf);
    if (!(c.ok && c.emitted > 0)) { free(c.codes); return; }
    if (sl > 32) { free(c.codes); return; } /* seed.go:202-207: first add fails */
    seed_payload pl = {(int32_t)pair_idx, (uint8_t)which, (int32_t)pat->len, (int32_t)off};
    for (int64_t i = 0; i < c.ncodes; i++) builder_add(cp, m, c.codes[i], (uint8_t)sl, pl);
    free(c.codes);
    int s = which_slot(which);
    cp->have[pair_idx] |= (uint8_t)(1u << s);
    cp->seed_off[s][pair_idx] = off;
    cp->seed_len[s][pair_idx] = sl;
}

/* ---- ac.go:64-139 -------------------------------------------------------- */
typedef struct {
    uint32_t next[4];
    uint32_t fail;
    uint32_t *out;
    uint32_t nout, ocap;
} build_node;

static void build_ac(orc_panel *cp) {
    int64_t nb = 1, bcap = 1024;
    build_node *b = (build_node *)calloc((size_t)bcap, sizeof *b);
    for (int64_t i = 0; i < cp->npat; i++) {
        const seed_pattern *p = &cp->patterns[i];
        uint32_t cur = 0;
        for (int k = p->len - 1; k >= 0; k--) {
            int code = (int)((p->code >> (2 * k)) & 3);
            if (b[cur].next[code] == 0) {
                if (nb == bcap) {
                    bcap *= 2;
                    b = (build_node *)realloc(b, (size_t)bcap * sizeof *b);
                    memset(b + nb, 0, (size_t)(bcap - nb) * sizeof *b);
                }
                b[cur].next[code] = (uint32_t)nb;
                nb++;
            }
            cur = b[cur].next[code];
        }
        VEC_PUSH(b[cur].out, b[cur].nout, b[cur].ocap, (uint32_t)i);
    }
    uint32_t *queue = (uint32_t *)malloc((size_t)nb * sizeof *queue);
    int64_t qh = 0, qt = 0;
    for (int c = 0; c < 4; c++) {
        uint32_t child = b[0].next[c];
        if (child) { b[child].fail = 0; queue[qt++] = child; }
    }
    while (qh < qt) {
        uint32_t r = queue[qh++];
        for (int c = 0; c < 4; c++) {
            uint32_t s = b[r].next[c];
            if (!s) continue;
            queue[qt++] = s;
            uint32_t f = b[r].fail;
            while (f > 0 && b[f].next[c] == 0) f = b[f].fail;
            if (b[f].next[c] != 0) f = b[f].next[c];
            b[s].fail = f;
            for (uint32_t k = 0; k < b[f].nout; k++) VEC_PUSH(b[s].out, b[s].nout, b[s].ocap, b[f].out[k]);
        }
    }
    free(queue);
    int64_t out_count = 0;
    for (int64_t i = 0; i < nb; i++) out_count += b[i].nout;
    cp->nodes = (ac_node *)malloc((size_t)nb * sizeof *cp->nodes);
    cp->out = (uint32_t *)malloc((size_t)(out_count ? out_count : 1) * sizeof *cp->out);
    cp->nnodes = nb;
    cp->nout = 0;
    for (int64_t i = 0; i < nb; i++) {
        ac_node *n = &cp->nodes[i];
        memcpy(n->next, b[i].next, sizeof n->next);
        n->fail = b[i].fail;
        n->out_start = (uint32_t)cp->nout;
        n->out_len = b[i].nout;
        for (uint32_t k = 0; k < b[i].nout; k++) cp->out[cp->nout++] = b[i].out[k];
        free(b[i].out);
    }
    free(b);
}

/* ---- compiled.go:96-136 --------------------------------------------------- */
static int make_oseq(oseq *o, const uint8_t *s, int64_t n) {
    o->len = n;
    o->bytes = (uint8_t *)malloc((size_t)(n > 0 ? n : 1));
    if (n > 0) memcpy(o->bytes, s, (size_t)n);
    return 0;
}

orc_panel *orc_compile_panel(const orc_cfg *cfg, const orc_pair *pairs, int64_t npairs) {
    init_base_code();
    orc_panel *cp = (orc_panel *)calloc(1, sizeof *cp);
    cp->cfg = *cfg;
    cp->npairs = npairs;
    if (npairs == 0) return cp;
    cp->pairs = (orc_pair *)calloc((size_t)npairs, sizeof *cp->pairs);
    cp->fwdA = (oseq *)calloc((size_t)npairs, sizeof(oseq));
    cp->fwdB = (oseq *)calloc((size_t)npairs, sizeof(oseq));
    cp->rcA = (oseq *)calloc((size_t)npairs, sizeof(oseq));
    cp->rcB = (oseq *)calloc((size_t)npairs, sizeof(oseq));
    cp->have = (uint8_t *)calloc((size_t)npairs, 1);
    for (int s = 0; s < 4; s++) {
        cp->seed_off[s] = (int64_t *)malloc((size_t)npairs * sizeof(int64_t));
        cp->seed_len[s] = (int64_t *)malloc((size_t)npairs * sizeof(int64_t));
        for (int64_t i = 0; i < npairs; i++) { cp->seed_off[s][i] = -1; cp->seed_len[s][i] = 0; }
    }
    int bad = 0;
    for (int64_t i = 0; i < npairs; i++) {
        make_oseq(&cp->fwdA[i], pairs[i].fwd, pairs[i].fwd_len);
        make_oseq(&cp->fwdB[i], pairs[i].rev, pairs[i].rev_len);
        make_oseq(&cp->rcA[i], pairs[i].fwd, pairs[i].fwd_len);
        make_oseq(&cp->rcB[i], pairs[i].rev, pairs[i].rev_len);
        if (orc_revcomp(pairs[i].fwd, pairs[i].fwd_len, cp->rcA[i].bytes) != 0) bad = 1;
        if (orc_revcomp(pairs[i].rev, pairs[i].rev_len, cp->rcB[i].bytes) != 0) bad = 1;
        cp->pairs[i] = pairs[i];
        cp->pairs[i].fwd = cp->fwdA[i].bytes;
        cp->pairs[i].rev = cp->fwdB[i].bytes;
    }
    if (bad) { orc_panel_free(cp); return NULL; } /* RevComp panics (rc.go:28-34) */

    /* buildSeedPatterns (seed.go:152-231) */
    if (cfg->seed_len >= 0) {
        pat_map m;
        memset(&m, 0, sizeof m);
        int64_t tw = cfg->terminal_window;
        for (int64_t i = 0; i < npairs; i++) {
            add_orientation(cp, &m, i, 'A', &cp->fwdA[i], 1, 0, tw);
            add_orientation(cp, &m, i, 'B', &cp->fwdB[i], 1, 0, tw);
            add_orientation(cp, &m, i, 'a', &cp->rcA[i], 0, tw, 0);
            add_orientation(cp, &m, i, 'b', &cp->rcB[i], 0, tw, 0);
        }
        free(m.slots);
    }
    if (cp->npat > 0) build_ac(cp);
    return cp;
}

void orc_panel_free(orc_panel *cp) {
    if (!cp) return;
    for (int64_t i = 0; i < cp->npairs; i++) {
        if (cp->fwdA) free(cp->fwdA[i].bytes);
        if (cp->fwdB) free(cp->fwdB[i].bytes);
        if (cp->rcA) free(cp->rcA[i].bytes);
        if (cp->rcB) free(cp->rcB[i].bytes);
    }
    free(cp->fwdA); free(cp->fwdB); free(cp->rcA); free(cp->rcB);
    free(cp->pairs); free(cp->have);
    for (int s = 0; s < 4; s++) { free(cp->seed_off[s]); free(cp->seed_len[s]); }
    for (int64_t i = 0; i < cp->npat; i++) free(cp->patterns[i].payloads);
    free(cp->patterns); free(cp->nodes); free(cp->out);
    free(cp);
}

int orc_panel_have(const orc_panel *cp, int64_t pair_idx, int which) {
    int s = which_slot(which);
    if (s < 0 || pair_idx < 0 || pair_idx >= cp->npairs) return 0;
    return (cp->have[pair_idx] >> s) & 1;
}
int64_t orc_panel_num_patterns(const orc_panel *cp) { return cp->npat; }
int64_t orc_panel_num_nodes(const orc_panel *cp) { return cp->nnodes; }
int orc_panel_seed_span(const orc_panel *cp, int64_t pair_idx, int which, int64_t *off, int64_t *len) {
    int s = which_slot(which);
    if (s < 0 || pair_idx < 0 || pair_idx >= cp->npairs || cp->seed_off[s][pair_idx] < 0) return 0;
    *off = cp->seed_off[s][pair_idx];
    *len = cp->seed_len[s][pair_idx];
    return 1;
}

/* ---- hit_collect.go:73-148 -------------------------------------------------- */
typedef struct {
    orc_matchlist matches;
    int64_t starts[8];
    int nstarts;
    int64_t *set; /* open addressing, value+1; NULL until promoted */
    int64_t nset, setcap;
} collector;

static void coll_reset(collector *c) {
    c->matches.n = 0;
    c->matches.pn = 0;
    c->nstarts = 0;
    free(c->set);
    c->set = NULL;
    c->nset = c->setcap = 0;
}

static int set_has(const collector *c, int64_t v) {
    uint64_t h = mix64((uint64_t)v) & (uint64_t)(c->setcap - 1);
    while (c->set[h]) {
        if (c->set[h] == v + 1) return 1;
        h = (h + 1) & (uint64_t)(c->setcap - 1);
    }
    return 0;
}
static void set_add(collector *c, int64_t v) {
    if ((c->nset + 1) * 2 > c->setcap) {
        int64_t nc = c->setcap ? c->setcap * 2 : 64;
        int64_t *ns = (int64_t *)calloc((size_t)nc, sizeof *ns);
        for (int64_t i = 0; i < c->setcap; i++) {
            if (!c->set[i]) continue;
            uint64_t h = mix64((uint64_t)(c->set[i] - 1)) & (uint64_t)(nc - 1);
            while (ns[h]) h = (h + 1) & (uint64_t)(nc - 1);
            ns[h] = c->set[i];
        }
        free(c->set);
        c->set = ns;
        c->setcap = nc;
    }
    uint64_t h = mix64((uint64_t)v) & (uint64_t)(c->setcap - 1);
    while (c->set[h]) h = (h + 1) & (uint64_t)(c->setcap - 1);
    c->set[h] = v + 1;
    c->nset++;
}

static int coll_seen(const collector *c, int64_t start) { /* hit_collect.go:100-111 */
    if (c->set) return set_has(c, start);
    for (int i = 0; i < c->nstarts; i++)
        if (c->starts[i] == start) return 1;
    return 0;
}
static void coll_mark(collector *c, int64_t start) { /* hit_collect.go:113-129 */
    if (c->set) { set_add(c, start); return; }
    if (c->nstarts < 8) { c->starts[c->nstarts++] = start; return; }
    for (int i = 0; i < c->nstarts; i++) set_add(c, c->starts[i]);
    c->nstarts = 0;
    set_add(c, start);
}

/* local copy of the matchlist push (layer 1 keeps its helper static) */
static void ml_push2(orc_matchlist *m, int64_t pos, int32_t mm, int32_t len, const int32_t *idx) {
    if (m->n == m->cap) {
        m->cap = m->cap ? m->cap * 2 : 16;
        m->v = (orc_match *)realloc(m->v, m->cap * sizeof *m->v);
    }
    if (m->pn + (size_t)mm > m->pcap) {
        while (m->pn + (size_t)mm > m->pcap) m->pcap = m->pcap ? m->pcap * 2 : 64;
        m->pool = (int32_t *)realloc(m->pool, m->pcap * sizeof *m->pool);
    }
    orc_match *x = &m->v[m->n++];
    x->pos = pos; x->mm = mm; x->len = len; x->idx_off = (int64_t)m->pn; x->idx_n = mm; x->pad = 0;
    if (mm > 0) memcpy(m->pool + m->pn, idx, (size_t)mm * sizeof *idx);
    m->pn += (size_t)mm;
}

/* ac.go:186-213 verifyAt + hit_collect.go:79-98 addVerified */
static void add_verified(collector *c, const uint8_t *seq, int64_t n, int64_t start, const oseq *pat,
                         int64_t max_mm, int64_t left_tw, int64_t right_tw, int64_t hit_cap,
                         int32_t *idx_scratch) {
    if (hit_cap > 0 && (int64_t)c->matches.n >= hit_cap) return;
    if (coll_seen(c, start)) return;
    coll_mark(c, start);
    int64_t pl = pat->len;
    if (start < 0 || start + pl > n) return;
    int64_t mm = 0, left_cut = left_tw, right_cut = pl - right_tw;
    for (int64_t j = 0; j < pl; j++) {
        if (!orc_base_match(seq[start + j], pat->bytes[j])) {
            if (j < left_cut || j >= right_cut) return;
            idx_scratch[mm] = (int32_t)j;
            mm++;
            if (max_mm >= 0 && mm > max_mm) return;
        }
    }
    ml_push2(&c->matches, start, (int32_t)mm, (int32_t)pl, idx_scratch);
}

/* ---- scratch (hit_collect.go:11-66) --------------------------------------- */
typedef struct {
    collector *coll; /* 4 per pair: fwdA, fwdB, revA, revB */
    orc_matchlist tmp[4];
    int32_t *idx_scratch;
    int64_t npairs;
} scratch_t;

static scratch_t *scratch_new(const orc_panel *cp) {
    scratch_t *s = (scratch_t *)calloc(1, sizeof *s);
    s->npairs = cp->npairs;
    s->coll = (collector *)calloc((size_t)(cp->npairs * 4 + 1), sizeof(collector));
    int64_t maxl = 1;
    for (int64_t i = 0; i < cp->npairs; i++) {
        if (cp->fwdA[i].len > maxl) maxl = cp->fwdA[i].len;
        if (cp->fwdB[i].len > maxl) maxl = cp->fwdB[i].len;
    }
    s->idx_scratch = (int32_t *)malloc((size_t)maxl * sizeof(int32_t));
    return s;
}
static void scratch_free(scratch_t *s) {
    for (int64_t i = 0; i < s->npairs * 4; i++) { coll_reset(&s->coll[i]); orc_matchlist_free(&s->coll[i].matches); }
    for (int k = 0; k < 4; k++) orc_matchlist_free(&s->tmp[k]);
    free(s->coll);
    free(s->idx_scratch);
    free(s);
}

/* ---- compiled.go:162-267 --------------------------------------------------- */
typedef struct {
    const orc_panel *cp;
    scratch_t *s;
    const uint8_t *seq;
    int64_t n;
} scan_ctx;

static void add_hit(scan_ctx *x, int64_t pair_idx, int which, int64_t start) { /* compiled.go:192-207 */
    const orc_panel *cp = x->cp;
    if (pair_idx < 0 || pair_idx >= cp->npairs) return;
    int64_t mm = cp->cfg.max_mm, tw = cp->cfg.terminal_window, hc = cp->cfg.hit_cap;
    collector *c = &x->s->coll[pair_idx * 4];
    switch (which) {
    case 'A': add_verified(&c[0], x->seq, x->n, start, &cp->fwdA[pair_idx], mm, 0, tw, hc, x->s->idx_scratch); break;
    case 'B': add_verified(&c[1], x->seq, x->n, start, &cp->fwdB[pair_idx], mm, 0, tw, hc, x->s->idx_scratch); break;
    case 'a': add_verified(&c[2], x->seq, x->n, start, &cp->rcA[pair_idx], mm, tw, 0, hc, x->s->idx_scratch); break;
    case 'b': add_verified(&c[3], x->seq, x->n, start, &cp->rcB[pair_idx], mm, tw, 0, hc, x->s->idx_scratch); break;
    }
}

/* halo.go:35-74 */
static void halo_starts(scan_ctx *x, int64_t pair_idx, int which, int64_t primer_len,
                        const int64_t *ranges, int64_t nranges) {
    int64_t seq_len = x->n;
    if (seq_len <= 0 || primer_len <= 0 || primer_len > seq_len || nranges == 0) return;
    int64_t limit = seq_len - primer_len;
    int64_t cur_lo = -1, cur_hi = -1;
    for (int64_t r = 0; r < nranges; r++) {
        int64_t rs = ranges[2 * r], re = ranges[2 * r + 1];
        if (re <= rs) continue;
        int64_t lo = rs - primer_len + 1;
        if (lo < 0) lo = 0;
        int64_t hi = re - 1;
        if (hi > limit) hi = limit;
        if (hi < lo) continue;
        if (cur_lo < 0) { cur_lo = lo; cur_hi = hi; continue; }
        if (lo <= cur_hi + 1) { if (hi > cur_hi) cur_hi = hi; continue; }
        for (int64_t st = cur_lo; st <= cur_hi; st++) add_hit(x, pair_idx, which, st);
        cur_lo = lo; cur_hi = hi;
    }
    if (cur_lo >= 0)
        for (int64_t st = cur_lo; st <= cur_hi; st++) add_hit(x, pair_idx, which, st);
}

static void simulate_compiled_scratch(const orc_panel *cp, scratch_t *s, const uint8_t *seq, int64_t n,
                                      orc_products *out) {
    if (cp->npairs == 0) return;
    for (int64_t i = 0; i < cp->npairs * 4; i++) coll_reset(&s->coll[i]);
    const orc_cfg *cfg = &cp->cfg;
    int64_t max_mm = cfg->max_mm, tw = cfg->terminal_window, hit_cap = cfg->hit_cap;

    int has_reset = 0; /* compiled.go:189, ac.go:141-148 */
    if (max_mm > 0)
        for (int64_t i = 0; i < n; i++)
            if (g_base_code[seq[i]] < 0) { has_reset = 1; break; }
    int force_fallback = has_reset && hit_cap > 0;

    scan_ctx x = {cp, s, seq, n};

    if (cp->npat > 0 && cp->nnodes > 0 && !force_fallback) { /* compiled.go:211-223, ac.go:151-180 */
        const ac_node *nodes = cp->nodes;
        uint32_t state = 0;
        for (int64_t i = 0; i < n; i++) {
            int code = g_base_code[seq[i]];
            if (code < 0) { state = 0; continue; }
            while (state > 0 && nodes[state].next[code] == 0) state = nodes[state].fail;
            uint32_t nx = nodes[state].next[code];
            if (nx != 0) state = nx;
            ac_node node = nodes[state]; /* the reference copies the node (ac.go:170) */
            if (node.out_len == 0) continue;
            for (uint32_t k = node.out_start; k < node.out_start + node.out_len; k++) {
                const seed_pattern *pat = &cp->patterns[cp->out[k]];
                for (int64_t q = 0; q < pat->np; q++) {
                    const seed_payload *pl = &pat->payloads[q];
                    int64_t start = i - pl->seed_offset - ((int64_t)pat->len - 1);
                    add_hit(&x, pl->pair_idx, pl->which, start);
                }
            }
        }
    }

    if (has_reset && !force_fallback) { /* compiled.go:230-232, halo.go:76-108 */
        int64_t *ranges = NULL, nr = 0, rcap = 0;
        for (int64_t i = 0; i < n;) {
            if (g_base_code[seq[i]] >= 0) { i++; continue; }
            int64_t st = i;
            while (i < n && g_base_code[seq[i]] < 0) i++;
            VEC_PUSH(ranges, nr, rcap, st);
            VEC_PUSH(ranges, nr, rcap, i);
        }
        int64_t nranges = nr / 2;
        if (nranges > 0 && max_mm > 0) {
            for (int64_t i = 0; i < cp->npairs; i++) {
                if (orc_panel_have(cp, i, 'A')) halo_starts(&x, i, 'A', cp->fwdA[i].len, ranges, nranges);
                if (orc_panel_have(cp, i, 'B')) halo_starts(&x, i, 'B', cp->fwdB[i].len, ranges, nranges);
                if (orc_panel_have(cp, i, 'a')) halo_starts(&x, i, 'a', cp->rcA[i].len, ranges, nranges);
                if (orc_panel_have(cp, i, 'b')) halo_starts(&x, i, 'b', cp->rcB[i].len, ranges, nranges);
            }
        }
        free(ranges);
    }

    /* fallback + join (compiled.go:238-265) */
    for (int64_t i = 0; i < cp->npairs; i++) {
        collector *c = &s->coll[i * 4];
        orc_matchlist *fA = &c[0].matches, *fB = &c[1].matches, *rA = &c[2].matches, *rB = &c[3].matches;
        if (force_fallback || !orc_panel_have(cp, i, 'A')) {
            orc_find_matches(seq, n, cp->fwdA[i].bytes, cp->fwdA[i].len, max_mm, hit_cap, tw, &s->tmp[0]);
            fA = &s->tmp[0];
        }
        if (force_fallback || !orc_panel_have(cp, i, 'B')) {
            orc_find_matches(seq, n, cp->fwdB[i].bytes, cp->fwdB[i].len, max_mm, hit_cap, tw, &s->tmp[1]);
            fB = &s->tmp[1];
        }
        if (force_fallback || !orc_panel_have(cp, i, 'a')) {
            orc_find_matches(seq, n, cp->rcA[i].bytes, cp->rcA[i].len, max_mm, hit_cap, 0, &s->tmp[2]);
            orc_filter_left_tw(&s->tmp[2], tw);
            rA = &s->tmp[2];
        }
        if (force_fallback || !orc_panel_have(cp, i, 'b')) {
            orc_find_matches(seq, n, cp->rcB[i].bytes, cp->rcB[i].len, max_mm, hit_cap, 0, &s->tmp[3]);
            orc_filter_left_tw(&s->tmp[3], tw);
            rB = &s->tmp[3];
        }
        /* NOTE: the reference joins after ALL pairs' fallbacks ran (two loops);
         * per-pair lists are independent so joining here is equivalent. */
        orc_join(cfg, n, &cp->pairs[i], (int32_t)i, fA, fB, rA, rB, out);
    }
}

void orc_simulate_compiled(const orc_panel *cp, const uint8_t *seq, int64_t n, orc_products *out) {
    init_base_code();
    orc_products_clear(out);
    scratch_t *s = scratch_new(cp);
    simulate_compiled_scratch(cp, s, seq, n, out);
    scratch_free(s);
}

/* ---- worker pool for timing (pipeline.go:60-125; path_ctx.go:83-179) ------ */
typedef struct {
    const uint8_t *seq;
    int64_t len;
} job_t;

typedef struct {
    const orc_panel *cp;
    const job_t *jobs;
    int64_t njobs;
    int64_t *next; /* shared counter */
    pthread_mutex_t *mu;
    int64_t nproducts;
} worker_t;

static void *worker_main(void *arg) {
    worker_t *w = (worker_t *)arg;
    scratch_t *s = scratch_new(w->cp);
    orc_products out;
    orc_products_init(&out);
    for (;;) {
        pthread_mutex_lock(w->mu);
        int64_t j = (*w->next)++;
        pthread_mutex_unlock(w->mu);
        if (j >= w->njobs) break;
        orc_products_clear(&out);
        /* the reference copies each record/chunk before handing it to a worker
         * (path_ctx.go:53,117); the copy is part of its cost */
        uint8_t *copy = (uint8_t *)malloc((size_t)(w->jobs[j].len > 0 ? w->jobs[j].len : 1));
        memcpy(copy, w->jobs[j].seq, (size_t)w->jobs[j].len);
        simulate_compiled_scratch(w->cp, s, copy, w->jobs[j].len, &out);
        free(copy);
        w->nproducts += (int64_t)out.n;
    }
    orc_products_free(&out);
    scratch_free(s);
    return NULL;
}

int64_t orc_scan_threaded(const orc_panel *cp, const uint8_t *const *seqs, const int64_t *lens,
                          int64_t nrec, int64_t chunk_size, int64_t overlap, int threads) {
    init_base_code();
    if (threads < 1) threads = 1;
    if (overlap < 0) overlap = 0;
    job_t *jobs = NULL;
    int64_t nj = 0, jcap = 0;
    for (int64_t r = 0; r < nrec; r++) {
        int64_t step = chunk_size - overlap;
        if (chunk_size <= 0 || overlap >= chunk_size || step <= 0 || lens[r] <= chunk_size) {
            job_t j = {seqs[r], lens[r]};
            VEC_PUSH(jobs, nj, jcap, j);
            continue;
        }
        /* rolling window: chunks [k*step, k*step+chunk_size), final remainder */
        int64_t start = 0, last_end = 0;
        while (lens[r] - start > chunk_size) {
            job_t j = {seqs[r] + start, chunk_size};
            VEC_PUSH(jobs, nj, jcap, j);
            last_end = start + chunk_size;
            start += step;
        }
        if (last_end < lens[r]) {
            job_t j = {seqs[r] + start, lens[r] - start};
            VEC_PUSH(jobs, nj, jcap, j);
        }
    }
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    int64_t next = 0;
    worker_t *ws = (worker_t *)calloc((size_t)threads, sizeof *ws);
    pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof *th);
    for (int t = 0; t < threads; t++) {
        ws[t].cp = cp; ws[t].jobs = jobs; ws[t].njobs = nj; ws[t].next = &next; ws[t].mu = &mu;
        pthread_create(&th[t], NULL, worker_main, &ws[t]);
    }
    int64_t total = 0;
    for (int t = 0; t < threads; t++) { pthread_join(th[t], NULL); total += ws[t].nproducts; }
    free(ws); free(th); free(jobs);
    return total;
}
