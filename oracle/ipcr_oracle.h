/*
 * ipcr_oracle.h -- CPU restatement of KPU-AGC/ipcr's core/engine scan path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference leg may load this library, and there only as the checker or as the
 * timed CPU baseline -- never as a fallback for the CUDA path.
 *
 * Two layers (all citations are file:line under /root/reference):
 *   1. "truth"  : core/primer/{iupac,rc,match}.go + core/engine/{engine,bruteforce}.go
 *                 == SimulateBatchBruteForce, the reference's own correctness oracle.
 *   2. "seeded" : core/engine/{seed,ac,halo,hit_collect,compiled}.go
 *                 == the production ForEachCompiledProduct path (used as the CPU
 *                 timing baseline and to pin emission order / hit-cap semantics).
 *
 * Parity pinning: the Go reference cannot be built in this image (no Go
 * toolchain), so this restatement is pinned against every known-answer test the
 * reference holds for the path (tests/test_oracle_golden.py lists them with
 * file:line) and layer 2 is cross-checked against layer 1 on the reference's own
 * oracle-equivalence inputs (core/engine/approx_seed_oracle_test.go:88-259).
 */
#ifndef IPCR_ORACLE_H
#define IPCR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* engine.Config (core/engine/engine.go:10-19) */
typedef struct {
    int64_t max_mm;
    int64_t terminal_window;
    int64_t min_len;
    int64_t max_len;
    int64_t hit_cap;
    int64_t seed_len;
    int32_t circular;
    int32_t need_sites; /* unused by the oracle: sites are slices of seq */
} orc_cfg;

/* primer.Pair (core/primer/pair.go:4-10); sequences are NOT NUL-terminated */
typedef struct {
    const uint8_t *fwd;
    int64_t fwd_len;
    const uint8_t *rev;
    int64_t rev_len;
    int64_t min_product;
    int64_t max_product;
} orc_pair;

/* primer.Match (core/primer/match.go:8-13); idx list lives in orc_matchlist.pool */
typedef struct {
    int64_t pos;
    int32_t mm;
    int32_t len;
    int64_t idx_off;
    int32_t idx_n;
    int32_t pad;
} orc_match;

typedef struct {
    orc_match *v;
    size_t n, cap;
    int32_t *pool;
    size_t pn, pcap;
} orc_matchlist;

/* engine.Product (core/engine/product.go:4-35), printable fields only.
 * type: 0 = "forward", 1 = "revcomp". idx lists live in orc_products.pool. */
typedef struct {
    int32_t pair_idx;
    int32_t type;
    int64_t start, end, length;
    int32_t fwd_mm, rev_mm;
    int64_t fwd_idx_off, rev_idx_off;
    int32_t fwd_idx_n, rev_idx_n;
} orc_product;

typedef struct {
    orc_product *v;
    size_t n, cap;
    int32_t *pool;
    size_t pn, pcap;
} orc_products;

/* ---- layer 1: truth ---------------------------------------------------- */
uint8_t orc_iupac_mask(uint8_t c);                       /* iupac.go:6-58 */
int orc_base_match(uint8_t g, uint8_t p);                /* iupac.go:62-67 */
int orc_mismatch_count(const uint8_t *g, const uint8_t *p, int64_t n); /* mismatch.go:3-14 */
/* returns 0, or 1+position of the first invalid byte (rc.go:41-56) */
int64_t orc_revcomp(const uint8_t *in, int64_t n, uint8_t *out);

void orc_matchlist_init(orc_matchlist *m);
void orc_matchlist_free(orc_matchlist *m);
void orc_products_init(orc_products *p);
void orc_products_free(orc_products *p);
void orc_products_clear(orc_products *p);

/* primer.FindMatches (match.go:30-90) */
void orc_find_matches(const uint8_t *seq, int64_t n, const uint8_t *primer, int64_t pl,
                      int64_t max_mm, int64_t cap_hits, int64_t terminal_window,
                      orc_matchlist *out);
/* filterLeftTW (engine.go:53-68); in-place */
void orc_filter_left_tw(orc_matchlist *ms, int64_t tw);
/* forEachJoinedProduct (engine.go:108-404) */
void orc_join(const orc_cfg *cfg, int64_t seq_len, const orc_pair *p, int32_t pair_idx,
              const orc_matchlist *fwdA, const orc_matchlist *fwdB,
              orc_matchlist *revA, orc_matchlist *revB, orc_products *out);
/* SimulateBatchBruteForce (bruteforce.go:11-38). returns 0 or -1 (invalid primer) */
int orc_simulate_bruteforce(const orc_cfg *cfg, const orc_pair *pairs, int64_t npairs,
                            const uint8_t *seq, int64_t n, orc_products *out);

/* ---- layer 2: seeded production path ----------------------------------- */
typedef struct orc_panel orc_panel;
/* CompilePanel (compiled.go:96-136). NULL on invalid primer. */
orc_panel *orc_compile_panel(const orc_cfg *cfg, const orc_pair *pairs, int64_t npairs);
void orc_panel_free(orc_panel *cp);
/* compiledHas (compiled.go:35-37); which in 'A','B','a','b' */
int orc_panel_have(const orc_panel *cp, int64_t pair_idx, int which);
int64_t orc_panel_num_patterns(const orc_panel *cp);
int64_t orc_panel_num_nodes(const orc_panel *cp);
/* seed span chosen for an orientation (seed.go:260-276): returns 1 if ok */
int orc_panel_seed_span(const orc_panel *cp, int64_t pair_idx, int which,
                        int64_t *off, int64_t *len);
/* ForEachCompiledProduct (compiled.go:162-267), products in emission order */
void orc_simulate_compiled(const orc_panel *cp, const uint8_t *seq, int64_t n,
                           orc_products *out);

/* pipeline.ForEachProduct worker model (internal/pipeline/pipeline.go:60-125)
 * restated for TIMING: one job per record, or per rolling overlapped chunk
 * (core/fasta/path_ctx.go:83-179) when chunk_size > overlap > =0; `threads`
 * pthreads pull jobs from a shared counter.  Returns the number of products
 * (after chunk-local emission; no cross-chunk dedupe).  */
int64_t orc_scan_threaded(const orc_panel *cp, const uint8_t *const *seqs,
                          const int64_t *lens, int64_t nrec, int64_t chunk_size,
                          int64_t overlap, int threads);

#ifdef __cplusplus
}
#endif
#endif
