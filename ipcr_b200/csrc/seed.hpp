// seed.hpp -- host side of the SEED-INDEX filter: the whole panel becomes one k-mer index, so
// one pass over the genome serves every orientation (the bit-parallel filter of plan.hpp /
// jit.hpp needs one pass per group of four).
//
// Same idea as the reference's candidate generator (buildSeedPatterns + enumerateSeedVariants,
// /root/reference/core/engine/seed.go:152-367, fed to an Aho-Corasick automaton, ac.go:64-139),
// re-designed for the GPU: no automaton (a dependent load per base), but
//   * ONE seed window per unique orientation, S <= 16 columns anchored at the protected (3')
//     end, S uniform over the panel so that a genome position has exactly one k-mer;
//   * every concrete k-mer a true site can show in that window (<= max_mm mismatches in the
//     free columns, none in the protected ones, IUPAC columns expanded) is enumerated;
//   * the k-mers go into an open-addressing hash table (8-byte slots: k-mer, orientation,
//     seed offset) that stays L2-resident, fronted by a 224 KB blocked Bloom bitmap that the
//     scan kernel keeps in SHARED memory: one LDS per genome position decides whether the
//     table is consulted at all.
// Completeness: a site with <= max_mm mismatches and a clean protected window shows one of the
// enumerated k-mers in its seed window unless that window holds a non-ACGT byte; those "dirty"
// windows are verified directly (listed by k_seed_scan, expanded by k_seed_dirty_expand), like the reference's halo pass
// (halo.go:47-108).  Exactness still lives in the verify kernel.
#pragma once

#include <cstdio>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <utility>
#include <vector>

#include "plan.hpp"

namespace ipcr {

constexpr uint32_t kSeedBloomWords = 57344;  // 224 KB of shared memory
constexpr uint32_t kSeedBloomBits = kSeedBloomWords * 32u;
constexpr uint32_t kSeedHalfWords = 2048;  // pigeonhole tables: 65536 bits per half window
constexpr uint32_t kSeedBloomMul = 0x85EBCA6Bu;
constexpr uint32_t kSeedHashMul = 0x9E3779B1u;
constexpr int kSeedMaxLen = 16;                    // 2 bits per base in a 32-bit k-mer
constexpr uint64_t kSeedMaxEntries = 16ull << 20;  // table <= 512 MB

// 2-bit code of a base = (ascii >> 1) & 3: A 0, C 1, T 2, G 3; index = one-hot bit (A,C,G,T)
constexpr uint32_t kSeedCodeOfBit[4] = {0u, 1u, 3u, 2u};

// slot = occupied(63) | seed offset(55:48) | orientation(47:32) | k-mer(31:0)
inline uint64_t seed_slot(uint32_t kmer, uint32_t orient, uint32_t col_off) {
    return (1ull << 63) | ((uint64_t)col_off << 48) | ((uint64_t)orient << 32) | kmer;
}

// bitmap bits of a k-mer: the word comes from the top of a multiplicative hash of the k-mer,
// the first bit position from the k-mer's low five bits, the second (k = 2) from hash bits
// 11..15 (a function of the k-mer's low 16 bits, so not determined by the first).
// Returns the word index, sets mask.  Needs S >= 3.
inline uint32_t seed_bloom_mul(int S) { return kSeedBloomMul << (32 - 2 * S); }

inline uint32_t seed_bloom_word(uint32_t kmer, int S, int k, uint32_t &mask) {
    const uint32_t hw = kmer * seed_bloom_mul(S);
    mask = 1u << (kmer & 31u);
    if (k >= 2) mask |= 1u << ((hw >> 11) & 31u);
    return (uint32_t)(((uint64_t)hw * kSeedBloomWords) >> 32);
}

struct SeedIndex {
    int S = 0;
    // first level: the shared-memory bitmap is keyed by the S1 bases at the HIGH end of the S-base seed
    // window (key1 = k-mer >> 2*(S - S1)).  Fewer free columns => far fewer distinct keys than S-mers
    // => a sparser bitmap; the exact (or hashed) S-mer test stays in the second level.
    int S1 = 0;
    // anchor stride of the first level (1, 2 or 4; needs S - S1 >= stride - 1): the bitmap holds, per S-mer, its
    // S1-mers at the window offsets d-stride+1 .. d (d = S - S1), and the kernel looks only at every stride-th
    // position; a hit there stands for the `stride` windows that hold the anchor at one of those offsets.
    int stride = 1;
    uint32_t kmer_mask = 0;
    uint32_t table_bits = 0;
    std::vector<uint64_t> slots;
    std::vector<uint32_t> bloom;
    // second level, L2-resident: one bit per k-mer value (2S <= 28: exact) or per hashed k-mer
    std::vector<uint32_t> bitmap2;
    // h = k-mer * mul (32-bit wrap): word = h >> shift, bit = h & 31.  Exact (mul 1, shift 5) when 2S <= 28.
    uint32_t bitmap2_mul = 1, bitmap2_shift = 5;
    bool use_bloom = false;
    int bloom_k = 1;  // bits per k-mer
    // PIGEONHOLE first level (S = 16; replaces the bitmap when the panel's keys would saturate it, e.g. mm 3 x 400
    // orientations): a window with <= m mismatches has <= m / 2 of them in its low or in its high 8 bases.
    // bloom[0, 2048) = one bit per low 8-mer within m / 2 mismatches of some orientation's low half (protected
    // columns exact, IUPAC expanded), bloom[2048, 4096) the same for the high halves.  pass = either half hits.
    bool pigeon = false;
    double pigeon_pass = 0.0;  // modelled share of positions that go on to the second level
    uint64_t n_entries = 0;
    uint64_t n_keys1 = 0;     // distinct first-level keys
    double bloom_fill = 0.0;  // fraction of bitmap bits set
    double cand_prob = 0.0;   // expected candidates per genome position
    std::vector<uint16_t> col_off;  // per orientation: site start = k-mer start - col_off
};

inline int seed_col_off(const Orient &o, int S) { return o.rc_type ? 0 : o.L - S; }

inline bool seed_col_protected(const Orient &o, int col, int tw, bool raw) {
    if (tw <= 0 || (raw && o.rc_type)) return false;
    return o.rc_type ? col < tw : col >= o.L - tw;
}

// number of k-mers seed_enumerate() emits for one orientation (saturating)
inline uint64_t seed_variant_count(const Orient &o, int S, int max_mm, int tw, bool raw) {
    const uint64_t kSat = 1ull << 40;
    const int off = seed_col_off(o, S);
    std::vector<uint64_t> ways((size_t)max_mm + 1, 0);
    ways[0] = 1;
    for (int j = 0; j < S; j++) {
        const uint64_t km = (uint64_t)class_popcount(o.cls[off + j]);
        const uint64_t kx = seed_col_protected(o, off + j, tw, raw) ? 0 : 4 - km;
        std::vector<uint64_t> nw((size_t)max_mm + 1, 0);
        for (int m = 0; m <= max_mm; m++) {
            if (!ways[m]) continue;
            nw[m] = std::min(kSat, nw[m] + ways[m] * km);
            if (m + 1 <= max_mm) nw[m + 1] = std::min(kSat, nw[m + 1] + ways[m] * kx);
        }
        ways.swap(nw);
    }
    uint64_t total = 0;
    for (uint64_t w : ways) total = std::min(kSat, total + w);
    return total;
}

inline void seed_enumerate(const Orient &o, int S, int max_mm, int tw, bool raw, std::vector<uint32_t> &out) {
    const int off = seed_col_off(o, S);
    struct Frame { int j; uint32_t kmer; int mm; };
    std::vector<Frame> stack;
    stack.push_back({0, 0u, 0});
    while (!stack.empty()) {
        const Frame f = stack.back();
        stack.pop_back();
        if (f.j == S) {
            out.push_back(f.kmer);
            continue;
        }
        const uint8_t cls = o.cls[off + f.j];
        const bool prot = seed_col_protected(o, off + f.j, tw, raw);
        for (int b = 0; b < 4; b++) {
            const bool match = (cls >> b) & 1;
            if (!match && (prot || f.mm >= max_mm)) continue;
            stack.push_back({f.j + 1, f.kmer | (kSeedCodeOfBit[b] << (2 * f.j)), f.mm + (match ? 0 : 1)});
        }
    }
}

// every 8-mer within t mismatches of columns [c0, c0 + 8) of orientation o's S = 16 seed window -> one bit each
inline void seed_half_enumerate(const Orient &o, int c0, int t, int tw, bool raw, std::vector<uint32_t> &table) {
    const int off = seed_col_off(o, 16) + c0;
    struct Frame { int j; uint32_t v; int mm; };
    std::vector<Frame> stack;
    stack.push_back({0, 0u, 0});
    while (!stack.empty()) {
        const Frame f = stack.back();
        stack.pop_back();
        if (f.j == 8) {
            table[f.v >> 5] |= 1u << (f.v & 31u);
            continue;
        }
        const uint8_t cls = o.cls[off + f.j];
        const bool prot = seed_col_protected(o, off + f.j, tw, raw);
        for (int b = 0; b < 4; b++) {
            const bool match = (cls >> b) & 1;
            if (!match && (prot || f.mm >= t)) continue;
            stack.push_back({f.j + 1, f.v | (kSeedCodeOfBit[b] << (2 * f.j)), f.mm + (match ? 0 : 1)});
        }
    }
}

// Modelled cost of the seed filter per genome position, in units of the scan kernel's unstrided fast path (one
// shared-memory bitmap lookup = ~10 warp instructions per 32 positions; 1.29 ms per 3 Gbp on B200).
//  * fast path at anchor stride t: 0.1 + 0.92 / t (1.05 / t with the second bitmap bit);
//  * an anchor that passes the first level sends its t windows through a DIVERGENT second-level lookup (an L2 round
//    trip; the t lookups of one anchor are issued together) while the other lanes of the warp wait: a warp works on
//    32 lanes x 32 / t anchors at a time and loops max-over-lanes(passing anchors) times, so the cost is
//    c * E[max of 32 Binomial(32 / t, pass)], c = 0.1 * (1 + (t - 1) * (0.5 + 8 * pass)): linear in `pass` while passes
//    are rare (a single false positive stalls the whole warp: this is why a strided panel wants the second bitmap
//    bit), flatter once every lane has some; with a stride the shorter fast path hides less of the round trips, the
//    more so the more there are.  Fitted to B200 within ~15 %: profiles/r02_seed_sweep.txt (100 pairs, S1 10..16 x
//    1..2 bits) and profiles/r02_seed_stride.txt (1..100 pairs x stride 1 / 2 / 4).  Beyond pass = 0.08 a stride
//    lost in every measurement (100 pairs, stride 2: 4.04 ms against 2.26) and is not considered;
//  * a candidate costs ~675 (table walk, append, verify).
inline double seed_bloom_fp(double keys, int k) {
    const double fill = 1.0 - std::exp(-(double)k * keys / (double)kSeedBloomBits);
    return std::pow(fill, k);
}

inline double seed_pass_rate(double keys1, int S1, int k) {
    return std::min(1.0, seed_bloom_fp(keys1, k) + keys1 / std::pow(4.0, S1));
}

// E[max over `lanes` independent Binomial(n, p)], n <= 32
inline double seed_expected_max(int n, double p, int lanes = 32) {
    if (p <= 0.0 || n <= 0) return 0.0;
    if (p >= 1.0) return (double)n;
    double pmf = std::pow(1.0 - p, n), cdf = 0.0, prev = 0.0, e = 0.0;
    for (int m = 0; m <= n; m++) {
        cdf = std::min(1.0, cdf + pmf);
        const double f = std::pow(cdf, lanes);
        e += (double)m * (f - prev);
        prev = f;
        pmf *= (double)(n - m) / (double)(m + 1) * p / (1.0 - p);
    }
    return e;
}

inline double seed_walk_cost(double pass, int t) {
    if (pass > 0.5) return 15.0;  // no bitmap: every position asks the second level
    if (t > 1 && pass > 0.08) return 1e9;
    return 0.1 * (1.0 + (double)(t - 1) * (0.5 + 8.0 * pass)) * seed_expected_max(32 / t, pass);
}

inline double seed_cost(double entries, int S, double keys1, int S1, int k, int t = 1) {
    const double prob = entries / std::pow(4.0, S);
    const double pass = seed_pass_rate(keys1, S1, k);
    const double fast = 0.1 + (k == 2 ? 1.05 : 0.92) / (double)t;
    return fast + seed_walk_cost(pass, t) + prob * 675.0;
}

// variants of the S1 columns at the high end of orientation o's S-column seed window (an upper bound on
// its distinct first-level keys)
inline uint64_t seed_key1_count(const Orient &o, int S, int S1, int max_mm, int tw, bool raw, int e = -1) {
    const uint64_t kSat = 1ull << 40;
    const int off = seed_col_off(o, S) + (e < 0 ? S - S1 : e);  // e: offset of the S1 columns inside the window
    std::vector<uint64_t> ways((size_t)max_mm + 1, 0);
    ways[0] = 1;
    for (int j = 0; j < S1; j++) {
        const uint64_t km = (uint64_t)class_popcount(o.cls[off + j]);
        const uint64_t kx = seed_col_protected(o, off + j, tw, raw) ? 0 : 4 - km;
        std::vector<uint64_t> nw((size_t)max_mm + 1, 0);
        for (int m = 0; m <= max_mm; m++) {
            if (!ways[m]) continue;
            nw[m] = std::min(kSat, nw[m] + ways[m] * km);
            if (m + 1 <= max_mm) nw[m + 1] = std::min(kSat, nw[m + 1] + ways[m] * kx);
        }
        ways.swap(nw);
    }
    uint64_t total = 0;
    for (uint64_t w : ways) total = std::min(kSat, total + w);
    return total;
}

inline SeedIndex seed_build(const Panel &h, int variant) {
    SeedIndex ix;
    const bool raw = variant != 0;
    const int mm = std::max(0, h.cfg.max_mm), tw = std::max(0, h.cfg.terminal_window);
    if (h.orients.empty()) return ix;
    int min_len = 1 << 30;
    for (const Orient &o : h.orients)
        if (o.L > 0) min_len = std::min(min_len, o.L);
    if (min_len == 1 << 30) min_len = 1;
    const int s_max = std::max(1, std::min(kSeedMaxLen, min_len));
    int best_s = 0, best_s1 = 0, best_k = 2, best_t = 1;
    double best_cost = 1e300;
    int max_stride = 4;
    if (const char *e = getenv("IPCR_SEED_STRIDE")) max_stride = atoi(e) >= 4 ? 4 : (atoi(e) >= 2 ? 2 : 1);
    for (int S = s_max; S >= 1; S--) {
        uint64_t e = 0;
        for (const Orient &o : h.orients)
            if (o.L >= S) e = std::min<uint64_t>(1ull << 40, e + seed_variant_count(o, S, mm, tw, raw));
        if (e > kSeedMaxEntries) continue;
        for (int S1 = S; S1 >= std::min(S, 3); S1--) {
            uint64_t k1 = 0;
            int offs = 0;  // offsets d, d-1, .. counted so far
            for (int t = 1; t <= max_stride && S - S1 >= t - 1; t *= 2) {
                for (; offs < t; offs++)
                    for (const Orient &o : h.orients)
                        if (o.L >= S)
                            k1 = std::min<uint64_t>(1ull << 40, k1 + seed_key1_count(o, S, S1, mm, tw, raw, S - S1 - offs));
                const uint64_t kk = std::min<uint64_t>(k1, e * (uint64_t)t);
                for (int k = 1; k <= 2; k++) {
                    const double c = seed_cost((double)e, S, (double)kk, S1, k, t);
                    if (c < best_cost) { best_cost = c; best_s = S; best_s1 = S1; best_k = k; best_t = t; }
                }
            }
        }
    }
    if (best_s == 0) best_s = best_s1 = std::min(s_max, 6);  // 4^6 x 2048 orientations always fits
    if (const char *e = getenv("IPCR_SEED_LEN")) {
        best_s = std::max(1, std::min(atoi(e), s_max));
        best_s1 = std::min(best_s1, best_s);
    }
    if (const char *e = getenv("IPCR_SEED_LEN1")) best_s1 = std::max(std::min(best_s, 3), std::min(atoi(e), best_s));
    if (const char *e = getenv("IPCR_SEED_STRIDE_FORCE")) {  // tests: this stride whatever the model says
        best_t = atoi(e) >= 4 ? 4 : (atoi(e) >= 2 ? 2 : 1);
        best_s1 = std::max(std::min(best_s, 3), std::min(best_s1, best_s - (best_t - 1)));
    }
    while (best_t > 1 && best_s - best_s1 < best_t - 1) best_t /= 2;
    ix.S = best_s;
    ix.S1 = best_s1;
    ix.stride = best_t;
    ix.kmer_mask = best_s == 16 ? 0xFFFFFFFFu : ((1u << (2 * best_s)) - 1u);
    std::vector<std::pair<uint32_t, uint32_t>> entries;  // (k-mer, orientation)
    std::vector<uint32_t> tmp;
    for (size_t q = 0; q < h.orients.size(); q++) {
        const Orient &o = h.orients[q];
        ix.col_off.push_back((uint16_t)seed_col_off(o, ix.S));
        tmp.clear();
        if (o.L >= ix.S) seed_enumerate(o, ix.S, mm, tw, raw, tmp);
        for (uint32_t k : tmp) entries.emplace_back(k, (uint32_t)q);
    }
    ix.n_entries = entries.size();
    ix.table_bits = 10;
    while ((1ull << ix.table_bits) < ix.n_entries * 5 / 2) ix.table_bits++;
    ix.slots.assign((size_t)1 << ix.table_bits, 0ull);
    ix.bloom.assign(kSeedBloomWords, 0u);
    const uint32_t tmask = (uint32_t)((1ull << ix.table_bits) - 1);
    ix.bloom_k = best_k;
    {
        const int bits2 = std::max(10, std::min(2 * ix.S, 28));
        if (2 * ix.S <= 28) { ix.bitmap2_mul = 1u; ix.bitmap2_shift = 5u; }
        else { ix.bitmap2_mul = kSeedHashMul; ix.bitmap2_shift = 32u - (uint32_t)(bits2 - 5); }
        ix.bitmap2.assign((size_t)1 << (bits2 - 5), 0u);
    }
    if (const char *e = getenv("IPCR_SEED_BLOOM_K")) ix.bloom_k = std::max(1, std::min(atoi(e), 2));
    const int d2 = 2 * (ix.S - ix.S1);
    const uint32_t key1_mask = ix.S1 >= 16 ? 0xFFFFFFFFu : ((1u << (2 * ix.S1)) - 1u);
    std::vector<uint32_t> keys1;
    keys1.reserve(entries.size() * (size_t)ix.stride);
    for (const auto &e : entries) {
        uint32_t hh = (uint32_t)(e.first * kSeedHashMul) >> (32 - ix.table_bits);
        while (ix.slots[hh]) hh = (hh + 1) & tmask;
        ix.slots[hh] = seed_slot(e.first, e.second, ix.col_off[e.second]);
        for (int i = 0; i < ix.stride; i++) keys1.push_back((e.first >> (d2 - 2 * i)) & key1_mask);
        const uint32_t h2 = e.first * ix.bitmap2_mul;
        ix.bitmap2[h2 >> ix.bitmap2_shift] |= 1u << (h2 & 31u);
    }
    std::sort(keys1.begin(), keys1.end());
    keys1.erase(std::unique(keys1.begin(), keys1.end()), keys1.end());
    ix.n_keys1 = keys1.size();
    if (ix.S1 >= 3) {
        for (uint32_t k1 : keys1) {
            uint32_t m;
            const uint32_t w = seed_bloom_word(k1, ix.S1, ix.bloom_k, m);
            ix.bloom[w] |= m;
        }
    }
    uint64_t bits_set = 0;
    for (uint32_t w : ix.bloom) bits_set += (uint64_t)__builtin_popcount(w);
    ix.bloom_fill = (double)bits_set / (double)kSeedBloomBits;
    ix.use_bloom = ix.S1 >= 3 && seed_pass_rate((double)ix.n_keys1, ix.S1, ix.bloom_k) <= 0.5;
    if (const char *e = getenv("IPCR_SEED_BLOOM")) ix.use_bloom = atoi(e) != 0 && ix.S1 >= 3;
    if (!ix.use_bloom) ix.stride = 1;  // without the bitmap every position goes to the second level
    // pigeonhole first level: worth it when the bitmap is (nearly) useless and the half-window tables are selective
    {
        bool all16 = ix.S == 16;
        for (const Orient &o : h.orients) all16 = all16 && o.L >= 16;
        int want = -1;  // -1 = decide by the model
        if (const char *e = getenv("IPCR_SEED_PIGEON")) want = atoi(e) != 0 ? 1 : 0;
        if (all16 && want != 0) {
            std::vector<uint32_t> lo(kSeedHalfWords, 0u), hi(kSeedHalfWords, 0u);
            const int t = mm / 2;
            for (const Orient &o : h.orients) {
                seed_half_enumerate(o, 0, t, tw, raw, lo);
                seed_half_enumerate(o, 8, t, tw, raw, hi);
            }
            uint64_t nlo = 0, nhi = 0;
            for (uint32_t w : lo) nlo += (uint64_t)__builtin_popcount(w);
            for (uint32_t w : hi) nhi += (uint64_t)__builtin_popcount(w);
            const double pass = 1.0 - (1.0 - (double)nlo / 65536.0) * (1.0 - (double)nhi / 65536.0);
            // fast path costs ~2.2 bitmap lookups (two LDS, more shifts); the second level as in seed_cost()
            const double prob = (double)ix.n_entries / std::pow(4.0, ix.S);
            const double cost_pigeon = 2.2 + seed_walk_cost(pass, 1) + prob * 675.0;
            const double cost_now = ix.use_bloom ? seed_cost((double)ix.n_entries, ix.S, (double)ix.n_keys1, ix.S1, ix.bloom_k, ix.stride)
                                                 : 1.0 + 15.0 + prob * 675.0;
            if (want == 1 || cost_pigeon < cost_now) {
                ix.pigeon = true;
                ix.pigeon_pass = pass;
                ix.use_bloom = true;
                ix.bloom_k = 3;  // reported by the debug tap: 3 = pigeonhole tables
                ix.S1 = ix.S;  // the kernel looks at the window where it starts (d = 0)
                ix.stride = 1;
                std::fill(ix.bloom.begin(), ix.bloom.end(), 0u);
                std::copy(lo.begin(), lo.end(), ix.bloom.begin());
                std::copy(hi.begin(), hi.end(), ix.bloom.begin() + kSeedHalfWords);
            }
        }
    }
    ix.cand_prob = (double)ix.n_entries / std::pow(4.0, ix.S);
    if (getenv("IPCR_SEED_DEBUG"))
        fprintf(stderr, "seed index[%d]: S=%d S1=%d bits=%d stride=%d entries=%llu keys1=%llu fill=%.4f bloom=%d pigeon=%d\n",
                variant, ix.S, ix.S1, ix.bloom_k, ix.stride, (unsigned long long)ix.n_entries,
                (unsigned long long)ix.n_keys1, ix.bloom_fill, ix.use_bloom ? 1 : 0, ix.pigeon ? 1 : 0);
    return ix;
}

// CPU evaluation of the index on one sequence (test tap; mirrors k_pack2 + k_seed_scan): every (orientation << 44 | start) the device filter would hand to verify
inline void seed_eval(const Panel &h, const SeedIndex &ix, const uint8_t *seq, uint64_t len,
                      std::vector<uint64_t> &out) {
    out.clear();
    const uint64_t S = (uint64_t)ix.S;
    if (S == 0 || len < S) return;
    const uint32_t tmask = (uint32_t)((1ull << ix.table_bits) - 1);
    for (uint64_t p = 0; p + S <= len; p++) {
        uint32_t kmer = 0, n_inv = 0;
        for (uint64_t j = 0; j < S; j++) {
            const uint8_t b = seq[p + j];
            if (!(b == 'A' || b == 'C' || b == 'G' || b == 'T')) n_inv++;
            kmer |= (uint32_t)((b >> 1) & 3u) << (2 * j);
        }
        if (n_inv == 0) {
            if (ix.pigeon) {
                const uint32_t lo = kmer & 0xFFFFu, hi = kmer >> 16;
                const bool in_lo = (ix.bloom[lo >> 5] >> (lo & 31u)) & 1u;
                const bool in_hi = (ix.bloom[kSeedHalfWords + (hi >> 5)] >> (hi & 31u)) & 1u;
                if (!in_lo && !in_hi) continue;
            } else if (ix.use_bloom) {
                // the anchor that answers for window p: the multiple of `stride` in [p + d - stride + 1, p + d]
                const uint64_t d = (uint64_t)(ix.S - ix.S1), q = (p + d) / (uint64_t)ix.stride * (uint64_t)ix.stride;
                const uint32_t key = (kmer >> (2 * (q - p))) & (ix.S1 >= 16 ? 0xFFFFFFFFu : ((1u << (2 * ix.S1)) - 1u));
                uint32_t m;
                const uint32_t w = seed_bloom_word(key, ix.S1, ix.bloom_k, m);
                if ((ix.bloom[w] & m) != m) continue;
            }
            const uint32_t h2 = kmer * ix.bitmap2_mul;
            if (!((ix.bitmap2[h2 >> ix.bitmap2_shift] >> (h2 & 31u)) & 1u)) continue;
            for (uint32_t hh = (uint32_t)(kmer * kSeedHashMul) >> (32 - ix.table_bits); ix.slots[hh]; hh = (hh + 1) & tmask) {
                const uint64_t slot = ix.slots[hh];
                const uint64_t c = (slot >> 48) & 0xFF;
                if ((uint32_t)slot == kmer && p >= c) out.push_back((((slot >> 32) & 0xFFFF) << 44) | (p - c));
            }
        } else if ((int)n_inv <= h.cfg.max_mm) {
            // dirty window: every replacement of the invalid bytes is looked up (k_seed_dirty_probe);
            // more than 16 distinct hits or more than 4 invalid bytes => all orientations
            std::vector<uint32_t> pos;
            uint32_t base = kmer;
            for (uint64_t j = 0; j < S; j++) {
                const uint8_t b = seq[p + j];
                if (!(b == 'A' || b == 'C' || b == 'G' || b == 'T')) {
                    pos.push_back((uint32_t)j);
                    base &= ~(3u << (2 * j));
                }
            }
            bool overflow = pos.size() > 4;
            std::vector<uint32_t> seen;
            for (uint32_t combo = 0; !overflow && combo < (1u << (2 * pos.size())); combo++) {
                uint32_t km = base;
                for (size_t u = 0; u < pos.size(); u++) km |= ((combo >> (2 * u)) & 3u) << (2 * pos[u]);
                const uint32_t h2 = km * ix.bitmap2_mul;
                if (!((ix.bitmap2[h2 >> ix.bitmap2_shift] >> (h2 & 31u)) & 1u)) continue;
                for (uint32_t hh = (uint32_t)(km * kSeedHashMul) >> (32 - ix.table_bits); ix.slots[hh] && !overflow;
                     hh = (hh + 1) & tmask) {
                    const uint64_t slot = ix.slots[hh];
                    if ((uint32_t)slot != km) continue;
                    const uint32_t key = (uint32_t)(slot >> 32) & 0x00FFFFFFu;
                    if (std::find(seen.begin(), seen.end(), key) != seen.end()) continue;
                    if (seen.size() == 16) overflow = true;
                    else seen.push_back(key);
                }
            }
            if (overflow) {
                for (size_t q = 0; q < h.orients.size(); q++) {
                    const uint64_t c = ix.col_off[q];
                    if (h.orients[q].L >= ix.S && p >= c) out.push_back(((uint64_t)q << 44) | (p - c));
                }
            } else {
                for (uint32_t key : seen) {
                    const uint64_t c = key >> 16;
                    if (p >= c) out.push_back(((uint64_t)(key & 0xFFFFu) << 44) | (p - c));
                }
            }
        }
    }
}

}  // namespace ipcr
