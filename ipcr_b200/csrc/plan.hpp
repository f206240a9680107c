// plan.hpp -- host-side panel compilation for the B200 scan.
//
// Replaces engine.CompilePanel (/root/reference/core/engine/compiled.go:96-136).
// The reference turns a primer panel into an Aho-Corasick automaton over
// enumerated seed variants (seed.go, ac.go).  On the GPU the panel becomes
//   * a set of UNIQUE orientation strings (A, B, rc(A), rc(B) de-duplicated
//     across pairs: `--self` panels reuse the same four strings three times,
//     internal/common/primers.go:11-38), each with its IUPAC class per column
//     (core/primer/iupac.go:6-58), and
//   * per orientation a bit-parallel FILTER PLAN: a boolean formula over
//     "column j matches" bits that every true site satisfies (pigeonhole over
//     disjoint column blocks + the 3' terminal window), evaluated for 32
//     positions per LOP3 on bit-planes of the genome.
// Exactness lives in the verify kernel (BaseMatch on the raw bytes); the plan
// only has to be a superset.  What the reference WOULD have seeded is still
// computed (Have / seed span) because hit-cap truncation and emission order of
// the reference depend on it (compiled.go:185-190,238-257; halo.go).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "../../include/ipcr_gpu.h"

namespace ipcr {

constexpr int kWindow = 32;       // rows of the genome a filter may look at
constexpr int kMaxOrient = 2048;  // sort-key budget (11 bits)

inline uint8_t iupac_class(uint8_t c) {  // iupac.go:21-38, uppercase only (RevComp rejects the rest)
    switch (c) {
    case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': return 8;
    case 'R': return 5; case 'Y': return 10; case 'S': return 6; case 'W': return 9;
    case 'K': return 12; case 'M': return 3; case 'B': return 14; case 'D': return 13;
    case 'H': return 11; case 'V': return 7; case 'N': return 15;
    }
    return 0;
}

inline char iupac_complement(char c) {  // rc.go:8-24
    switch (c) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    case 'R': return 'Y'; case 'Y': return 'R'; case 'S': return 'S'; case 'W': return 'W';
    case 'K': return 'M'; case 'M': return 'K'; case 'B': return 'V'; case 'V': return 'B';
    case 'D': return 'H'; case 'H': return 'D'; case 'N': return 'N';
    }
    return 0;
}

// rc.go:41-56; returns false and sets bad_pos (1-based) on an invalid byte
inline bool revcomp(const std::string &in, std::string &out, size_t &bad_pos) {
    out.resize(in.size());
    for (size_t i = 0; i < in.size(); i++) {
        size_t src = in.size() - 1 - i;
        char c = iupac_complement(in[src]);
        if (!c) { bad_pos = src + 1; return false; }
        out[i] = c;
    }
    return true;
}

struct Orient {
    std::string seq;
    int L = 0;
    bool rc_type = false;  // 'a'/'b': protected window is the LEFT tw columns
    bool seeded = false;   // what the reference's buildSeedPatterns decides (seed.go:170-213)
    int seed_off = 0, seed_len = 0;
    std::vector<uint8_t> cls;  // IUPAC class per column
};

struct PairRoles {
    int fwdA, fwdB, rcA, rcB;  // unique-orientation ids
    int alen, blen;
    int64_t min_len, max_len;  // resolved (engine.go:113-120)
};

struct ColRef {
    uint8_t row;  // j - j0, < kWindow
    uint8_t cls;  // IUPAC class (1..14)
};

struct OrientPlan {
    int j0 = 0;
    std::vector<ColRef> prot;                 // all must match
    std::vector<std::vector<ColRef>> blocks;  // >= need blocks must match entirely
    int need = 0;                             // 0 => block test is vacuous
    double cand_prob = 1.0;
    double cost = 0.0;
};

struct ScanPlan {
    bool raw_rc = false;  // rc-type orientations ignore the protected window (hit-cap parity)
    std::vector<OrientPlan> o;
    double cost_per_word = 0.0;  // modelled LOP3 per 32 positions, all orientations
};

// ---- what the reference would have seeded ----------------------------------
inline int configured_seed_len(int primer_len, int cfg) {  // seed.go:104-115
    if (cfg <= 0) return primer_len < 12 ? primer_len : 12;
    return cfg > primer_len ? primer_len : cfg;
}

inline int class_popcount(uint8_t c) { return __builtin_popcount(c); }

inline void choose_seed_span(const Orient &o, int seed_len_cfg, bool prefer_right, int &off, int &len) {
    // seed.go:260-283 with exactVariantEstimate (seed.go:285-302)
    len = configured_seed_len(o.L, seed_len_cfg);
    off = 0;
    uint64_t best = 0;
    for (int s = 0; s + len <= o.L; s++) {
        uint64_t score = 1;
        for (int i = 0; i < len; i++) {
            uint64_t k = (uint64_t)class_popcount(o.cls[s + i]);
            if (k == 0) k = 5;
            score *= k;  // wraps like Go's uint64
        }
        bool tie = prefer_right ? (s > off) : (s < off);
        if (s == 0 || score < best || (score == best && tie)) { off = s; best = score; }
    }
}

inline bool reference_would_seed(const Orient &o, const ipcr_cfg &cfg, int &off, int &len) {
    off = len = 0;
    if (cfg.seed_len < 0 || o.L == 0) return false;  // seed.go:156-159,171-173
    choose_seed_span(o, cfg.seed_len, /*prefer_right=*/!o.rc_type, off, len);
    if (len <= 0 || len > o.L) return false;
    int max_mm = cfg.max_mm < 0 ? 0 : cfg.max_mm;
    int tw = cfg.terminal_window < 0 ? 0 : cfg.terminal_window;
    int left_tw = o.rc_type ? tw : 0, right_tw = o.rc_type ? 0 : tw;
    // number of concrete variants enumerateSeedVariants would emit (seed.go:304-367)
    const uint64_t kSat = 1ull << 40;
    std::vector<uint64_t> ways((size_t)max_mm + 1, 0);
    ways[0] = 1;
    for (int i = 0; i < len; i++) {
        int full = off + i;
        bool prot = (left_tw > 0 && full < left_tw) || (right_tw > 0 && full >= o.L - right_tw);
        uint64_t km = (uint64_t)class_popcount(o.cls[full]);
        uint64_t kx = prot ? 0 : 4 - km;
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
    if (total == 0 || total > 50000) return false;  // seed.go:97,188-192
    if (len > 32) return false;                     // seed.go:202-207
    return true;
}

// ---- filter planning --------------------------------------------------------
inline double class_bits(uint8_t c) { return std::log2(4.0 / (double)class_popcount(c)); }

// probability that at least `need` of the blocks match entirely on a uniform genome
inline double at_least_prob(const std::vector<double> &p, int need) {
    std::vector<double> d(p.size() + 1, 0.0);
    d[0] = 1.0;
    for (size_t b = 0; b < p.size(); b++)
        for (int k = (int)b + 1; k >= 0; k--)
            d[k] = d[k] * (1.0 - p[b]) + (k > 0 ? d[k - 1] * p[b] : 0.0);
    double s = 0.0;
    for (size_t k = (size_t)std::max(need, 0); k < d.size(); k++) s += d[k];
    return s;
}

inline int and_inputs(const std::vector<ColRef> &cols) {
    int n = 0;
    for (const ColRef &c : cols) n += class_popcount(c.cls);
    return n;
}

inline double lop3_for_inputs(int inputs) { return inputs <= 1 ? 0.0 : std::ceil((inputs - 1) / 2.0); }

// modelled cost of one candidate position, in lane-level LOP3 (a candidate costs ~45 warp
// instructions between the filter's slow path and the verify kernel; measured round 1)
inline double verify_cost_lop3() {
    static const double v = [] {
        const char *e = getenv("IPCR_PLAN_VERIFY_COST");
        return e ? atof(e) : 2500.0;
    }();
    return v;
}

// columns of a primer the filter may look at (anchored at the 3' end).  23 keeps the
// register ring of the specialised kernel at 24 rows (96 registers) for every primer length;
// longer primers simply leave their 5'-most columns to the verify kernel.
inline int filter_window() {
    static const int v = [] {
        const char *e = getenv("IPCR_PLAN_WINDOW");
        const int w = e ? atoi(e) : 23;
        return std::max(4, std::min(w, kWindow));
    }();
    return v;
}

inline OrientPlan plan_orientation(const Orient &o, int max_mm, int tw, bool raw) {
    OrientPlan pl;
    if (tw < 0) tw = 0;
    const double kVerify = verify_cost_lop3();
    const int win = filter_window();
    int L = o.L;
    int lo = 0, hi = L;  // window [lo, hi)
    if (L > win) {
        if (o.rc_type) { lo = 0; hi = win; } else { lo = L - win; hi = L; }
    }
    pl.j0 = lo;
    int plo, phi;  // protected range
    if (o.rc_type) { plo = 0; phi = std::min(tw, L); } else { plo = std::max(L - tw, 0); phi = L; }
    if (raw || tw == 0) { plo = phi = 0; }
    std::vector<ColRef> free_cols;
    double prot_bits = 0.0;
    for (int j = lo; j < hi; j++) {
        uint8_t c = o.cls[j];
        if (c == 15) continue;  // N column: never a mismatch, carries no information
        ColRef r{(uint8_t)(j - lo), c};
        if (j >= plo && j < phi) { pl.prot.push_back(r); prot_bits += class_bits(c); }
        else free_cols.push_back(r);
    }
    double prot_p = std::pow(2.0, -prot_bits);
    double prot_cost = lop3_for_inputs(and_inputs(pl.prot) + 1);
    int m = max_mm < 0 ? 0 : max_mm;
    int nfree = (int)free_cols.size();
    if (nfree < m + 1) {  // cannot force any block to be exact
        pl.need = 0;
        pl.cand_prob = prot_p;
        pl.cost = prot_cost + pl.cand_prob * 32.0 * kVerify;
        return pl;
    }
    // Pigeonhole only needs DISJOINT blocks, not a partition: mismatches in unused columns can
    // only help.  So besides the number of blocks B we also choose how many columns to use
    // (the most informative ones), trading LOP3 per word against candidates to verify.
    std::vector<int> order((size_t)nfree);
    for (int i = 0; i < nfree; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        return class_bits(free_cols[a].cls) > class_bits(free_cols[b].cls);
    });
    double best_total = 1e300;
    const int bmax = std::min(std::min(nfree, m + 1 + 6), 15);
    for (int B = m + 1; B <= bmax; B++) {
        for (int used = B; used <= nfree; used++) {
            std::vector<int> pick(order.begin(), order.begin() + used);
            std::sort(pick.begin(), pick.end());
            double total_bits = 0.0;
            for (int i : pick) total_bits += class_bits(free_cols[i].cls);
            // contiguous split (in primer order) balanced by information content
            std::vector<std::vector<ColRef>> blocks((size_t)B);
            double acc = 0.0;
            int b = 0;
            for (int n = 0; n < used; n++) {
                const int remaining_cols = used - n, remaining_blocks = B - b;
                if (!blocks[b].empty() && b + 1 < B &&
                    (acc >= total_bits * (b + 1) / B || remaining_cols <= remaining_blocks - 1)) b++;
                blocks[b].push_back(free_cols[pick[n]]);
                acc += class_bits(free_cols[pick[n]].cls);
            }
            bool ok = true;
            for (auto &bl : blocks) if (bl.empty()) ok = false;
            if (!ok) continue;
            const int need = B - m;
            std::vector<double> bp;
            double cost = prot_cost;
            for (auto &bl : blocks) {
                double bits = 0.0;
                for (const ColRef &c : bl) bits += class_bits(c.cls);
                bp.push_back(std::pow(2.0, -bits));
                cost += lop3_for_inputs(and_inputs(bl));
            }
            cost += need == 1 ? std::ceil((B - 1) / 2.0) + 1.0 : (double)need * B;
            const double p = prot_p * at_least_prob(bp, need);
            const double total = cost + p * 32.0 * kVerify;
            if (total < best_total) {
                best_total = total;
                pl.blocks = blocks;
                pl.need = need;
                pl.cand_prob = p;
                pl.cost = total;
            }
        }
    }
    return pl;
}

struct Panel {
    ipcr_cfg cfg{};
    std::vector<std::string> fwd, rev;  // snapshot of the pairs (compiled.go:98)
    std::vector<int64_t> min_product, max_product;
    std::vector<Orient> orients;
    std::vector<PairRoles> roles;
    std::vector<uint8_t> have;  // bit0 A, bit1 B, bit2 a, bit3 b
    ScanPlan plan[2];           // [0] protected window everywhere, [1] raw rc orientations
    int max_primer_len = 0;
    int jit_mode = 1;

    int intern(const std::string &s, bool rc_type, std::map<std::pair<std::string, bool>, int> &idx) {
        auto key = std::make_pair(s, rc_type);
        auto it = idx.find(key);
        if (it != idx.end()) return it->second;
        Orient o;
        o.seq = s;
        o.L = (int)s.size();
        o.rc_type = rc_type;
        o.cls.resize(s.size());
        for (size_t i = 0; i < s.size(); i++) o.cls[i] = iupac_class((uint8_t)s[i]);
        o.seeded = reference_would_seed(o, cfg, o.seed_off, o.seed_len);
        int id = (int)orients.size();
        orients.push_back(std::move(o));
        idx.emplace(key, id);
        return id;
    }

    // returns "" or an error message
    std::string build(const ipcr_cfg &c, const ipcr_pair *pairs, size_t n) {
        cfg = c;
        if (cfg.max_mm < 0)
            return "max_mm < 0 is not supported: the reference is inconsistent there "
                   "(core/engine/ac.go:207 ignores the budget, core/primer/match.go:79 treats it as 0)";
        if (cfg.terminal_window < 0) cfg.terminal_window = 0;
        std::map<std::pair<std::string, bool>, int> idx;
        for (size_t i = 0; i < n; i++) {
            std::string a(pairs[i].fwd ? pairs[i].fwd : "", pairs[i].fwd_len);
            std::string b(pairs[i].rev ? pairs[i].rev : "", pairs[i].rev_len);
            std::string ra, rb;
            size_t bad = 0;
            if (!revcomp(a, ra, bad) || !revcomp(b, rb, bad)) {
                std::ostringstream e;  // the reference panics here (rc.go:28-34)
                e << "pair " << i << ": invalid reverse-complement base at position " << bad
                  << "; expected normalized uppercase IUPAC DNA";
                return e.str();
            }
            if (a.size() > IPCR_GPU_MAX_PRIMER_LEN || b.size() > IPCR_GPU_MAX_PRIMER_LEN) {
                std::ostringstream e;
                e << "pair " << i << ": primers longer than " << IPCR_GPU_MAX_PRIMER_LEN
                  << " nt are not supported by the GPU engine";
                return e.str();
            }
            fwd.push_back(a);
            rev.push_back(b);
            min_product.push_back(pairs[i].min_product);
            max_product.push_back(pairs[i].max_product);
            PairRoles r{};
            r.fwdA = intern(a, false, idx);
            r.fwdB = intern(b, false, idx);
            r.rcA = intern(ra, true, idx);
            r.rcB = intern(rb, true, idx);
            r.alen = (int)a.size();
            r.blen = (int)b.size();
            r.min_len = pairs[i].min_product ? pairs[i].min_product : cfg.min_len;
            r.max_len = pairs[i].max_product ? pairs[i].max_product : cfg.max_len;
            roles.push_back(r);
            uint8_t h = 0;
            if (orients[r.fwdA].seeded) h |= 1;
            if (orients[r.fwdB].seeded) h |= 2;
            if (orients[r.rcA].seeded) h |= 4;
            if (orients[r.rcB].seeded) h |= 8;
            have.push_back(h);
            max_primer_len = std::max(max_primer_len, std::max(r.alen, r.blen));
        }
        if ((int)orients.size() > kMaxOrient) {
            std::ostringstream e;
            e << "panel has " << orients.size() << " unique primer orientations; the GPU engine supports "
              << kMaxOrient << " per panel (split the panel)";
            return e.str();
        }
        for (int raw = 0; raw < 2; raw++) {
            plan[raw].raw_rc = raw != 0;
            plan[raw].o.clear();
            plan[raw].cost_per_word = 0.0;
            for (const Orient &o : orients) {
                OrientPlan pl = plan_orientation(o, cfg.max_mm, cfg.terminal_window, raw && o.rc_type);
                plan[raw].cost_per_word += pl.cost;
                plan[raw].o.push_back(std::move(pl));
            }
        }
        return "";
    }
};

// CPU evaluation of a plan at one window (test tap; mirrors the kernels' formula)
inline bool plan_eval(const Orient &o, const OrientPlan &pl, const uint8_t *seq, uint64_t len, int64_t start) {
    auto col_match = [&](const ColRef &c) -> bool {
        int64_t p = start + pl.j0 + c.row;
        if (p < 0 || (uint64_t)p >= len) return false;
        uint8_t g = seq[p], one = g == 'A' ? 1 : g == 'C' ? 2 : g == 'G' ? 4 : g == 'T' ? 8 : 0;
        return (one & c.cls) != 0;
    };
    (void)o;
    for (const ColRef &c : pl.prot) if (!col_match(c)) return false;
    if (pl.need <= 0) return true;
    int ok = 0;
    for (const auto &bl : pl.blocks) {
        bool all = true;
        for (const ColRef &c : bl) if (!col_match(c)) { all = false; break; }
        ok += all;
    }
    return ok >= pl.need;
}

}  // namespace ipcr
