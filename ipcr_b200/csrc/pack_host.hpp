// pack_host.hpp -- the host side of the packed genome format: ASCII -> 2-bit codes (16 bases per
// word: A 0, C 1, T 2, G 3 = (byte >> 1) & 3) + hard-mismatch bitmap (32 bases per word: bit = the
// byte is not an uppercase A/C/G/T, /root/reference/core/primer/iupac.go:62-67).  Bit-for-bit what
// k_pack2 (seed_kernels.cuh) writes on the device, produced by the native FASTA feeder so that
// 0.375 B/bp cross PCIe instead of 1 B/bp.  Padding bytes count as 0x00: code 0, invalid.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#if defined(__x86_64__)
#include <cpuid.h>
#include <immintrin.h>
#endif
#include <thread>
#include <utility>
#include <vector>

namespace ipcr {

constexpr uint64_t kPackChunk = 4096;  // = kSeedChunk: positions one scan warp takes at a time

// 32-base groups the scan kernels may touch for n bases: whole chunks + one look-ahead group
inline uint64_t packed_groups(uint64_t n_bases) {
    const uint64_t n_chunks = (n_bases + kPackChunk - 1) / kPackChunk;
    return n_chunks * (kPackChunk / 32u) + 1;
}

// four bytes -> 8 code bits + 4 invalid bits (the arithmetic of pack4 in seed_kernels.cuh)
inline void pack4_host(uint32_t w, uint32_t &code8, uint32_t &inv4) {
    const uint32_t t = (w >> 1) & 0x03030303u;
    code8 = (t * 0x01041040u) >> 24;
    const uint32_t e2 = (t >> 1) & ~t & 0x01010101u;
    const uint32_t expect = 0x41414141u + 2u * t + 15u * e2;
    const uint32_t d = expect ^ w;
    const uint32_t nz = (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;
    inv4 = ((nz >> 7) * 0x10204080u) >> 28;
}

#if defined(__x86_64__)
// AVX2 + BMI2 form of the same (chosen at run time): a byte is valid iff it equals one of 'A' 'C' 'G' 'T' (four
// compares, one movemask = the 32 bitmap bits); its code is bits 1..2 of the byte, gathered 8 bytes at a time by PEXT.
// ~20 instructions per 32 bases instead of ~110.
__attribute__((target("avx2,bmi2")))
inline uint32_t pack_groups_staged_avx2(const uint8_t *src, uint64_t g0, uint64_t ng, uint32_t *codes, uint32_t *inval) {
    uint32_t any = 0;
    const __m256i vA = _mm256_set1_epi8('A'), vC = _mm256_set1_epi8('C'), vG = _mm256_set1_epi8('G'), vT = _mm256_set1_epi8('T');
    const unsigned long long m = 0x0606060606060606ull;
    for (uint64_t i = 0; i < ng; i++) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + 32 * i));
        const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(v, vA), _mm256_cmpeq_epi8(v, vC)),
                                           _mm256_or_si256(_mm256_cmpeq_epi8(v, vG), _mm256_cmpeq_epi8(v, vT)));
        const uint32_t bad = ~(uint32_t)_mm256_movemask_epi8(ok);
        inval[g0 + i] = bad;
        any |= bad;
        unsigned long long q[4];
        memcpy(q, src + 32 * i, 32);
        codes[2 * (g0 + i)] = (uint32_t)(_pext_u64(q[0], m) | (_pext_u64(q[1], m) << 16));
        codes[2 * (g0 + i) + 1] = (uint32_t)(_pext_u64(q[2], m) | (_pext_u64(q[3], m) << 16));
    }
    return any;
}

inline bool pack_have_avx2() {
    static const bool have = [] {
        unsigned a = 0, b = 0, c = 0, d = 0;
        if (!__get_cpuid_count(7, 0, &a, &b, &c, &d)) return false;
        const bool avx2 = (b >> 5) & 1u, bmi2 = (b >> 8) & 1u;
        if (!avx2 || !bmi2) return false;
        unsigned a1 = 0, b1 = 0, c1 = 0, d1 = 0;  // the OS must save the YMM state: OSXSAVE + XCR0[2:1] = 11
        if (!__get_cpuid(1, &a1, &b1, &c1, &d1) || !((c1 >> 27) & 1u)) return false;
        unsigned lo, hi;
        __asm__ volatile("xgetbv" : "=a"(lo), "=d"(hi) : "c"(0));
        return ((lo | (hi & 0u)) & 6u) == 6u && getenv("IPCR_FEEDER_NO_AVX2") == nullptr;
    }();
    return have;
}
#endif

// ng whole groups starting at group g0, their 32 * ng bytes at `src` (a staging buffer: the feeder's pack-only mode
// never materialises the ASCII buffer)
// returns the OR of the bitmap words it wrote (0: no hard-mismatch byte in there)
inline uint32_t pack_groups_staged(const uint8_t *src, uint64_t g0, uint64_t ng, uint32_t *codes, uint32_t *inval) {
#if defined(__x86_64__)
    if (pack_have_avx2()) return pack_groups_staged_avx2(src, g0, ng, codes, inval);
#endif
    uint32_t any = 0;
    for (uint64_t i = 0; i < ng; i++) {
        uint32_t w[8];
        memcpy(w, src + 32 * i, 32);
        uint32_t c0 = 0, c1 = 0, inv = 0;
        for (int k = 0; k < 8; k++) {
            uint32_t c8, i4;
            pack4_host(w[k], c8, i4);
            if (k < 4) c0 |= c8 << (8 * k); else c1 |= c8 << (8 * (k - 4));
            inv |= i4 << (4 * k);
        }
        codes[2 * (g0 + i)] = c0;
        codes[2 * (g0 + i) + 1] = c1;
        inval[g0 + i] = inv;
        any |= inv;
    }
    return any;
}

// the ASCII a packed buffer stands for, positions [a, b): A/C/G/T, every hard-mismatch byte as 'N' (for the engine
// all non-ACGT bytes are the same: a mismatch at their column and an automaton reset, core/engine/ac.go:141-148)
inline void unpack_host(const uint32_t *codes, const uint32_t *inval, uint64_t a, uint64_t b, uint8_t *out) {
    static const char kBase[4] = {'A', 'C', 'T', 'G'};
    for (uint64_t p = a; p < b; p++) {
        const bool bad = (inval[p >> 5] >> (p & 31u)) & 1u;
        out[p - a] = bad ? (uint8_t)'N' : (uint8_t)kBase[(codes[p >> 4] >> (2u * (uint32_t)(p & 15u))) & 3u];
    }
}

// groups [g0, g1) of a buffer of n bases; returns true if any byte of [32*g0, min(n, 32*g1)) is invalid
inline bool pack_groups_host(const uint8_t *ascii, uint64_t n, uint64_t g0, uint64_t g1, uint32_t *codes, uint32_t *inval) {
    uint32_t any = 0;
#if defined(__x86_64__)
    if (pack_have_avx2() && g0 < std::min<uint64_t>(g1, n / 32)) {  // the whole groups
        const uint64_t gf = std::min<uint64_t>(g1, n / 32);
        pack_groups_staged_avx2(ascii + 32 * g0, g0, gf - g0, codes, inval);
        for (uint64_t g = g0; g < gf; g++) any |= inval[g];
        g0 = gf;
    }
#endif
    for (uint64_t g = g0; g < g1; g++) {
        const uint64_t P0 = g * 32ull;
        uint32_t w[8];
        if (P0 + 32ull <= n) {
            memcpy(w, ascii + P0, 32);
        } else {
            memset(w, 0, sizeof w);
            if (P0 < n) memcpy(w, ascii + P0, (size_t)(n - P0));
        }
        uint32_t c0 = 0, c1 = 0, inv = 0;
        for (int k = 0; k < 8; k++) {
            uint32_t c8, i4;
            pack4_host(w[k], c8, i4);
            if (k < 4) c0 |= c8 << (8 * k); else c1 |= c8 << (8 * (k - 4));
            inv |= i4 << (4 * k);
        }
        codes[2 * g] = c0;
        codes[2 * g + 1] = c1;
        inval[g] = inv;
        if (P0 + 32ull <= n) any |= inv;
        else if (P0 < n) any |= inv & ((1u << (uint32_t)(n - P0)) - 1u);
    }
    return any != 0;
}

// whole buffer, multi-threaded; codes[2 * packed_groups(n)], inval[packed_groups(n)]
inline void pack_host(const uint8_t *ascii, uint64_t n, uint32_t *codes, uint32_t *inval, int threads) {
    const uint64_t groups = packed_groups(n);
    if (threads < 1) threads = 1;
    if (groups < (1u << 16)) threads = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) {
        const uint64_t g0 = groups * (uint64_t)t / (uint64_t)threads, g1 = groups * (uint64_t)(t + 1) / (uint64_t)threads;
        if (t + 1 < threads) th.emplace_back([=] { pack_groups_host(ascii, n, g0, g1, codes, inval); });
        else pack_groups_host(ascii, n, g0, g1, codes, inval);
    }
    for (auto &x : th) x.join();
}

// The bitmap of a real genome is almost empty (N runs are few and long): as a list of runs [start, end) it is
// a few KB instead of 0.125 B/bp, and the device rebuilds the words (k_inval_fill).  Includes the padding
// positions behind n (the packer marks them invalid).  Returns false when there are more than max_runs runs
// (scattered IUPAC codes): the caller then ships the bitmap itself.
inline bool packed_invalid_runs(const uint32_t *inval, uint64_t groups, uint64_t max_runs, std::vector<uint64_t> &runs) {
    runs.clear();
    bool open = false;
    uint64_t start = 0;
    for (uint64_t g = 0; g < groups; g++) {
        const uint32_t w = inval[g];
        if (w == 0u) {
            if (open) {
                runs.push_back(start);
                runs.push_back(g * 32ull);
                open = false;
                if (runs.size() / 2 > max_runs) return false;
            }
            continue;
        }
        if (w == 0xFFFFFFFFu) {
            if (!open) {
                open = true;
                start = g * 32ull;
            }
            continue;
        }
        for (uint32_t i = 0; i < 32u; i++) {
            const bool bad = (w >> i) & 1u;
            if (bad && !open) {
                open = true;
                start = g * 32ull + i;
            } else if (!bad && open) {
                runs.push_back(start);
                runs.push_back(g * 32ull + i);
                open = false;
            }
        }
        if (runs.size() / 2 > max_runs) return false;
    }
    if (open) {
        runs.push_back(start);
        runs.push_back(groups * 32ull);
    }
    return runs.size() / 2 <= max_runs;
}

// the same over several threads: each takes a range of groups, runs that touch at a range boundary are merged
inline bool packed_invalid_runs_mt(const uint32_t *inval, uint64_t groups, uint64_t max_runs, std::vector<uint64_t> &runs,
                                   int threads) {
    if (threads < 2 || groups < (1u << 20)) return packed_invalid_runs(inval, groups, max_runs, runs);
    std::vector<std::vector<uint64_t>> part((size_t)threads);
    std::vector<char> ok((size_t)threads, 1);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) {
        const uint64_t g0 = groups * (uint64_t)t / (uint64_t)threads, g1 = groups * (uint64_t)(t + 1) / (uint64_t)threads;
        th.emplace_back([&, t, g0, g1] {
            ok[(size_t)t] = packed_invalid_runs(inval + g0, g1 - g0, max_runs, part[(size_t)t]) ? 1 : 0;
            for (uint64_t &v : part[(size_t)t]) v += g0 * 32ull;
        });
    }
    for (auto &x : th) x.join();
    runs.clear();
    for (int t = 0; t < threads; t++) {
        if (!ok[(size_t)t]) return false;
        for (size_t i = 0; i + 1 < part[(size_t)t].size(); i += 2) {
            if (!runs.empty() && runs.back() == part[(size_t)t][i]) runs.back() = part[(size_t)t][i + 1];
            else {
                runs.push_back(part[(size_t)t][i]);
                runs.push_back(part[(size_t)t][i + 1]);
            }
        }
        if (runs.size() / 2 > max_runs) return false;
    }
    return true;
}

// the same from a sorted list of disjoint group ranges [g0, g1) that hold every set bit (the one-trip feeder knows
// which of its blocks wrote a non-zero bitmap word): a clean genome costs nothing, one with a few N runs a few blocks
inline bool packed_invalid_runs_ranges(const uint32_t *inval, const std::vector<std::pair<uint64_t, uint64_t>> &ranges,
                                       uint64_t max_runs, std::vector<uint64_t> &runs) {
    runs.clear();
    std::vector<uint64_t> part;
    for (const auto &r : ranges) {
        if (r.second <= r.first) continue;
        if (!packed_invalid_runs(inval + r.first, r.second - r.first, max_runs, part)) return false;
        for (size_t i = 0; i + 1 < part.size(); i += 2) {
            const uint64_t a = part[i] + r.first * 32ull, b = part[i + 1] + r.first * 32ull;
            if (!runs.empty() && runs.back() == a) runs.back() = b;
            else {
                runs.push_back(a);
                runs.push_back(b);
            }
        }
        if (runs.size() / 2 > max_runs) return false;
    }
    return true;
}

// does [a, b) hold an invalid base?  (per-record reset flags of the feeder: it upper-cases a-z, so an
// invalid byte IS a reset byte, core/engine/ac.go:141-148)
inline bool packed_any_invalid(const uint32_t *inval, uint64_t a, uint64_t b) {
    if (a >= b) return false;
    uint64_t w0 = a >> 5, w1 = (b - 1) >> 5;
    for (uint64_t w = w0; w <= w1; w++) {
        uint32_t m = 0xFFFFFFFFu;
        if (w == w0) m &= 0xFFFFFFFFu << (a & 31u);
        if (w == w1 && ((b & 31u) != 0)) m &= (1u << (b & 31u)) - 1u;
        if (inval[w] & m) return true;
    }
    return false;
}

}  // namespace ipcr
