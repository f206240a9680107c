// seed_kernels.cuh -- kernels of the seed-index filter (host side: seed.hpp).
//   K1s k_pack2        ASCII -> 2-bit codes (16 bases / word) + hard-mismatch bitmap (32 / word)
//   K2s k_seed_scan    one k-mer per genome position: Bloom bitmap in shared memory, then an
//                      L2-resident bitmap and hash table; hits become (orientation, start)
//                      candidates; seed windows holding 1..max_mm non-ACGT bytes are listed
//   K2d k_seed_dirty_probe   dirty windows: 4^t base replacements looked up in the index
//   K2e k_seed_dirty_expand  (overflow only) dirty windows x orientations -> candidates
// All of it is integer/byte work; exactness is in k_verify (kernels.cuh).
#pragma once

#include "device_types.cuh"
#include "kernels.cuh"

namespace ipcr {

constexpr uint32_t kSeedChunk = 4096;  // positions one warp takes at a time (128 per lane)
constexpr uint32_t kSeedBloomWordsDev = 57344;  // 224 KB
constexpr uint32_t kSeedHalfWordsDev = 2048;    // pigeonhole first level: one bit per 8-mer, 8 KB per half
constexpr uint32_t kSeedBloomMulDev = 0x85EBCA6Bu;
constexpr uint32_t kSeedHashMulDev = 0x9E3779B1u;

// ---------------------------------------------------------------------------
// K1s.  Four bytes at a time, no table and no per-byte branch:
//   code  = (byte >> 1) & 3                       A 0, C 1, T 2, G 3
//   valid = byte == 0x41 + 2*code + 15*[code==2]  (re-encode and compare: only uppercase
//           A/C/G/T survive, core/primer/iupac.go:62-67)
// and two multiplies gather the 4x2 code bits / 4x1 invalid bits.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void pack4(uint32_t w, uint32_t &code8, uint32_t &inv4) {
    const uint32_t t = (w >> 1) & 0x03030303u;
    code8 = (t * 0x01041040u) >> 24;
    const uint32_t e2 = (t >> 1) & ~t & 0x01010101u;
    const uint32_t expect = 0x41414141u + 2u * t + 15u * e2;
    const uint32_t d = expect ^ w;
    const uint32_t nz = (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;
    inv4 = ((nz >> 7) * 0x10204080u) >> 28;
}

struct Pack2Args {
    const uint8_t *ascii;
    uint64_t total_len;
    uint64_t n_groups;  // 32-base groups to write (covers total_len plus padding)
    uint32_t *codes;    // [2 * n_groups]
    uint32_t *inval;    // [n_groups]
    const SegmentDev *segs;
    uint32_t n_segs;
    uint32_t *seg_reset;
    uint8_t *rec_reset;  // optional: per-record flag, raised here so that the asynchronous chain never reads seg_reset on the host
    uint32_t aligned;   // ascii is 16-byte aligned
};

__global__ void __launch_bounds__(256)
k_pack2(Pack2Args a) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t gidx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; gidx < a.n_groups; gidx += stride) {
        const uint64_t P0 = gidx * 32ull;
        uint32_t w[8];
        if (a.aligned && P0 + 32ull <= a.total_len) {
            const uint4 q0 = __ldg(reinterpret_cast<const uint4 *>(a.ascii + P0));
            const uint4 q1 = __ldg(reinterpret_cast<const uint4 *>(a.ascii + P0) + 1);
            w[0] = q0.x; w[1] = q0.y; w[2] = q0.z; w[3] = q0.w;
            w[4] = q1.x; w[5] = q1.y; w[6] = q1.z; w[7] = q1.w;
        } else {  // ragged tail / unaligned buffer: byte loads, zero (= invalid) padding
#pragma unroll
            for (int k = 0; k < 8; k++) w[k] = 0u;
            for (uint32_t k = 0; k < 32u && P0 + k < a.total_len; k++)
                w[k >> 2] |= (uint32_t)a.ascii[P0 + k] << (8u * (k & 3u));
        }
        uint32_t c0 = 0, c1 = 0, inv = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            uint32_t c8, i4;
            pack4(w[k], c8, i4);
            if (k < 4) c0 |= c8 << (8 * k); else c1 |= c8 << (8 * (k - 4));
            inv |= i4 << (4 * k);
        }
        reinterpret_cast<uint2 *>(a.codes)[gidx] = make_uint2(c0, c1);
        a.inval[gidx] = inv;
        if (inv != 0u) {  // rare: is one of them a reset byte (not even lowercase acgt)?
            uint32_t first = 32u;
            for (uint32_t k = 0; k < 32u && first == 32u; k++) {
                if (!((inv >> k) & 1u) || P0 + k >= a.total_len) continue;
                const uint32_t b = (w[k >> 2] >> (8u * (k & 3u))) & 0xFFu;
                const uint32_t x = (b | 0x20u) - 'a';  // a,c,g,t -> 0,2,6,19
                if (!(x < 32u && ((0x00080045u >> x) & 1u))) first = k;
            }
            if (first < 32u) {
                const uint32_t s0 = find_segment(a.segs, a.n_segs, P0 + first);
                const uint64_t seg_end = a.segs[s0].global_off + a.segs[s0].len;
                if (P0 + first < seg_end && a.seg_reset[s0] == 0u) {
                    a.seg_reset[s0] = 1u;
                    if (a.rec_reset) a.rec_reset[a.segs[s0].rec_idx] = 1u;
                }
                if (seg_end < P0 + 32u) {  // the group straddles a segment boundary: look at every byte
                    for (uint32_t k = first + 1; k < 32u; k++) {
                        if (!((inv >> k) & 1u) || P0 + k >= a.total_len) continue;
                        const uint32_t b = (w[k >> 2] >> (8u * (k & 3u))) & 0xFFu;
                        const uint32_t x = (b | 0x20u) - 'a';
                        if (x < 32u && ((0x00080045u >> x) & 1u)) continue;
                        const uint32_t s1 = find_segment(a.segs, a.n_segs, P0 + k);
                        if (P0 + k - a.segs[s1].global_off < a.segs[s1].len) {
                            a.seg_reset[s1] = 1u;
                            if (a.rec_reset) a.rec_reset[a.segs[s1].rec_idx] = 1u;
                        }
                    }
                }
            }
        }
    }
}

// The hard-mismatch bitmap from its run list (pack_host.hpp: packed_invalid_runs): one warp per run, the lanes
// stride over the run's words; partial first / last words are OR-ed in atomically (two runs may share a word).
// The bitmap was zeroed before.
__global__ void __launch_bounds__(256)
k_inval_fill(const unsigned long long *__restrict__ runs, uint64_t n_runs, uint32_t *__restrict__ inval) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < n_runs; r += n_warps) {
        const uint64_t a = runs[2 * r], b = runs[2 * r + 1];  // [a, b)
        if (a >= b) continue;
        const uint64_t w0 = a >> 5, w1 = (b - 1) >> 5;
        for (uint64_t w = w0 + lane; w <= w1; w += 32u) {
            uint32_t m = 0xFFFFFFFFu;
            if (w == w0) m &= 0xFFFFFFFFu << (a & 31u);
            if (w == w1 && (b & 31u) != 0u) m &= (1u << (b & 31u)) - 1u;
            if (m == 0xFFFFFFFFu) inval[w] = m;  // interior words belong to this run alone
            else atomicOr(&inval[w], m);
        }
    }
}

struct SeedExpandArgs {
    const unsigned long long *dirty;
    uint64_t n_dirty;
    const uint16_t *col_off;
    const OrientDev *orients;
    uint32_t n_orient;
    uint32_t S;
    unsigned long long *cand;  // already offset to the first free entry
};

__global__ void __launch_bounds__(256)
k_seed_dirty_expand(SeedExpandArgs a) {
    const uint64_t n = a.n_dirty * a.n_orient;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t q = (uint32_t)(i % a.n_orient);
        const uint64_t p = a.dirty[i / a.n_orient];
        const uint32_t c = a.col_off[q];
        const bool ok = a.orients[q].L >= a.S && p >= c;
        a.cand[i] = ((unsigned long long)q << kCandPosBits) | (ok ? p - c : kCandPosMask);  // else: never a site
    }
}

// ---------------------------------------------------------------------------
// K2s.  block = 32 warps, one CTA per SM (224 KB Bloom bitmap in shared memory).  A warp takes
// 4096 consecutive positions; lane l owns the 128 k-mer starts at +128*l: its 8 code words
// (+1 look-ahead) and 4 bitmap words (+1) sit in registers, every shift is an immediate.
// Per position: funnel-shift + mask -> k-mer; two IMADs -> bitmap index; ONE LDS; the answer
// bit is funnel-shifted into a 32-position hit word.  Set bits (and nothing else) walk the
// hash table.  Windows holding an invalid byte are masked out and, if still viable, listed as dirty.
// ---------------------------------------------------------------------------
struct SeedScanArgs {
    const uint32_t *codes;
    const uint32_t *inval;
    uint64_t n_chunks;
    uint32_t S;
    uint32_t d;          // S - S1: the first-level key is the high S1 bases of the seed window
    uint32_t kmer_mask;
    uint32_t bloom_mul;  // kSeedBloomMul << (32 - 2*S1)
    const uint32_t *bloom;
    const uint32_t *bitmap2;  // second level (global, L2-resident)
    uint32_t bitmap2_mul, bitmap2_shift;  // h = k-mer * mul: word = h >> shift, bit = h & 31
    const unsigned long long *table;
    uint32_t table_shift;  // 32 - table_bits
    uint32_t table_mask;
    unsigned long long *cand;
    unsigned long long *cand_count;
    uint64_t cand_cap;
    // dirty windows (k-mer starts whose window holds 1..max_mm invalid bytes)
    uint64_t total_len;
    uint32_t max_mm;
    unsigned long long *dirty;
    unsigned long long *dirty_count;
    uint64_t dirty_cap;
    unsigned long long *chunk_counter;  // VAR & 4: next unclaimed chunk (zeroed before the launch)
    uint32_t prefetch;  // != 0: a warp asks L2 for its NEXT chunk's sectors before it works on this one
};

// L2 residency: the second-level bitmap (<= 32 MB) and the hash table are hit at random by
// every SM and must stay in the 126 MB L2 while 1+ GB of packed genome streams through it.
// Lookups carry an evict_last policy, the streaming loads an evict_first one.
__device__ __forceinline__ uint64_t l2_policy_keep() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_stream() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint32_t ld_u32_policy(const uint32_t *p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ unsigned long long ld_u64_policy(const unsigned long long *p, uint64_t pol) {
    unsigned long long v;
    asm volatile("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint4 ld_u128_policy(const uint4 *p, uint64_t pol) {
    uint4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}

// candidates are rare and arrive in divergent code: the lanes that are here together elect a
// leader and take ONE atomic for all of them (warp-aggregated append)
__device__ __forceinline__ void seed_emit(const SeedScanArgs &a, unsigned long long slot, uint64_t p) {
    const uint64_t c = (slot >> 48) & 0xFFull;
    if (p >= c) {
        const uint32_t peers = __activemask();
        const uint32_t leader = (uint32_t)__ffs(peers) - 1u;
        unsigned long long base = 0ull;
        if ((threadIdx.x & 31u) == leader) base = atomicAdd(a.cand_count, (unsigned long long)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        const unsigned long long i = base + __popc(peers & ((1u << (threadIdx.x & 31u)) - 1u));
        if (i < a.cand_cap) a.cand[i] = (((slot >> 32) & 0xFFFFull) << kCandPosBits) | (p - c);
    }
}

__device__ __forceinline__ void seed_walk(const SeedScanArgs &a, uint32_t kmer, uint64_t p, uint64_t keep) {
    uint32_t h = (kmer * kSeedHashMulDev) >> a.table_shift;
    for (;;) {
        const unsigned long long slot = ld_u64_policy(a.table + h, keep);
        if (slot == 0ull) break;
        if ((uint32_t)slot == kmer) seed_emit(a, slot, p);
        h = (h + 1u) & a.table_mask;
    }
}

// BLOOM: 0 = no shared-memory bitmap (every position goes to the second level), k = 1..2 bits per k-mer, both
// in the same word, 3 = pigeonhole half-window tables (see below).
// VAR (tuning switches, IPCR_SEED_VAR; measured on B200, profiles/r02_seed_variants.txt: +-0.5 %, default 0):
// 1 = the look-ahead / look-behind words come from the neighbouring lane by shuffle instead of five extra strided
// loads per chunk; 4 = warps claim chunks from a global counter (no tail, no imbalance between SMs) instead of a
// static stride.  (Forming the bitmap address with a multiply-add to move it off the ALU pipe does not survive
// ptxas: a multiply by 4 becomes LEA again.)
// T (BLOOM 1 / 2 only): anchor stride of the first level.  A panel with few keys does not need a lookup per position:
// the index holds, per S-mer, its S1-mers at the T window offsets d-T+1 .. d, so the S1-mer at every T-th
// position ("anchor") answers for the T windows that contain it at one of those offsets.  A set anchor bit is
// then spread over its T windows and the second level sorts them out.  T lookups fewer per T positions; the
// key set grows T-fold, so seed.hpp's cost model only picks T > 1 while the bitmap stays sparse.
template <int BLOOM, int VAR = 0, int T = 1>
__global__ void __launch_bounds__(1024, 1)
k_seed_scan(SeedScanArgs a) {
    extern __shared__ uint32_t sbloom[];
    if (BLOOM) {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.bloom);
        uint4 *dst = reinterpret_cast<uint4 *>(sbloom);
        const uint32_t n16 = BLOOM == 3 ? kSeedHalfWordsDev * 2u / 4u : kSeedBloomWordsDev / 4u;
        for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
        __syncthreads();
    }
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t n_warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    const uint64_t keep = l2_policy_keep(), stream = l2_policy_stream();
    // The first level looks at the S1 bases at position q; a hit means "the S-mer starting at
    // p = q - d may be in the index" (d = S - S1 >= 0: the first-level key is the HIGH end of the
    // seed window).  The lane therefore keeps one code word and one bitmap word to the LEFT of its
    // 128 positions: c[0] / iv[0].
    const uint32_t d = a.d;
    uint64_t ch = warp;
    if (VAR & 4) {
        if (lane == 0) ch = atomicAdd(a.chunk_counter, 1ull);
        ch = __shfl_sync(0xFFFFFFFFu, ch, 0);
    }
    while (ch < a.n_chunks) {
        const uint64_t P0 = ch * kSeedChunk + (uint64_t)lane * 128u;
        uint32_t c[10], iv[6];
        if (!(VAR & 4) && a.prefetch != 0u && ch + n_warps < a.n_chunks) {
            // a warp's loads and its lookups are one dependent chain per chunk; with a strided first level the
            // chain is short and the DRAM latency shows.  The next chunk (one lane = one 32-byte sector of codes,
            // every other lane one sector of the bitmap) is on its way into L2 while this one is worked on.
            const uint64_t Pn = P0 + n_warps * kSeedChunk;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.codes + (Pn >> 4)));
            if ((lane & 1u) == 0u) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.inval + (Pn >> 5)));
        }
        {
            const uint4 *cp = reinterpret_cast<const uint4 *>(a.codes + (P0 >> 4));
            const uint4 q0 = ld_u128_policy(cp, stream), q1 = ld_u128_policy(cp + 1, stream);
            c[1] = q0.x; c[2] = q0.y; c[3] = q0.z; c[4] = q0.w;
            c[5] = q1.x; c[6] = q1.y; c[7] = q1.z; c[8] = q1.w;
            const uint4 qi = ld_u128_policy(reinterpret_cast<const uint4 *>(a.inval + (P0 >> 5)), stream);
            iv[1] = qi.x; iv[2] = qi.y; iv[3] = qi.z; iv[4] = qi.w;
            c[0] = 0u;
            iv[0] = 0xFFFFFFFFu;  // before the buffer: no window starts there
            if (VAR & 1) {
                // lane l's look-ahead word is lane l+1's first word, its look-behind word lane l-1's last:
                // four shuffles, and real loads only at the two ends of the warp's 4096 positions
                c[9] = __shfl_down_sync(0xFFFFFFFFu, c[1], 1);
                iv[5] = __shfl_down_sync(0xFFFFFFFFu, iv[1], 1);
                const uint32_t cb = __shfl_up_sync(0xFFFFFFFFu, c[8], 1), ib = __shfl_up_sync(0xFFFFFFFFu, iv[4], 1);
                if (lane == 31u) {
                    c[9] = ld_u32_policy(a.codes + (P0 >> 4) + 8, stream);
                    iv[5] = ld_u32_policy(a.inval + (P0 >> 5) + 4, stream);
                }
                if (d != 0u) {
                    if (lane != 0u) {
                        c[0] = cb;
                        iv[0] = ib;
                    } else if (P0 != 0ull) {
                        c[0] = ld_u32_policy(a.codes + (P0 >> 4) - 1, stream);
                        iv[0] = ld_u32_policy(a.inval + (P0 >> 5) - 1, stream);
                    }
                }
            } else {
                c[9] = ld_u32_policy(a.codes + (P0 >> 4) + 8, stream);
                iv[5] = ld_u32_policy(a.inval + (P0 >> 5) + 4, stream);
                if (d != 0u && P0 != 0ull) {
                    c[0] = ld_u32_policy(a.codes + (P0 >> 4) - 1, stream);
                    iv[0] = ld_u32_policy(a.inval + (P0 >> 5) - 1, stream);
                }
            }
        }
#pragma unroll
        for (int blk = 0; blk < 4; blk++) {
            const uint32_t wm = c[2 * blk], w0 = c[2 * blk + 1], w1 = c[2 * blk + 2], w2 = c[2 * blk + 3];
            uint32_t hits = 0u;
            if (BLOOM == 3) {
                // PIGEONHOLE first level (S = 16): the key set of a big panel at mm 3 saturates any bitmap that
                // fits shared memory, but a window with <= m mismatches has <= m / 2 of them in its low or in its
                // high 8 bases.  Two exact 64 Kbit tables (8 KB each) hold every 8-mer within m / 2 mismatches of
                // some orientation's low / high half; a position goes on iff EITHER half is in its table.
                const char *sb = reinterpret_cast<const char *>(sbloom);
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const uint32_t x = j < 16 ? __funnelshift_r(w0, w1, 2 * j) : __funnelshift_r(w1, w2, 2 * (j - 16));
                    const uint32_t wl = *reinterpret_cast<const uint32_t *>(sb + ((x >> 3) & 0x1FFCu));
                    const uint32_t wh = *reinterpret_cast<const uint32_t *>(sb + kSeedHalfWordsDev * 4u + ((x >> 19) & 0x1FFCu));
                    const uint32_t bit = __funnelshift_r(wl, 0u, x) | __funnelshift_r(wh, 0u, x >> 16);
                    hits = __funnelshift_r(hits, bit, 1);
                }
            } else if (BLOOM) {
                const char *sb = reinterpret_cast<const char *>(sbloom);
#pragma unroll
                for (int j = 0; j < 32; j += T) {
                    // x: 32-bit window whose low 2*S1 bits are the first-level key.  The multiplier is
                    // pre-shifted by 32-2*S1, so the product ignores the bits above the key: no mask needed.
                    const uint32_t x = j < 16 ? __funnelshift_r(w0, w1, 2 * j) : __funnelshift_r(w1, w2, 2 * (j - 16));
                    // The ALU pipe (SHF/LOP3/LEA, half rate) is the busy one: the byte offset and the
                    // second bit position are formed with multiplies on the FMA pipe instead.
                    const uint32_t hw = x * a.bloom_mul;
                    uint32_t off, b2;
                    off = __umulhi(hw, kSeedBloomWordsDev) * 4u;
                    const uint32_t w = *reinterpret_cast<const uint32_t *>(sb + off);
                    uint32_t bit = __funnelshift_r(w, 0u, x);  // bit (key & 31)
                    if (BLOOM == 2) {
                        asm("mul.hi.u32 %0, %1, 2097152;" : "=r"(b2) : "r"(hw));  // hw >> 11: hash bits 11..15
                        bit &= __funnelshift_r(w, 0u, b2);
                    }
                    if (T > 1) bit &= 1u;  // T bits enter the hit word per step: only the answer may be set
                    hits = __funnelshift_r(hits, bit, T);
                }
                if (T > 1) {
                    // anchor bit j answers for the windows whose hit bits are j .. j+T-1 (j + T <= 32).  Their T
                    // second-level lookups are independent: issued together they cost one L2 round trip, not T
                    // (this code is divergent, the rest of the warp waits for it).
                    uint32_t anchors = hits;
                    hits = 0u;
                    while (anchors) {
                        uint32_t j;
                        asm("bfind.u32 %0, %1;" : "=r"(j) : "r"(anchors));
                        anchors ^= 1u << j;
                        uint32_t wd[T], h2[T];
#pragma unroll
                        for (int i = 0; i < T; i++) {
                            const int32_t jp = (int32_t)j + i - (int32_t)d;  // window start relative to the block
                            const uint32_t lo = jp < 0 ? wm : (jp < 16 ? w0 : w1), hi = jp < 0 ? w0 : (jp < 16 ? w1 : w2);
                            h2[i] = (__funnelshift_r(lo, hi, 2u * (uint32_t)jp) & a.kmer_mask) * a.bitmap2_mul;
                            wd[i] = ld_u32_policy(a.bitmap2 + (h2[i] >> a.bitmap2_shift), keep);
                        }
#pragma unroll
                        for (int i = 0; i < T; i++) hits |= (__funnelshift_r(wd[i], 0u, h2[i]) & 1u) << (j + i);
                    }
                }
            } else if (d == 0u) {
                // No first level (the panel's keys saturate any bitmap that fits shared memory, e.g. mm 3 x 400
                // orientations): EVERY position asks the L2-resident second-level bitmap.  The 32 lookups of a
                // block are independent: issued back to back they overlap their L2 latency instead of paying it
                // one by one in the walk loop below, which then only sees the ~1.5 % that pass.
                uint32_t wd[32], sh[32];
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const uint32_t x = j < 16 ? __funnelshift_r(w0, w1, 2 * j) : __funnelshift_r(w1, w2, 2 * (j - 16));
                    const uint32_t h2 = (x & a.kmer_mask) * a.bitmap2_mul;
                    sh[j] = h2;
                    wd[j] = ld_u32_policy(a.bitmap2 + (h2 >> a.bitmap2_shift), keep);
                }
#pragma unroll
                for (int j = 0; j < 32; j++) hits = __funnelshift_r(hits, __funnelshift_r(wd[j], 0u, sh[j]), 1);
            } else {
                hits = 0xFFFFFFFFu;
            }
            const uint32_t ipr = iv[blk], ilo = iv[blk + 1], ihi = iv[blk + 2];
            // windows are addressed by their START p.  A lane lists the dirty windows that start in
            // ITS 128 positions, and masks the hits q whose window p = q - d is not clean.
            if ((ilo | ihi) != 0u || (d != 0u && (ipr >> (32u - d)) != 0u)) {
                // rare: a window holding an invalid byte has no k-mer.  Drop it here; if it holds
                // 1..max_mm of them it may still be a site (genome N = one mismatch): those "dirty"
                // windows are listed and resolved by k_seed_dirty_probe (halo.go:47-108).
                const uint64_t v = (uint64_t)ilo | ((uint64_t)ihi << 32);
                const uint64_t smask = (1ull << a.S) - 1ull;
                uint64_t spread = 0ull;
                for (uint32_t k = 0; k < a.S; k++) spread |= v >> k;
                const uint32_t dr = (uint32_t)spread;  // windows [j, j+S) of this block that hold an invalid byte
                if (a.max_mm > 0u && ((ilo & ihi) != 0xFFFFFFFFu || a.S <= a.max_mm)) {
                    uint32_t elig = 0u;  // dirty windows that can still be sites (inside an N run none can)
                    for (uint32_t todo = dr; todo; todo &= todo - 1u) {
                        const uint32_t j = (uint32_t)(__ffs(todo) - 1);
                        const uint32_t n = (uint32_t)__popcll((v >> j) & smask);
                        if (n <= a.max_mm && P0 + (uint64_t)(blk * 32) + j + a.S <= a.total_len) elig |= 1u << j;
                    }
                    if (elig) {  // one atomic per lane and block, not per window
                        unsigned long long i = atomicAdd(a.dirty_count, (unsigned long long)__popc(elig));
                        for (; elig; elig &= elig - 1u, i++)
                            if (i < a.dirty_cap) a.dirty[i] = P0 + (uint64_t)(blk * 32) + (uint32_t)(__ffs(elig) - 1);
                    }
                }
                // hit bit j stands for the window starting d positions earlier
                uint32_t df = dr;
                if (d != 0u) {
                    const uint64_t vf = (v << d) | (uint64_t)(ipr >> (32u - d));
                    uint64_t sf = 0ull;
                    for (uint32_t k = 0; k < a.S; k++) sf |= vf >> k;
                    df = (uint32_t)sf;
                }
                hits &= ~df;
            }
            // second level for the set bits: one bit per (hashed) S-mer in an L2-resident bitmap;
            // only its hits walk the hash table.  The loop is written for instruction count: highest
            // set bit first = one FLO, the funnel shift wraps its amount mod 32, the bitmap word index
            // and bit position come straight from the hash, the bit test sets a predicate.
            constexpr int NB = (VAR & 2) ? 2 : ((VAR & 8) ? 3 : 1);
            if (NB > 1 && T == 1 && !(BLOOM == 0)) {
                // A lane's passes are a chain of dependent L2 round trips (the warp loops max-over-lanes times and
                // nothing else of this warp runs meanwhile): NB lookups per trip shorten the chain NB-fold.
                while (hits) {
                    uint32_t jj[NB], km[NB], hh[NB], wd[NB];
#pragma unroll
                    for (int b = 0; b < NB; b++) {
                        jj[b] = 0xFFFFFFFFu;
                        wd[b] = 0u;
                        if (hits) {
                            uint32_t j;
                            asm("bfind.u32 %0, %1;" : "=r"(j) : "r"(hits));
                            hits ^= 1u << j;
                            jj[b] = j;
                            const int32_t jp = (int32_t)j - (int32_t)d;
                            const uint32_t lo = jp < 0 ? wm : (jp < 16 ? w0 : w1), hi = jp < 0 ? w0 : (jp < 16 ? w1 : w2);
                            km[b] = __funnelshift_r(lo, hi, 2u * (uint32_t)jp) & a.kmer_mask;
                            hh[b] = km[b] * a.bitmap2_mul;
                            wd[b] = ld_u32_policy(a.bitmap2 + (hh[b] >> a.bitmap2_shift), keep);
                        }
                    }
#pragma unroll
                    for (int b = 0; b < NB; b++)
                        if (jj[b] != 0xFFFFFFFFu && (wd[b] & __funnelshift_l(0u, 1u, hh[b])))
                            seed_walk(a, km[b], P0 + (uint64_t)(blk * 32) + jj[b] - d, keep);
                }
            }
            while (hits) {
                uint32_t j;
                asm("bfind.u32 %0, %1;" : "=r"(j) : "r"(hits));  // position of the highest set bit
                hits ^= 1u << j;
                const int32_t jp = (int32_t)j - (int32_t)d;  // window start relative to the block: -d .. 31
                const uint32_t lo = jp < 0 ? wm : (jp < 16 ? w0 : w1), hi = jp < 0 ? w0 : (jp < 16 ? w1 : w2);
                const uint32_t kmer = __funnelshift_r(lo, hi, 2u * (uint32_t)jp) & a.kmer_mask;
                const uint32_t h2 = kmer * a.bitmap2_mul;
                if ((BLOOM == 0 && d == 0u) || T > 1 ||  // already tested above
                    (ld_u32_policy(a.bitmap2 + (h2 >> a.bitmap2_shift), keep) & __funnelshift_l(0u, 1u, h2)))
                    seed_walk(a, kmer, P0 + (uint64_t)(blk * 32) + j - d, keep);
            }
        }
        if (VAR & 4) {
            if (lane == 0) ch = atomicAdd(a.chunk_counter, 1ull);
            ch = __shfl_sync(0xFFFFFFFFu, ch, 0);
        } else {
            ch += n_warps;
        }
    }
}

// ---------------------------------------------------------------------------
// K2d.  Dirty windows (1..max_mm invalid bytes): an invalid genome byte is a mismatch at its
// column whatever the primer says, so a site through this window shows, after replacing each
// invalid byte by SOME base, a k-mer of the index.  One thread per window tries all 4^t
// replacements (t <= 4), collects the distinct (orientation, seed offset) hits and appends
// them as candidates.  Windows with more than kDirtySeen hits or t > 4 go to the overflow
// list and are expanded against every orientation (k_seed_dirty_expand).
// Replaces the direct verification of scanNonACGTHalos (core/engine/halo.go:47-108).
// ---------------------------------------------------------------------------
constexpr int kDirtySeen = 16;
constexpr uint32_t kDirtyMaxInvalid = 4;

struct SeedDirtyProbeArgs {
    const unsigned long long *dirty;
    uint64_t n_dirty;
    const uint32_t *codes;
    const uint32_t *inval;
    uint32_t S;
    uint32_t kmer_mask;
    const uint32_t *bitmap2;
    uint32_t bitmap2_mul, bitmap2_shift;
    const unsigned long long *table;
    uint32_t table_shift, table_mask;
    unsigned long long *cand;
    unsigned long long *cand_count;
    uint64_t cand_cap;
    unsigned long long *over;  // capacity n_dirty
    unsigned long long *over_count;
};

__global__ void __launch_bounds__(128)
k_seed_dirty_probe(SeedDirtyProbeArgs a) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_dirty; i += stride) {
        const uint64_t p = a.dirty[i];
        const uint32_t sh = (uint32_t)p & 15u;
        const uint32_t kmer0 = __funnelshift_r(a.codes[p >> 4], a.codes[(p >> 4) + 1], 2u * sh) & a.kmer_mask;
        const uint64_t ivw = (uint64_t)a.inval[p >> 5] | ((uint64_t)a.inval[(p >> 5) + 1] << 32);
        uint32_t iv = (uint32_t)(ivw >> (p & 31u)) & ((1u << a.S) - 1u);  // S <= 16
        const uint32_t t = __popc(iv);
        bool overflow = t > kDirtyMaxInvalid;
        uint32_t seen[kDirtySeen];
        int n_seen = 0;
        if (!overflow) {
            uint32_t pos[kDirtyMaxInvalid] = {0u, 0u, 0u, 0u};
            uint32_t base = kmer0;
            for (uint32_t u = 0; u < t; u++) {
                pos[u] = (uint32_t)(__ffs(iv) - 1);
                iv &= iv - 1u;
                base &= ~(3u << (2u * pos[u]));
            }
            for (uint32_t combo = 0; combo < (1u << (2u * t)) && !overflow; combo++) {
                uint32_t km = base;
                for (uint32_t u = 0; u < t; u++) km |= ((combo >> (2u * u)) & 3u) << (2u * pos[u]);
                const uint32_t h2 = km * a.bitmap2_mul;
                if (!((a.bitmap2[h2 >> a.bitmap2_shift] >> (h2 & 31u)) & 1u)) continue;
                for (uint32_t h = (km * kSeedHashMulDev) >> a.table_shift;; h = (h + 1u) & a.table_mask) {
                    const unsigned long long slot = a.table[h];
                    if (slot == 0ull) break;
                    if ((uint32_t)slot != km) continue;
                    const uint32_t key = (uint32_t)(slot >> 32) & 0x00FFFFFFu;  // orientation | seed offset << 16
                    bool dup = false;
                    for (int k = 0; k < n_seen; k++) dup |= seen[k] == key;
                    if (dup) continue;
                    if (n_seen == kDirtySeen) { overflow = true; break; }
                    seen[n_seen++] = key;
                }
            }
        }
        if (overflow) {
            a.over[atomicAdd(a.over_count, 1ull)] = p;
            continue;
        }
        for (int k = 0; k < n_seen; k++) {
            const uint64_t c = seen[k] >> 16;
            if (p < c) continue;
            const unsigned long long idx = atomicAdd(a.cand_count, 1ull);
            if (idx < a.cand_cap) a.cand[idx] = ((unsigned long long)(seen[k] & 0xFFFFu) << kCandPosBits) | (p - c);
        }
    }
}

// ---------------------------------------------------------------------------
// K2r.  Dirty windows, resolved in ONE launch of the asynchronous chain: the replacement
// lookups of k_seed_dirty_probe, followed directly by the exact verify of every distinct
// (orientation, seed offset) hit; windows the probe cannot resolve are verified against every
// orientation.  The window count is read from the device, so the host never waits for it.
// ---------------------------------------------------------------------------
struct SeedDirtyResolveArgs {
    SeedDirtyProbeArgs p;              // cand/over fields unused
    const unsigned long long *n_dirty_ptr;
    uint64_t dirty_cap;
    const uint16_t *col_off;
    uint32_t n_orient;
    VerifyArgs v;                      // cand fields unused
};

__device__ __forceinline__ void dirty_verify_emit(const VerifyArgs &v, uint32_t q, uint64_t start) {
    ipcr_site s;
    if (verify_one(v, q, start, &s)) {
        const unsigned long long dst = atomicAdd(v.site_count, 1ull);
        if (dst < v.site_cap) v.sites[dst] = s;
    }
}

__global__ void __launch_bounds__(128)
k_seed_dirty_resolve(SeedDirtyResolveArgs r) {
    const SeedDirtyProbeArgs &a = r.p;
    const uint64_t n_dirty = min((uint64_t)*r.n_dirty_ptr, r.dirty_cap);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_dirty; i += stride) {
        const uint64_t p = a.dirty[i];
        const uint32_t sh = (uint32_t)p & 15u;
        const uint32_t kmer0 = __funnelshift_r(a.codes[p >> 4], a.codes[(p >> 4) + 1], 2u * sh) & a.kmer_mask;
        const uint64_t ivw = (uint64_t)a.inval[p >> 5] | ((uint64_t)a.inval[(p >> 5) + 1] << 32);
        uint32_t iv = (uint32_t)(ivw >> (p & 31u)) & ((1u << a.S) - 1u);  // S <= 16
        const uint32_t t = __popc(iv);
        bool overflow = t > kDirtyMaxInvalid;
        uint32_t seen[kDirtySeen];
        int n_seen = 0;
        if (!overflow) {
            uint32_t pos[kDirtyMaxInvalid] = {0u, 0u, 0u, 0u};
            uint32_t base = kmer0;
            for (uint32_t u = 0; u < t; u++) {
                pos[u] = (uint32_t)(__ffs(iv) - 1);
                iv &= iv - 1u;
                base &= ~(3u << (2u * pos[u]));
            }
            for (uint32_t combo = 0; combo < (1u << (2u * t)) && !overflow; combo++) {
                uint32_t km = base;
                for (uint32_t u = 0; u < t; u++) km |= ((combo >> (2u * u)) & 3u) << (2u * pos[u]);
                const uint32_t h2 = km * a.bitmap2_mul;
                if (!((a.bitmap2[h2 >> a.bitmap2_shift] >> (h2 & 31u)) & 1u)) continue;
                for (uint32_t h = (km * kSeedHashMulDev) >> a.table_shift;; h = (h + 1u) & a.table_mask) {
                    const unsigned long long slot = a.table[h];
                    if (slot == 0ull) break;
                    if ((uint32_t)slot != km) continue;
                    const uint32_t key = (uint32_t)(slot >> 32) & 0x00FFFFFFu;  // orientation | seed offset << 16
                    bool dup = false;
                    for (int k = 0; k < n_seen; k++) dup |= seen[k] == key;
                    if (dup) continue;
                    if (n_seen == kDirtySeen) { overflow = true; break; }
                    seen[n_seen++] = key;
                }
            }
        }
        if (overflow) {  // rare: every orientation whose seed window can sit here
            for (uint32_t q = 0; q < r.n_orient; q++) {
                const uint32_t c = r.col_off[q];
                if (r.v.orients[q].L >= a.S && p >= c) dirty_verify_emit(r.v, q, p - c);
            }
            continue;
        }
        for (int k = 0; k < n_seen; k++) {
            const uint64_t c = seen[k] >> 16;
            if (p >= c) dirty_verify_emit(r.v, seen[k] & 0xFFFFu, p - c);
        }
    }
}

}  // namespace ipcr
