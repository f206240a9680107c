// kernels.cuh -- hand-written sm_100a kernels of the scan path (everything except
// the NVRTC-specialised filter, which is generated in jit.hpp):
//   K1  k_pack            ASCII -> striped one-hot bit-planes (+ reset-byte flags)
//   K2g k_filter_generic  table-driven bit-parallel candidate filter
//   K3  k_verify          exact IUPAC verify of candidates + warp-ballot compaction
//   K4  radix sort / scan (sortscan.cuh)
//   K5  k_site_keys .. k_join  hit cap, segmenting and amplicon pairing
#pragma once

#include "device_types.cuh"

namespace ipcr {

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// ---------------------------------------------------------------------------
// K1: pack.  Replaces the genome side of primer.BaseMatch
// (/root/reference/core/primer/iupac.go:62-67): only uppercase A/C/G/T set a bit.
// grid = (n_tiles, ceil(rows/32)), block = (32, 32): x = row in chunk (consecutive
// bytes -> coalesced reads), y = lane.  The 32x32 uint4 results are transposed
// through shared memory so that each warp writes 512 contiguous bytes.
// Also raises seg_reset[s] when segment s holds a byte outside ACGTacgt
// (sequenceHasAutomatonReset, core/engine/ac.go:141-148).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t find_segment(const SegmentDev *segs, uint32_t n_segs, uint64_t P) {
    uint32_t lo = 0, hi = n_segs;  // last segment with global_off <= P
    while (lo + 1 < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (segs[mid].global_off <= P) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(1024)
k_pack_bytes(const uint8_t *__restrict__ ascii, GenomeLayout lay, uint4 *__restrict__ planes,
       const SegmentDev *__restrict__ segs, uint32_t n_segs, uint32_t *__restrict__ seg_reset,
       uint8_t *__restrict__ rec_reset) {
    __shared__ uint4 tile[32][33];
    const uint64_t t = blockIdx.x;
    const uint32_t r = blockIdx.y * 32u + threadIdx.x;
    const uint32_t l = threadIdx.y;
    uint32_t a = 0, c = 0, g = 0, tt = 0, rs = 0;
    if (r < lay.rows) {
        const uint64_t stripe0 = t * kStripesPerTile + (uint64_t)l * 32u;
#pragma unroll 8
        for (int i = 0; i < 32; i++) {
            const uint64_t P = (stripe0 + i) * lay.W + r;
            if (P < lay.total_len) {
                const uint32_t b = ascii[P];
                a |= (uint32_t)(b == 'A') << i;
                c |= (uint32_t)(b == 'C') << i;
                g |= (uint32_t)(b == 'G') << i;
                tt |= (uint32_t)(b == 'T') << i;
                const uint32_t u = b & 0xDFu;
                rs |= (uint32_t)!(u == 'A' || u == 'C' || u == 'G' || u == 'T') << i;
            }
        }
        if (rs != 0 && r < lay.W) {  // rare path: attribute reset bytes to their segment
            for (int i = 0; i < 32; i++) {
                if (!((rs >> i) & 1u)) continue;
                const uint64_t P = (stripe0 + i) * lay.W + r;
                const uint32_t s = find_segment(segs, n_segs, P);
                if (P - segs[s].global_off < segs[s].len) {
                    seg_reset[s] = 1u;
                    if (rec_reset) rec_reset[segs[s].rec_idx] = 1u;
                }
            }
        }
    }
    tile[threadIdx.x][threadIdx.y] = make_uint4(a, c, g, tt);
    __syncthreads();
    const uint32_t rr = blockIdx.y * 32u + threadIdx.y;
    if (rr < lay.rows)
        planes[(t * lay.rows + rr) * kLanes + threadIdx.x] = tile[threadIdx.y][threadIdx.x];
}

// K1 fast path (16-byte aligned input, W % 16 == 0).  grid = n_tiles, block = 32 warps:
// warp y packs lane y of the tile, thread i owns stripe (tile*1024 + y*32 + i).  Each
// thread streams ITS stripe with 16-byte loads (16 rows per load; a 128-B line is used up
// by 8 consecutive loads of the same thread), and the byte -> bit-plane TRANSPOSE is done
// by the warp vote unit: __ballot_sync(byte == 'A') over the 32 stripes of a lane IS the
// A-plane word of that (row, lane).  16 rows x 32 lanes are staged in shared memory and
// written out as 8 KB fully coalesced.
__global__ void __launch_bounds__(1024)
k_pack(const uint8_t *__restrict__ ascii, GenomeLayout lay, uint4 *__restrict__ planes,
       const SegmentDev *__restrict__ segs, uint32_t n_segs, uint32_t *__restrict__ seg_reset,
       uint8_t *__restrict__ rec_reset) {
    __shared__ uint4 stage[2][16][32];
    const uint64_t t = blockIdx.x;
    const uint32_t i = threadIdx.x & 31u, y = threadIdx.x >> 5;
    const uint64_t stripe = t * kStripesPerTile + (uint64_t)y * 32u + i;
    const uint64_t base = stripe * lay.W;  // position of row 0 of this thread's stripe
    const uint32_t n_chunks = (lay.rows + 15u) / 16u;
    // the whole tile (incl. repeated rows) lies inside the buffer => no per-byte bounds checks
    const bool interior = (t + 1) * kStripesPerTile * (uint64_t)lay.W + 64u <= lay.total_len;
    for (uint32_t c = 0; c < n_chunks; c++) {
        const uint64_t P0 = base + (uint64_t)c * 16u;
        uint32_t v[4] = {0u, 0u, 0u, 0u};
        if (interior || P0 + 16u <= lay.total_len) {
            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(ascii + P0));
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else if (P0 < lay.total_len) {  // ragged tail: byte loads, zero (= hard mismatch) padding
            for (uint32_t k = 0; k < 16u && P0 + k < lay.total_len; k++)
                v[k >> 2] |= (uint32_t)ascii[P0 + k] << (8u * (k & 3u));
        }
        uint4 *out = &stage[c & 1u][0][y];
#pragma unroll
        for (uint32_t k = 0; k < 16u; k++) {
            const uint32_t b = (v[k >> 2] >> (8u * (k & 3u))) & 0xFFu;
            const uint32_t A = __ballot_sync(0xFFFFFFFFu, b == 'A');
            const uint32_t C = __ballot_sync(0xFFFFFFFFu, b == 'C');
            const uint32_t G = __ballot_sync(0xFFFFFFFFu, b == 'G');
            const uint32_t T = __ballot_sync(0xFFFFFFFFu, b == 'T');
            if (i == 0) out[k * 32u] = make_uint4(A, C, G, T);
            if ((A | C | G | T) != 0xFFFFFFFFu) {  // warp-uniform: some byte of this row is not ACGT
                const uint32_t x = (b | 0x20u) - 'a';  // a,c,g,t -> 0,2,6,19
                const bool acgt = x < 32u && ((0x00080045u >> x) & 1u);
                const uint64_t P = P0 + k;
                if (!acgt && P < lay.total_len && c * 16u + k < lay.W) {
                    const uint32_t s = find_segment(segs, n_segs, P);
                    if (P - segs[s].global_off < segs[s].len) {
                        seg_reset[s] = 1u;
                        if (rec_reset) rec_reset[segs[s].rec_idx] = 1u;
                    }
                }
            }
        }
        __syncthreads();  // stage[c&1] complete; the other buffer was drained two syncs ago
        if (threadIdx.x < 512u) {
            const uint32_t row = c * 16u + (threadIdx.x >> 5);
            if (row < lay.rows)
                planes[(t * lay.rows + row) * kLanes + (threadIdx.x & 31u)] = stage[c & 1u][threadIdx.x >> 5][threadIdx.x & 31u];
        }
    }
}

// ---------------------------------------------------------------------------
// K2g: table-driven filter.  One warp per tile; lane l walks rows 0..W-1 of its
// own 32 stripes.  The last 32 rows live in a per-lane ring in shared memory
// (one word per used IUPAC class), so "column j of the primer at start row w" is
// ring[(w + j) & 31] -- a plain load, never a shift.  For every orientation the
// plan (plan.hpp) is evaluated on 32 positions at once:
//     cand = AND(protected columns) & [ #(blocks fully matching) >= need ]
// Replaces the candidate generation of scanACEach + scanNonACGTHalos
// (/root/reference/core/engine/ac.go:151-180, halo.go:76-108).
// ---------------------------------------------------------------------------
struct FilterArgs {
    const uint4 *planes;
    GenomeLayout lay;
    const OrientProg *progs;
    const ColDev *cols;
    const uint8_t *blk_len;
    const uint8_t *slot_cls;  // IUPAC class of each ring slot
    uint32_t n_orient;
    uint32_t n_slots;
    unsigned long long *cand;
    unsigned long long *cand_count;
    uint64_t cand_cap;
    uint64_t tile_begin, tile_end;
};

__device__ __forceinline__ void ring_store(uint32_t *my, uint32_t slot_row, uint4 v, const uint8_t *slot_cls,
                                           uint32_t n_slots) {
    for (uint32_t s = 0; s < n_slots; s++) {
        const uint32_t cl = slot_cls[s];
        uint32_t w = 0;
        if (cl & 1u) w |= v.x;
        if (cl & 2u) w |= v.y;
        if (cl & 4u) w |= v.z;
        if (cl & 8u) w |= v.w;
        my[(slot_row * n_slots + s) * 32u] = w;
    }
}

__device__ __forceinline__ void emit_candidates(const FilterArgs &a, uint32_t word, uint32_t q, uint32_t j0,
                                                uint64_t tile, uint32_t w) {
    const uint64_t stripe0 = tile * kStripesPerTile + (uint64_t)lane_id() * 32u;
    while (word) {
        const int i = __ffs(word) - 1;
        word &= word - 1;
        const uint64_t P = (stripe0 + i) * a.lay.W + w;
        if (P >= a.lay.total_len || P < j0) continue;
        const unsigned long long slot = atomicAdd(a.cand_count, 1ull);
        if (slot < a.cand_cap) a.cand[slot] = ((unsigned long long)q << kCandPosBits) | (P - j0);
    }
}

__global__ void __launch_bounds__(128)
k_filter_generic(FilterArgs a) {
    extern __shared__ uint32_t ring_smem[];
    const uint32_t warps_per_block = blockDim.x >> 5;
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    uint32_t *my = ring_smem + (size_t)warp * 32u * a.n_slots * 32u + lane;
    const uint64_t gw = (uint64_t)blockIdx.x * warps_per_block + warp;
    const uint64_t nw = (uint64_t)gridDim.x * warps_per_block;
    const uint32_t W = a.lay.W, rows = a.lay.rows, ns = a.n_slots;

    for (uint64_t tile = a.tile_begin + gw; tile < a.tile_end; tile += nw) {
        const uint4 *base = a.planes + tile * rows * kLanes + lane;
        for (uint32_t r = 0; r < 31; r++) ring_store(my, r, base[(size_t)r * kLanes], a.slot_cls, ns);
        uint4 nxt = base[(size_t)31 * kLanes];
        for (uint32_t w = 0; w < W; w++) {
            ring_store(my, (w + 31u) & 31u, nxt, a.slot_cls, ns);
            if (w + 1 < W) nxt = base[(size_t)(w + 32u) * kLanes];
            for (uint32_t q = 0; q < a.n_orient; q++) {
                const OrientProg pg = a.progs[q];
                const ColDev *c = a.cols + pg.col_begin;
                uint32_t acc = 0xFFFFFFFFu;
                for (uint32_t k = 0; k < pg.n_prot; k++)
                    acc &= my[((((w + c[k].row) & 31u) * ns) + c[k].slot) * 32u];
                c += pg.n_prot;
                if (pg.need > 0 && acc != 0) {
                    uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;  // bit-sliced count of matching blocks
                    for (uint32_t b = 0; b < pg.n_blocks; b++) {
                        const uint32_t len = a.blk_len[pg.blk_begin + b];
                        uint32_t blk = 0xFFFFFFFFu;
                        for (uint32_t k = 0; k < len; k++)
                            blk &= my[((((w + c[k].row) & 31u) * ns) + c[k].slot) * 32u];
                        c += len;
                        uint32_t carry = c0 & blk;
                        c0 ^= blk;
                        uint32_t t1 = c1 & carry;
                        c1 ^= carry;
                        uint32_t t2 = c2 & t1;
                        c2 ^= t1;
                        c3 ^= t2;
                    }
                    const uint32_t cnt[4] = {c0, c1, c2, c3};
                    uint32_t gt = 0, eq = 0xFFFFFFFFu;
#pragma unroll
                    for (int bit = 3; bit >= 0; bit--) {
                        const uint32_t nb = ((pg.need >> bit) & 1u) ? 0xFFFFFFFFu : 0u;
                        gt |= eq & cnt[bit] & ~nb;
                        eq &= ~(cnt[bit] ^ nb);
                    }
                    acc &= (gt | eq);
                }
                if (acc != 0) emit_candidates(a, acc, q, pg.j0, tile, w);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// K3: verify.  One thread per candidate re-checks the FULL primer on the raw
// bytes with the reference's predicate (BaseMatch, iupac.go:62-67; verifyAt,
// ac.go:186-213; FindMatches, match.go:68-89): mismatch mask, mismatch budget,
// protected 3' window, segment bounds.  Survivors are compacted with a warp
// ballot + prefix popcount and ONE atomic per warp (matchCollector,
// hit_collect.go:73-98).
// ---------------------------------------------------------------------------
struct VerifyArgs {
    const unsigned long long *cand;
    uint64_t n_cand;
    // asynchronous chain: the count is still on the device (n = min(*n_cand_ptr, n_cand)); NULL = n_cand
    const unsigned long long *n_cand_ptr;
    const uint8_t *ascii;
    const SegmentDev *segs;
    uint32_t n_segs;
    const OrientDev *orients;
    const uint8_t *cls;
    ScanParams prm;
    ipcr_site *sites;
    unsigned long long *site_count;
    uint64_t site_cap;
    // packed genome (seed-index path): when set, the window is compared 64 columns at a time against
    // the orientation's four acceptance masks instead of byte by byte on the ASCII
    unsigned long long *publish_seq_ptr;  // peer exchange: where the last block stores publish_seq (or NULL)
    unsigned long long publish_seq;
    unsigned int *blocks_done;            // zeroed counter of finished blocks
    const uint32_t *codes;     // 2-bit codes, 16 bases per word (A 0, C 1, T 2, G 3)
    const uint32_t *inval;     // bit = byte is not an uppercase A/C/G/T
    const uint64_t *oplanes;   // [orientation][4]: columns accepting A, C, G, T
};

__device__ __forceinline__ uint32_t onehot_of(uint32_t b) {
    return b == 'A' ? 1u : b == 'C' ? 2u : b == 'G' ? 4u : b == 'T' ? 8u : 0u;
}

// bits 0, 2, 4, .. of x packed into the low 32 bits
__device__ __forceinline__ uint32_t even_bits(uint64_t x) {
    x &= 0x5555555555555555ull;
    x = (x | (x >> 1)) & 0x3333333333333333ull;
    x = (x | (x >> 2)) & 0x0F0F0F0F0F0F0F0Full;
    x = (x | (x >> 4)) & 0x00FF00FF00FF00FFull;
    x = (x | (x >> 8)) & 0x0000FFFF0000FFFFull;
    x = (x | (x >> 16)) & 0x00000000FFFFFFFFull;
    return (uint32_t)x;
}

// The same predicate on the packed genome: the 64 bases from `start` become two bit planes (low / high
// code bit) and an invalid mask; column j mismatches iff its byte is invalid or the primer's class at j
// does not accept the base.  ~10 word loads on 2-3 cache lines instead of 2 byte loads per column.
__device__ __forceinline__ bool verify_one_codes(const VerifyArgs &a, uint32_t q, uint64_t start, ipcr_site *s) {
    const uint32_t si = find_segment(a.segs, a.n_segs, start);
    const SegmentDev seg = a.segs[si];
    const OrientDev o = a.orients[q];
    const uint64_t local = start - seg.global_off;
    if (o.L == 0) return false;  // match.go:32-34
    if (!(start >= seg.global_off && local + o.L <= seg.len && local < seg.own_len)) return false;
    const uint32_t *cw = a.codes + (start >> 4);
    const uint32_t sh = 2u * ((uint32_t)start & 15u);
    const uint32_t w0 = cw[0], w1 = cw[1], w2 = cw[2], w3 = cw[3], w4 = cw[4];
    const uint64_t c01 = (uint64_t)__funnelshift_r(w0, w1, sh) | ((uint64_t)__funnelshift_r(w1, w2, sh) << 32);
    const uint64_t c23 = (uint64_t)__funnelshift_r(w2, w3, sh) | ((uint64_t)__funnelshift_r(w3, w4, sh) << 32);
    const uint64_t lo = (uint64_t)even_bits(c01) | ((uint64_t)even_bits(c23) << 32);
    const uint64_t hi = (uint64_t)even_bits(c01 >> 1) | ((uint64_t)even_bits(c23 >> 1) << 32);
    const uint32_t *iw = a.inval + (start >> 5);
    const uint32_t ish = (uint32_t)start & 31u;
    const uint32_t i0 = iw[0], i1 = iw[1], i2 = iw[2];
    const uint64_t inv = (uint64_t)__funnelshift_r(i0, i1, ish) | ((uint64_t)__funnelshift_r(i1, i2, ish) << 32);
    const uint64_t *pl = a.oplanes + 4u * q;
    const uint64_t match = (~hi & ~lo & pl[0]) | (~hi & lo & pl[1]) | (hi & lo & pl[2]) | (hi & ~lo & pl[3]);
    const uint64_t lenmask = o.L >= 64 ? ~0ull : ((1ull << o.L) - 1ull);
    const uint64_t mask = (~match | inv) & lenmask;
    const int mm = __popcll(mask);
    if (mm > a.prm.max_mm) return false;
    const bool viol = (mask & o.prot_mask) != 0;
    if (viol && !(a.prm.raw_rc && o.rc_type)) return false;
    uint32_t flags = viol ? IPCR_SITE_TWVIOL : 0u;
    if (o.seeded) {  // would the reference's AC pass have missed it? (ac.go:159-162)
        const uint64_t seedmask = (o.seed_len >= 64 ? ~0ull : ((1ull << o.seed_len) - 1ull)) << o.seed_off;
        uint64_t bad = inv & seedmask & lenmask;
        if (bad) {
            if (a.ascii) {  // lowercase acgt are invalid bases but do not reset the automaton (ac.go:25-28)
                bool reset = false;
                for (; bad; bad &= bad - 1ull) {
                    const uint32_t u = a.ascii[start + (uint32_t)(__ffsll((long long)bad) - 1)] & 0xDFu;
                    if (!(u == 'A' || u == 'C' || u == 'G' || u == 'T')) reset = true;
                }
                if (reset) flags |= IPCR_SITE_HALO;
            } else {
                flags |= IPCR_SITE_HALO;
            }
        }
    }
    s->pos = seg.rec_off + local;
    s->mm_mask = mask;
    s->rec_idx = seg.rec_idx;
    s->orient = (uint16_t)q;
    s->mm = (uint8_t)mm;
    s->flags = (uint8_t)flags;
    return true;
}

// exact check of orientation q at global start `start`; true => *s is a site
__device__ __forceinline__ bool verify_one(const VerifyArgs &a, uint32_t q, uint64_t start, ipcr_site *s) {
    if (a.codes) return verify_one_codes(a, q, start, s);
    const uint32_t si = find_segment(a.segs, a.n_segs, start);
    const SegmentDev seg = a.segs[si];
    const OrientDev o = a.orients[q];
    const uint64_t local = start - seg.global_off;
    if (o.L == 0) return false;  // match.go:32-34
    if (!(start >= seg.global_off && local + o.L <= seg.len && local < seg.own_len)) return false;
    const uint8_t *g = a.ascii + start;
    const uint8_t *pc = a.cls + o.cls_off;
    uint64_t mask = 0;
    int mm = 0;
    for (int j = 0; j < o.L; j++) {
        if (!(onehot_of(g[j]) & pc[j])) {
            mask |= 1ull << j;
            if (++mm > a.prm.max_mm) return false;
        }
    }
    const bool viol = (mask & o.prot_mask) != 0;
    if (viol && !(a.prm.raw_rc && o.rc_type)) return false;
    uint32_t flags = viol ? IPCR_SITE_TWVIOL : 0u;
    if (o.seeded) {  // would the reference's AC pass have missed it? (ac.go:159-162)
        for (int j = o.seed_off; j < o.seed_off + o.seed_len; j++) {
            const uint32_t u = g[j] & 0xDFu;
            if (!(u == 'A' || u == 'C' || u == 'G' || u == 'T')) { flags |= IPCR_SITE_HALO; break; }
        }
    }
    s->pos = seg.rec_off + local;
    s->mm_mask = mask;
    s->rec_idx = seg.rec_idx;
    s->orient = (uint16_t)q;
    s->mm = (uint8_t)mm;
    s->flags = (uint8_t)flags;
    return true;
}

// warp-ballot compaction: one atomic per warp (all 32 lanes must call)
__device__ __forceinline__ void compact_site(const VerifyArgs &a, bool ok, const ipcr_site &s) {
    const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, ok);
    if (!ballot) return;
    unsigned long long base = 0;
    if (lane_id() == 0) base = atomicAdd(a.site_count, (unsigned long long)__popc(ballot));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (ok) {
        const unsigned long long dst = base + __popc(ballot & ((1u << lane_id()) - 1u));
        if (dst < a.site_cap) a.sites[dst] = s;
    }
}

// candidates as a position list (table-driven filter)
__global__ void __launch_bounds__(256)
k_verify(VerifyArgs a) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t n_cand = a.n_cand;
    if (a.n_cand_ptr) n_cand = min((uint64_t)*a.n_cand_ptr, n_cand);
    const uint64_t n_round = (n_cand + 31ull) & ~31ull;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        bool ok = false;
        ipcr_site s;
        if (i < n_cand) {
            const unsigned long long cw = a.cand[i];
            ok = verify_one(a, (uint32_t)(cw >> kCandPosBits), cw & kCandPosMask, &s);
        }
        compact_site(a, ok, s);
    }
    // peer exchange: the LAST block to finish publishes the step number of the (now complete) site block
    // system-wide -- what would otherwise be one more tiny kernel launch at the end of stage 1
    if (a.publish_seq_ptr) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(a.blocks_done, 1u) == gridDim.x - 1u) {
                *a.blocks_done = 0u;  // ready for the next launch (stream-ordered)
                __threadfence_system();
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.publish_seq_ptr), "l"(a.publish_seq) : "memory");
            }
        }
    }
}

// ---- candidates as per-thread lists of 32-position words (specialised filter) ----
// entry = {word, orientation, row, tile}; the owning thread's lane is (thread & 31)
struct WordListArgs {
    const uint4 *lists;       // [n_threads][K]
    const uint32_t *counts;   // [n_threads] entries the thread wanted to write
    uint32_t *clamped;        // [n_threads] min(counts, K)
    const uint32_t *offs;     // [n_threads + 1] exclusive scan of clamped
    uint32_t n_threads;
    uint32_t K;
    unsigned long long *stats;  // [0] max count, [1] total candidate bits
    GenomeLayout lay;
    uint32_t variant;
};

__global__ void k_list_prepare(WordListArgs w) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= w.n_threads) return;
    const uint32_t c = w.counts[t];
    const uint32_t k = c < w.K ? c : w.K;
    w.clamped[t] = k;
    unsigned long long bits = 0;
    for (uint32_t e = 0; e < k; e++) bits += __popc(w.lists[(uint64_t)t * w.K + e].x);
    if (c) atomicMax(&w.stats[0], (unsigned long long)c);
    if (bits) atomicAdd(&w.stats[1], bits);
}

// one THREAD per staged word (most words carry a single candidate bit); the warp
// iterates while any lane still has a bit left so that the ballot compaction stays convergent
__global__ void __launch_bounds__(256)
k_verify_words(VerifyArgs a, WordListArgs w, uint64_t n_entries) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_round = (n_entries + 31ull) & ~31ull;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_round; e += stride) {
        uint32_t word = 0, q = 0, j0 = 0;
        uint64_t stripe0 = 0, row = 0;
        if (e < n_entries) {
            // owning thread: last t with offs[t] <= e
            uint32_t lo = 0, hi = w.n_threads;
            while (lo + 1 < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if ((uint64_t)w.offs[mid] <= e) lo = mid; else hi = mid;
            }
            const uint4 ent = w.lists[(uint64_t)lo * w.K + (e - w.offs[lo])];
            word = ent.x;
            q = ent.y;
            row = ent.z;
            stripe0 = (uint64_t)ent.w * kStripesPerTile + (uint64_t)(lo & 31u) * 32u;
            j0 = a.orients[q].j0[w.variant];
        }
        while (__any_sync(0xFFFFFFFFu, word != 0)) {
            bool ok = false;
            ipcr_site s;
            if (word) {
                const int i = __ffs(word) - 1;
                word &= word - 1;
                const uint64_t P = (stripe0 + i) * w.lay.W + row;
                if (P < w.lay.total_len && P >= j0) ok = verify_one(a, q, P - j0, &s);
            }
            compact_site(a, ok, s);
        }
    }
}

// ---------------------------------------------------------------------------
// K5: hit cap, segmenting, pairing.
// ---------------------------------------------------------------------------
struct KeyLayout {
    uint32_t pos_bits;    // bits for the record-local position
    uint32_t n_orient;    // U
    uint32_t key_bits;    // pos_bits + 1 + prefix bits (invalid flag is bit key_bits)
    uint32_t pad;
};

__device__ __forceinline__ uint64_t make_prefix(const KeyLayout &kl, uint32_t rec, uint32_t q) {
    return (uint64_t)rec * kl.n_orient + q;
}

// sort key: ((rec*U + q) << 1 | halo) << pos_bits | pos ; invalid sites sort last
// rec_reset (optional): per-record flags raised on the device by the pack kernels of the
// asynchronous chain; OR-ed with the host-supplied recs[].has_reset_byte
__device__ __forceinline__ uint64_t site_key(const ipcr_site &s, const OrientDev *__restrict__ orients,
                                             const ipcr_record_info *__restrict__ recs,
                                             const uint8_t *__restrict__ rec_reset, const ScanParams &prm,
                                             const KeyLayout &kl) {
    const OrientDev o = orients[s.orient];
    const bool reset = recs[s.rec_idx].has_reset_byte != 0 || (rec_reset && rec_reset[s.rec_idx] != 0);
    const bool has_reset = reset && prm.max_mm > 0;  // compiled.go:189
    // rc orientations served by primer.FindMatches are capped BEFORE filterLeftTW
    // (compiled.go:250-257): forced fallback, or an orientation the reference cannot seed
    const bool raw_mode = prm.hit_cap > 0 && o.rc_type && (has_reset || !o.seeded);
    bool valid = true;
    if ((s.flags & IPCR_SITE_TWVIOL) && !raw_mode) valid = false;
    // the reference appends halo-found sites after the AC-found ones (compiled.go:211-232)
    const uint64_t halo = ((s.flags & IPCR_SITE_HALO) && !o.rc_type && o.seeded && has_reset && prm.hit_cap == 0) ? 1u : 0u;
    uint64_t key = (((make_prefix(kl, s.rec_idx, s.orient) << 1) | halo) << kl.pos_bits) | s.pos;
    if (!valid) key = 1ull << kl.key_bits;
    return key;
}

__global__ void k_site_keys(const ipcr_site *__restrict__ sites, uint64_t n, const OrientDev *__restrict__ orients,
                            const ipcr_record_info *__restrict__ recs, const uint8_t *__restrict__ rec_reset,
                            ScanParams prm, KeyLayout kl, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = site_key(sites[i], orients, recs, rec_reset, prm, kl);
    vals[i] = (uint32_t)i;
}

__device__ __forceinline__ uint64_t lower_bound_u64(const uint64_t *a, uint64_t n, uint64_t v) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// keep[k] = 1 iff sorted site k survives the hit cap (first hit_cap per
// (record, orientation) in list order) and the left-window filter
__global__ void k_site_keep(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n,
                            const ipcr_site *__restrict__ sites, ScanParams prm, KeyLayout kl,
                            uint32_t *__restrict__ keep) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint64_t key = keys[k];
    uint32_t kp = 0;
    if (key < (1ull << kl.key_bits)) {
        const uint64_t prefix = key >> (kl.pos_bits + 1);
        const uint64_t first = lower_bound_u64(keys, n, prefix << (kl.pos_bits + 1));
        const uint64_t rank = k - first;
        const ipcr_site s = sites[vals[k]];
        kp = 1;
        if (prm.hit_cap > 0 && rank >= (uint64_t)prm.hit_cap) kp = 0;
        if (s.flags & IPCR_SITE_TWVIOL) kp = 0;
    }
    keep[k] = kp;
}

__global__ void k_site_compact(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n,
                               const ipcr_site *__restrict__ sites, const uint32_t *__restrict__ keep,
                               const uint32_t *__restrict__ dst, uint64_t *__restrict__ fkeys,
                               FinalSite *__restrict__ fs) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n || !keep[k]) return;
    const ipcr_site s = sites[vals[k]];
    const uint32_t d = dst[k];
    fkeys[d] = keys[k];
    FinalSite f;
    f.pos = s.pos;
    f.mask = s.mm_mask;
    f.mm = s.mm;
    f.pad = 0;
    fs[d] = f;
}

struct JoinArgs {
    const uint64_t *fkeys;
    const FinalSite *fs;
    uint64_t n_sites;
    const PairDev *pairs;
    uint32_t n_pairs;
    uint32_t n_recs;
    const ipcr_record_info *recs;
    ScanParams prm;
    KeyLayout kl;
    uint32_t *entry_tasks;      // [n_entries] number of forward sites per (rec, pair, block)
    const uint32_t *task_off;   // [n_entries + 1]
    uint64_t n_tasks;
    uint32_t *task_cnt;         // [n_tasks]
    const uint32_t *prod_off;   // [n_tasks + 1]
    ipcr_product *products;
    uint64_t prod_cap;
};

__device__ __forceinline__ void segment_of(const JoinArgs &a, uint32_t rec, uint32_t q, uint64_t &b, uint64_t &e) {
    const uint64_t prefix = make_prefix(a.kl, rec, q);
    b = lower_bound_u64(a.fkeys, a.n_sites, prefix << (a.kl.pos_bits + 1));
    e = lower_bound_u64(a.fkeys, a.n_sites, (prefix + 1) << (a.kl.pos_bits + 1));
}

// entry e = (rec * n_pairs + pair) * 2 + block ; block 0 = A x rc(B) "forward", 1 = B x rc(A) "revcomp"
__global__ void k_join_entries(JoinArgs a) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t n_entries = (uint64_t)a.n_recs * a.n_pairs * 2ull;
    if (e >= n_entries) return;
    const uint32_t blk = (uint32_t)(e & 1ull);
    const uint32_t pair = (uint32_t)((e >> 1) % a.n_pairs);
    const uint32_t rec = (uint32_t)((e >> 1) / a.n_pairs);
    const PairDev p = a.pairs[pair];
    uint64_t b, en;
    segment_of(a, rec, blk == 0 ? p.fwdA : p.fwdB, b, en);
    a.entry_tasks[e] = (uint32_t)(en - b);
}

// Partner lists are rc-type orientations: never halo-flagged, so inside their segment the sort
// keys ascend with the position held in their low pos_bits -- the searches read the 8-byte keys
// (shared memory in the single-CTA pairing kernel), not the 24-byte site records.
// first index in [b, e) with pos >= v  (lowerBoundMatchPos, engine.go:87-89)
__device__ __forceinline__ int64_t lb_pos(const uint64_t *fkeys, uint64_t pos_mask, int64_t b, int64_t e, int64_t v) {
    while (b < e) {
        const int64_t mid = b + ((e - b) >> 1);
        if ((int64_t)(fkeys[mid] & pos_mask) >= v) e = mid; else b = mid + 1;
    }
    return b;
}
// first index in [b, e) with pos > v  (upperBoundMatchPos, engine.go:91-93)
__device__ __forceinline__ int64_t ub_pos(const uint64_t *fkeys, uint64_t pos_mask, int64_t b, int64_t e, int64_t v) {
    while (b < e) {
        const int64_t mid = b + ((e - b) >> 1);
        if ((int64_t)(fkeys[mid] & pos_mask) > v) e = mid; else b = mid + 1;
    }
    return b;
}

__device__ __forceinline__ uint64_t flip_mask(uint64_t m, int n) {  // engine.go:127-136
    return n <= 0 ? 0ull : (__brevll(m) >> (64 - n));
}

// One forward site of one (record, pair, direction) against its partner list [pb, pe): a verbatim
// restatement of one iteration of the outer loops of forEachJoinedProduct
// (core/engine/engine.go:144-272 / :276-402).  FILL = false counts, true writes from products[out].
// fidx = index of the forward site in the final (sorted, capped) list.  Returns the product count.
template <bool FILL>
__device__ __forceinline__ uint32_t join_core(const JoinArgs &a, uint32_t rec, uint32_t pair, uint32_t blk,
                                              const PairDev &p, uint64_t fidx, int64_t pb, int64_t pe, uint64_t out) {
    const int64_t seq_len = (int64_t)a.recs[rec].len;
    const int64_t minL = p.min_len, maxL = p.max_len;
    const int64_t plen = blk == 0 ? p.blen : p.alen;  // partner primer length
    const uint64_t pos_mask = (1ull << a.kl.pos_bits) - 1ull;
    const int64_t apos = (int64_t)(a.fkeys[fidx] & pos_mask);
    uint32_t cnt = 0;
    FinalSite ma;
    bool have_ma = false;

    auto emit = [&](int64_t j, int64_t end, int64_t length) {
        if (FILL) {
            if (out < a.prod_cap) {
                if (!have_ma) {
                    ma = a.fs[fidx];
                    have_ma = true;
                }
                const FinalSite mb = a.fs[j];
                ipcr_product pr;
                pr.start = apos;
                pr.end = end;
                pr.length = length;
                pr.fwd_mm_mask = ma.mask;
                pr.rev_mm_mask = flip_mask(mb.mask, (int)plen);
                pr.pair_idx = pair;
                pr.rec_idx = rec;
                pr.fwd_mm = (uint16_t)ma.mm;
                pr.rev_mm = (uint16_t)mb.mm;
                pr.type = (uint8_t)blk;
                for (int z = 0; z < 11; z++) pr.reserved[z] = 0;
                a.products[out] = pr;
            }
            out++;
        }
        cnt++;
    };

    {  // linear partners, right-most first (engine.go:145-206)
        const int64_t last = seq_len - plen;
        int64_t lo2 = apos + 1;
        if (minL > 0) {
            lo2 = apos + minL - plen;
            if (lo2 <= apos) lo2 = apos + 1;
        }
        if (lo2 < 0) lo2 = 0;
        int64_t hi2 = last;
        if (maxL > 0) {
            hi2 = apos + maxL - plen;
            if (hi2 > last) hi2 = last;
        }
        if (hi2 >= lo2) {
            const int64_t iMin = lb_pos(a.fkeys, pos_mask, pb, pe, lo2);
            const int64_t iMax = ub_pos(a.fkeys, pos_mask, pb, pe, hi2) - 1;
            for (int64_t j = iMax; j >= iMin; j--) {
                const int64_t end = (int64_t)(a.fkeys[j] & pos_mask) + plen;
                const int64_t length = end - apos;
                if ((minL != 0 && length < minL) || (maxL != 0 && length > maxL)) continue;
                emit(j, end, length);
            }
        }
    }
    if (a.prm.circular) {  // wrap-around partners (engine.go:209-271)
        const int64_t X = seq_len - apos;
        int64_t loW = 0;
        if (minL > 0) {
            int64_t needed = minL - X - plen;
            if (needed < 0) needed = 0;
            loW = needed;
        }
        int64_t hiW = apos - 1;
        if (maxL > 0) {
            const int64_t allowed = maxL - X - plen;
            if (allowed < hiW) hiW = allowed;
        }
        if (hiW >= loW) {
            const int64_t iMinW = lb_pos(a.fkeys, pos_mask, pb, pe, loW);
            const int64_t iMaxW = ub_pos(a.fkeys, pos_mask, pb, pe, hiW) - 1;
            for (int64_t j = iMaxW; j >= iMinW; j--) {
                const int64_t mpos = (int64_t)(a.fkeys[j] & pos_mask);
                if (mpos >= apos) continue;
                const int64_t end = mpos + plen;
                const int64_t length = X + end;
                if ((minL != 0 && length < minL) || (maxL != 0 && length > maxL)) continue;
                emit(j, end, length);
            }
        }
    }
    return cnt;
}

// task t -> (entry, forward site) through the scanned entry sizes
template <bool FILL>
__device__ __forceinline__ void join_task(const JoinArgs &a, uint64_t t) {
    const uint64_t n_entries = (uint64_t)a.n_recs * a.n_pairs * 2ull;
    // entry = last e with task_off[e] <= t
    uint64_t lo = 0, hi = n_entries;
    while (lo + 1 < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if ((uint64_t)a.task_off[mid] <= t) lo = mid; else hi = mid;
    }
    const uint64_t e = lo;
    const uint32_t k = (uint32_t)(t - a.task_off[e]);
    const uint32_t blk = (uint32_t)(e & 1ull);
    const uint32_t pair = (uint32_t)((e >> 1) % a.n_pairs);
    const uint32_t rec = (uint32_t)((e >> 1) / a.n_pairs);
    const PairDev p = a.pairs[pair];
    uint64_t fb, fe, pb64, pe64;
    segment_of(a, rec, blk == 0 ? p.fwdA : p.fwdB, fb, fe);
    segment_of(a, rec, blk == 0 ? p.rcB : p.rcA, pb64, pe64);
    const uint32_t cnt = join_core<FILL>(a, rec, pair, blk, p, fb + k, (int64_t)pb64, (int64_t)pe64,
                                         FILL ? a.prod_off[t] : 0);
    if (!FILL) a.task_cnt[t] = cnt;
}

// 64-bit total of a u32 array (guards the 32-bit product scan against wrap-around)
__global__ void __launch_bounds__(256)
k_sum_u32(const uint32_t *__restrict__ v, uint64_t n, unsigned long long *__restrict__ out) {
    unsigned long long acc = 0ull;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) acc += v[i];
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xFFFFFFFFu, acc, d);
    if ((threadIdx.x & 31u) == 0u && acc) atomicAdd(out, acc);
}

template <bool FILL>
__global__ void __launch_bounds__(128)
k_join(JoinArgs a) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < a.n_tasks) join_task<FILL>(a, t);
}

// ---------------------------------------------------------------------------
// K6: visitors on the device.
// ---------------------------------------------------------------------------
// amplicon byte i of a product: record bytes [start, end) or, for a circular wrap
// (start > end), seq[start:] + seq[:end] (internal/pipeline/pipeline.go:81-88)
struct AmpliconArgs {
    const uint8_t *ascii;
    const uint64_t *rec_base;  // concatenated coordinate of position 0 of each record
    const uint64_t *rec_len;
    const ipcr_product *products;
    uint64_t n_products;
};

__device__ __forceinline__ uint8_t amplicon_byte(const AmpliconArgs &a, const ipcr_product &p, uint64_t i) {
    const uint64_t rl = a.rec_len[p.rec_idx];
    uint64_t pos = (uint64_t)p.start + i;
    if (pos >= rl) pos -= rl;
    return a.ascii[a.rec_base[p.rec_idx] + pos];
}

__device__ __forceinline__ uint32_t to_upper(uint32_t c) { return (c - 'a') < 26u ? c - 32u : c; }

// oligo.BestHit (/root/reference/core/oligo/oligo.go:19-77): best probe hit on either strand
// of every amplicon.  One CTA per product; threads stride over start positions and count
// IUPAC mismatches with early exit; a block-wide min over (mm, pos) picks the best per strand.
struct ProbeArgs {
    AmpliconArgs amp;
    const uint8_t *cls_plus;   // IUPAC class per probe column
    const uint8_t *cls_minus;  // of the reverse complement
    uint32_t L;
    int32_t max_mm;
    uint32_t exact_strict;     // maxMM == 0 and strict-ACGT probe: '+' anywhere beats '-' (oligo.go:33-42)
    ipcr_probe_hit *out;
};

__global__ void __launch_bounds__(128)
k_probe_annotate(ProbeArgs a) {
    __shared__ unsigned long long best[2];
    const uint64_t pi = blockIdx.x;
    if (pi >= a.amp.n_products) return;
    const ipcr_product p = a.amp.products[pi];
    if (threadIdx.x < 2) best[threadIdx.x] = ~0ull;
    __syncthreads();
    const uint64_t alen = (uint64_t)p.length;
    if (alen >= a.L && a.L > 0) {
        unsigned long long mine[2] = {~0ull, ~0ull};
        for (uint64_t s0 = threadIdx.x; s0 + a.L <= alen; s0 += blockDim.x) {
            int mp = 0, mn = 0;
            for (uint32_t j = 0; j < a.L; j++) {
                const uint32_t one = onehot_of(to_upper(amplicon_byte(a.amp, p, s0 + j)));
                mp += !(one & a.cls_plus[j]);
                mn += !(one & a.cls_minus[j]);
                if (mp > a.max_mm && mn > a.max_mm) break;
            }
            if (mp <= a.max_mm) mine[0] = min(mine[0], ((unsigned long long)mp << 48) | s0);
            if (mn <= a.max_mm) mine[1] = min(mine[1], ((unsigned long long)mn << 48) | s0);
        }
        if (mine[0] != ~0ull) atomicMin(&best[0], mine[0]);
        if (mine[1] != ~0ull) atomicMin(&best[1], mine[1]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ipcr_probe_hit h;
        h.pos = 0; h.mm = 0; h.found = 0; h.strand = 0;
        for (int z = 0; z < 4; z++) h.reserved[z] = 0;
        const unsigned long long bp = best[0], bm = best[1];
        int pick = -1;
        if (a.exact_strict) pick = bp != ~0ull ? 0 : (bm != ~0ull ? 1 : -1);
        else if (bp != ~0ull && (bm == ~0ull || bp <= bm)) pick = 0;  // equal (mm, pos): '+' wins (oligo.go:54)
        else if (bm != ~0ull) pick = 1;
        if (pick >= 0) {
            const unsigned long long v = pick == 0 ? bp : bm;
            h.found = 1;
            h.strand = pick == 0 ? '+' : '-';
            h.mm = (uint16_t)(v >> 48);
            h.pos = (int64_t)(v & ((1ull << 48) - 1));
        }
        a.out[pi] = h;
    }
}

// amplicon sequences of a product list, back to back (offs[i] = start of amplicon i)
__global__ void k_gather_amplicons(AmpliconArgs a, const uint64_t *__restrict__ offs, uint8_t *__restrict__ out) {
    const uint64_t pi = blockIdx.x;
    if (pi >= a.n_products) return;
    const ipcr_product p = a.products[pi];
    uint8_t *dst = out + offs[pi];
    for (uint64_t i = threadIdx.x; i < (uint64_t)p.length; i += blockDim.x) dst[i] = amplicon_byte(a, p, i);
}

// ---------------------------------------------------------------------------
// synthetic input: uniform ACGT from splitmix64(seed, absolute position / 32)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void k_synth_fill(uint8_t *dst, uint64_t n, uint64_t first_pos, uint64_t seed) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t p = first_pos + i;
        const uint64_t h = splitmix64(seed ^ ((p >> 5) * 0xD6E8FEB86659FD93ull));
        const uint32_t two = (uint32_t)(h >> ((p & 31ull) * 2ull)) & 3u;
        dst[i] = (uint8_t)("ACGT"[two]);
    }
}

}  // namespace ipcr
