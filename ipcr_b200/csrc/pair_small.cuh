// pair_small.cuh -- the asynchronous chain: stage-1 site blocks and stage 2 (hit cap, sort,
// amplicon pairing) in ONE single-CTA launch.
//
// A scan usually leaves a few thousand primer sites.  The multi-kernel stage 2 (k_site_keys ->
// sort -> keep -> scan -> compact -> entries -> scan -> join<count> -> scan -> join<fill>) costs
// nine launches and three host round trips for ~20 us of work.  Here stage 1 writes its sites,
// counters and per-record reset flags into a SITE BLOCK in device memory, and one CTA of 1024
// threads does all of stage 2 out of shared memory over any number of such blocks (one per GPU
// shard or per streamed slab): it reads every count from the DEVICE, so the host never waits
// between kernels, and leaves a header the host reads once, together with the products.
// Same arithmetic and same emission order as the multi-kernel path: it calls the same device
// functions (site_key, join_task).  Replaces matchCollector's cap, sortMatchesByPos, filterLeftTW
// and forEachJoinedProduct (/root/reference/core/engine/hit_collect.go:73-98, engine.go:53-76,
// engine.go:108-404).
#pragma once

#include "kernels.cuh"
#include "sortscan.cuh"

namespace ipcr {

constexpr uint32_t kPairSmallMax = kSortSmallMax;  // 8192 sites
constexpr uint32_t kPairMaxBlocks = 64;

// ---- site block: what stage 1 leaves behind and what multi-GPU shards exchange -------------
// block = [Stage1Ctl | rec_reset bytes, padded to 16 | sites]; all blocks of a call have one size.
struct Stage1Ctl {
    unsigned long long n_sites;     // sites this shard found (may exceed site_cap: then some are lost)
    unsigned long long n_cand, n_dirty;
    unsigned long long cand_cap, dirty_cap, site_cap;
    uint32_t variant;               // plan variant stage 1 ran with
    uint32_t n_recs;
    unsigned long long seq;         // peer exchange: step number published once the block is complete
};
static_assert(sizeof(Stage1Ctl) == 64, "Stage1Ctl is one 64-byte block");

__host__ __device__ __forceinline__ uint64_t block_ctl_bytes(uint64_t n_recs) {
    return sizeof(Stage1Ctl) + ((n_recs + 15ull) & ~15ull);
}
__host__ __device__ __forceinline__ uint64_t block_bytes_for(uint64_t n_recs, uint64_t site_cap) {
    return (block_ctl_bytes(n_recs) + site_cap * sizeof(ipcr_site) + 15ull) & ~15ull;
}

// first kernel of the chain: clears the counters and flags, records the capacities
__global__ void __launch_bounds__(256)
k_block_init(uint8_t *block, Stage1Ctl init, uint32_t *seg_reset, uint32_t n_segs, const uint8_t *rr_init) {
    if (threadIdx.x == 0) {
        Stage1Ctl *ctl = reinterpret_cast<Stage1Ctl *>(block);
        init.seq = ctl->seq;  // the peer exchange's sequence number outlives the re-initialisation
        *ctl = init;
    }
    uint8_t *rr = block + sizeof(Stage1Ctl);
    const uint32_t rr_bytes = (uint32_t)(block_ctl_bytes(init.n_recs) - sizeof(Stage1Ctl));
    // rr_init: per-record flags the host already knows (packed input: the feeder saw every byte)
    for (uint32_t i = threadIdx.x; i < rr_bytes; i += blockDim.x) rr[i] = (rr_init && i < init.n_recs) ? rr_init[i] : 0u;
    for (uint32_t i = threadIdx.x; i < n_segs; i += blockDim.x) seg_reset[i] = 0u;
}

// last kernel of stage 1 in the peer exchange: every write of the chain is complete (stream order); make the
// block visible system-wide and publish the step number the peers' pairing kernels are spinning on
__global__ void k_block_publish(uint8_t *block, unsigned long long seq) {
    __threadfence_system();
    unsigned long long *p = &reinterpret_cast<Stage1Ctl *>(block)->seq;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(seq) : "memory");
}

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// flags of PairHeader.flags
constexpr uint32_t kPairNotSmall = 1u;        // more than kPairSmallMax sites: run the multi-kernel path
constexpr uint32_t kPairSitesLost = 2u;       // a block's stage 1 produced more sites than the block holds
constexpr uint32_t kPairTaskOverflow = 4u;    // task_cap too small (n_tasks is valid)
constexpr uint32_t kPairProdOverflow = 8u;    // prod_cap too small (n_prod is valid)
constexpr uint32_t kPairStage1Overflow = 16u; // a block's candidate or dirty-window list overflowed
constexpr uint32_t kPairAnyReset = 32u;       // some record holds a reset byte
constexpr uint32_t kPairMixedVariant = 64u;   // blocks ran different plan variants (informational)
constexpr uint32_t kPairPeerTimeout = 128u;   // a peer never published its block (peer exchange)

struct PairHeader {
    unsigned long long n_sites, n_final, n_tasks, n_prod;
    unsigned long long n_cand, n_dirty;  // largest per-block counts (capacity hints)
    uint32_t flags;
    uint32_t min_variant;                // smallest plan variant among the blocks
    uint32_t pad[2];
};
static_assert(sizeof(PairHeader) == 64, "PairHeader is read back as one 64-byte block");
static_assert(sizeof(PairHeader) == sizeof(ipcr_product), "products follow the header in one buffer");

struct PairSmallArgs {
    // blocks back to back in local memory (blocks + b * block_bytes), or -- peer exchange -- one pointer per rank
    // into that rank's own device memory (mapped through CUDA IPC / peer access, read over NVLink)
    const uint8_t *blocks;
    const uint8_t *const *peer_blocks;  // device array of n_blocks pointers, or NULL
    unsigned long long wait_seq;        // peer exchange: spin until every block's seq == wait_seq (0 = do not wait)
    uint8_t *stage;                     // peer exchange: local copy of the published blocks (n_blocks * block_bytes)
    uint32_t n_blocks;
    uint64_t block_bytes, ctl_bytes, block_site_cap;
    const OrientDev *orients;
    uint8_t *rec_reset;  // [n_recs] scratch: OR of the blocks' per-record flags
    JoinArgs j;  // pairs, recs, prm, kl; fs / entry_tasks / task_off / task_cnt / prod_off / products are scratch
    const uint32_t *use_off;   // [n_orient + 1]: (pair, direction) slots in which an orientation is the forward primer
    const uint32_t *use_list;  // pair << 1 | direction
    uint32_t task_cap;
    uint32_t smem_words;  // dynamic shared memory of this launch, in 32-bit words
    unsigned long long *dbg;  // optional: 32 clock64 stamps of thread 0 (IPCR_CHAIN_DEBUG)
    uint32_t bucket_cap;  // (record, orientation) buckets the dynamic shared memory has room for (0 = general mode only)
    PairHeader *header;
};

// exclusive scan of in[0, n) -> out[0, n], out[n] = total, by one CTA of 1024 threads; returns the total
// to every thread.  in and out may not alias.  wsum: 33 words of shared memory.
__device__ __forceinline__ uint32_t cta_scan_global(const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *wsum) {
    const uint32_t per = (n + blockDim.x - 1) / blockDim.x;
    const uint32_t b = min(n, threadIdx.x * per), e = min(n, b + per);
    uint32_t tsum = 0;
    for (uint32_t i = b; i < e; i++) tsum += in[i];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t x = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
        if (lane >= (uint32_t)d) x += y;
    }
    __syncthreads();  // wsum may still be read by a previous call
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = lane < (blockDim.x >> 5) ? wsum[lane] : 0u;
        uint32_t xi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, xi, d);
            if (lane >= (uint32_t)d) xi += y;
        }
        wsum[lane] = xi - w;
        if (lane == 31) wsum[32] = xi;
    }
    __syncthreads();
    uint32_t run = x - tsum + wsum[warp];
    for (uint32_t i = b; i < e; i++) {
        const uint32_t v = in[i];
        out[i] = run;
        run += v;
    }
    const uint32_t total = wsum[32];
    if (threadIdx.x == 0) out[n] = total;
    __syncthreads();
    return total;
}

// block-wide exclusive scan of one value per thread (1024 threads); returns the exclusive prefix, *total to all
__device__ __forceinline__ uint32_t cta_scan_value(uint32_t v, uint32_t *wsum, uint32_t *total) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
        if (lane >= (uint32_t)d) x += y;
    }
    __syncthreads();
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = wsum[lane];
        uint32_t xi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, xi, d);
            if (lane >= (uint32_t)d) xi += y;
        }
        wsum[lane] = xi - w;
        if (lane == 31) wsum[32] = xi;
    }
    __syncthreads();
    *total = wsum[32];
    return x - v + wsum[warp];
}

// in-place exclusive scan of arr[0, n) in shared memory, arr[n] = total; returns the total
__device__ __forceinline__ uint32_t cta_scan_inplace(uint32_t *arr, uint32_t n, uint32_t *wsum) {
    const uint32_t per = (n + blockDim.x - 1) / blockDim.x;
    const uint32_t b = min(n, threadIdx.x * per), e = min(n, b + per);
    uint32_t tsum = 0;
    for (uint32_t i = b; i < e; i++) tsum += arr[i];
    uint32_t total;
    uint32_t run = cta_scan_value(tsum, wsum, &total);
    for (uint32_t i = b; i < e; i++) {
        const uint32_t v = arr[i];
        arr[i] = run;
        run += v;
    }
    if (threadIdx.x == 0) arr[n] = total;
    __syncthreads();
    return total;
}

constexpr uint32_t kPairMaxProductive = 1024;  // tasks with products the bucket mode ranks (one per thread)
constexpr uint32_t kBucketMaxSize = 64;  // larger (record, orientation) lists go through the general mode
constexpr int kPairPerThread = kPairSmallMax / 1024;

// BUCKET MODE.  Sites fall into (record, orientation) buckets that are tiny in a sparse scan, so the
// sort is a counting sort over the buckets plus a rank inside each bucket, the scanned bucket sizes ARE
// the segment table of the join (no binary searches), and every thread walks a contiguous range of
// (record, pair, direction) entries -- counting, then writing products in the reference's order with
// no per-entry or per-task arrays.  Shared memory: sk[8192] words, off[nb + 1], foff[nb + 1].
// Returns false (nothing written) when a bucket holds more than kBucketMaxSize sites.
__device__ __forceinline__ bool pair_bucket_mode(const PairSmallArgs &a, JoinArgs j, PairHeader &h, uint32_t n,
                                                 unsigned long long *sk, const unsigned long long *s_boff,
                                                 const uint8_t *const *s_blk, uint32_t *wsum) {
    const KeyLayout kl = j.kl;
    const uint32_t nb = j.n_recs * kl.n_orient;
    uint32_t *off = reinterpret_cast<uint32_t *>(sk + kPairSmallMax);
    uint32_t *foff = off + nb + 1;
    __shared__ uint32_t s_max, s_tasks;
    const long long t0 = clock64();
    int dbg_n = 0;
    auto mark = [&]() {  // IPCR_CHAIN_DEBUG: cycle count of thread 0 at the end of every phase, into a.dbg
        if (a.dbg && threadIdx.x == 0 && dbg_n < 31) a.dbg[1 + dbg_n] = (unsigned long long)(clock64() - t0);
        dbg_n++;
    };
    for (uint32_t i = threadIdx.x; i <= nb; i += blockDim.x) off[i] = foff[i] = 0u;
    if (threadIdx.x == 0) s_max = s_tasks = 0u;
    __syncthreads();
    mark();  // 1 zeroed
    auto site_at = [&](uint32_t i) -> const ipcr_site & {
        uint32_t b = 0;
        while (b + 1 < a.n_blocks && (unsigned long long)i >= s_boff[b + 1]) b++;
        return reinterpret_cast<const ipcr_site *>(s_blk[b] + a.ctl_bytes)[i - (uint32_t)s_boff[b]];
    };
    // ---- 1. keys; slot inside the bucket (arrival order, fixed by the rank pass below) ----
    unsigned long long word[kPairPerThread];
    uint32_t slot[kPairPerThread];
#pragma unroll
    for (int c = 0; c < kPairPerThread; c++) {
        const uint32_t i = threadIdx.x + 1024u * c;
        word[c] = ~0ull;
        slot[c] = 0u;
        if (i < n) {
            const ipcr_site &s = site_at(i);
            const uint64_t key = site_key(s, a.orients, j.recs, a.rec_reset, j.prm, kl);
            if (key < (1ull << kl.key_bits)) {  // invalid sites never count
                word[c] = (key << 14) | ((unsigned long long)i << 1) | ((s.flags & IPCR_SITE_TWVIOL) ? 1ull : 0ull);
                slot[c] = atomicAdd(&off[(uint32_t)(key >> (kl.pos_bits + 1))], 1u);
            }
        }
    }
    __syncthreads();
    mark();  // 2 keys
    {  // largest bucket
        uint32_t mx = 0;
        for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) mx = max(mx, off[i]);
        mx = __reduce_max_sync(0xFFFFFFFFu, mx);
        if ((threadIdx.x & 31u) == 0u) atomicMax(&s_max, mx);
    }
    __syncthreads();
    if (s_max > kBucketMaxSize) return false;
    mark();  // 3 max
    const uint32_t n_valid = cta_scan_inplace(off, nb, wsum);
    mark();  // 4 scan 1
    // ---- 2. scatter into bucket order, then rank inside each bucket ----
#pragma unroll
    for (int c = 0; c < kPairPerThread; c++)
        if (word[c] != ~0ull) sk[off[(uint32_t)(word[c] >> (14 + kl.pos_bits + 1))] + slot[c]] = word[c];
    __syncthreads();
    uint32_t bk[kPairPerThread], rank[kPairPerThread];
#pragma unroll
    for (int c = 0; c < kPairPerThread; c++) {
        const uint32_t k = threadIdx.x + 1024u * c;
        word[c] = ~0ull;
        if (k < n_valid) {
            const unsigned long long w = sk[k];
            const uint32_t b = (uint32_t)(w >> (14 + kl.pos_bits + 1));
            uint32_t r = 0;
            for (uint32_t x = off[b]; x < off[b + 1]; x++) r += sk[x] < w ? 1u : 0u;
            word[c] = w;
            bk[c] = b;
            rank[c] = r;
        }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < kPairPerThread; c++)
        if (word[c] != ~0ull) sk[off[bk[c]] + rank[c]] = word[c];
    __syncthreads();
    mark();  // 5 sorted
    // ---- 3. hit cap (rank in the sorted bucket) + left-window filter; final position of the kept ones ----
    uint32_t prior[kPairPerThread];
    const uint64_t cap = j.prm.hit_cap > 0 ? (uint64_t)j.prm.hit_cap : ~0ull;
#pragma unroll
    for (int c = 0; c < kPairPerThread; c++) {
        const uint32_t k = threadIdx.x + 1024u * c;
        word[c] = ~0ull;
        if (k < n_valid) {
            const unsigned long long w = sk[k];
            const uint32_t b = (uint32_t)(w >> (14 + kl.pos_bits + 1));
            const uint32_t b0 = off[b];
            if ((w & 1ull) == 0ull && (uint64_t)(k - b0) < cap) {
                uint32_t pr = 0;
                for (uint32_t x = b0; x < k; x++) pr += (sk[x] & 1ull) ? 0u : 1u;  // x - b0 < k - b0 < cap
                word[c] = w;
                bk[c] = b;
                prior[c] = pr;
                atomicAdd(&foff[b], 1u);
            }
        }
    }
    __syncthreads();
    mark();  // 6 keep
    const uint32_t n_final = cta_scan_inplace(foff, nb, wsum);
    mark();  // 7 scan 2
    h.n_final = n_final;
    j.n_sites = n_final;
    if (n_final == 0u) return true;
    {
        FinalSite *fs = const_cast<FinalSite *>(j.fs);
#pragma unroll
        for (int c = 0; c < kPairPerThread; c++) {
            if (word[c] == ~0ull) continue;
            const uint32_t dst = foff[bk[c]] + prior[c];
            const ipcr_site s = site_at((uint32_t)(word[c] >> 1) & (kPairSmallMax - 1u));
            sk[dst] = word[c] >> 14;  // every word was read into registers before the scan's barriers
            FinalSite f;
            f.pos = s.pos;
            f.mask = s.mm_mask;
            f.mm = s.mm;
            f.pad = 0;
            fs[dst] = f;
        }
    }
    __syncthreads();
    j.fkeys = reinterpret_cast<const uint64_t *>(sk);
    mark();  // 8 final written
    // ---- 4. join, driven by the sites instead of the (record, pair, direction) entries: a final site of
    // orientation q is a forward site of every (pair, direction) that uses q (use_off / use_list, built
    // with the panel).  One thread per site counts its products per use; the few PRODUCTIVE tasks are
    // ranked by (entry, forward site) = the reference's emission order, their counts scanned in that
    // order, and one thread per productive task writes its products.  No per-entry arrays, no walk over
    // the (mostly empty) entries. ----
    __shared__ uint32_t s_npt;
    if (threadIdx.x == 0) s_npt = 0u;
    // The panel tables the join walks (uses per orientation, pairs, record lengths) move into what is left of the
    // dynamic shared memory when they fit: every task otherwise pays four dependent global loads before it has
    // looked at a single site.
    const uint32_t *use_off = a.use_off, *use_list = a.use_list;
    {
        const uint32_t n_uses = 2u * j.n_pairs;
        uint32_t *base = foff + nb + 1;
        base += ((uintptr_t)base & 15u) ? (16u - ((uintptr_t)base & 15u)) / 4u : 0u;  // 16-byte aligned (PairDev copies)
        const uint64_t need = (uint64_t)(kl.n_orient + 1) + n_uses + 8ull * j.n_pairs + 4ull * j.n_recs + 8;
        const uint64_t have = a.smem_words > (uint64_t)(base - reinterpret_cast<uint32_t *>(sk)) ?
                                  a.smem_words - (uint64_t)(base - reinterpret_cast<uint32_t *>(sk)) : 0;
        if (need <= have) {
            PairDev *s_pairs = reinterpret_cast<PairDev *>(base);
            ipcr_record_info *s_recs = reinterpret_cast<ipcr_record_info *>(base + 8u * j.n_pairs);
            uint32_t *s_uoff = base + 8u * j.n_pairs + 4u * j.n_recs, *s_ulist = s_uoff + kl.n_orient + 1;
            for (uint32_t i = threadIdx.x; i < 8u * j.n_pairs; i += blockDim.x) base[i] = reinterpret_cast<const uint32_t *>(j.pairs)[i];
            for (uint32_t i = threadIdx.x; i < 4u * j.n_recs; i += blockDim.x)
                reinterpret_cast<uint32_t *>(s_recs)[i] = reinterpret_cast<const uint32_t *>(j.recs)[i];
            for (uint32_t i = threadIdx.x; i <= kl.n_orient; i += blockDim.x) s_uoff[i] = a.use_off[i];
            for (uint32_t i = threadIdx.x; i < n_uses; i += blockDim.x) s_ulist[i] = a.use_list[i];
            j.pairs = s_pairs;
            j.recs = s_recs;
            use_off = s_uoff;
            use_list = s_ulist;
        }
    }
    __syncthreads();
    mark();  // 9 tables staged
    unsigned long long *pt_key = reinterpret_cast<unsigned long long *>(j.entry_tasks);  // [kPairMaxProductive] scratch
    uint32_t *pt_cnt = j.task_cnt, *pt_off = const_cast<uint32_t *>(j.prod_off);
    // pass A: (site, use) pairs whose partner list is not empty -- a few hundred of several thousand -- are LISTED;
    // pass B then runs the partner search once per listed task with every lane busy (doing the search inside
    // pass A made each warp execute it in nearly every iteration for one or two of its lanes)
    __shared__ uint32_t s_nlive;
    if (threadIdx.x == 0) s_nlive = 0u;
    __syncthreads();
    uint32_t *live = const_cast<uint32_t *>(j.task_off);  // scratch, task_cap entries: site << 19 | use
    uint32_t my_tasks = 0;
    for (uint32_t f = threadIdx.x; f < n_final; f += blockDim.x) {
        const uint32_t prefix = (uint32_t)(sk[f] >> (kl.pos_bits + 1));
        const uint32_t rec = prefix / kl.n_orient, q = prefix - rec * kl.n_orient;
        for (uint32_t u = use_off[q]; u < use_off[q + 1]; u++) {
            const uint32_t use = use_list[u], pair = use >> 1, blk = use & 1u;
            const uint32_t bp = rec * kl.n_orient + (blk == 0 ? j.pairs[pair].rcB : j.pairs[pair].rcA);
            my_tasks++;
            if (foff[bp] == foff[bp + 1]) continue;
            const uint32_t slot = atomicAdd(&s_nlive, 1u);
            if (slot < a.task_cap) live[slot] = (f << 19) | use;
        }
    }
    __syncthreads();
    const uint32_t n_live = s_nlive;
    if (n_live > a.task_cap) {
        h.n_tasks = n_live;
        h.flags |= kPairTaskOverflow;
        return true;
    }
    for (uint32_t t = threadIdx.x; t < n_live; t += blockDim.x) {
        const uint32_t f = live[t] >> 19, use = live[t] & ((1u << 19) - 1u), pair = use >> 1, blk = use & 1u;
        const uint32_t prefix = (uint32_t)(sk[f] >> (kl.pos_bits + 1));
        const uint32_t rec = prefix / kl.n_orient;
        const PairDev p = j.pairs[pair];
        const uint32_t bp = rec * kl.n_orient + (blk == 0 ? p.rcB : p.rcA);
        const uint32_t cnt = join_core<false>(j, rec, pair, blk, p, f, foff[bp], foff[bp + 1], 0);
        if (cnt == 0u) continue;
        const uint32_t slot = atomicAdd(&s_npt, 1u);
        if (slot < kPairMaxProductive) {
            pt_key[slot] = ((unsigned long long)((rec * j.n_pairs + pair) * 2u + blk) << 13) | f;
            pt_cnt[slot] = cnt;
        }
    }
    my_tasks = __reduce_add_sync(0xFFFFFFFFu, my_tasks);
    if ((threadIdx.x & 31u) != 0u) my_tasks = 0u;
    if (my_tasks) atomicAdd(&s_tasks, my_tasks);
    __syncthreads();
    h.n_tasks = s_tasks;
    const uint32_t n_pt = s_npt;
    mark();  // 10 counted
    if (n_pt == 0u) return true;
    if (n_pt > kPairMaxProductive) return false;  // dense output: the general mode scales with it
    // rank = number of productive tasks with a smaller key; the counts land in rank order and are scanned
    uint32_t my_rank = 0, my_cnt = 0;
    unsigned long long my_key = 0ull;
    if (threadIdx.x < n_pt) {
        my_key = pt_key[threadIdx.x];
        my_cnt = pt_cnt[threadIdx.x];
        for (uint32_t x = 0; x < n_pt; x++) my_rank += pt_key[x] < my_key ? 1u : 0u;
    }
    __syncthreads();
    if (threadIdx.x < n_pt) pt_off[my_rank] = my_cnt;
    __syncthreads();
    uint32_t v = threadIdx.x < n_pt ? pt_off[threadIdx.x] : 0u, n_prod;
    const uint32_t excl = cta_scan_value(v, wsum, &n_prod);
    h.n_prod = n_prod;
    if ((uint64_t)n_prod > j.prod_cap) {
        h.flags |= kPairProdOverflow;
        return true;
    }
    if (threadIdx.x < n_pt) pt_off[threadIdx.x] = excl;  // offset of the task with rank = threadIdx.x
    __syncthreads();
    if (threadIdx.x < n_pt) {
        const uint32_t f = (uint32_t)my_key & (kPairSmallMax - 1u), en = (uint32_t)(my_key >> 13);
        const uint32_t blk = en & 1u, pair = (en >> 1) % j.n_pairs, rec = (en >> 1) / j.n_pairs;
        const PairDev p = j.pairs[pair];
        const uint32_t bp = rec * kl.n_orient + (blk == 0 ? p.rcB : p.rcA);
        join_core<true>(j, rec, pair, blk, p, f, foff[bp], foff[bp + 1], pt_off[my_rank]);
    }
    mark();  // 11 filled
    if (a.dbg && threadIdx.x == 0) a.dbg[0] = (unsigned long long)dbg_n;
    return true;
}

// dynamic shared memory: m sort words (key << 13 | index), then kPairSmallMax final keys
__global__ void __launch_bounds__(1024)
k_pair_small(PairSmallArgs a) {
    extern __shared__ unsigned long long sk[];
    __shared__ uint32_t wsum[33];
    __shared__ uint32_t s_total, s_flags, s_minvar;
    __shared__ unsigned long long s_boff[kPairMaxBlocks + 1], s_ncand, s_ndirty;
    __shared__ const uint8_t *s_blk[kPairMaxBlocks];
    __shared__ unsigned long long s_ns[kPairMaxBlocks];
    __shared__ uint32_t s_timeout;
    JoinArgs j = a.j;
    if (threadIdx.x == 0) s_timeout = 0u;
    if (threadIdx.x < a.n_blocks)
        s_blk[threadIdx.x] = a.peer_blocks ? a.peer_blocks[threadIdx.x] : a.blocks + (uint64_t)threadIdx.x * a.block_bytes;
    __syncthreads();
    // ---- peer exchange: the collective is fused into this kernel.  Thread b waits until rank b has published
    // the block of this step (ld.acquire.sys on its control word, over NVLink for a remote rank); everything read
    // afterwards comes straight out of the owners' memories. ----
    if (a.wait_seq != 0ull) {
        if (threadIdx.x < a.n_blocks) {
            const unsigned long long *p = &reinterpret_cast<const Stage1Ctl *>(s_blk[threadIdx.x])->seq;
            unsigned long long t0, t1;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
            while (ld_acquire_sys_u64(p) != a.wait_seq) {
                __nanosleep(200);
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
                if (t1 - t0 > 4000000000ull) {  // 4 s: a peer died or never ran stage 1
                    atomicOr(&s_timeout, 1u);
                    break;
                }
            }
            s_ns[threadIdx.x] = reinterpret_cast<const Stage1Ctl *>(s_blk[threadIdx.x])->n_sites;  // every rank in parallel
        }
        __syncthreads();
        if (s_timeout) {
            if (threadIdx.x == 0) {
                PairHeader h{};
                h.flags = kPairPeerTimeout;
                *a.header = h;
            }
            return;
        }
        // Pull what was published -- control words, record flags, the sites actually present -- into local
        // memory with ONE wave of independent 16-byte loads over NVLink (every thread has a few in flight), instead
        // of paying the remote latency once per dependent access in the phases below.
        if (a.stage) {
            __shared__ uint32_t s_words[kPairMaxBlocks + 1];
            if (threadIdx.x == 0) {
                uint32_t acc = 0;
                for (uint32_t b = 0; b < a.n_blocks; b++) {
                    s_words[b] = acc;
                    const unsigned long long ns = s_ns[b];
                    const uint64_t have = ns < a.block_site_cap ? ns : a.block_site_cap;
                    acc += (uint32_t)((a.ctl_bytes + have * sizeof(ipcr_site) + 15ull) >> 4);
                }
                s_words[a.n_blocks] = acc;
            }
            __syncthreads();
            const uint32_t total = s_words[a.n_blocks];
            for (uint32_t base = 0; base < total; base += 8u * blockDim.x) {  // eight remote loads in flight per thread
                uint4 v[8];
                uint4 *dst[8];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint32_t g = base + (uint32_t)k * blockDim.x + threadIdx.x;
                    dst[k] = nullptr;
                    if (g < total) {
                        uint32_t b = 0;
                        while (b + 1 < a.n_blocks && g >= s_words[b + 1]) b++;
                        const uint32_t w = g - s_words[b];
                        dst[k] = reinterpret_cast<uint4 *>(a.stage + (uint64_t)b * a.block_bytes) + w;
                        v[k] = reinterpret_cast<const uint4 *>(s_blk[b])[w];
                    }
                }
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (dst[k]) *dst[k] = v[k];
            }
            __syncthreads();
            if (threadIdx.x < a.n_blocks) s_blk[threadIdx.x] = a.stage + (uint64_t)threadIdx.x * a.block_bytes;
            __syncthreads();
        }
    }
    // ---- block headers: counts, overflow flags, merged per-record reset flags ----
    if (threadIdx.x == 0) {
        unsigned long long off = 0ull, nc = 0ull, nd = 0ull;
        uint32_t fl = 0u, vmin = 0xFFFFFFFFu, vmax = 0u;
        for (uint32_t b = 0; b < a.n_blocks; b++) {
            const Stage1Ctl *ctl = reinterpret_cast<const Stage1Ctl *>(s_blk[b]);
            s_boff[b] = off;
            off += ctl->n_sites;
            if (ctl->n_sites > a.block_site_cap || ctl->n_sites > ctl->site_cap) fl |= kPairSitesLost;
            if (ctl->n_cand > ctl->cand_cap || ctl->n_dirty > ctl->dirty_cap) fl |= kPairStage1Overflow;
            nc = max(nc, ctl->n_cand);
            nd = max(nd, ctl->n_dirty);
            vmin = min(vmin, ctl->variant);
            vmax = max(vmax, ctl->variant);
        }
        s_boff[a.n_blocks] = off;
        if (vmin != vmax) fl |= kPairMixedVariant;
        s_flags = fl;
        s_minvar = a.n_blocks ? vmin : 0u;
        s_ncand = nc;
        s_ndirty = nd;
    }
    uint32_t my_reset = 0u;
    for (uint32_t r = threadIdx.x; r < j.n_recs; r += blockDim.x) {
        uint8_t v = 0u;
        for (uint32_t b = 0; b < a.n_blocks; b++) v |= s_blk[b][sizeof(Stage1Ctl) + r];
        a.rec_reset[r] = v;
        my_reset |= v;
    }
    const int any_reset = __syncthreads_or((int)my_reset);
    PairHeader h;
    h.n_sites = s_boff[a.n_blocks];
    h.n_final = h.n_tasks = h.n_prod = 0ull;
    h.n_cand = s_ncand;
    h.n_dirty = s_ndirty;
    h.flags = s_flags | (any_reset ? kPairAnyReset : 0u);
    h.min_variant = s_minvar;
    h.pad[0] = h.pad[1] = 0u;
    if (!(h.flags & (kPairSitesLost | kPairStage1Overflow)) && h.n_sites > kPairSmallMax) h.flags |= kPairNotSmall;
    const uint32_t n = (uint32_t)h.n_sites;
    if ((h.flags & (kPairSitesLost | kPairStage1Overflow | kPairNotSmall)) != 0u || n == 0u) {
        if (threadIdx.x == 0) *a.header = h;
        return;
    }
    if (j.n_recs * j.kl.n_orient <= a.bucket_cap) {
        if (pair_bucket_mode(a, j, h, n, sk, s_boff, s_blk, wsum)) {
            if (threadIdx.x == 0) *a.header = h;
            return;
        }
        __syncthreads();
    }
    auto site_at = [&](uint32_t i) -> const ipcr_site & {
        uint32_t b = 0;
        while (b + 1 < a.n_blocks && (unsigned long long)i >= s_boff[b + 1]) b++;
        return reinterpret_cast<const ipcr_site *>(s_blk[b] + a.ctl_bytes)[i - (uint32_t)s_boff[b]];
    };
    const KeyLayout kl = j.kl;
    uint32_t m = 2;
    while (m < n) m <<= 1;
    unsigned long long *fk = sk + m;  // final keys (shared memory), read by the join through a generic pointer
    // ---- keys + bitonic sort of (key << 13 | index): equal keys keep their input order ----
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x)
        sk[i] = i < n ? ((site_key(site_at(i), a.orients, j.recs, a.rec_reset, j.prm, kl) << kSortSmallIdxBits) | i) : ~0ull;
    __syncthreads();
    for (uint32_t k = 2; k <= m; k <<= 1) {
        for (uint32_t jj = k >> 1; jj > 0; jj >>= 1) {
            for (uint32_t t = threadIdx.x; t < (m >> 1); t += blockDim.x) {
                const uint32_t lo = ((t & ~(jj - 1u)) << 1) | (t & (jj - 1u)), hi = lo + jj;
                const unsigned long long x = sk[lo], y = sk[hi];
                if ((x > y) == ((lo & k) == 0u)) {
                    sk[lo] = y;
                    sk[hi] = x;
                }
            }
            __syncthreads();
        }
    }
    // ---- hit cap + left-window filter: thread t owns the sorted run [t*per, t*per + per) ----
    const uint32_t per = (n + blockDim.x - 1) / blockDim.x;  // <= 8
    const uint32_t b = min(n, threadIdx.x * per), e = min(n, b + per);
    uint32_t keepbits = 0u, kept = 0u;
    for (uint32_t k = b; k < e; k++) {
        const uint64_t key = sk[k] >> kSortSmallIdxBits;
        if (key >= (1ull << kl.key_bits)) continue;  // invalid sites sort last
        const uint64_t first_key = (key >> (kl.pos_bits + 1)) << (kl.pos_bits + 1);
        uint32_t lo = 0, hi = k;  // first sorted index whose key >= first_key
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if ((sk[mid] >> kSortSmallIdxBits) < first_key) lo = mid + 1; else hi = mid;
        }
        const uint32_t rank = k - lo;
        bool kp = true;
        if (j.prm.hit_cap > 0 && (uint64_t)rank >= (uint64_t)j.prm.hit_cap) kp = false;
        if (kp && (site_at((uint32_t)sk[k] & ((1u << kSortSmallIdxBits) - 1u)).flags & IPCR_SITE_TWVIOL)) kp = false;
        if (kp) {
            keepbits |= 1u << (k - b);
            kept++;
        }
    }
    {  // block-wide exclusive scan of `kept`
        const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
        uint32_t x = kept;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
            if (lane >= (uint32_t)d) x += y;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = wsum[lane];
            uint32_t xi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, xi, d);
                if (lane >= (uint32_t)d) xi += y;
            }
            wsum[lane] = xi - w;
            if (lane == 31) s_total = xi;
        }
        __syncthreads();
        uint32_t dst = x - kept + wsum[warp];
        FinalSite *fs = const_cast<FinalSite *>(j.fs);
        for (uint32_t k = b; k < e; k++) {
            if (!((keepbits >> (k - b)) & 1u)) continue;
            const ipcr_site s = site_at((uint32_t)sk[k] & ((1u << kSortSmallIdxBits) - 1u));
            fk[dst] = sk[k] >> kSortSmallIdxBits;
            FinalSite f;
            f.pos = s.pos;
            f.mask = s.mm_mask;
            f.mm = s.mm;
            f.pad = 0;
            fs[dst] = f;
            dst++;
        }
    }
    __syncthreads();
    h.n_final = s_total;
    j.n_sites = s_total;
    j.fkeys = reinterpret_cast<const uint64_t *>(fk);
    if (s_total == 0u) {
        if (threadIdx.x == 0) *a.header = h;
        return;
    }
    // ---- forward sites per (record, pair, direction), then one task per forward site ----
    const uint32_t n_entries = j.n_recs * j.n_pairs * 2u;
    for (uint32_t en = threadIdx.x; en < n_entries; en += blockDim.x) {
        const uint32_t blk = en & 1u, pair = (en >> 1) % j.n_pairs, rec = (en >> 1) / j.n_pairs;
        const PairDev p = j.pairs[pair];
        uint64_t sb, se;
        segment_of(j, rec, blk == 0 ? p.fwdA : p.fwdB, sb, se);
        j.entry_tasks[en] = (uint32_t)(se - sb);
    }
    __syncthreads();
    const uint32_t n_tasks = cta_scan_global(j.entry_tasks, const_cast<uint32_t *>(j.task_off), n_entries, wsum);
    h.n_tasks = n_tasks;
    j.n_tasks = n_tasks;

    if (n_tasks > a.task_cap) h.flags |= kPairTaskOverflow;
    if (n_tasks == 0u || n_tasks > a.task_cap) {
        if (threadIdx.x == 0) *a.header = h;
        return;
    }
    for (uint32_t t = threadIdx.x; t < n_tasks; t += blockDim.x) join_task<false>(j, t);
    __syncthreads();
    const uint32_t n_prod = cta_scan_global(j.task_cnt, const_cast<uint32_t *>(j.prod_off), n_tasks, wsum);
    h.n_prod = n_prod;
    if ((uint64_t)n_prod > j.prod_cap) {
        h.flags |= kPairProdOverflow;
        if (threadIdx.x == 0) *a.header = h;
        return;
    }
    for (uint32_t t = threadIdx.x; t < n_tasks; t += blockDim.x) join_task<true>(j, t);
    if (threadIdx.x == 0) *a.header = h;
}

// the multi-kernel stage 2 wants one contiguous site list: copies the blocks' sites back to back
struct UnpackArgs {
    const uint8_t *blocks;
    const uint8_t *const *peer_blocks;
    uint32_t n_blocks;
    uint64_t block_bytes, ctl_bytes;
    ipcr_site *sites;
    uint64_t site_cap;
};

__global__ void __launch_bounds__(256)
k_unpack_blocks(UnpackArgs a) {
    unsigned long long off = 0ull;
    for (uint32_t bk = 0; bk < a.n_blocks; bk++) {
        const uint8_t *blk = a.peer_blocks ? a.peer_blocks[bk] : a.blocks + (uint64_t)bk * a.block_bytes;
        const unsigned long long n = reinterpret_cast<const Stage1Ctl *>(blk)->n_sites;
        const ipcr_site *ss = reinterpret_cast<const ipcr_site *>(blk + a.ctl_bytes);
        for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
             i += (unsigned long long)gridDim.x * blockDim.x)
            if (off + i < a.site_cap) a.sites[off + i] = ss[i];
        off += n;
    }
}

}  // namespace ipcr
