// sortscan.cuh -- device-wide exclusive scan (u32) and stable LSD radix sort
// (u64 keys, u32 values) used by the site pipeline: K4 `sort_sites` replaces
// sortMatchesByPos (/root/reference/core/engine/engine.go:70-76); the scan
// drives the warp/CTA compaction and the count->fill pairing passes.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ipcr {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;  // 1024

// out[i] = sum(in[0..i)) within the block's tile; bsum[block] = tile total
__global__ void __launch_bounds__(kScanThreads)
k_scan_blocks(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint64_t n, uint32_t *__restrict__ bsum) {
    __shared__ uint32_t wsum[kScanThreads / 32];
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t tsum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        v[k] = (base + k < n) ? in[base + k] : 0u;
        tsum += v[k];
    }
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t x = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
        if (lane >= (uint32_t)d) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < kScanThreads / 32 ? wsum[lane] : 0u;
        uint32_t xi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, xi, d);
            if (lane >= (uint32_t)d) xi += y;
        }
        if (lane < kScanThreads / 32) wsum[lane] = xi - w;  // exclusive
        if (lane == kScanThreads / 32 - 1) bsum[blockIdx.x] = xi;
    }
    __syncthreads();
    uint32_t run = x - tsum + wsum[warp];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}

// out[i] += boff[block(i)] ; out[n] = boff[nb] (the grand total)
__global__ void __launch_bounds__(kScanThreads)
k_scan_add(uint32_t *__restrict__ out, uint64_t n, const uint32_t *__restrict__ boff, uint64_t nb) {
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    const uint32_t add = boff[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; k++)
        if (base + k < n) out[base + k] += add;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = boff[nb];
}

__global__ void k_scan_total1(uint32_t *out, uint64_t n, const uint32_t *bsum) { out[n] = bsum[0]; }

// Small inputs (the usual case: a scan leaves a few thousand sites): ONE block scans the whole
// array -- each thread sums a contiguous chunk, the 1024 partials are scanned with shuffles.
constexpr int kScanSmallThreads = 1024;
constexpr uint64_t kScanSmallMax = 1ull << 16;

__global__ void __launch_bounds__(kScanSmallThreads)
k_scan_small(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint32_t n) {
    __shared__ uint32_t wsum[kScanSmallThreads / 32];
    const uint32_t per = (n + kScanSmallThreads - 1) / kScanSmallThreads;
    const uint32_t b = threadIdx.x * per, e = min(n, b + per);
    uint32_t tsum = 0;
    for (uint32_t i = b; i < e; i++) tsum += in[i];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t x = tsum;
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
        if (lane == 31) out[n] = xi;
    }
    __syncthreads();
    uint32_t run = x - tsum + wsum[warp];
    for (uint32_t i = b; i < e; i++) {
        const uint32_t v = in[i];  // read before the write: in and out never alias in this file
        out[i] = run;
        run += v;
    }
}

// words of temporary storage exclusive_scan needs for n items
inline uint64_t scan_tmp_words(uint64_t n) {
    uint64_t words = 0;
    while (true) {
        const uint64_t nb = (n + kScanTile - 1) / kScanTile;
        words += nb + (nb + 1) + 8;
        if (nb <= 1) break;
        n = nb;
    }
    return words + 8;
}

// out must hold n + 1 words (out[n] = total).  n == 0 is allowed.
inline void exclusive_scan(cudaStream_t st, const uint32_t *in, uint32_t *out, uint64_t n, uint32_t *tmp,
                           uint32_t *launches) {
    if (n == 0) {
        cudaMemsetAsync(out, 0, sizeof(uint32_t), st);
        return;
    }
    if (n <= kScanSmallMax) {
        k_scan_small<<<1, kScanSmallThreads, 0, st>>>(in, out, (uint32_t)n);
        if (launches) (*launches)++;
        return;
    }
    const uint64_t nb = (n + kScanTile - 1) / kScanTile;
    uint32_t *bsum = tmp;
    k_scan_blocks<<<(unsigned)nb, kScanThreads, 0, st>>>(in, out, n, bsum);
    if (launches) (*launches)++;
    if (nb == 1) {
        k_scan_total1<<<1, 1, 0, st>>>(out, n, bsum);
        if (launches) (*launches)++;
        return;
    }
    uint32_t *boff = tmp + nb;  // nb + 1 words
    exclusive_scan(st, bsum, boff, nb, tmp + nb + (nb + 1) + 8, launches);
    k_scan_add<<<(unsigned)nb, kScanThreads, 0, st>>>(out, n, boff, nb);
    if (launches) (*launches)++;
}

// ---- radix sort ---------------------------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortChunks = 8;
constexpr int kSortTile = kSortThreads * kSortChunks;  // 2048 keys per block

// hist[d * nb + block] = number of keys of the block's tile with digit d
__global__ void __launch_bounds__(kSortThreads)
k_sort_hist(const uint64_t *__restrict__ keys, uint64_t n, uint32_t shift, uint32_t *__restrict__ hist, uint64_t nb) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * kSortTile;
    for (int c = 0; c < kSortChunks; c++) {
        const uint64_t i = base + (uint64_t)c * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(uint64_t)threadIdx.x * nb + blockIdx.x] = h[threadIdx.x];
}

// stable scatter: order inside a tile is (chunk, thread) order
__global__ void __launch_bounds__(kSortThreads)
k_sort_scatter(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t *__restrict__ keys_out,
               uint32_t *__restrict__ vals_out, uint64_t n, uint32_t shift, const uint32_t *__restrict__ hoff,
               uint64_t nb) {
    __shared__ uint32_t run[256];                       // keys of each digit already placed by earlier chunks
    __shared__ uint32_t wcnt[kSortThreads / 32][256];   // per-warp digit counts of the current chunk
    run[threadIdx.x] = hoff[(uint64_t)threadIdx.x * nb + blockIdx.x];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t base = (uint64_t)blockIdx.x * kSortTile;
    for (int c = 0; c < kSortChunks; c++) {
        for (int w = 0; w < kSortThreads / 32; w++) wcnt[w][threadIdx.x] = 0;
        __syncthreads();
        const uint64_t i = base + (uint64_t)c * kSortThreads + threadIdx.x;
        const bool valid = i < n;
        uint64_t key = 0;
        uint32_t val = 0, d = 256u + lane;  // invalid lanes never share a digit
        if (valid) {
            key = keys[i];
            val = vals[i];
            d = (uint32_t)(key >> shift) & 255u;
        }
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
        const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank_in_warp == 0) wcnt[warp][d] = __popc(peers);
        __syncthreads();
        if (valid) {
            uint32_t off = run[d] + rank_in_warp;
            for (uint32_t w = 0; w < warp; w++) off += wcnt[w][d];
            keys_out[off] = key;
            vals_out[off] = val;
        }
        __syncthreads();
        uint32_t tot = 0;
        for (int w = 0; w < kSortThreads / 32; w++) tot += wcnt[w][threadIdx.x];
        run[threadIdx.x] += tot;
        __syncthreads();
    }
}

// Small inputs: ONE block sorts in shared memory (bitonic network on key << 13 | index, so the
// order of equal keys is the input order, like the stable radix sort).
constexpr uint32_t kSortSmallMax = 8192;
constexpr uint32_t kSortSmallIdxBits = 13;

__global__ void __launch_bounds__(1024)
k_sort_small(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t *__restrict__ keys_out,
             uint32_t *__restrict__ vals_out, uint32_t n, uint32_t m, uint64_t key_mask) {
    extern __shared__ unsigned long long sk[];
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x)
        sk[i] = i < n ? (((keys[i] & key_mask) << kSortSmallIdxBits) | i) : ~0ull;
    __syncthreads();
    for (uint32_t k = 2; k <= m; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < (m >> 1); t += blockDim.x) {
                const uint32_t lo = ((t & ~(j - 1u)) << 1) | (t & (j - 1u)), hi = lo + j;
                const unsigned long long a = sk[lo], b = sk[hi];
                if ((a > b) == ((lo & k) == 0u)) {
                    sk[lo] = b;
                    sk[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t src = (uint32_t)sk[i] & ((1u << kSortSmallIdxBits) - 1u);
        keys_out[i] = keys[src];
        vals_out[i] = vals[src];
    }
}

inline uint64_t sort_tmp_words(uint64_t n) {
    const uint64_t nb = (n + kSortTile - 1) / kSortTile;
    const uint64_t hn = 256 * nb;
    return hn + (hn + 1) + scan_tmp_words(hn) + 16;
}

// Sorts (keys, vals) by the low `key_bits` bits.  Ping-pongs between (k0, v0) and
// (k1, v1); returns 0 if the result is in (k0, v0), 1 if in (k1, v1).
inline int radix_sort(cudaStream_t st, uint64_t *k0, uint32_t *v0, uint64_t *k1, uint32_t *v1, uint64_t n,
                      uint32_t key_bits, uint32_t *tmp, uint32_t *launches) {
    if (n <= 1) return 0;
    if (n <= kSortSmallMax && key_bits + kSortSmallIdxBits <= 64) {
        uint32_t m = 2;
        while (m < n) m <<= 1;
        if (m * sizeof(unsigned long long) > 48 * 1024)  // per device: cheap enough to repeat
            cudaFuncSetAttribute(k_sort_small, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(kSortSmallMax * sizeof(unsigned long long)));
        const uint64_t key_mask = key_bits >= 64 ? ~0ull : ((1ull << key_bits) - 1ull);
        k_sort_small<<<1, 1024, m * sizeof(unsigned long long), st>>>(k0, v0, k1, v1, (uint32_t)n, m, key_mask);
        if (launches) (*launches)++;
        return 1;
    }
    const uint64_t nb = (n + kSortTile - 1) / kSortTile;
    const uint64_t hn = 256 * nb;
    uint32_t *hist = tmp, *hoff = tmp + hn, *stmp = tmp + hn + (hn + 1);
    int cur = 0;
    for (uint32_t shift = 0; shift < key_bits; shift += 8) {
        const uint64_t *ki = cur ? k1 : k0;
        const uint32_t *vi = cur ? v1 : v0;
        uint64_t *ko = cur ? k0 : k1;
        uint32_t *vo = cur ? v0 : v1;
        k_sort_hist<<<(unsigned)nb, kSortThreads, 0, st>>>(ki, n, shift, hist, nb);
        exclusive_scan(st, hist, hoff, hn, stmp, launches);
        k_sort_scatter<<<(unsigned)nb, kSortThreads, 0, st>>>(ki, vi, ko, vo, n, shift, hoff, nb);
        if (launches) *launches += 2;
        cur ^= 1;
    }
    return cur;
}

}  // namespace ipcr
