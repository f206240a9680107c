// sort_products.cuh -- `--sort` on the device.  A dense-hit panel returns 10^6+ products and the
// reference's collector then spends its time in sort.SliceStable with string keys
// (/root/reference/internal/common/sort.go:34-74).  Here the comparator becomes integer keys -- the strings
// enter as RANKS computed once per run on the host (equal strings, equal rank) -- and a stable LSD radix sort
// (sortscan.cuh) orders the products:
//   LessProduct = (SourceFile, base SequenceID) -> rec_rank ; Start + off ; End + off ; Length ; Type ;
//                 ExperimentID -> pair_rank ; FwdMM ; RevMM ; [mismatch lists, SequenceID, Seq: host tie-break]
// Four key words, least significant sorted first; every pass is stable, so the last pass leaves LessProduct order.
#pragma once

#include "device_types.cuh"
#include "sortscan.cuh"

namespace ipcr {

struct ProductKeyArgs {
    const ipcr_product *products;
    uint64_t n;
    const uint32_t *rec_rank;   // [n_recs]
    const long long *rec_off;   // [n_recs] chunk offset added to start / end, or NULL
    const uint32_t *pair_rank;  // [n_pairs]
    uint64_t *k1, *k2, *k3, *k4;  // most .. least significant
    uint32_t *idx;
};

__global__ void __launch_bounds__(256)
k_product_keys(ProductKeyArgs a) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const ipcr_product p = a.products[i];
    const long long off = a.rec_off ? a.rec_off[p.rec_idx] : 0ll;
    a.k1[i] = ((uint64_t)a.rec_rank[p.rec_idx] << 44) | (uint64_t)(p.start + off);
    a.k2[i] = (uint64_t)(p.end + off);
    a.k3[i] = ((uint64_t)p.length << 20) | ((uint64_t)p.type << 19) | (uint64_t)a.pair_rank[p.pair_idx];
    a.k4[i] = ((uint64_t)p.fwd_mm << 8) | (uint64_t)p.rev_mm;
    a.idx[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256)
k_gather_u64(const uint64_t *__restrict__ src, const uint32_t *__restrict__ idx, uint64_t n, uint64_t *__restrict__ dst) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

__global__ void __launch_bounds__(256)
k_gather_products(const ipcr_product *__restrict__ src, const uint32_t *__restrict__ idx, uint64_t n,
                  ipcr_product *__restrict__ dst, uint8_t *__restrict__ tie, const uint64_t *__restrict__ k1,
                  const uint64_t *__restrict__ k2, const uint64_t *__restrict__ k3, const uint64_t *__restrict__ k4) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = idx[i];
    // 64-byte records: four 16-byte loads / stores per product
    const uint4 *sp = reinterpret_cast<const uint4 *>(src + s);
    uint4 *dp = reinterpret_cast<uint4 *>(dst + i);
#pragma unroll
    for (int k = 0; k < 4; k++) dp[k] = sp[k];
    // tie[i] = product i compares equal to product i-1 on every device key: the host finishes those runs
    uint8_t t = 0;
    if (i > 0) {
        const uint32_t q = idx[i - 1];
        t = (k1[s] == k1[q] && k2[s] == k2[q] && k3[s] == k3[q] && k4[s] == k4[q]) ? 1 : 0;
    }
    tie[i] = t;
}

}  // namespace ipcr
