// meta.cuh — the rest of an output sstable, produced on the device while the merged stream is still in HBM:
//   Filter.db     BloomFilter.add of every written key            S/utils/BloomFilter.java:79-122 (hash3_x64_128, S/utils/MurmurHash.java:178-260)
//   Summary.db    every min_index_interval-th Index.db entry      S/io/sstable/indexsummary/IndexSummaryBuilder.java:200-228
//   Statistics.db the MetadataCollector reductions                S/io/sstable/metadata/MetadataCollector.java:107-147,208-270
// The per-cell / per-row part of the statistics is gathered inside K4 where the rows are written (StatAcc in partition.cuh) and folded
// into StatGlobal with a handful of atomics per block; everything per KEY or per PARTITION (bloom bits, HyperLogLog registers, the two
// EstimatedHistograms, index-summary samples, first / last key) is one thread per output partition in k_meta_keys after the positions are known.
#pragma once
#include "partition.cuh"

namespace b200c {

enum { META_PSIZE = B200C_PSIZE_BUCKETS, META_CELLS = B200C_CELLS_BUCKETS, META_HLL = 1 << B200C_HLL_P };

struct StatGlobal {
    long long min_ts, max_ts, min_ldt, max_ldt; int min_ttl, max_ttl; unsigned int seen, pdel;
    unsigned long long rows, cols, cells, tombs;
    unsigned long long psize[META_PSIZE], cells_hist[META_CELLS];
    long long psize_off[META_PSIZE], cells_off[META_CELLS];       // EstimatedHistogram.newOffsets, filled by the host
    unsigned int hll[META_HLL];
    unsigned long long nsamples, sum_bytes;                       // index-summary samples / entry bytes emitted so far
    unsigned int first_len, last_len;
    unsigned long long tdrop_n, tdrop_overflow;
    long long tdrop_point[B200C_TDROP_CAP]; unsigned long long tdrop_count[B200C_TDROP_CAP];
};

__device__ __forceinline__ void stat_flush(StatGlobal* g, const StatAcc& a) {      // one thread's (or one warp's reduced) accumulator -> global
    if (a.seen & 1) { atomicMin(&g->min_ts, (long long)a.min_ts); atomicMax(&g->max_ts, (long long)a.max_ts); }
    if (a.seen & 2) { atomicMin(&g->min_ldt, (long long)a.min_ldt); atomicMax(&g->max_ldt, (long long)a.max_ldt); }
    if (a.seen & 4) { atomicMin(&g->min_ttl, a.min_ttl); atomicMax(&g->max_ttl, a.max_ttl); }
    if (a.seen) atomicOr(&g->seen, a.seen);
    if (a.pdel) atomicOr(&g->pdel, 1u);
    if (a.rows) atomicAdd(&g->rows, a.rows);
    if (a.cols) atomicAdd(&g->cols, a.cols);
    if (a.cells) atomicAdd(&g->cells, a.cells);
    if (a.tombs) atomicAdd(&g->tombs, a.tombs);
}
// warp-wide fold of the accumulators (all 32 lanes call it), lane 0 flushes
__device__ __forceinline__ void stat_flush_warp(StatGlobal* g, StatAcc a) {
    if (!(a.seen & 1)) { a.min_ts = I64_MAX; a.max_ts = I64_MIN; }
    if (!(a.seen & 2)) { a.min_ldt = I64_MAX; a.max_ldt = I64_MIN; }
    if (!(a.seen & 4)) { a.min_ttl = 0x7FFFFFFF; a.max_ttl = (int32_t)0x80000000; }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        int64_t o;
        o = __shfl_xor_sync(FULL_MASK, a.min_ts, d); a.min_ts = o < a.min_ts ? o : a.min_ts;
        o = __shfl_xor_sync(FULL_MASK, a.max_ts, d); a.max_ts = o > a.max_ts ? o : a.max_ts;
        o = __shfl_xor_sync(FULL_MASK, a.min_ldt, d); a.min_ldt = o < a.min_ldt ? o : a.min_ldt;
        o = __shfl_xor_sync(FULL_MASK, a.max_ldt, d); a.max_ldt = o > a.max_ldt ? o : a.max_ldt;
        int32_t t;
        t = __shfl_xor_sync(FULL_MASK, a.min_ttl, d); a.min_ttl = t < a.min_ttl ? t : a.min_ttl;
        t = __shfl_xor_sync(FULL_MASK, a.max_ttl, d); a.max_ttl = t > a.max_ttl ? t : a.max_ttl;
        a.seen |= __shfl_xor_sync(FULL_MASK, a.seen, d); a.pdel |= __shfl_xor_sync(FULL_MASK, a.pdel, d);
        a.rows += __shfl_xor_sync(FULL_MASK, a.rows, d); a.cols += __shfl_xor_sync(FULL_MASK, a.cols, d);
        a.cells += __shfl_xor_sync(FULL_MASK, a.cells, d); a.tombs += __shfl_xor_sync(FULL_MASK, a.tombs, d);
    }
    if ((threadIdx.x & 31) == 0) stat_flush(g, a);
}

// MurmurHash.hash3_x64_128(key, seed 0) -> (h1, h2)   (the token function of compact.cu returns h1 only)
__device__ void murmur3_x64_128_dev(const uint8_t* key, uint32_t len, uint64_t& o1, uint64_t& o2) {
    const uint32_t nblocks = len >> 4;
    uint64_t h1 = 0, h2 = 0;
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    auto rotl = [](uint64_t v, int n) { return (v << n) | (v >> (64 - n)); };
    auto fmix = [](uint64_t k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k; };
    for (uint32_t i = 0; i < nblocks; i++) {
        uint64_t k1 = 0, k2 = 0;
        for (int b = 0; b < 8; b++) { k1 |= (uint64_t)key[i * 16 + b] << (8 * b); k2 |= (uint64_t)key[i * 16 + 8 + b] << (8 * b); }
        k1 *= c1; k1 = rotl(k1, 31); k1 *= c2; h1 ^= k1; h1 = rotl(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl(k2, 33); k2 *= c1; h2 ^= k2; h2 = rotl(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const uint8_t* t = key + nblocks * 16; uint64_t k1 = 0, k2 = 0; const int rem = len & 15;
    for (int i = rem - 1; i >= 8; i--) k2 ^= (uint64_t)(int64_t)(int8_t)t[i] << (8 * (i - 8));
    if (rem > 8) { k2 *= c2; k2 = rotl(k2, 33); k2 *= c1; h2 ^= k2; }
    for (int i = (rem < 8 ? rem : 8) - 1; i >= 0; i--) k1 ^= (uint64_t)(int64_t)(int8_t)t[i] << (8 * i);
    if (rem > 0) { k1 *= c1; k1 = rotl(k1, 31); k1 *= c2; h1 ^= k1; }
    h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
    h1 += h2; h2 += h1; h1 = fmix(h1); h2 = fmix(h2); h1 += h2; h2 += h1;
    o1 = h1; o2 = h2;
}
// MurmurHash.hash2_64(key, seed 0)  S/utils/MurmurHash.java:94-152 (tail bytes sign-extended)
__device__ uint64_t murmur2_64_dev(const uint8_t* key, uint32_t len) {
    const uint64_t m = 0xc6a4a7935bd1e995ULL; const int r = 47;
    uint64_t h = m * (uint64_t)len;
    const uint32_t nl = len >> 3;
    for (uint32_t i = 0; i < nl; i++) {
        uint64_t k = 0; for (int b = 0; b < 8; b++) k += (uint64_t)key[i * 8 + b] << (8 * b);
        k *= m; k ^= k >> r; k *= m; h ^= k; h *= m;
    }
    const int rem = len & 7; const uint8_t* t = key + len - rem;
    for (int i = rem - 1; i >= 1; i--) h ^= (uint64_t)(int64_t)(int8_t)t[i] << (8 * i);
    if (rem >= 1) { h ^= (uint64_t)(int64_t)(int8_t)t[0]; h *= m; }
    h ^= h >> r; h *= m; h ^= h >> r;
    return h;
}
__device__ __forceinline__ int hist_index(const long long* offs, int n, long long v) {        // EstimatedHistogram.findIndex: first offset >= v
    int a = 0, b = n; while (a < b) { int mid = (a + b) >> 1; if (offs[mid] < v) a = mid + 1; else b = mid; } return a;
}

struct MetaArgs {
    const CParams* P; const uint64_t* contrib; const uint64_t* op_first; const uint64_t* upos; const uint64_t* pbase;
    const uint64_t* dsize; const uint64_t* ipos; const uint32_t* ihead; const uint32_t* ccount; const uint64_t* wrank;   // wrank: exclusive scan of (dsize > 0)
    uint64_t nparts; uint64_t index_base;         // Index.db bytes written by earlier pieces
    uint64_t written_base;                        // partitions written by earlier pieces
    StatGlobal* sg; uint32_t* bloom; uint64_t bloom_bits; int bloom_k; uint32_t interval;
    uint32_t* sample_j;                           // out: partition of every summary sample taken in this piece, in order
    uint8_t* first_key; uint8_t* last_key;        // 65535-byte device buffers
};

// one thread per output partition of the piece
__global__ void __launch_bounds__(256) k_meta_keys(const MetaArgs a) {
    __shared__ unsigned int s_ps[META_PSIZE], s_cs[META_CELLS];
    for (int i = threadIdx.x; i < META_PSIZE; i += blockDim.x) s_ps[i] = 0;
    for (int i = threadIdx.x; i < META_CELLS; i += blockDim.x) s_cs[i] = 0;
    __syncthreads();
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < a.nparts && a.dsize[j]) {
        StatGlobal* g = a.sg;
        const uint64_t e = a.contrib[a.op_first[j]]; const uint64_t gi = a.pbase[(e >> 56) & 0x7F] + (e & 0xFFFFFFFFFFull);
        const uint8_t* key = a.P->U + a.upos[gi] + 2; const uint32_t kl = a.ihead[j] - 2;
        atomicAdd(&s_ps[hist_index(g->psize_off, META_PSIZE - 1, (long long)a.dsize[j])], 1u);          // addPartitionSizeInBytes
        atomicAdd(&s_cs[hist_index(g->cells_off, META_CELLS - 1, (long long)a.ccount[j])], 1u);         // addCellPerPartitionCount
        {   // addKey -> HyperLogLogPlus(13, 25).offerHashed, dense registers
            const uint64_t h = murmur2_64_dev(key, kl); const uint32_t idx = (uint32_t)(h >> (64 - B200C_HLL_P));
            const uint64_t w = (h << B200C_HLL_P) | (1ull << (B200C_HLL_P - 1));
            atomicMax(&g->hll[idx], (unsigned int)(__clzll((long long)w) + 1));
        }
        if (a.bloom_bits) {                                                     // BloomFilter.add: indexes = |(h2 + i * h1) % capacity|
            uint64_t h1, h2; murmur3_x64_128_dev(key, kl, h1, h2);
            long long base = (long long)h2; const long long inc = (long long)h1, cap = (long long)a.bloom_bits;
            for (int i = 0; i < a.bloom_k; i++) {
                long long r = base % cap; const uint64_t idx = (uint64_t)(r < 0 ? -r : r);
                atomicOr(&a.bloom[idx >> 5], 1u << (idx & 31));                  // byte idx >> 3, bit idx & 7 of a little-endian word
                base = (long long)((unsigned long long)base + (unsigned long long)inc);
            }
        }
        const uint64_t rank = a.written_base + a.wrank[j];
        if (rank % a.interval == 0) a.sample_j[rank / a.interval - (a.written_base + a.interval - 1) / a.interval] = (uint32_t)j;   // maybeAddEntry at full sampling
        if (rank == 0) { for (uint32_t i = 0; i < kl; i++) a.first_key[i] = key[i]; g->first_len = kl; }
        if (a.wrank[j] + 1 == a.wrank[a.nparts]) { for (uint32_t i = 0; i < kl; i++) a.last_key[i] = key[i]; g->last_len = kl; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < META_PSIZE; i += blockDim.x) if (s_ps[i]) atomicAdd(&a.sg->psize[i], (unsigned long long)s_ps[i]);
    for (int i = threadIdx.x; i < META_CELLS; i += blockDim.x) if (s_cs[i]) atomicAdd(&a.sg->cells_hist[i], (unsigned long long)s_cs[i]);
}
__global__ void __launch_bounds__(256) k_written_flags(uint64_t nparts, const uint64_t* __restrict__ dsize, uint32_t* __restrict__ flag) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nparts) flag[j] = dsize[j] ? 1u : 0u;
}
// index-summary entries of this piece's samples: size pass, then emit `key | i64 Index.db position (native order)` (IndexSummaryBuilder :204-209)
__global__ void __launch_bounds__(256) k_summary_sizes(const MetaArgs a, uint64_t nsamples, uint32_t* __restrict__ esize) {
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < nsamples) esize[s] = a.ihead[a.sample_j[s]] - 2 + 8;
}
__global__ void __launch_bounds__(256) k_summary_emit(const MetaArgs a, uint64_t nsamples, const uint64_t* __restrict__ epos, uint8_t* __restrict__ entries /* + sum_bytes */,
                                                      uint64_t* __restrict__ eoffs /* + nsamples so far: offset of every entry in the entries region */) {
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsamples) return;
    const uint32_t j = a.sample_j[s];
    const uint64_t e = a.contrib[a.op_first[j]]; const uint64_t gi = a.pbase[(e >> 56) & 0x7F] + (e & 0xFFFFFFFFFFull);
    const uint8_t* key = a.P->U + a.upos[gi] + 2; const uint32_t kl = a.ihead[j] - 2;
    const uint64_t base = a.sg->sum_bytes + epos[s];
    uint8_t* d = entries + base;
    for (uint32_t i = 0; i < kl; i++) d[i] = key[i];
    const uint64_t pos = a.index_base + a.ipos[j];
    for (int b = 0; b < 8; b++) d[kl + b] = (uint8_t)(pos >> (8 * b));
    eoffs[a.sg->nsamples + s] = base;
}
__global__ void k_meta_advance(StatGlobal* g, uint64_t nsamples, const uint64_t* __restrict__ epos) {
    g->nsamples += nsamples; g->sum_bytes += nsamples ? epos[nsamples] : 0;
}
// exact tombstone drop-time histogram: hash table -> dense list -> the B200C_TDROP_CAP smallest points, ascending (rank by counting)
struct TdropDense { unsigned long long n; unsigned long long key[TDROP_SLOTS], cnt[TDROP_SLOTS]; };
__global__ void __launch_bounds__(1024) k_tdrop_compact(const TdropTable* __restrict__ td, TdropDense* __restrict__ d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= TDROP_SLOTS) return;
    const unsigned long long k = td->key[i];
    if (k == ~0ull) return;
    const unsigned long long at = atomicAdd(&d->n, 1ull);
    d->key[at] = k; d->cnt[at] = td->cnt[i];
}
__global__ void __launch_bounds__(1024) k_tdrop_final(const TdropTable* __restrict__ td, const TdropDense* __restrict__ d, StatGlobal* g) {
    const unsigned long long n = d->n; const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { g->tdrop_n = n; if (td->overflow || n > B200C_TDROP_CAP) g->tdrop_overflow = 1; }
    if (i >= n) return;
    const unsigned long long k = d->key[i];
    unsigned int rank = 0;
    for (unsigned long long q = 0; q < n; q++) rank += d->key[q] < k ? 1u : 0u;
    if (rank < B200C_TDROP_CAP) { g->tdrop_point[rank] = (long long)k; g->tdrop_count[rank] = d->cnt[i]; }
}

} // namespace b200c
