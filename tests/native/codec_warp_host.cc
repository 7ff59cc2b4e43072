// Host build of the warp-per-chunk compressors K5 runs (cassandra_b200/csrc/lz4.cuh: lz4_compress_warp, snappy.cuh: snappy_compress_warp) on the
// 32-fiber warp emulator of warp_emu.h. TEST INFRASTRUCTURE: the same source the GPU compiles, byte-compared with the oracle / golden vectors on the CPU.
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "warp_emu.h"
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    unsigned long long pool = ((unsigned long long)b << 32) | a; unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= (unsigned)((pool >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) { sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline unsigned atomicOr(unsigned* p, unsigned v) { unsigned o = *p; *p = o | v; return o; }       // (one fiber runs at a time)
static inline unsigned atomicAnd(unsigned* p, unsigned v) { unsigned o = *p; *p = o & v; return o; }
#include "../../cassandra_b200/csrc/lz4.cuh"
#include "../../cassandra_b200/csrc/lz4_chain.cuh"
#include "../../cassandra_b200/csrc/snappy.cuh"
#include "../../cassandra_b200/csrc/snappy_chain.cuh"

using namespace b200c;

// mode 0: LZ4, chunk copy in "shared memory"; 1: LZ4 reading the chunk in place (the L1 variant); 2 / 3: Snappy with max_bits 14 / 15;
// 4: LZ4 in place with the distinct-hash fast path; 7: LZ4 in two passes; 8 / 9: Snappy in two passes (max_bits 14 / 15, chunks up to 32 KiB); [7] (lz4_chain.cuh: predecessor links, then the parse over an insertion bitmap)
extern "C" int warp_compress(int mode, const uint8_t* in, int n, uint8_t* out) {
    std::vector<uint8_t> s_in((size_t)n + 64, 0); memcpy(s_in.data(), in, n);
    std::vector<uint16_t> tab(1 << 15, 0xDEAD);                       // the kernels zero what they use
    std::vector<uint8_t> dup(8192, 0xEE);
    std::vector<uint32_t> ent((size_t)n + 64, 0xABABABABu), bm((size_t)n / 32 + 2, 0xCDCDCDCDu);
    std::vector<uint16_t> tab2(2 << 15, 0xDEAD);
    int result = -1;
    // 4-byte aligned base as the kernel guarantees
    warp_emu::run([&](int lane) {
        int r;
        if (mode == 8 || mode == 9) {                                        // Snappy in two passes (snappy_chain.cuh), max_bits 14 / 15
            const int mb = mode == 9 ? 15 : 14;
            snappy_chain_build_warp<true>(s_in.data(), n, mb, tab2.data(), ent.data(), lane);
            __syncwarp();
            r = snappy_compress_warp_chain(s_in.data(), n, ent.data(), bm.data(), out, lane);
        }
        else if (mode == 7) {
            lz4_chain_build_warp<true>(s_in.data(), n, tab.data(), ent.data(), lane);
            __syncwarp();
            r = lz4_compress_warp_chain(s_in.data(), n, ent.data(), bm.data(), out, lane);
        }
        else if (mode == 0) r = lz4_compress_warp<false>(s_in.data(), n, tab.data(), out, lane);
        else if (mode == 1) r = lz4_compress_warp<true>(s_in.data(), n, tab.data(), out, lane);
        else if (mode == 4) r = lz4_compress_warp<true>(s_in.data(), n, tab.data(), out, lane, dup.data());
        else if (mode == 5 || mode == 6) r = snappy_compress_warp<true>(s_in.data(), n, tab.data(), mode == 6 ? 15 : 14, out, lane);     // Snappy reading the chunk in place
        else r = snappy_compress_warp(s_in.data(), n, tab.data(), mode == 3 ? 15 : 14, out, lane);
        if (lane == 0) result = r;
    });
    return result;
}
