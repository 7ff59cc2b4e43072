// lz4_thread.cuh — the per-thread (scalar) LZ4 block decoder of K1's thread-per-chunk mapping and its 8-byte write-combining sink.
// Plain C++ (no CUDA intrinsics) marked B200C_HD so that the very same source also compiles with g++: tests/test_lz4_thread_host.py
// fuzzes it on the CPU under AddressSanitizer against the oracle (valid streams must decode exactly, damaged ones must fail or
// stay inside the buffers) — the memcheck the GPU kernel cannot get on a box without a debugger budget.
#pragma once
#include <cstdint>
#include <cstddef>
#ifdef __CUDACC__
#define B200C_HD __host__ __device__ __forceinline__
#define B200C_HD_NOINLINE __host__ __device__
#else
#define B200C_HD inline
#define B200C_HD_NOINLINE inline
#endif

namespace b200c {

enum { LZ4T_MINMATCH = 4 };

B200C_HD uint64_t ld_le64(const uint8_t* p) {
    uintptr_t a = (uintptr_t)p; const uint64_t* q = (const uint64_t*)(a & ~(uintptr_t)7); uint32_t sh = (uint32_t)(a & 7) * 8;
    uint64_t lo = q[0];
    if (sh) lo = (lo >> sh) | (q[1] << (64 - sh));
    return lo;
}
struct WordSink {                       // acc = the bytes of the aligned word around dst + op that lie before dst + op
    uint8_t* dst; int op; uint64_t acc;
    B200C_HD void put(uint64_t v, int n) {        // 1 <= n <= 8, bytes of v above n are zero
        int k = (int)((uintptr_t)(dst + op) & 7);
        acc |= v << (8 * k);
        if (k + n >= 8) { *(uint64_t*)(dst + op - k) = acc; acc = k ? (v >> (8 * (8 - k))) : 0ull; }
        op += n;
    }
    B200C_HD void flush_bytes() {                 // make the pending bytes visible in memory (acc stays valid)
        int k = (int)((uintptr_t)(dst + op) & 7);
        for (int j = 0; j < k; j++) dst[op - k + j] = (uint8_t)(acc >> (8 * j));
    }
    B200C_HD void reload() {                      // after byte-wise stores: pick the partial word up again
        int k = (int)((uintptr_t)(dst + op) & 7);
        acc = k ? (*(const uint64_t*)(dst + op - k) & ((1ull << (8 * k)) - 1ull)) : 0ull;
    }
};
B200C_HD uint64_t low_bytes(uint64_t v, int n) { return n >= 8 ? v : (v & ((1ull << (8 * n)) - 1ull)); }

// LZ4_decompress_safe semantics; dst is 8-byte aligned, src arbitrary (>= 16 readable bytes of slack behind both buffers)
B200C_HD_NOINLINE int lz4_decompress_thread(const uint8_t* __restrict__ src, int n, uint8_t* dst, int cap) {
    if (n == 0) return cap == 0 ? 0 : -1;
    WordSink w{dst, 0, 0ull};
    int ip = 0;
    for (;;) {
        if (ip >= n) return -1;
        uint32_t token = src[ip++];
        int len = (int)(token >> 4);
        if (len == 15) { uint32_t s; do { if (ip >= n) return -1; s = src[ip++]; len += (int)s; } while (s == 255 && len < (1 << 24)); }
        if (n - ip < len || cap - w.op < len) return -1;
        for (; len >= 8; len -= 8, ip += 8) w.put(ld_le64(src + ip), 8);
        if (len) { w.put(low_bytes(ld_le64(src + ip), len), len); ip += len; }
        if (ip == n) break;
        if (n - ip < 2) return -1;
        int offset = (int)src[ip] | ((int)src[ip + 1] << 8); ip += 2;
        if (offset == 0 || offset > w.op) return -1;
        int ml = (int)(token & 15);
        if (ml == 15) { uint32_t s; do { if (ip >= n) return -1; s = src[ip++]; ml += (int)s; } while (s == 255 && ml < (1 << 24)); }
        ml += LZ4T_MINMATCH;
        if (cap - w.op < ml) return -1;
        if (offset >= 16) {                 // the 8 source bytes end at least 8 bytes before op: all of them are in memory already
            for (; ml >= 8; ml -= 8) w.put(ld_le64(dst + w.op - offset), 8);
            if (ml) w.put(low_bytes(ld_le64(dst + w.op - offset), ml), ml);
        } else {                            // short period: byte by byte through memory
            w.flush_bytes();
            for (int i = 0; i < ml; i++) dst[w.op + i] = dst[w.op - offset + i];
            w.op += ml;
            w.reload();
        }
    }
    w.flush_bytes();
    return w.op;
}

} // namespace b200c
