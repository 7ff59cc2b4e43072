// Host build of K1's two-pass LZ4 decoder (cassandra_b200/csrc/lz4_batch.cuh): the walk is plain C++, the copy runs on the 32-fiber warp
// emulator of warp_emu.h. TEST INFRASTRUCTURE: the same source the GPU compiles, compared with the oracle's decoder on the CPU.
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "warp_emu.h"
static inline int __ffs(int x) { return __builtin_ffs(x); }
#define FULL_MASK 0xFFFFFFFFu
#define B200C_WARP_EMU 1
#include "../../cassandra_b200/csrc/lz4_batch.cuh"

using namespace b200c;

// decodes block src[0..n) into dst (cap = expected size): returns the decoded size, -1 when the walk refuses the block.
// src and dst carry 16 bytes of slack like every engine buffer.
extern "C" int batch_decompress(const uint8_t* src, int n, uint8_t* dst, int cap, int* nseq_out) {
    std::vector<uint16_t> rec((size_t)(n > 0 ? (n - 1) / 3 + 1 : 0) + 4, 0xBEEF);
    const int ns = lz4_walk_thread(src, n, cap, rec.data(), (int)rec.size() - 4);
    if (nseq_out) *nseq_out = ns;
    if (ns < 0) return -1;
    int result = -2;
    warp_emu::run([&](int lane) { int r = lz4_copy_warp(src, n, rec.data(), ns, dst, lane); if (lane == 0) result = r; });
    return result;
}
