// lz4_batch.cuh — LZ4 block decoding in two kernels (K1, B200C_K1=2).
//
// Decoding an LZ4 block is a chain: where sequence i + 1 starts is only known once sequence i's token (and length bytes) are read. The
// thread-per-chunk decoder (lz4_thread.cuh) walks that chain and copies as it goes; with ~10^5 chunks in flight every match copy is a
// random read of a 16 KiB output window that has long left L2 (ncu: 5.6 x the algorithmic bytes from DRAM). Here the chain and the copies
// are separated:
//   walk (lz4_walk_thread, one THREAD per chunk): reads only tokens, length bytes and offsets, checks everything LZ4_decompress_safe checks
//        (literal / match lengths against both buffers, offset != 0 and inside the output, exact output size) and records where every
//        sequence's token starts (u16). Sequential reads of a private stream, two-byte writes: no random access at all.
//   copy (lz4_copy_warp, one WARP per chunk, few chunks in flight): 32 sequences per step, one per lane. The output positions are a
//        prefix sum; all literals of the step are copied at once; the matches go in as few rounds as their dependencies allow (a match
//        waits only for the earlier matches of the same step whose destination overlaps its source — found with two binary searches over
//        the lanes' output positions). Long literals and matches are copied by the whole warp.
// Both compile for the host: the walk is plain C++ (fuzzed under ASan like lz4_thread.cuh), the copy runs on the warp emulator.
#pragma once
#include "lz4_thread.cuh"

namespace b200c {

enum { LZ4B_COOP = 48 };          // literal / match runs longer than this are copied by the whole warp

// Returns the number of sequences (>= 1), or -1 for a malformed block or one that does not decode to exactly `cap` bytes.
// rec[k] = offset of sequence k's token in src. n == 0 decodes to nothing (cap must be 0): 0 sequences.
B200C_HD_NOINLINE int lz4_walk_thread(const uint8_t* __restrict__ src, int n, int cap, uint16_t* __restrict__ rec, int rec_cap) {
    if (n == 0) return cap == 0 ? 0 : -1;
    int ip = 0, op = 0, k = 0;
    for (;;) {
        if (ip >= n || k >= rec_cap) return -1;
        const int tpos = ip;
        const uint32_t token = src[ip++];
        int len = (int)(token >> 4);
        if (len == 15) { uint32_t s; do { if (ip >= n) return -1; s = src[ip++]; len += (int)s; } while (s == 255 && len < (1 << 24)); }
        if (n - ip < len || cap - op < len) return -1;
        ip += len; op += len;
        rec[k++] = (uint16_t)tpos;
        if (ip == n) break;
        if (n - ip < 2) return -1;
        const int offset = (int)src[ip] | ((int)src[ip + 1] << 8); ip += 2;
        if (offset == 0 || offset > op) return -1;
        int ml = (int)(token & 15);
        if (ml == 15) { uint32_t s; do { if (ip >= n) return -1; s = src[ip++]; ml += (int)s; } while (s == 255 && ml < (1 << 24)); }
        ml += LZ4T_MINMATCH;
        if (cap - op < ml) return -1;
        op += ml;
    }
    return op == cap ? k : -1;
}

#if defined(__CUDACC__) || defined(B200C_WARP_EMU)
// count of lanes whose (non-decreasing) value is <= key (le) or < key (!le), capped at 31
__device__ __forceinline__ int lz4b_count(int arr, int key, bool le) {
    int pos = 0;
#pragma unroll
    for (int step = 16; step; step >>= 1) {
        const int v = __shfl_sync(FULL_MASK, arr, pos + step - 1);
        if (le ? (v <= key) : (v < key)) pos += step;
    }
    return pos;
}

// Copies what a block the walk accepted decodes to: src/n the block, rec/nseq the walk's result, dst the output (global memory; a warp
// reads back what its own lanes stored after __syncwarp()). Returns the decoded size (warp-uniform).
__device__ __forceinline__ int lz4_copy_warp(const uint8_t* __restrict__ src, int n, const uint16_t* __restrict__ rec, int nseq, uint8_t* dst, int lane) {
    const uint32_t lt_mask = (1u << lane) - 1u;
    int base = 0;
    for (int b0 = 0; b0 < nseq; b0 += 32) {
        const int i = b0 + lane;
        int lit = 0, ml = 0, off = 0, ls = 0;
        if (i < nseq) {
            int q = (int)rec[i];
            const uint32_t t = src[q++];
            lit = (int)(t >> 4);
            if (lit == 15) { uint32_t s; do { s = src[q++]; lit += (int)s; } while (s == 255); }
            ls = q; q += lit;
            if (q < n) {
                off = (int)src[q] | ((int)src[q + 1] << 8); q += 2;
                ml = (int)(t & 15);
                if (ml == 15) { uint32_t s; do { s = src[q++]; ml += (int)s; } while (s == 255); }
                ml += LZ4T_MINMATCH;
            }
        }
        // output positions: exclusive prefix sum of the sequence sizes
        const int tot = lit + ml;
        int incl = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t2 = __shfl_sync(FULL_MASK, incl, lane - d); if (lane >= d) incl += t2; }
        const int my_op = base + incl - tot, mop = my_op + lit;
        base += __shfl_sync(FULL_MASK, incl, 31);
        // ---- literals: all of the step at once -----------------------------------------------------------------------------------
        if (lit <= LZ4B_COOP) for (int k = 0; k < lit; k++) dst[my_op + k] = src[ls + k];
        for (uint32_t big = __ballot_sync(FULL_MASK, lit > LZ4B_COOP); big; big &= big - 1) {
            const int l = __ffs(big) - 1;
            const int bs = __shfl_sync(FULL_MASK, ls, l), bo = __shfl_sync(FULL_MASK, my_op, l), bn = __shfl_sync(FULL_MASK, lit, l);
            for (int k = lane; k < bn; k += 32) dst[bo + k] = src[bs + k];
        }
        __syncwarp();
        // ---- matches: which earlier matches of this step write into my source? their destinations are [mop_k, mop_k + ml_k), both ends
        //      non-decreasing in k; my source is [s, e) -------------------------------------------------------------------------------
        const int s = mop - off, e = (off < ml) ? mop : s + ml;
        const int k_lo = lz4b_count(mop + ml, s, true);         // lanes whose destination ends at or before s
        const int k_hi = lz4b_count(mop, e, false);             // lanes whose destination starts before e
        const uint32_t upto_hi = k_hi >= 32 ? FULL_MASK : ((1u << k_hi) - 1u);
        const uint32_t dep = upto_hi & ~((1u << k_lo) - 1u) & lt_mask;
        bool pending = ml > 0;
        for (;;) {
            const uint32_t pm = __ballot_sync(FULL_MASK, pending);
            if (!pm) break;
            const bool can = pending && !(pm & dep);
            if (can && ml <= LZ4B_COOP) {
                if (off >= ml) for (int k = 0; k < ml; k++) dst[mop + k] = dst[s + k];
                else for (int k = 0; k < ml; k++) dst[mop + k] = dst[s + k % off];          // overlapping: the period is the last `off` bytes
            }
            for (uint32_t big = __ballot_sync(FULL_MASK, can && ml > LZ4B_COOP); big; big &= big - 1) {
                const int l = __ffs(big) - 1;
                const int bs = __shfl_sync(FULL_MASK, s, l), bo = __shfl_sync(FULL_MASK, mop, l), bn = __shfl_sync(FULL_MASK, ml, l), bf = __shfl_sync(FULL_MASK, off, l);
                for (int k = lane; k < bn; k += 32) dst[bo + k] = dst[bs + (k < bf ? k : k % bf)];
            }
            if (can) pending = false;
            __syncwarp();
        }
    }
    return base;
}
#endif

} // namespace b200c
