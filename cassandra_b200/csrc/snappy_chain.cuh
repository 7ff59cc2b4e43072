// snappy_chain.cuh — Snappy compression of one chunk in two passes, the scheme of lz4_chain.cuh applied to snappy.cuh's matcher: pass A
// (chain_build_warp with Snappy's hash) links every position to its two nearest earlier positions with the same table index, pass B parses
// with one "inserted" bit per position instead of the 32 KiB (64 KiB for the 15-bit generation) table that limits snappy.cuh to 6 (3) chunks
// per SM. Byte-identical to snappy.cuh, i.e. to Google's library (tests/test_snappy_golden.py, tests/test_codec_warp_host.py).
// CompressFragment inserts positions in increasing order (search attempts move forward; after a copy it inserts ip - 1, then tests and
// inserts ip), so table[h] is the largest inserted position with index h, or 0 while there is none — exactly the first inserted position
// along the chain. One fragment: chunks of up to 32 KiB (15-bit links; a fragment is 64 KiB).
#pragma once
#include "snappy.cuh"
#include "lz4_chain.cuh"

namespace b200c {

struct SnappyHash { uint32_t tmask; int max_bits; __device__ __forceinline__ uint32_t operator()(uint32_t bytes) const { return snappy_tidx(bytes, tmask, max_bits); } };
__host__ __device__ __forceinline__ int snappy_table_size(int frag, int max_bits) { int t = 256; while (t < frag && t < (1 << max_bits)) t <<= 1; return t; }
// positions the matcher can look up or insert: everything below ip_limit = n - 15
__host__ __device__ __forceinline__ int snappyc_positions(int n) { return n >= 15 ? n - 14 : 0; }

template <bool GLOBAL> __device__ __forceinline__ void snappy_chain_build_warp(const uint8_t* s_in, int n, int max_bits, uint16_t* s_t1, uint32_t* __restrict__ ent, int lane) {
    const int tsz = snappy_table_size(n, max_bits);
    chain_build_warp<GLOBAL>(s_in, snappyc_positions(n), tsz, 31 - __clz(tsz), SnappyHash{(uint32_t)tsz - 1u, max_bits}, s_t1, ent, lane);
}

// s_in: the chunk in global memory (4-byte aligned, >= 8 readable bytes behind it), n <= 32768; ent: pass A's output; s_bm: (n + 31) / 32 words.
__device__ int snappy_compress_warp_chain(const uint8_t* s_in, int n, const uint32_t* __restrict__ ent, uint32_t* s_bm, uint8_t* out, int lane) {
    const uint32_t* in32 = (const uint32_t*)s_in;
    int op = 0;
    {   uint32_t v = (uint32_t)n; uint8_t pre[5]; int k = 0;
        while (v >= 0x80) { pre[k++] = (uint8_t)(v | 0x80); v >>= 7; } pre[k++] = (uint8_t)v;
        if (lane == 0) for (int i = 0; i < k; i++) out[i] = pre[i];
        op = k; }
    for (int i = lane; i < ((n + 31) >> 5); i += 32) s_bm[i] = (i == 0) ? 1u : 0u;       // position 0: what an empty table entry points at
    __syncwarp();
    int ip = 0; const int ip_end = n;
    if (n >= 15) {
        const int ip_limit = n - 15;
        bool have_prefix = false;
        for (;;) {
            const int next_emit = ip;
            const int start = ip + 1;                                   // `next_emit = ip++`
            const bool unrolled = ip_limit - start >= 16;
            int hit_ip = 0, candidate = 0; bool ended = false, immediate = false;
            if (lane < 4) LZ4C_PREFETCH(ent + start + 64 + 32 * lane);
            int q0 = start; uint32_t skip0 = 32; int a0 = 0;            // state of the first attempt of the current window
            for (bool first = true;; first = false) {
                const bool prefixed = first && have_prefix;
                // the first window of a search is contiguous (the first 32 attempts advance by one byte; lanes 0 and 1 of a prefixed window
                // are ip - 1 and ip): its positions and their "inserted by an earlier lane" status are arithmetic. Later windows are in
                // the accelerated regime: they mark their positions in the bitmap first and take the unused marks back (lz4_chain.cuh).
                const bool contiguous = first;
                const int w_lo = prefixed ? ip - 1 : start;
                int q; bool valid, putonly = false; int qn = 0; uint32_t skn = 0;
                if (contiguous) {
                    q = w_lo + lane;
                    const int l = lane - (prefixed ? 2 : 0);                // attempt number (prefix lanes: negative)
                    putonly = prefixed && lane == 0;
                    qn = q + 1; skn = 32u + (uint32_t)(l + 1);              // (attempts 0..31 step by one: skip_l = 32 + l)
                    valid = l < 0 || (unrolled && l < 16) || qn <= ip_limit;
                } else {
                    const int l = lane;
                    q = q0; uint32_t sk = skip0;
                    for (int t = 0; t < l; t++) { const uint32_t st = sk >> 5; q += (int)st; sk += st; }
                    const uint32_t st = sk >> 5; qn = q + (int)st; skn = sk + st;
                    valid = (unrolled && a0 + l < 16) || qn <= ip_limit;
                }
                const uint32_t inval = __ballot_sync(FULL_MASK, !valid);
                const int first_inv = inval ? (__ffs(inval) - 1) : 32;
                if (!contiguous) { if (valid) atomicOr(&s_bm[q >> 5], 1u << (q & 31)); __syncwarp(); }
                const uint32_t e = valid ? ent[q] : 0u;
                const int q1 = (int)(e & 0x7FFFu), q2 = (int)((e >> 16) & 0x7FFFu);
                const bool in1 = contiguous && q1 >= w_lo, in2 = contiguous && q2 >= w_lo;
                const bool ins1 = in1 || ((s_bm[q1 >> 5] >> (q1 & 31)) & 1u), ins2 = in2 || ((s_bm[q2 >> 5] >> (q2 & 31)) & 1u);
                int cand = ins1 ? q1 : q2;
                bool hit = valid && !putonly && (ins1 ? ((e >> 15) & 1u) : (ins2 ? (e >> 31) : 0u));
                const bool deeper = valid && !ins1 && !ins2;
                uint32_t hits = __ballot_sync(FULL_MASK, hit);
                const uint32_t dmask = __ballot_sync(FULL_MASK, deeper);
                if (dmask) {                                          // (see lz4_chain.cuh: only lanes in front of the first known hit walk)
                    int limit = hits ? (__ffs(hits) - 1) : 32; if (first_inv < limit) limit = first_inv;
                    if ((int)(__ffs(dmask) - 1) < limit) {
                        bool walking = deeper; int c = q2;                     // c: known not to be inserted; its entry names the next two
                        for (;;) {
                            walking = walking && lane < limit;
                            const uint32_t eq_ = walking ? ent[c] : 0u;
                            const int c1 = (int)(eq_ & 0x7FFFu), c2 = (int)((eq_ >> 16) & 0x7FFFu);
                            const bool i1 = walking && ((contiguous && c1 >= w_lo) || ((s_bm[c1 >> 5] >> (c1 & 31)) & 1u));
                            const bool i2 = walking && !i1 && ((contiguous && c2 >= w_lo) || ((s_bm[c2 >> 5] >> (c2 & 31)) & 1u));
                            const bool ins = i1 || i2;
                            if (ins) { walking = false; cand = i1 ? c1 : c2; hit = !putonly && (lz4_rd32<true>(in32, cand) == lz4_rd32<true>(in32, q)); }
                            else if (walking) c = c2;
                            const uint32_t nh = __ballot_sync(FULL_MASK, ins && hit);
                            if (nh && (int)(__ffs(nh) - 1) < limit) limit = __ffs(nh) - 1;
                            if (!__any_sync(FULL_MASK, walking && lane < limit)) break;
                        }
                        hits = __ballot_sync(FULL_MASK, hit);
                    }
                }
                const int first_hit = hits ? (__ffs(hits) - 1) : 32;
                const bool found = first_hit < first_inv;
                const int last_ins = found ? first_hit : (first_inv - 1);
                if (contiguous) {
                    if (last_ins >= 0 && lane < 2) {
                        const int w = (w_lo >> 5) + lane;
                        const uint32_t mbits = lz4c_range_bits(w_lo, w_lo + last_ins + 1, w);
                        if (mbits) s_bm[w] |= mbits;
                    }
                } else {
                    __syncwarp();
                    if (valid && lane > last_ins) atomicAnd(&s_bm[q >> 5], ~(1u << (q & 31)));
                }
                __syncwarp();
                if (found) {
                    hit_ip = contiguous ? w_lo + first_hit : __shfl_sync(FULL_MASK, q, first_hit);
                    candidate = __shfl_sync(FULL_MASK, cand, first_hit);
                    immediate = prefixed && first_hit == 1;
                    break;
                }
                if (first_inv < 32) { ended = true; break; }
                q0 = __shfl_sync(FULL_MASK, qn, 31); skip0 = __shfl_sync(FULL_MASK, skn, 31);                   // the attempt after lane 31's
                a0 += prefixed ? 30 : 32;
            }
            if (ended) { ip = next_emit; break; }
            ip = hit_ip;
            if (!immediate) op = snappy_emit_literal_warp<true>(out, op, s_in + next_emit, ip - next_emit, lane);
            // FindMatchLength(candidate + 4, ip + 4, ip_end): 32 bytes per step
            int matched = 4;
            for (;;) {
                const int i = matched + lane;
                const bool eq = (ip + i < ip_end) && (lz4_rd8<true>(s_in, candidate + i) == lz4_rd8<true>(s_in, ip + i));
                const uint32_t b = __ballot_sync(FULL_MASK, eq);
                if (b == FULL_MASK) { matched += 32; continue; }
                matched += __ffs(~b) - 1;
                break;
            }
            op = snappy_emit_copy_warp(out, op, ip - candidate, matched, (matched - 4) < 8, lane);
            ip += matched;
            if (ip >= ip_limit) break;                                  // emit_remainder from here
            have_prefix = true;
        }
    }
    if (ip < ip_end) op = snappy_emit_literal_warp<true>(out, op, s_in + ip, ip_end - ip, lane);
    return op;
}

} // namespace b200c
