// snappy.cuh — Snappy raw-format codec for one chunk per warp (S/io/compress/SnappyCompressor.java:77-105).
// Third-party algorithm: Google snappy (snappy-java 1.1.10.4 bundles 1.1.10), portable multiply hash. The reference holds no
// Snappy-compressed fixture; byte parity is pinned against the library itself in its >= 1.2.0 generation (max_bits = 15,
// tests/golden/snappy), the 1.1.x generation (max_bits = 14) differing only in kMaxHashTableBits.
#pragma once
#include "common.cuh"
#include "lz4.cuh"          // lz4_rd32 / lz4_rd8: chunk bytes from a shared-memory copy or, GLOBAL, through the read-only L1 path

namespace b200c {

__host__ __device__ __forceinline__ int snappy_max_compressed_length(int n) { return 32 + n + n / 6; }

__device__ __forceinline__ uint32_t snappy_tidx(uint32_t bytes, uint32_t tmask, int max_bits) { return ((0x1e35a7bdu * bytes) >> (32 - max_bits)) & tmask; }

__device__ __forceinline__ int snappy_emit_literal(uint8_t* out, int op, const uint8_t* lit, int len) {
    int n = len - 1;
    if (n < 60) out[op++] = (uint8_t)(n << 2);
    else {
        int count = ((31 - __clz(n)) >> 3) + 1;
        out[op++] = (uint8_t)((59 + count) << 2);
        for (int i = 0; i < count; i++) out[op++] = (uint8_t)(n >> (8 * i));
    }
    for (int i = 0; i < len; i++) out[op + i] = lit[i];
    return op + len;
}
__device__ __forceinline__ int snappy_emit_copy64(uint8_t* out, int op, int offset, int len, bool lt12) {
    if (lt12 && offset < 2048) {
        out[op++] = (uint8_t)(1 + ((len - 4) << 2) + ((offset >> 3) & 0xe0));
        out[op++] = (uint8_t)(offset & 0xff);
    } else {
        uint32_t u = 2u + (uint32_t)((len - 1) << 2) + ((uint32_t)offset << 8);
        out[op] = (uint8_t)u; out[op + 1] = (uint8_t)(u >> 8); out[op + 2] = (uint8_t)(u >> 16);
        op += 3;
    }
    return op;
}
__device__ __forceinline__ int snappy_emit_copy(uint8_t* out, int op, int offset, int len, bool lt12) {
    if (lt12) return snappy_emit_copy64(out, op, offset, len, true);
    while (len >= 68) { op = snappy_emit_copy64(out, op, offset, 64, false); len -= 64; }
    if (len > 64) { op = snappy_emit_copy64(out, op, offset, 60, false); len -= 60; }
    return snappy_emit_copy64(out, op, offset, len, len < 12);
}

// EmitLiteral, warp wide: tag (+ length bytes) by lane 0, the bytes 32 per step. Returns the new output position (warp-uniform).
template <bool GLOBAL = false> __device__ __forceinline__ int snappy_emit_literal_warp(uint8_t* out, int op, const uint8_t* lit, int len, int lane) {
    const int n = len - 1;
    int hdr = 1;
    if (n < 60) { if (lane == 0) out[op] = (uint8_t)(n << 2); }
    else {
        const int count = ((31 - __clz(n)) >> 3) + 1; hdr = 1 + count;
        if (lane == 0) { out[op] = (uint8_t)((59 + count) << 2); for (int i = 0; i < count; i++) out[op + 1 + i] = (uint8_t)(n >> (8 * i)); }
    }
    for (int i = lane; i < len; i += 32) out[op + hdr + i] = (uint8_t)lz4_rd8<GLOBAL>(lit, i);
    return op + hdr + len;
}
// EmitCopy: 64-byte copies while len >= 68 (3 bytes each, one per lane), then one or two short ones by lane 0
__device__ __forceinline__ int snappy_emit_copy_warp(uint8_t* out, int op, int offset, int len, bool lt12, int lane) {
    if (lt12) { if (lane == 0) snappy_emit_copy64(out, op, offset, len, true); return op + ((offset < 2048) ? 2 : 3); }
    int nfull = 0;
    if (len >= 68) { nfull = (len - 68) / 64 + 1; for (int i = lane; i < nfull; i += 32) snappy_emit_copy64(out, op + 3 * i, offset, 64, false); op += 3 * nfull; len -= 64 * nfull; }
    int end = op;
    if (len > 64) { if (lane == 0) snappy_emit_copy64(out, end, offset, 60, false); end += 3; len -= 60; }
    const bool l12 = len < 12;
    if (lane == 0) snappy_emit_copy64(out, end, offset, len, l12);
    return end + ((l12 && offset < 2048) ? 2 : 3);
}

// One fragment (<= 64 KiB), whole warp. CompressFragment's greedy matcher is sequential by definition (what position i finds in the hash table
// depends on every earlier lookup-and-insert); as in lz4.cuh the warp SPECULATES 32 consecutive search attempts at once — hash, table probe and
// 4-byte compare in parallel, lanes with equal hashes resolved in attempt order with per-bit ballots — and takes the first hit, which is what
// the sequential loop would have found. After a match the library (1) inserts position ip - 1, (2) tests position ip (lookup + insert: a hit is
// an immediate copy with no literal), (3) resumes the search at ip + 1: those three ride in the first window as lane 0 (insert only), lane 1
// (the test) and lanes 2.. (search attempts 0..29), lane order being the sequential order.
// Attempt k of a search looks at q_k with q_0 = start, q_(k+1) = q_k + (skip_k >> 5), skip_(k+1) = skip_k + (skip_k >> 5), skip_0 = 32; the first 16
// attempts are unconditional when at least 16 bytes remain before ip_limit (the library's unrolled prologue), every other attempt ends the
// fragment when its successor would pass ip_limit.
template <bool GLOBAL = false> __device__ int snappy_compress_fragment_warp(const uint8_t* s_in, int base, int input_size, uint16_t* s_tab, int table_size, int max_bits, uint8_t* out, int op, int lane) {
    const uint32_t* in32 = (const uint32_t*)s_in;
    const uint32_t tmask = (uint32_t)table_size - 1;
    const uint32_t lt_mask = (1u << lane) - 1u;
    int ip = base; const int ip_end = base + input_size;
    if (input_size >= 15) {
        const int ip_limit = base + input_size - 15;
        bool have_prefix = false;
        for (;;) {
            const int next_emit = ip;
            const int start = ip + 1;                                   // `next_emit = ip++`
            const bool unrolled = ip_limit - start >= 16;
            int hit_ip = 0, candidate = 0; bool ended = false, immediate = false;
            int q0 = start; uint32_t skip0 = 32; int a0 = 0;            // state of the first attempt of the current window
            for (bool first = true;; first = false) {
                const bool prefixed = first && have_prefix;
                int q; bool valid, putonly = false; int qn = 0; uint32_t skn = 0;
                if (prefixed && lane < 2) { q = lane == 0 ? ip - 1 : ip; valid = true; putonly = lane == 0; }
                else {
                    const int l = lane - (prefixed ? 2 : 0);
                    q = q0; uint32_t sk = skip0;
                    for (int t = 0; t < l; t++) { const uint32_t st = sk >> 5; q += (int)st; sk += st; }       // (one add per earlier attempt; 31 at most)
                    const uint32_t st = sk >> 5; qn = q + (int)st; skn = sk + st;
                    valid = (unrolled && a0 + l < 16) || qn <= ip_limit;
                }
                const uint32_t dword = valid ? lz4_rd32<GLOBAL>(in32, q) : 0u;
                const uint32_t h = snappy_tidx(dword, tmask, max_bits);
                int cand = valid ? base + (int)s_tab[h] : 0;
                uint32_t same = FULL_MASK;
                for (int b = 0; b < max_bits; b++) { const uint32_t mb = __ballot_sync(FULL_MASK, (h >> b) & 1u); same &= ((h >> b) & 1u) ? mb : ~mb; }
                const uint32_t prev = same & lt_mask;
                const int src = prev ? (31 - __clz(prev)) : lane;
                const int pq = __shfl_sync(FULL_MASK, q, src);
                if (prev) cand = pq;
                const bool hit = valid && !putonly && lz4_rd32<GLOBAL>(in32, cand) == dword;
                const uint32_t hits = __ballot_sync(FULL_MASK, hit);
                const uint32_t inval = __ballot_sync(FULL_MASK, !valid);
                const int first_hit = hits ? (__ffs(hits) - 1) : 32;
                const int first_inv = inval ? (__ffs(inval) - 1) : 32;
                if (first_hit < first_inv) {
                    const uint32_t le = (first_hit == 31) ? FULL_MASK : ((2u << first_hit) - 1u);
                    const uint32_t later_same = same & le & ~lt_mask & ~(1u << lane);
                    if (lane <= first_hit && !later_same) s_tab[h] = (uint16_t)(q - base);
                    hit_ip = __shfl_sync(FULL_MASK, q, first_hit);
                    candidate = __shfl_sync(FULL_MASK, cand, first_hit);
                    immediate = prefixed && first_hit == 1;
                    break;
                }
                if (first_inv < 32) { ended = true; break; }
                { const uint32_t later_same = same & ~lt_mask & ~(1u << lane); if (!later_same) s_tab[h] = (uint16_t)(q - base); }
                __syncwarp();
                q0 = __shfl_sync(FULL_MASK, qn, 31); skip0 = __shfl_sync(FULL_MASK, skn, 31);                   // the attempt after lane 31's
                a0 += prefixed ? 30 : 32;
            }
            if (ended) { ip = next_emit; break; }
            __syncwarp();
            ip = hit_ip;
            if (!immediate) op = snappy_emit_literal_warp<GLOBAL>(out, op, s_in + next_emit, ip - next_emit, lane);
            // FindMatchLength(candidate + 4, ip + 4, ip_end): 32 bytes per step
            int matched = 4;
            for (;;) {
                const int i = matched + lane;
                const bool eq = (ip + i < ip_end) && (lz4_rd8<GLOBAL>(s_in, candidate + i) == lz4_rd8<GLOBAL>(s_in, ip + i));
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
    if (ip < ip_end) op = snappy_emit_literal_warp<GLOBAL>(out, op, s_in + ip, ip_end - ip, lane);
    return op;
}

// s_tab: (1 << max_bits) x u16. Returns compressed size (warp-uniform).
// GLOBAL: s_in is the chunk where it lies in global memory (4-byte aligned, >= 8 readable bytes behind it), see lz4_compress_warp
template <bool GLOBAL = false> __device__ int snappy_compress_warp(const uint8_t* s_in, int n, uint16_t* s_tab, int max_bits, uint8_t* out, int lane) {
    int op = 0;
    {   uint32_t v = (uint32_t)n; uint8_t pre[5]; int k = 0;
        while (v >= 0x80) { pre[k++] = (uint8_t)(v | 0x80); v >>= 7; } pre[k++] = (uint8_t)v;
        if (lane == 0) for (int i = 0; i < k; i++) out[i] = pre[i];
        op = k; }
    for (int pos = 0; pos < n; pos += 65536) {
        int frag = min(n - pos, 65536);
        int table_size = frag > (1 << max_bits) ? (1 << max_bits) : (frag < 256 ? 256 : (2 << (31 - __clz(frag - 1))));
        for (int i = lane; i < table_size / 2; i += 32) ((uint32_t*)s_tab)[i] = 0;
        __syncwarp();
        op = snappy_compress_fragment_warp<GLOBAL>(s_in, pos, frag, s_tab, table_size, max_bits, out, op, lane);
        __syncwarp();
    }
    return op;
}

// Returns decoded size or -1. Literal/copy bodies are spread across the warp.
__device__ int snappy_decompress_warp(const uint8_t* __restrict__ src, int n, uint8_t* s_out, int cap, int lane) {
    int ip = 0; uint32_t ulen = 0; int shift = 0; bool ok = false;
    while (ip < n && shift < 35) { uint32_t b = src[ip++]; ulen |= (b & 0x7f) << shift; if (!(b & 0x80)) { ok = true; break; } shift += 7; }
    if (!ok || ulen > (uint32_t)cap) return -1;
    int op = 0; const int oend = (int)ulen;
    while (ip < n) {
        uint32_t tag = src[ip++];
        int len, offset;
        if ((tag & 3) == 0) {
            len = (int)(tag >> 2) + 1;
            if (len > 60) { int cnt = len - 60; if (n - ip < cnt) return -1; uint32_t l = 0; for (int i = 0; i < cnt; i++) l |= (uint32_t)src[ip + i] << (8 * i); ip += cnt; if (l >= (1u << 24)) return -1; len = (int)l + 1; }
            if (n - ip < len || oend - op < len) return -1;
            for (int i = lane; i < len; i += 32) s_out[op + i] = src[ip + i];
            op += len; ip += len;
            continue;
        } else if ((tag & 3) == 1) {
            if (n - ip < 1) return -1; len = (int)((tag >> 2) & 7) + 4; offset = (int)((tag >> 5) << 8) | src[ip]; ip += 1;
        } else if ((tag & 3) == 2) {
            if (n - ip < 2) return -1; len = (int)(tag >> 2) + 1; offset = (int)src[ip] | ((int)src[ip + 1] << 8); ip += 2;
        } else {
            if (n - ip < 4) return -1; len = (int)(tag >> 2) + 1;
            uint32_t o = (uint32_t)src[ip] | ((uint32_t)src[ip + 1] << 8) | ((uint32_t)src[ip + 2] << 16) | ((uint32_t)src[ip + 3] << 24); ip += 4;
            if (o > (uint32_t)op) return -1; offset = (int)o;
        }
        if (offset == 0 || offset > op || oend - op < len) return -1;
        __syncwarp();
        const uint8_t* m = s_out + op - offset;
        for (int i = lane; i < len; i += 32) { int j = i; if (j >= offset) j %= offset; s_out[op + i] = m[j]; }
        op += len;
        __syncwarp();
    }
    __syncwarp();
    return op == oend ? op : -1;
}

} // namespace b200c
