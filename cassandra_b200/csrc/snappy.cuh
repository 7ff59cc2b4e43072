// snappy.cuh — Snappy raw-format codec for one chunk per warp (S/io/compress/SnappyCompressor.java:77-105).
// Third-party algorithm: Google snappy (snappy-java 1.1.10.4 bundles 1.1.10), portable multiply hash. The reference holds no
// Snappy-compressed fixture; byte parity is pinned against the library itself in its >= 1.2.0 generation (max_bits = 15,
// tests/golden/snappy), the 1.1.x generation (max_bits = 14) differing only in kMaxHashTableBits.
// Round-1 implementation: the greedy matcher runs on lane 0 (sequential), copies and decompression use the whole warp.
#pragma once
#include "common.cuh"

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

// one fragment (<= 64 KiB). s_in 4-byte aligned with >= 8 bytes of zeroed slack. Sequential; call from a single lane.
__device__ int snappy_compress_fragment_seq(const uint8_t* s_in, int base, int input_size, uint16_t* s_tab, int table_size, int max_bits, uint8_t* out, int op) {
    const uint32_t* in32 = (const uint32_t*)s_in;
    const uint32_t tmask = (uint32_t)table_size - 1;
    int ip = base; const int ip_end = base + input_size;
    if (input_size >= 15) {
        const int ip_limit = base + input_size - 15;
        for (;;) {
            int next_emit = ip++;
            uint32_t skip = 32;
            int candidate = 0; bool found = false;
            if (ip_limit - ip >= 16) {
                int delta = ip - base;
                for (int i = 0; i < 16; i++) {
                    uint32_t dword = rd32_at(in32, ip + i);
                    uint32_t e = snappy_tidx(dword, tmask, max_bits);
                    candidate = base + s_tab[e];
                    s_tab[e] = (uint16_t)(delta + i);
                    if (rd32_at(in32, candidate) == dword) {
                        out[op] = (uint8_t)(i << 2);
                        for (int k = 0; k <= i; k++) out[op + 1 + k] = s_in[next_emit + k];
                        ip += i; op += i + 2; found = true; break;
                    }
                }
                if (!found) { ip += 16; skip += 16; }
            }
            if (!found) {
                for (;;) {
                    uint32_t data = rd32_at(in32, ip);
                    uint32_t e = snappy_tidx(data, tmask, max_bits);
                    uint32_t between = skip >> 5; skip += between;
                    int next_ip = ip + (int)between;
                    if (next_ip > ip_limit) { ip = next_emit; goto emit_remainder; }
                    candidate = base + s_tab[e];
                    s_tab[e] = (uint16_t)(ip - base);
                    if (data == rd32_at(in32, candidate)) break;
                    ip = next_ip;
                }
                op = snappy_emit_literal(out, op, s_in + next_emit, ip - next_emit);
            }
            do {
                int b0 = ip; int matched = 4;
                while (ip + matched < ip_end && s_in[candidate + matched] == s_in[ip + matched]) matched++;
                bool lt12 = (matched - 4) < 8;
                ip += matched;
                op = snappy_emit_copy(out, op, b0 - candidate, matched, lt12);
                if (ip >= ip_limit) goto emit_remainder;
                s_tab[snappy_tidx(rd32_at(in32, ip - 1), tmask, max_bits)] = (uint16_t)(ip - base - 1);
                uint32_t e = snappy_tidx(rd32_at(in32, ip), tmask, max_bits);
                candidate = base + s_tab[e];
                s_tab[e] = (uint16_t)(ip - base);
            } while (rd32_at(in32, ip) == rd32_at(in32, candidate));
        }
    }
emit_remainder:
    if (ip < ip_end) op = snappy_emit_literal(out, op, s_in + ip, ip_end - ip);
    return op;
}

// s_tab: (1 << max_bits) x u16. Returns compressed size (warp-uniform).
__device__ int snappy_compress_warp(const uint8_t* s_in, int n, uint16_t* s_tab, int max_bits, uint8_t* out, int lane) {
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
        if (lane == 0) op = snappy_compress_fragment_seq(s_in, pos, frag, s_tab, table_size, max_bits, out, op);
        op = __shfl_sync(FULL_MASK, op, 0);
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
