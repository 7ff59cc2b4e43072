// oracle/capi_codec.cc — TEST INFRASTRUCTURE. extern "C" shims over oracle/codec.h for ctypes (tests, bench cpu_baseline).
#include "codec.h"
using namespace oracle;
extern "C" {
int      orc_vint_size(uint64_t v) { return vint_size(v); }
int      orc_vint_write(uint8_t* out, uint64_t v) { return vint_write(out, v); }
int      orc_vint_read(const uint8_t* in, int n, uint64_t* v) { return vint_read(in, in + n, v); }
uint32_t orc_crc32(uint32_t crc, const uint8_t* p, uint64_t n) { return crc32_ieee(crc, p, n); }
uint32_t orc_crc32_combine(uint32_t a, uint32_t b, uint64_t lenB) { return crc32_combine(a, b, lenB); }
int64_t  orc_murmur3_token(const uint8_t* k, uint64_t n) { return murmur3_token(k, n); }
void     orc_murmur3_x64_128(const uint8_t* k, uint64_t n, uint64_t seed, uint64_t* out) { murmur3_x64_128(k, n, seed, out); }
int      orc_lz4_compress_bound(int n) { return lz4_compress_bound(n); }
int      orc_lz4_compress_block(const uint8_t* s, int n, uint8_t* d, int cap) { return lz4_compress_block(s, n, d, cap); }
int      orc_lz4_decompress_block(const uint8_t* s, int n, uint8_t* d, int cap) { return lz4_decompress_block(s, n, d, cap); }
int      orc_snappy_max_compressed_length(int n) { return snappy_max_compressed_length(n); }
int      orc_snappy_compress(const uint8_t* s, int n, uint8_t* d) { return snappy_compress(s, n, d, 14); }
int      orc_snappy_decompress(const uint8_t* s, int n, uint8_t* d, int cap) { return snappy_decompress(s, n, d, cap); }
int      orc_chunk_max_compressed(int c, int n) { return chunk_max_compressed(c, n); }
int      orc_chunk_compress(int c, const uint8_t* s, int n, uint8_t* d) { return chunk_compress(c, s, n, d); }
int      orc_chunk_decompress(int c, const uint8_t* s, int n, uint8_t* d, int cap) { return chunk_decompress(c, s, n, d, cap); }
}

// Whole-stream CompressedSequentialWriter on the CPU (S/io/compress/CompressedSequentialWriter.java:140-206 with min_compress_ratio 0):
// chunk + big-endian CRC32 per chunk, chunk offsets, returns the image length. out_cap >= nchunks * (chunk_max_compressed + 4).
extern "C" uint64_t orc_compress_stream(int comp, const uint8_t* in, uint64_t n, int chunk_len, uint8_t* out, uint64_t* offs) {
    uint64_t o = 0, k = 0;
    for (uint64_t i = 0; i < n; i += chunk_len, k++) {
        int len = (int)((n - i) < (uint64_t)chunk_len ? (n - i) : (uint64_t)chunk_len);
        int c = chunk_compress(comp, in + i, len, out + o);
        uint32_t crc = crc32_ieee(0, out + o, c);
        offs[k] = o; o += c;
        out[o] = (uint8_t)(crc >> 24); out[o + 1] = (uint8_t)(crc >> 16); out[o + 2] = (uint8_t)(crc >> 8); out[o + 3] = (uint8_t)crc; o += 4;
    }
    return o;
}
