// oracle/codec.h — TEST INFRASTRUCTURE (CPU oracle). Not part of the product path.
//
// CPU restatement of the byte-level codecs on Cassandra's SSTable compaction hot path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
// may link or call anything under oracle/. The product (cassandra_b200/csrc) never does.
//
// Every function cites the reference file:line it follows (S/ = src/java/org/apache/cassandra/).
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>
#include <string>

namespace oracle {

// ---- vint: S/utils/vint/VIntCoding.java:303-327 (write), :535-540 (size), :281-293 (prefix), :507-525 (zigzag)
int      vint_size(uint64_t v);
int      vint_write(uint8_t* out, uint64_t v);                 // returns bytes written
int      vint_read(const uint8_t* in, const uint8_t* end, uint64_t* v);   // returns bytes read, -1 on overrun
static inline uint64_t zigzag_enc(int64_t n) { return ((uint64_t)n << 1) ^ (uint64_t)(n >> 63); }
static inline int64_t  zigzag_dec(uint64_t n) { return (int64_t)(n >> 1) ^ -(int64_t)(n & 1); }

// ---- CRC32 (IEEE / zlib polynomial 0xEDB88320): java.util.zip.CRC32 as used by S/io/util/ChecksumWriter.java:34-37
uint32_t crc32_ieee(uint32_t crc, const uint8_t* p, size_t n);   // same contract as zlib crc32()
uint32_t crc32_combine(uint32_t crcA, uint32_t crcB, uint64_t lenB);

// ---- Murmur3 x64 128 (Cassandra variant, sign-extended tail): S/utils/MurmurHash.java:178-260
void     murmur3_x64_128(const uint8_t* key, size_t len, uint64_t seed, uint64_t out[2]);
// S/dht/Murmur3Partitioner.java:256-296: token = h1, MIN→MAX, empty key → MIN
int64_t  murmur3_token(const uint8_t* key, size_t len);

// ---- LZ4 block (third-party: org.lz4:lz4-java 1.8.0 → liblz4 1.9.x LZ4_compress_default, byU16 table for <64KiB+11)
// call site S/io/compress/LZ4Compressor.java:113-134.  Restated from the published algorithm (lz4.c, LZ4_compress_generic).
int      lz4_compress_bound(int n);
int      lz4_compress_block(const uint8_t* src, int n, uint8_t* dst, int cap);   // returns compressed size, 0 on failure
int      lz4_decompress_block(const uint8_t* src, int n, uint8_t* dst, int cap); // returns size, <0 on malformed input

// ---- Snappy raw format (third-party: org.xerial.snappy:snappy-java 1.1.10.4 → snappy 1.1.10). PARITY UNPINNED (no golden, no lib).
int      snappy_max_compressed_length(int n);
uint64_t murmur2_64(const uint8_t* key, int length, uint64_t seed);
int      snappy_compress(const uint8_t* src, int n, uint8_t* dst, int max_table_bits = 14);
int      snappy_uncompressed_length(const uint8_t* src, int n);
int      snappy_decompress(const uint8_t* src, int n, uint8_t* dst, int cap);

// ---- chunk codec as ICompressor sees it: S/io/compress/LZ4Compressor.java:113-190, SnappyCompressor.java:77-105
enum Compressor { COMP_LZ4 = 1, COMP_SNAPPY = 2, COMP_SNAPPY15 = 3, COMP_NONE = 0 };
int      chunk_max_compressed(int compressor, int chunk_len);
int      chunk_compress(int compressor, const uint8_t* src, int n, uint8_t* dst);
int      chunk_decompress(int compressor, const uint8_t* src, int n, uint8_t* dst, int cap);

} // namespace oracle
