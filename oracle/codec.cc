// oracle/codec.cc — TEST INFRASTRUCTURE (CPU oracle). See codec.h for the rules on who may call this.
//
// Byte-level codecs of the SSTable compaction hot path, restated for the CPU.
// S/ = /root/reference/src/java/org/apache/cassandra/
#include "codec.h"
#include <cstring>
#include <cstdlib>

namespace oracle {

// ------------------------------------------------------------------------------------------------
// vint — S/utils/vint/VIntCoding.java
// ------------------------------------------------------------------------------------------------

// computeUnsignedVIntSize, VIntCoding.java:535-540: (639 - 9*nlz(v|1)) >> 6
int vint_size(uint64_t v) {
    int magnitude = __builtin_clzll(v | 1);
    return (639 - magnitude * 9) >> 6;
}

// writeUnsignedVInt, VIntCoding.java:303-327: size-1 leading one bits, then the value big-endian; 9-byte form = 0xFF + 8 raw bytes
int vint_write(uint8_t* out, uint64_t v) {
    int size = vint_size(v);
    if (size == 1) { out[0] = (uint8_t)v; return 1; }
    if (size < 9) {
        int shift = (8 - size) << 3;
        int extra = size - 1;
        uint64_t mask = (uint64_t)(uint8_t)(~(0xff >> extra)) << 56;   // encodeExtraBytesToRead, :286-290
        uint64_t reg = (v << shift) | mask;
        for (int i = 0; i < size; i++) out[i] = (uint8_t)(reg >> (56 - 8 * i));
        return size;
    }
    out[0] = 0xFF;
    for (int i = 0; i < 8; i++) out[1 + i] = (uint8_t)(v >> (56 - 8 * i));
    return 9;
}

// readUnsignedVInt, VIntCoding.java:77-117 + numberOfExtraBytesToRead :292-298
int vint_read(const uint8_t* in, const uint8_t* end, uint64_t* v) {
    if (in >= end) return -1;
    uint8_t first = in[0];
    if (first < 0x80) { *v = first; return 1; }
    int extra = (first == 0xFF) ? 8 : __builtin_clz((uint32_t)(uint8_t)~first) - 24;   // count of leading 1 bits
    if (in + 1 + extra > end) return -1;
    uint64_t r = first & (0xff >> extra);
    for (int i = 0; i < extra; i++) r = (r << 8) | in[1 + i];
    *v = r;
    return 1 + extra;
}

// ------------------------------------------------------------------------------------------------
// CRC32 IEEE — java.util.zip.CRC32 (S/io/util/ChecksumWriter.java:34-37, S/utils/ChecksumType.java)
// ------------------------------------------------------------------------------------------------
static uint32_t g_crc_tab[8][256];
static bool g_crc_init = false;
static void crc_init() {
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
        g_crc_tab[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
        for (int t = 1; t < 8; t++)
            g_crc_tab[t][i] = (g_crc_tab[t - 1][i] >> 8) ^ g_crc_tab[0][g_crc_tab[t - 1][i] & 0xff];
    g_crc_init = true;
}
static struct CrcInit { CrcInit() { crc_init(); } } g_crc_initializer;

uint32_t crc32_ieee(uint32_t crc, const uint8_t* p, size_t n) {
    if (!g_crc_init) crc_init();
    uint32_t c = ~crc;
    while (n >= 8) {
        uint32_t a, b;
        memcpy(&a, p, 4); memcpy(&b, p + 4, 4);
        a ^= c;
        c = g_crc_tab[7][a & 0xff] ^ g_crc_tab[6][(a >> 8) & 0xff] ^ g_crc_tab[5][(a >> 16) & 0xff] ^ g_crc_tab[4][a >> 24] ^
            g_crc_tab[3][b & 0xff] ^ g_crc_tab[2][(b >> 8) & 0xff] ^ g_crc_tab[1][(b >> 16) & 0xff] ^ g_crc_tab[0][b >> 24];
        p += 8; n -= 8;
    }
    while (n--) c = g_crc_tab[0][(c ^ *p++) & 0xff] ^ (c >> 8);
    return ~c;
}

// GF(2) polynomial multiply mod P (reflected representation), used to append zero bytes to a CRC.
static uint32_t gf2_mulmod(uint32_t a, uint32_t b) {
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) {
        if (a & 0x80000000u) r ^= b;       // reflected: bit31 is x^0
        a <<= 1;
        b = (b & 1) ? (b >> 1) ^ 0xEDB88320u : (b >> 1);
    }
    return r;
}
// x^(8*n) mod P in reflected form
static uint32_t gf2_xpow8n(uint64_t nbytes) {
    uint32_t result = 0x80000000u;         // x^0
    uint32_t sq = 0x00800000u;             // x^8 (reflected: bit (31-8))
    while (nbytes) {
        if (nbytes & 1) result = gf2_mulmod(result, sq);
        sq = gf2_mulmod(sq, sq);
        nbytes >>= 1;
    }
    return result;
}
// crc(A||B) from crc(A), crc(B), len(B): same contract as zlib crc32_combine
uint32_t crc32_combine(uint32_t crcA, uint32_t crcB, uint64_t lenB) {
    return gf2_mulmod(crcA, gf2_xpow8n(lenB)) ^ crcB;
}

// ------------------------------------------------------------------------------------------------
// Murmur3 — S/utils/MurmurHash.java:152-260
// ------------------------------------------------------------------------------------------------
static inline uint64_t rotl64(uint64_t v, int n) { return (v << n) | (v >> (64 - n)); }
static inline uint64_t fmix(uint64_t k) {   // MurmurHash.java:167-176
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k;
}
void murmur3_x64_128(const uint8_t* key, size_t len, uint64_t seed, uint64_t out[2]) {
    const size_t nblocks = len >> 4;
    uint64_t h1 = seed, h2 = seed;
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    for (size_t i = 0; i < nblocks; i++) {          // body: getBlock is unsigned little-endian, :152-160
        uint64_t k1 = 0, k2 = 0;
        for (int b = 0; b < 8; b++) { k1 |= (uint64_t)key[i * 16 + b] << (8 * b); k2 |= (uint64_t)key[i * 16 + 8 + b] << (8 * b); }
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const uint8_t* t = key + nblocks * 16;
    uint64_t k1 = 0, k2 = 0;
    // tail: "(long) key.get(i)" sign-extends each byte before the shift, :214-233
    #define SB(i) ((uint64_t)(int64_t)(int8_t)t[i])
    switch (len & 15) {
        case 15: k2 ^= SB(14) << 48; /* fallthrough */
        case 14: k2 ^= SB(13) << 40; /* fallthrough */
        case 13: k2 ^= SB(12) << 32; /* fallthrough */
        case 12: k2 ^= SB(11) << 24; /* fallthrough */
        case 11: k2 ^= SB(10) << 16; /* fallthrough */
        case 10: k2 ^= SB(9) << 8;   /* fallthrough */
        case 9:  k2 ^= SB(8) << 0;
                 k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; /* fallthrough */
        case 8:  k1 ^= SB(7) << 56; /* fallthrough */
        case 7:  k1 ^= SB(6) << 48; /* fallthrough */
        case 6:  k1 ^= SB(5) << 40; /* fallthrough */
        case 5:  k1 ^= SB(4) << 32; /* fallthrough */
        case 4:  k1 ^= SB(3) << 24; /* fallthrough */
        case 3:  k1 ^= SB(2) << 16; /* fallthrough */
        case 2:  k1 ^= SB(1) << 8;  /* fallthrough */
        case 1:  k1 ^= SB(0);
                 k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    }
    #undef SB
    h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
    h1 += h2; h2 += h1;
    h1 = fmix(h1); h2 = fmix(h2);
    h1 += h2; h2 += h1;
    out[0] = h1; out[1] = h2;
}

// Murmur3Partitioner.getToken, S/dht/Murmur3Partitioner.java:256-296
int64_t murmur3_token(const uint8_t* key, size_t len) {
    if (len == 0) return INT64_MIN;                     // MINIMUM
    uint64_t h[2];
    murmur3_x64_128(key, len, 0, h);
    int64_t v = (int64_t)h[0];
    return v == INT64_MIN ? INT64_MAX : v;              // normalize :291-295
}

// ------------------------------------------------------------------------------------------------
// LZ4 block format. Third-party arithmetic: lz4-java 1.8.0's JNI binding calls liblz4's LZ4_compress_default
// (fast mode, acceleration 1). For inputs < 64 KiB + 11 liblz4 uses the 16-bit-index hash table ("byU16", 13-bit hash).
// This restates LZ4_compress_generic for {notLimited, byU16, noDict, noDictIssue, acceleration=1}.
// Pinned by tests/test_oracle_codec.py against the reference's golden SSTables and against the system liblz4.so.1.
// ------------------------------------------------------------------------------------------------
enum { MINMATCH = 4, MFLIMIT = 12, LASTLITERALS = 5, LZ4_MINLENGTH = MFLIMIT + 1, ML_BITS = 4, ML_MASK = 15, RUN_MASK = 15,
       LZ4_HASHLOG_U16 = 13, LZ4_SKIPTRIGGER = 6, LZ4_64KLIMIT = 65536 + (MFLIMIT - 1) };

int lz4_compress_bound(int n) { return n > 0x7E000000 ? 0 : n + n / 255 + 16; }

static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint32_t lz4_hash_u16(uint32_t seq) { return (seq * 2654435761u) >> (32 - LZ4_HASHLOG_U16); }

int lz4_compress_block(const uint8_t* src, int n, uint8_t* dst, int cap) {
    if (n < 0 || n >= LZ4_64KLIMIT) return 0;          // this oracle covers the byU16 regime only (chunk_length <= 64 KiB)
    if (cap < lz4_compress_bound(n)) return 0;
    uint16_t table[1 << LZ4_HASHLOG_U16];
    memset(table, 0, sizeof(table));
    const uint8_t* ip = src;
    const uint8_t* anchor = src;
    const uint8_t* const iend = src + n;
    const uint8_t* const mflimitPlusOne = iend - MFLIMIT + 1;
    const uint8_t* const matchlimit = iend - LASTLITERALS;
    uint8_t* op = dst;
    uint32_t forwardH;

    if (n < LZ4_MINLENGTH) goto last_literals;

    table[lz4_hash_u16(rd32(ip))] = (uint16_t)(ip - src);     // first byte
    ip++; forwardH = lz4_hash_u16(rd32(ip));

    for (;;) {
        const uint8_t* match;
        uint8_t* token;
        {   // find a match
            const uint8_t* forwardIp = ip;
            int step = 1;
            int searchMatchNb = 1 << LZ4_SKIPTRIGGER;
            do {
                uint32_t h = forwardH;
                uint32_t current = (uint32_t)(forwardIp - src);
                uint32_t matchIndex = table[h];
                ip = forwardIp;
                forwardIp += step;
                step = (searchMatchNb++ >> LZ4_SKIPTRIGGER);
                if (forwardIp > mflimitPlusOne) goto last_literals;
                match = src + matchIndex;
                forwardH = lz4_hash_u16(rd32(forwardIp));
                table[h] = (uint16_t)current;
            } while (rd32(match) != rd32(ip));
        }
        // catch up
        while ((ip > anchor) && (match > src) && (ip[-1] == match[-1])) { ip--; match--; }
        {   // encode literal length + literals
            unsigned litLength = (unsigned)(ip - anchor);
            token = op++;
            if (litLength >= RUN_MASK) {
                int len = (int)(litLength - RUN_MASK);
                *token = (RUN_MASK << ML_BITS);
                for (; len >= 255; len -= 255) *op++ = 255;
                *op++ = (uint8_t)len;
            } else *token = (uint8_t)(litLength << ML_BITS);
            memcpy(op, anchor, litLength);
            op += litLength;
        }
    next_match:
        // offset (little-endian)
        { uint16_t off = (uint16_t)(ip - match); op[0] = (uint8_t)off; op[1] = (uint8_t)(off >> 8); op += 2; }
        {   // match length
            const uint8_t* pi = ip + MINMATCH; const uint8_t* pm = match + MINMATCH;
            unsigned matchCode = 0;
            while (pi + matchCode < matchlimit && pi[matchCode] == pm[matchCode]) matchCode++;   // == LZ4_count
            ip += matchCode + MINMATCH;
            if (matchCode >= ML_MASK) {
                *token += ML_MASK;
                matchCode -= ML_MASK;
                while (matchCode >= 255) { *op++ = 255; matchCode -= 255; }
                *op++ = (uint8_t)matchCode;
            } else *token += (uint8_t)matchCode;
        }
        anchor = ip;
        if (ip >= mflimitPlusOne) break;
        table[lz4_hash_u16(rd32(ip - 2))] = (uint16_t)(ip - 2 - src);    // fill table
        {   // test next position
            uint32_t h = lz4_hash_u16(rd32(ip));
            uint32_t current = (uint32_t)(ip - src);
            uint32_t matchIndex = table[h];
            match = src + matchIndex;
            table[h] = (uint16_t)current;
            if (rd32(match) == rd32(ip)) { token = op++; *token = 0; goto next_match; }
        }
        forwardH = lz4_hash_u16(rd32(++ip));
    }
last_literals:
    {
        size_t lastRun = (size_t)(iend - anchor);
        if (lastRun >= RUN_MASK) {
            size_t acc = lastRun - RUN_MASK;
            *op++ = RUN_MASK << ML_BITS;
            for (; acc >= 255; acc -= 255) *op++ = 255;
            *op++ = (uint8_t)acc;
        } else *op++ = (uint8_t)(lastRun << ML_BITS);
        memcpy(op, anchor, lastRun);
        op += lastRun;
    }
    return (int)(op - dst);
}

// LZ4_decompress_safe semantics: exact-size decode with full bounds checks (lz4-java safeDecompressor,
// S/io/compress/LZ4Compressor.java:136-190).
int lz4_decompress_block(const uint8_t* src, int n, uint8_t* dst, int cap) {
    const uint8_t* ip = src; const uint8_t* const iend = src + n;
    uint8_t* op = dst; uint8_t* const oend = dst + cap;
    if (n == 0) return (cap == 0) ? 0 : -1;
    for (;;) {
        if (ip >= iend) return -1;
        unsigned token = *ip++;
        size_t len = token >> ML_BITS;
        if (len == RUN_MASK) { unsigned s; do { if (ip >= iend) return -1; s = *ip++; len += s; } while (s == 255); }
        if ((size_t)(iend - ip) < len || (size_t)(oend - op) < len) return -1;
        memcpy(op, ip, len); op += len; ip += len;
        if (ip == iend) break;                                   // last sequence has no match part
        if (iend - ip < 2) return -1;
        size_t offset = ip[0] | ((size_t)ip[1] << 8); ip += 2;
        if (offset == 0 || offset > (size_t)(op - dst)) return -1;
        size_t ml = token & ML_MASK;
        if (ml == ML_MASK) { unsigned s; do { if (ip >= iend) return -1; s = *ip++; ml += s; } while (s == 255); }
        ml += MINMATCH;
        if ((size_t)(oend - op) < ml) return -1;
        const uint8_t* m = op - offset;
        for (size_t i = 0; i < ml; i++) op[i] = m[i];            // byte-wise: overlapping copy repeats the pattern
        op += ml;
    }
    return (int)(op - dst);
}

// ------------------------------------------------------------------------------------------------
// Snappy raw format. Third-party arithmetic: snappy-java 1.1.10.4 → snappy 1.1.10 (portable multiply hash); a restatement of the
// published CompressFragment / EmitLiteral / EmitCopy / FindMatchLength algorithm. The reference holds no Snappy-compressed fixture
// (SURVEY §8c), so parity is pinned against Google's library itself: with kMaxHashTableBits = 15 (snappy >= 1.2.0, the generation
// bundled in this image's pyarrow) this code reproduces the library byte for byte on tests/golden/snappy/vectors.json and in a live
// differential test (tests/test_snappy_golden.py); the 1.1.10 generation differs only in that constant (14, snappy.h of 1.1.10).
// Call site S/io/compress/SnappyCompressor.java:77-105.
// ------------------------------------------------------------------------------------------------
// MurmurHash.hash2_64 (S/utils/MurmurHash.java:94-152): MurmurHash2 64-bit; the tail bytes are SIGN-extended (Java bytes), the block bytes are not
uint64_t murmur2_64(const uint8_t* key, int length, uint64_t seed) {
    const uint64_t m64 = 0xc6a4a7935bd1e995ull; const int r64 = 47;
    uint64_t h64 = (seed & 0xffffffffull) ^ (m64 * (uint64_t)length);
    int lenLongs = length >> 3;
    for (int i = 0; i < lenLongs; i++) {
        uint64_t k64 = 0; for (int b = 0; b < 8; b++) k64 += (uint64_t)key[i * 8 + b] << (8 * b);
        k64 *= m64; k64 ^= k64 >> r64; k64 *= m64;
        h64 ^= k64; h64 *= m64;
    }
    int rem = length & 7; const uint8_t* t = key + length - rem;
    for (int i = rem - 1; i >= 1; i--) h64 ^= (uint64_t)(int64_t)(int8_t)t[i] << (8 * i);
    if (rem >= 1) { h64 ^= (uint64_t)(int64_t)(int8_t)t[0]; h64 *= m64; }
    h64 ^= h64 >> r64; h64 *= m64; h64 ^= h64 >> r64;
    return h64;
}

int snappy_max_compressed_length(int n) { return 32 + n + n / 6; }

static inline uint8_t* snappy_emit_literal(uint8_t* op, const uint8_t* lit, size_t len) {
    size_t n = len - 1;
    if (n < 60) { *op++ = (uint8_t)(n << 2); }
    else {
        int count = ((31 - __builtin_clz((uint32_t)n)) >> 3) + 1;
        *op++ = (uint8_t)((59 + count) << 2);
        for (int i = 0; i < count; i++) *op++ = (uint8_t)(n >> (8 * i));
    }
    memcpy(op, lit, len);
    return op + len;
}
static inline uint8_t* snappy_emit_copy_upto64(uint8_t* op, size_t offset, size_t len, bool lt12) {
    if (lt12 && offset < 2048) {
        *op++ = (uint8_t)(1 + ((len - 4) << 2) + ((offset >> 3) & 0xe0));
        *op++ = (uint8_t)(offset & 0xff);
    } else {
        uint32_t u = 2 + (uint32_t)((len - 1) << 2) + (uint32_t)(offset << 8);
        op[0] = (uint8_t)u; op[1] = (uint8_t)(u >> 8); op[2] = (uint8_t)(u >> 16);
        op += 3;
    }
    return op;
}
static inline uint8_t* snappy_emit_copy(uint8_t* op, size_t offset, size_t len, bool lt12) {
    if (lt12) return snappy_emit_copy_upto64(op, offset, len, true);
    while (len >= 68) { op = snappy_emit_copy_upto64(op, offset, 64, false); len -= 64; }
    if (len > 64) { op = snappy_emit_copy_upto64(op, offset, 60, false); len -= 60; }
    return snappy_emit_copy_upto64(op, offset, len, len < 12);
}
// HashBytes: ((kMagic * bytes) >> (32 - kMaxHashTableBits)) & mask. kMaxHashTableBits is 14 in snappy 1.1.x (what snappy-java 1.1.10.4
// bundles: COMP_SNAPPY) and 15 from snappy 1.2.0 on (COMP_SNAPPY15) — the only difference between the two generations' compressors.
static thread_local int t_snappy_max_bits = 14;
static inline uint32_t snappy_table_index(uint32_t bytes, uint32_t tmask) {
    const uint32_t kMagic = 0x1e35a7bd;
    return ((kMagic * bytes) >> (32 - t_snappy_max_bits)) & tmask;
}
static uint8_t* snappy_compress_fragment(const uint8_t* input, size_t input_size, uint8_t* op, uint16_t* table, int table_size) {
    const uint8_t* ip = input;
    const uint32_t tmask = (uint32_t)table_size - 1;
    const uint8_t* ip_end = input + input_size;
    const uint8_t* base_ip = ip;
    const size_t kInputMarginBytes = 15;
    if (input_size >= kInputMarginBytes) {
        const uint8_t* ip_limit = input + input_size - kInputMarginBytes;
        for (;;) {
            const uint8_t* next_emit = ip++;
            uint32_t skip = 32;
            const uint8_t* candidate = nullptr;
            bool found = false;
            if (ip_limit - ip >= 16) {
                size_t delta = ip - base_ip;
                for (int i = 0; i < 16; i++) {
                    uint32_t dword = rd32(ip + i);
                    uint16_t* e = &table[snappy_table_index(dword, tmask)];
                    candidate = base_ip + *e;
                    *e = (uint16_t)(delta + i);
                    if (rd32(candidate) == dword) {
                        // literal [next_emit, ip+i) of length i+1
                        *op = (uint8_t)(i << 2);
                        memcpy(op + 1, next_emit, i + 1);
                        ip += i;
                        op = op + i + 2;
                        found = true;
                        break;
                    }
                }
                if (!found) { ip += 16; skip += 16; }
            }
            if (!found) {
                for (;;) {
                    uint32_t data = rd32(ip);
                    uint16_t* e = &table[snappy_table_index(data, tmask)];
                    uint32_t bytes_between = skip >> 5;
                    skip += bytes_between;
                    const uint8_t* next_ip = ip + bytes_between;
                    if (next_ip > ip_limit) { ip = next_emit; goto emit_remainder; }
                    candidate = base_ip + *e;
                    *e = (uint16_t)(ip - base_ip);
                    if (data == rd32(candidate)) break;
                    ip = next_ip;
                }
                op = snappy_emit_literal(op, next_emit, ip - next_emit);
            }
            // emit_match
            do {
                const uint8_t* base = ip;
                size_t matched = 4;
                while (ip + matched < ip_end && candidate[matched] == ip[matched]) matched++;
                bool lt12 = (matched - 4) < 8;
                ip += matched;
                size_t offset = base - candidate;
                op = snappy_emit_copy(op, offset, matched, lt12);
                if (ip >= ip_limit) goto emit_remainder;
                table[snappy_table_index(rd32(ip - 1), tmask)] = (uint16_t)(ip - base_ip - 1);
                uint16_t* e = &table[snappy_table_index(rd32(ip), tmask)];
                candidate = base_ip + *e;
                *e = (uint16_t)(ip - base_ip);
            } while (rd32(ip) == rd32(candidate));
        }
    }
emit_remainder:
    if (ip < ip_end) op = snappy_emit_literal(op, ip, ip_end - ip);
    return op;
}
int snappy_compress(const uint8_t* src, int n, uint8_t* dst, int max_table_bits) {
    t_snappy_max_bits = max_table_bits;
    uint8_t* op = dst;
    uint32_t v = (uint32_t)n;                                   // varint32 preamble (little-endian base-128)
    while (v >= 0x80) { *op++ = (uint8_t)(v | 0x80); v >>= 7; }
    *op++ = (uint8_t)v;
    static thread_local uint16_t table[1 << 15];
    size_t pos = 0;
    while (pos < (size_t)n) {
        size_t frag = (size_t)n - pos; if (frag > 65536) frag = 65536;
        int table_size;
        if (frag > (1u << max_table_bits)) table_size = 1 << max_table_bits;
        else if (frag < (1u << 8)) table_size = 1 << 8;
        else table_size = 2 << (31 - __builtin_clz((uint32_t)(frag - 1)));
        memset(table, 0, table_size * sizeof(uint16_t));
        op = snappy_compress_fragment(src + pos, frag, op, table, table_size);
        pos += frag;
    }
    return (int)(op - dst);
}
int snappy_uncompressed_length(const uint8_t* src, int n) {
    uint32_t v = 0; int shift = 0;
    for (int i = 0; i < n && i < 5; i++) {
        v |= (uint32_t)(src[i] & 0x7f) << shift;
        if (!(src[i] & 0x80)) return (int)v;
        shift += 7;
    }
    return -1;
}
int snappy_decompress(const uint8_t* src, int n, uint8_t* dst, int cap) {
    const uint8_t* ip = src; const uint8_t* iend = src + n;
    uint32_t ulen = 0; int shift = 0; bool ok = false;
    while (ip < iend && shift < 35) { uint8_t b = *ip++; ulen |= (uint32_t)(b & 0x7f) << shift; if (!(b & 0x80)) { ok = true; break; } shift += 7; }
    if (!ok || (int)ulen > cap) return -1;
    uint8_t* op = dst; uint8_t* oend = dst + ulen;
    while (ip < iend) {
        uint8_t tag = *ip++;
        size_t len, offset;
        switch (tag & 3) {
        case 0: {
            len = (tag >> 2) + 1;
            if (len > 60) { int cnt = (int)len - 60; if (iend - ip < cnt) return -1; len = 0; for (int i = 0; i < cnt; i++) len |= (size_t)ip[i] << (8 * i); len += 1; ip += cnt; }
            if ((size_t)(iend - ip) < len || (size_t)(oend - op) < len) return -1;
            memcpy(op, ip, len); op += len; ip += len;
            continue; }
        case 1: if (iend - ip < 1) return -1; len = ((tag >> 2) & 7) + 4; offset = ((size_t)(tag >> 5) << 8) | ip[0]; ip += 1; break;
        case 2: if (iend - ip < 2) return -1; len = (tag >> 2) + 1; offset = ip[0] | ((size_t)ip[1] << 8); ip += 2; break;
        default: if (iend - ip < 4) return -1; len = (tag >> 2) + 1; offset = rd32(ip); ip += 4; break;
        }
        if (offset == 0 || offset > (size_t)(op - dst) || (size_t)(oend - op) < len) return -1;
        for (size_t i = 0; i < len; i++) op[i] = op[i - offset];
        op += len;
    }
    return op == oend ? (int)ulen : -1;
}

// ------------------------------------------------------------------------------------------------
// ICompressor view — S/io/compress/LZ4Compressor.java:108-134 (4-byte LE length prefix + block), SnappyCompressor.java:72-88
// ------------------------------------------------------------------------------------------------
int chunk_max_compressed(int compressor, int chunk_len) {
    if (compressor == COMP_LZ4) return 4 + lz4_compress_bound(chunk_len);       // initialCompressedBufferLength :108-111
    if (compressor == COMP_SNAPPY || compressor == COMP_SNAPPY15) return snappy_max_compressed_length(chunk_len);
    return chunk_len;
}
int chunk_compress(int compressor, const uint8_t* src, int n, uint8_t* dst) {
    if (compressor == COMP_LZ4) {
        dst[0] = (uint8_t)n; dst[1] = (uint8_t)(n >> 8); dst[2] = (uint8_t)(n >> 16); dst[3] = (uint8_t)(n >> 24);
        int c = lz4_compress_block(src, n, dst + 4, lz4_compress_bound(n));
        return c <= 0 ? -1 : 4 + c;
    }
    if (compressor == COMP_SNAPPY) return snappy_compress(src, n, dst, 14);
    if (compressor == COMP_SNAPPY15) return snappy_compress(src, n, dst, 15);
    memcpy(dst, src, n); return n;
}
int chunk_decompress(int compressor, const uint8_t* src, int n, uint8_t* dst, int cap) {
    if (compressor == COMP_LZ4) {
        if (n < 4) return -1;
        int ulen = (int)(src[0] | (src[1] << 8) | (src[2] << 16) | ((uint32_t)src[3] << 24));
        if (ulen < 0 || ulen > cap) return -1;
        int w = lz4_decompress_block(src + 4, n - 4, dst, ulen);
        return (w == ulen) ? ulen : -1;                                          // "Decompressed lengths mismatch" :161-164
    }
    if (compressor == COMP_SNAPPY || compressor == COMP_SNAPPY15) return snappy_decompress(src, n, dst, cap);
    if (n > cap) return -1;
    memcpy(dst, src, n); return n;
}

} // namespace oracle
