// CPU fuzz of the engine's per-thread LZ4 decoder (cassandra_b200/csrc/lz4_thread.cuh, the code K1's thread-per-chunk kernel runs)
// against the oracle, meant to be built with -fsanitize=address,undefined. Test infrastructure only.
//   valid streams   : oracle::lz4_compress_block output must decode to the original bytes, for every source alignment
//   damaged streams : flipped / truncated input must either fail (-1) or decode — never touch memory outside
//                     [src, src + n + 16) and [dst, dst + cap + 16)  (the slack every engine buffer carries)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <random>
#include "../../cassandra_b200/csrc/lz4_thread.cuh"
#include "../../cassandra_b200/csrc/lz4_batch.cuh"      // the walk of the two-pass decoder: same acceptance set, exact-size record buffer
#include "../../oracle/codec.h"

static std::vector<uint8_t> make_data(std::mt19937_64& rng, int n, int kind) {
    std::vector<uint8_t> d(n);
    switch (kind) {
    case 0: for (auto& b : d) b = (uint8_t)rng(); break;                                           // incompressible
    case 1: { uint8_t v = (uint8_t)rng(); for (int i = 0; i < n; i++) { if (rng() % 97 == 0) v = (uint8_t)rng(); d[i] = v; } } break;   // runs (offset 1)
    case 2: { int period = 2 + (int)(rng() % 14); for (int i = 0; i < n; i++) d[i] = i < period ? (uint8_t)rng() : (rng() % 53 ? d[i - period] : (uint8_t)rng()); } break;  // short periods
    case 3: { for (int i = 0; i < n; i++) { int r = i % 27; d[i] = r == 0 ? 0x24 : (r < 9 ? (uint8_t)(rng() >> (8 * (r & 3))) : (r < 17 ? (uint8_t)(i / 27 >> (r & 7)) : (uint8_t)(r * 7))); } } break;  // row-like
    default: { std::vector<std::vector<uint8_t>> words(48); for (auto& w : words) { w.resize(3 + rng() % 20); for (auto& b : w) b = (uint8_t)rng(); }
               int i = 0; while (i < n) { auto& w = words[rng() % words.size()]; for (uint8_t b : w) { if (i < n) d[i++] = b; } } } break;                  // dictionary text
    }
    return d;
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 400;
    std::mt19937_64 rng(0xB200C);
    long ok = 0, rejected = 0, survived = 0;
    { uint8_t* e = (uint8_t*)malloc(1 + 16); e[0] = 0; uint8_t* d = nullptr; if (posix_memalign((void**)&d, 8, 24)) return 3;      // the empty block: one token, no bytes
      if (b200c::lz4_decompress_thread(e, 1, d, 0) != 0 || b200c::lz4_decompress_thread(e, 0, d, 0) != 0 || b200c::lz4_decompress_thread(e, 0, d, 5) != -1) return 6; free(e); free(d); }
    for (int it = 0; it < iters; it++) {
        const int n = it < 8 ? it + 1 : 1 + (int)(rng() % 3 ? rng() % 16384 : rng() % 65535);
        std::vector<uint8_t> data = make_data(rng, n, it % 5);
        std::vector<uint8_t> comp(n + n / 255 + 32);
        int c = oracle::lz4_compress_block(data.data(), n, comp.data(), (int)comp.size());
        if (c <= 0) { fprintf(stderr, "oracle compress failed n=%d\n", n); return 2; }
        for (int align = 0; align < 8; align += (it % 3 ? 3 : 1)) {
            // exact-size heap blocks (+16 bytes of slack, as the engine guarantees): ASan traps anything beyond
            uint8_t* sraw = (uint8_t*)malloc(align + c + 16); uint8_t* src = sraw + align; memcpy(src, comp.data(), c); memset(src + c, 0xEE, 16);
            uint8_t* dst = nullptr; if (posix_memalign((void**)&dst, 8, n + 16 + 8)) return 3;
            memset(dst, 0xCD, n + 16);
            { const int rc = (c - 1) / 3 + 1; uint16_t* rec = (uint16_t*)malloc(sizeof(uint16_t) * rc); int ns = b200c::lz4_walk_thread(src, c, n, rec, rc);
              if (ns < 1 || ns > rc) { fprintf(stderr, "WALK refused a valid block it=%d n=%d ns=%d\n", it, n, ns); return 7; } free(rec); }
            int got = b200c::lz4_decompress_thread(src, c, dst, n);
            if (got != n || memcmp(dst, data.data(), n)) { fprintf(stderr, "MISMATCH it=%d n=%d align=%d got=%d\n", it, n, align, got); return 1; }
            ok++;
            // wrong capacity must be refused, not overrun
            if (n > 0) { int g2 = b200c::lz4_decompress_thread(src, c, dst, n - 1 - (int)(rng() % (n > 9 ? 9 : n))); if (g2 >= 0 && g2 > n) return 4; if (g2 < 0) rejected++; }
            // damage: bit flips and truncations
            for (int m = 0; m < 6 && c > 0; m++) {
                std::vector<uint8_t> bad(src, src + c);
                int cc = c;
                if (m & 1) cc = (int)(rng() % c); else for (int k = 0; k < 1 + (int)(rng() % 3); k++) bad[rng() % c] ^= (uint8_t)(1u << (rng() % 8));
                uint8_t* braw = (uint8_t*)malloc(align + cc + 16); uint8_t* b = braw + align; if (cc) memcpy(b, bad.data(), cc); memset(b + cc, 0x11, 16);
                int g3 = b200c::lz4_decompress_thread(b, cc, dst, n);
                { const int rc = cc > 0 ? (cc - 1) / 3 + 1 : 1; uint16_t* rec = (uint16_t*)malloc(sizeof(uint16_t) * rc); int ns = b200c::lz4_walk_thread(b, cc, n, rec, rc); free(rec);
                  if ((ns >= 0) != (g3 == n)) { fprintf(stderr, "WALK and decoder disagree it=%d m=%d ns=%d g3=%d n=%d\n", it, m, ns, g3, n); return 8; } }
                if (g3 > n) { fprintf(stderr, "OVERRUN reported it=%d\n", it); return 5; }
                if (g3 < 0) rejected++; else survived++;
                free(braw);
            }
            free(sraw); free(dst);
        }
    }
    printf("lz4_thread_fuzz ok: %ld exact decodes, %ld damaged inputs rejected, %ld decoded within bounds\n", ok, rejected, survived);
    return 0;
}
