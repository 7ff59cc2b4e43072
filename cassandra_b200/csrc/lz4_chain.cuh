// lz4_chain.cuh — LZ4 block compression in two kernels, bit-exact with liblz4's LZ4_compress_default like lz4.cuh, built so that the
// SEQUENTIAL part of the format no longer owns 16 KiB of shared memory per chunk.
//
// lz4.cuh keeps liblz4's hash table (8192 x u16) in shared memory for the whole life of a chunk: 14 chunks per SM, each of them one long
// chain of dependent instructions — that is what bounds K5 (ncu: 14 warps per SM, 40 % issue utilisation). What the table answers is
// "the most recent INSERTED position with this hash"; which positions get inserted depends on the parse, but "the most recent EARLIER
// positions with this hash" do not. So:
//
//   pass A (lz4_chain_build_warp, one warp per chunk, table in shared memory, no parse — 512 independent-looking steps per chunk):
//       for every position p the two nearest earlier positions with the same hash, q1 > q2, and whether their 4 bytes equal p's:
//       ent[p] = q1 | eq1 << 15 | q2 << 16 | eq2 << 31        (no earlier position: 0 — the zero-initialised table entry liblz4 reads)
//   pass B (lz4_compress_warp_chain, the parse): the candidate of a search attempt at p is the first INSERTED position along
//       q1, q2, ent[q2].q1, ...; "inserted" is one bit per position (2 KiB of shared memory per chunk instead of 16). Nine lookups in ten
//       end at q1 (measured on the benchmark's data), and the hit test is the precomputed bit: the search window reads no chunk bytes.
//
// Equivalence with the table: positions are inserted in increasing order (search attempts move forward; after a match liblz4 inserts ip-2,
// which lies behind every earlier insertion, then tests and inserts ip), so table[h] is always the LARGEST inserted position with hash h,
// or 0 while none was inserted — and the chain from p visits exactly the earlier positions with p's hash in decreasing order.
// Chunk lengths up to 32 KiB (15-bit positions); longer chunks keep the lz4.cuh kernel.
#pragma once
#include "lz4.cuh"

namespace b200c {

enum { LZ4C_MAX_CHUNK = 32768, LZ4C_NONE = 0xFFFF, LZ4C_DUP_ENTRIES = 8192 };      // dup table: one byte per LZ4 hash (no false sharing)

#if defined(__CUDA_ARCH__)
#define LZ4C_PREFETCH(p) asm volatile("prefetch.global.L2 [%0];" :: "l"(p))
#else
#define LZ4C_PREFETCH(p) ((void)0)
#endif

// number of positions that can ever be looked up or inserted: search attempts and the post-match test stay below mflimitPlusOne = n - 11
__host__ __device__ __forceinline__ int lz4c_positions(int n) { return n >= LZ4_MINLENGTH ? n - LZ4_MFLIMIT + 1 : 0; }

// ---- pass A ---------------------------------------------------------------------------------------------------------------------------
// s_in: the chunk (4-byte aligned, >= npos + 7 readable bytes) — GLOBAL: where it lies in global memory, else a copy in shared memory (the build
// pass is a chain of table updates: with the chunk's bytes a DRAM access away each of its 512 steps cost 1 350 cycles); s_t1: nent x u16 (last position per hash);
// s_dup: LZ4C_DUP_ENTRIES bytes; ent: npos words. HASH: 4 bytes -> table index (< nent = 1 << hbits).
template <bool GLOBAL, class HASH> __device__ __forceinline__ void chain_build_warp(const uint8_t* s_in, int npos, int nent, int hbits, HASH hash, uint16_t* s_t1, uint32_t* __restrict__ ent, int lane) {
    const uint32_t* in32 = (const uint32_t*)s_in;
    {
        uint4* a = (uint4*)s_t1;
        for (int i = lane; i < (nent * 2) / 16; i += 32) a[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
    }
    __syncwarp();
    const uint32_t lt_mask = (1u << lane) - 1u;
    // ---- phase 1 (sequential in the table): nearest earlier position with the same hash. The only shared memory is the table itself (16 KiB:
    //      14 chunks per SM — a warp alone on its scheduler runs this chain at ~11 cycles per instruction, residency is what it needs), so the
    //      chunk is read where it lies, four steps ahead of the table work (the bytes of a step are a DRAM access away).
    enum { AHEAD = 4 };
    uint32_t ring[AHEAD];
#pragma unroll
    for (int k = 0; k < AHEAD; k++) ring[k] = (32 * k + lane < npos) ? lz4_rd32<GLOBAL>(in32, 32 * k + lane) : 0u;
    for (int p0 = 0; p0 < npos; p0 += 32 * AHEAD) {
#pragma unroll
        for (int k = 0; k < AHEAD; k++) {
            const int p = p0 + 32 * k + lane; const bool valid = p < npos;
            if (p0 + 32 * k >= npos) break;
            const uint32_t seq = ring[k];
            ring[k] = (p + 32 * AHEAD < npos) ? lz4_rd32<GLOBAL>(in32, p + 32 * AHEAD) : 0u;
            const uint32_t h = hash(seq);
            // every lane reads its slot, then writes its position into it and reads back: a lane that reads another position shares its hash with
            // a lane of this step (runs and repeated row prefixes make that common in SSTable bytes). One round per shared hash puts the group in
            // order: nearest lower lane = predecessor, highest lane owns the slot.
            uint32_t q1 = valid ? (uint32_t)s_t1[h] : (uint32_t)LZ4C_NONE;
            __syncwarp();
            if (valid) s_t1[h] = (uint16_t)p;
            __syncwarp();
            const bool lost = valid && s_t1[h] != (uint16_t)p;
            uint32_t todo = __ballot_sync(FULL_MASK, lost);
            if (todo) {
                bool fix = false;
                while (todo) {
                    const int l = __ffs(todo) - 1;
                    const uint32_t hh = __shfl_sync(FULL_MASK, h, l);
                    const uint32_t same = __ballot_sync(FULL_MASK, valid && h == hh);
                    if (valid && h == hh) {
                        const uint32_t lower = same & lt_mask;
                        if (lower) q1 = (uint32_t)(p0 + 32 * k + 31 - __clz(lower));
                        fix = !(same & ~lt_mask & ~(1u << lane));                 // the group's highest lane
                    }
                    todo &= ~same;
                }
                if (fix) s_t1[h] = (uint16_t)p;
                __syncwarp();
            }
            if (valid) ent[p] = q1 == LZ4C_NONE ? 0u : q1;      // (nothing of this phase waits for a load behind the table: a warp issues in order,
                                                                //  the 4-byte compare of the candidate would stall the next step's table work)
        }
    }
    __syncwarp();
    // ---- phase 2 (parallel, no dependency between the steps: eight of them in flight hide the dependent reads): the predecessor's predecessor
    //      and both equality bits; position 0 ends every chain (its own link is 0)
#pragma unroll 8
    for (int p0 = 0; p0 < npos; p0 += 32) {
        const int p = p0 + lane;
        if (p < npos) {
            const uint32_t a1 = ent[p] & 0x7FFFu;
            const uint32_t a2 = ent[a1] & 0x7FFFu;
            const uint32_t sp = lz4_rd32<GLOBAL>(in32, p);
            const uint32_t e1 = lz4_rd32<GLOBAL>(in32, (int)a1) == sp, e2 = lz4_rd32<GLOBAL>(in32, (int)a2) == sp;
            ent[p] = a1 | (e1 << 15) | (a2 << 16) | (e2 << 31);
        }
    }
}
struct Lz4Hash { __device__ __forceinline__ uint32_t operator()(uint32_t seq) const { return lz4_hash_u16(seq); } };
template <bool GLOBAL> __device__ __forceinline__ void lz4_chain_build_warp(const uint8_t* s_in, int n, uint16_t* s_t1, uint32_t* __restrict__ ent, int lane) {
    chain_build_warp<GLOBAL>(s_in, lz4c_positions(n), LZ4_TABLE_ENTRIES, LZ4_HASHLOG_U16, Lz4Hash(), s_t1, ent, lane);
}

// bits [lo, hi) that fall into 32-bit word w of a bitmap
__device__ __forceinline__ uint32_t lz4c_range_bits(int lo, int hi, int w) {
    const int a = lo > w * 32 ? lo - w * 32 : 0, b = hi < w * 32 + 32 ? hi - w * 32 : 32;
    if (b <= a) return 0u;
    const uint32_t upto_b = b >= 32 ? FULL_MASK : ((1u << b) - 1u);
    return upto_b & ~((1u << a) - 1u);
}

// ---- pass B ---------------------------------------------------------------------------------------------------------------------------
// s_in as above; ent: pass A's output for this chunk; s_bm: (n + 31) / 32 words of shared memory (zeroed here); out: capacity >=
// lz4_compress_bound(n). Returns the compressed size (warp-uniform). The control flow is lz4_compress_warp's (32 speculated attempts per
// step, the post-match insert / test folded into the first window as lanes 0 and 1); only the candidate lookup and the hit test differ.
__device__ int lz4_compress_warp_chain(const uint8_t* s_in, int n, const uint32_t* __restrict__ ent, uint32_t* s_bm, uint8_t* out, int lane) {
    const uint32_t* in32 = (const uint32_t*)s_in;
    for (int i = lane; i < ((n + 31) >> 5); i += 32) s_bm[i] = (i == 0) ? 1u : 0u;       // "first byte": position 0 is inserted
    __syncwarp();
    int anchor = 0, op = 0;
    const int mfl1 = n - LZ4_MFLIMIT + 1;
    const int matchlimit = n - LZ4_LASTLITERALS;

    if (n >= LZ4_MINLENGTH) {
        int fwd = 1; bool have_prefix = false; int pre_ip = 0;
        for (;;) {
            int ip = 0, match = 0, token_pos; bool ended = false, immediate = false;
            if (lane < 4) LZ4C_PREFETCH(ent + fwd + 64 + 32 * lane);
            int a0 = 0;
            for (bool first = true;; first = false) {
                const bool prefixed = first && have_prefix;
                // The window's own positions count as inserted for the lanes behind them (they are, unless an earlier lane hits — and then
                // the later lanes' answers are dropped). Contiguous windows (every attempt within the first 65: step 1) know their positions
                // and that membership by arithmetic; windows in the accelerated regime mark them in the bitmap first and take the unused
                // marks back afterwards.
                const bool contiguous = a0 + 32 <= 66;
                const int w_lo = prefixed ? pre_ip - 2 : fwd + a0;               // first position of a contiguous window
                const int hole = prefixed ? pre_ip - 1 : -1;                      // the one position lanes 0 and 1 of a prefixed window leave out
                int p; bool valid, putonly = false;
                if (contiguous) {
                    p = w_lo + lane + ((prefixed && lane >= 1) ? 1 : 0);
                    putonly = prefixed && lane == 0;
                    valid = (prefixed && lane < 2) || p + 1 <= mfl1;
                } else {
                    const int a = a0 + lane;
                    p = fwd + lz4_attempt_offset(a);
                    valid = fwd + lz4_attempt_offset(a + 1) <= mfl1;
                }
                const uint32_t inval = __ballot_sync(FULL_MASK, !valid);
                const int first_inv = inval ? (__ffs(inval) - 1) : 32;
                if (!contiguous) { if (valid) atomicOr(&s_bm[p >> 5], 1u << (p & 31)); __syncwarp(); }
                const uint32_t e = valid ? ent[p] : 0u;
                const int q1 = (int)(e & 0x7FFFu), q2 = (int)((e >> 16) & 0x7FFFu);
                const bool in1 = contiguous && q1 >= w_lo && q1 != hole, in2 = contiguous && q2 >= w_lo && q2 != hole;
                const bool ins1 = in1 || ((s_bm[q1 >> 5] >> (q1 & 31)) & 1u), ins2 = in2 || ((s_bm[q2 >> 5] >> (q2 & 31)) & 1u);
                int cand = ins1 ? q1 : q2;
                bool hit = valid && !putonly && (ins1 ? ((e >> 15) & 1u) : (ins2 ? (e >> 31) : 0u));
                const bool deeper = valid && !ins1 && !ins2;
                uint32_t hits = __ballot_sync(FULL_MASK, hit);
                const uint32_t dmask = __ballot_sync(FULL_MASK, deeper);
                if (dmask) {
                    // further down the chain (one lookup in forty goes this deep): follow the first-level links, then compare the bytes. Only
                    // lanes in front of the first hit known so far can decide the window — the speculated lanes behind a hit sit INSIDE the
                    // match that follows, exactly the positions whose hash chains are longest (hundreds of never-inserted predecessors)
                    int limit = hits ? (__ffs(hits) - 1) : 32; if (first_inv < limit) limit = first_inv;
                    if ((int)(__ffs(dmask) - 1) < limit) {
                        bool walking = deeper; int q = q2;                     // q: a position known not to be inserted; its entry names the next two
                        for (;;) {
                            walking = walking && lane < limit;
                            const uint32_t eq_ = walking ? ent[q] : 0u;
                            const int c1 = (int)(eq_ & 0x7FFFu), c2 = (int)((eq_ >> 16) & 0x7FFFu);
                            const bool i1 = walking && ((contiguous && c1 >= w_lo && c1 != hole) || ((s_bm[c1 >> 5] >> (c1 & 31)) & 1u));
                            const bool i2 = walking && !i1 && ((contiguous && c2 >= w_lo && c2 != hole) || ((s_bm[c2 >> 5] >> (c2 & 31)) & 1u));
                            const bool ins = i1 || i2;
                            if (ins) { walking = false; cand = i1 ? c1 : c2; hit = !putonly && (lz4_rd32<true>(in32, cand) == lz4_rd32<true>(in32, p)); }
                            else if (walking) q = c2;
                            const uint32_t nh = __ballot_sync(FULL_MASK, ins && hit);
                            if (nh && (int)(__ffs(nh) - 1) < limit) limit = __ffs(nh) - 1;
                            if (!__any_sync(FULL_MASK, walking && lane < limit)) break;
                        }
                        hits = __ballot_sync(FULL_MASK, hit);
                    }
                }
                const int first_hit = hits ? (__ffs(hits) - 1) : 32;
                const bool found = first_hit < first_inv;
                // insertions: lanes up to the hit (all valid lanes when there is none)
                const int last_ins = found ? first_hit : (first_inv - 1);        // -1: no valid lane at all
                if (contiguous) {
                    if (last_ins >= 0 && lane < 2) {
                        const int p_last = w_lo + last_ins + ((prefixed && last_ins >= 1) ? 1 : 0);
                        const int w = (w_lo >> 5) + lane;
                        uint32_t mbits = lz4c_range_bits(w_lo, p_last + 1, w);
                        if (hole >= 0 && (hole >> 5) == w) mbits &= ~(1u << (hole & 31));
                        if (mbits) s_bm[w] |= mbits;
                    }
                } else {
                    __syncwarp();
                    if (valid && lane > last_ins) atomicAnd(&s_bm[p >> 5], ~(1u << (p & 31)));
                }
                __syncwarp();
                if (found) {
                    ip = contiguous ? (w_lo + first_hit + ((prefixed && first_hit >= 1) ? 1 : 0)) : __shfl_sync(FULL_MASK, p, first_hit);
                    match = __shfl_sync(FULL_MASK, cand, first_hit);
                    immediate = prefixed && first_hit == 1;
                    break;
                }
                if (first_inv < 32) { ended = true; break; }
                a0 += prefixed ? 30 : 32;
            }
            if (ended) break;

            // ---- catch up (backwards, not after an immediate match) and match length (forwards from ip + 4, up to matchlimit: LZ4_count):
            //      both first steps are loaded together — one memory latency; the total is back + forward because the bytes in between are
            //      the match itself
            int back = 0, mc = 0;
            {
                const int j = lane + 1;
                const bool okb = !immediate && (ip - j >= anchor) && (match - j >= 0) && (lz4_rd8<true>(s_in, ip - j) == lz4_rd8<true>(s_in, match - j));
                const bool eqf = (ip + LZ4_MINMATCH + lane < matchlimit) && (lz4_rd8<true>(s_in, ip + LZ4_MINMATCH + lane) == lz4_rd8<true>(s_in, match + LZ4_MINMATCH + lane));
                const uint32_t bb = __ballot_sync(FULL_MASK, okb), bf = __ballot_sync(FULL_MASK, eqf);
                back = (bb == FULL_MASK) ? 32 : (__ffs(~bb) - 1);
                mc = (bf == FULL_MASK) ? 32 : (__ffs(~bf) - 1);
                while (back && !(back & 31)) {                   // every lane matched: keep going, 32 bytes per step
                    const int jj = back + lane + 1;
                    const bool ok = (ip - jj >= anchor) && (match - jj >= 0) && (lz4_rd8<true>(s_in, ip - jj) == lz4_rd8<true>(s_in, match - jj));
                    const uint32_t b = __ballot_sync(FULL_MASK, ok);
                    const int steps = (b == FULL_MASK) ? 32 : (__ffs(~b) - 1);
                    back += steps;
                    if (steps < 32) break;
                }
                while (mc && !(mc & 31)) {
                    const int i = mc + lane;
                    const bool eq = (ip + LZ4_MINMATCH + i < matchlimit) && (lz4_rd8<true>(s_in, ip + LZ4_MINMATCH + i) == lz4_rd8<true>(s_in, match + LZ4_MINMATCH + i));
                    const uint32_t b = __ballot_sync(FULL_MASK, eq);
                    const int steps = (b == FULL_MASK) ? 32 : (__ffs(~b) - 1);
                    mc += steps;
                    if (steps < 32) break;
                }
            }
            const int ip_end = ip + LZ4_MINMATCH + mc;         // first byte behind the match
            ip -= back; match -= back; mc += back;

            int lit_nibble = 0;
            if (!immediate) {
                // ---- literals ---------------------------------------------------------------------------------------
                int lit = ip - anchor;
                token_pos = op++;
                if (lit >= 15) op += lz4_emit_len_ext(out + op, lit - 15, lane);
                for (int i = lane; i < lit; i += 32) out[op + i] = (uint8_t)lz4_rd8<true>(s_in, anchor + i);
                op += lit;
                lit_nibble = lit < 15 ? lit : 15;
            } else token_pos = op++;                      // immediate match: token with literal length 0

            // ---- the match: offset, length beyond MINMATCH ------------------------------------------------------------
            if (lane == 0) {
                const int off = ip - match; out[op] = (uint8_t)off; out[op + 1] = (uint8_t)(off >> 8);
                out[token_pos] = (uint8_t)((lit_nibble << 4) | (mc < 15 ? mc : 15));
            }
            op += 2;
            if (mc >= 15) op += lz4_emit_len_ext(out + op, mc - 15, lane);
            ip = ip_end;
            anchor = ip;
            if (ip >= mfl1) break;
            have_prefix = true; pre_ip = ip; fwd = ip + 1;
        }
    }
    // ---- last literals ------------------------------------------------------------------------------------------
    {
        int last = n - anchor;
        if (lane == 0) out[op] = (uint8_t)((last < 15 ? last : 15) << 4);
        op++;
        if (last >= 15) op += lz4_emit_len_ext(out + op, last - 15, lane);
        for (int i = lane; i < last; i += 32) out[op + i] = (uint8_t)lz4_rd8<true>(s_in, anchor + i);
        op += last;
    }
    return op;
}

} // namespace b200c
