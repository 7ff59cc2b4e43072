// Host build of the engine's K4 thread path (cassandra_b200/csrc/partition.cuh: process_partition — row merge, reconciliation,
// purge, big-format serialisation, promoted index) for CPU-side parity and memory-safety tests. Test infrastructure only: the very
// source the GPU runs is compiled with g++ behind a few intrinsic shims; K1 (chunk decode), K2 (Index.db walk) and K3 (partition
// merge) are replaced by straightforward host code built on the oracle's codecs, so that what is under test is K4 alone.
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <string>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    unsigned long long pool = ((unsigned long long)b << 32) | a; unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= (unsigned)((pool >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) { sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
static inline unsigned __shfl_xor_sync(unsigned, unsigned v, int) { return v; }
#include "../../cassandra_b200/csrc/partition.cuh"
#include "../../include/b200c.h"
#include "../../oracle/codec.h"

namespace b200c { int64_t murmur3_token(const uint8_t* key, uint32_t len) { return oracle::murmur3_token(key, len); } }
using namespace b200c;

namespace {
struct Part { int64_t tok; const uint8_t* key; uint32_t klen; int src; uint64_t local; };
int fail(char* err, int cap, const std::string& m) { if (err && cap > 0) snprintf(err, cap, "%s", m.c_str()); return -1; }
}

extern "C" int k4host_compact(const b200c_manifest* m, uint8_t* uout, uint64_t ucap, uint64_t* ulen, uint8_t* iout, uint64_t icap, uint64_t* ilen,
                              uint64_t* stats /*[0] merged unfiltereds, [1] rows out, [2] partitions out*/, char* err, int errcap) {
    const int K = m->ninputs;
    if (K <= 0 || K > MAXK) return fail(err, errcap, "inputs");
    // ---- "K1": decompress every input into one buffer (64 KiB-aligned bases, slack behind) ---------------------------------------
    std::vector<uint64_t> ubase(K + 1, 0);
    for (int i = 0; i < K; i++) ubase[i + 1] = ubase[i] + ((m->inputs[i].data_length + 64 + 65535) & ~65535ull);
    std::vector<uint8_t> Ubuf(ubase[K] + 256 + 16, 0);
    uint8_t* U = Ubuf.data() + ((16 - ((uintptr_t)Ubuf.data() & 15)) & 15);
    for (int i = 0; i < K; i++) {
        const b200c_input& in = m->inputs[i];
        for (uint64_t ch = 0; ch < in.nchunks; ch++) {
            uint64_t off = in.chunk_offsets[ch], next = ch + 1 < in.nchunks ? in.chunk_offsets[ch + 1] : in.data_len;
            int ul = (int)std::min<uint64_t>(in.chunk_len, in.data_length - ch * (uint64_t)in.chunk_len);
            if (next < off + 4 || next > in.data_len) return fail(err, errcap, "chunk offsets");
            int got = oracle::chunk_decompress(in.compressor, in.data + off, (int)(next - off - 4), U + ubase[i] + ch * (uint64_t)in.chunk_len, ul);
            if (got != ul) return fail(err, errcap, "chunk decode");
        }
    }
    // ---- "K2": walk Index.db --------------------------------------------------------------------------------------------------------
    std::vector<uint64_t> pbase(K + 1, 0);
    std::vector<int64_t> tok; std::vector<uint64_t> kp, upos; std::vector<uint16_t> klen; std::vector<const uint8_t*> keyp;
    std::vector<Part> parts;
    for (int i = 0; i < K; i++) {
        const b200c_input& in = m->inputs[i];
        pbase[i] = tok.size();
        uint64_t o = 0, n = 0;
        while (o < in.index_len) {
            if (o + 2 > in.index_len) return fail(err, errcap, "index");
            uint32_t kl = ((uint32_t)in.index[o] << 8) | in.index[o + 1];
            const uint8_t* key = in.index + o + 2; uint64_t p = o + 2 + kl;
            uint64_t pos, ps; int r = oracle::vint_read(in.index + p, in.index + in.index_len, &pos); if (r <= 0) return fail(err, errcap, "index"); p += r;
            r = oracle::vint_read(in.index + p, in.index + in.index_len, &ps); if (r <= 0) return fail(err, errcap, "index"); p += r;
            uint64_t pre = 0; for (uint32_t q = 0; q < 8; q++) pre = (pre << 8) | (q < kl ? key[q] : 0);
            int64_t t = m->partitioner ? (int64_t)(pre ^ 0x8000000000000000ull) : oracle::murmur3_token(key, kl);     // k_index_emit
            tok.push_back(t); kp.push_back(pre); klen.push_back((uint16_t)kl); upos.push_back(ubase[i] + pos); keyp.push_back(key);
            if ((m->token_lo == INT64_MIN || t > m->token_lo) && t <= m->token_hi) parts.push_back(Part{t, key, kl, i, n});
            o = p + ps; n++;
        }
        tok.push_back(0); kp.push_back(0); klen.push_back(0); upos.push_back(ubase[i] + in.data_length); keyp.push_back(nullptr);     // sentinel
    }
    pbase[K] = tok.size();
    // ---- "K3": partition merge: order by (token, unsigned key bytes), sources in input order ------------------------------------------
    std::stable_sort(parts.begin(), parts.end(), [](const Part& a, const Part& b) {
        if (a.tok != b.tok) return a.tok < b.tok;
        int c = memcmp(a.key, b.key, std::min(a.klen, b.klen)); if (c) return c < 0;
        if (a.klen != b.klen) return a.klen < b.klen;
        return a.src < b.src; });
    std::vector<uint64_t> contrib, op_first;
    for (size_t k = 0; k < parts.size(); k++) {
        bool head = k == 0 || parts[k].tok != parts[k - 1].tok || parts[k].klen != parts[k - 1].klen || memcmp(parts[k].key, parts[k - 1].key, parts[k].klen)
                    || parts[k].src == parts[k - 1].src;      // a source contributes one partition per merge round (MergeIterator), even to an (invalid) repeated key
        if (head) op_first.push_back(contrib.size());
        contrib.push_back(((uint64_t)head << 63) | ((uint64_t)parts[k].src << 56) | parts[k].local);
    }
    op_first.push_back(contrib.size());
    const uint64_t nparts = op_first.size() - 1;
    // ---- parameters, exactly as compact.cu fills them ----------------------------------------------------------------------------------
    CParams P; memset(&P, 0, sizeof(P));
    P.U = U; P.ninputs = K; P.nclust = m->nclustering; P.ncols = m->ncolumns; P.column_index_size = m->column_index_size > 0 ? m->column_index_size : 65536;
    for (int k = 0; k < m->nclustering; k++) { P.ctype[k] = m->clustering[k].type; P.cfix[k] = m->clustering[k].fixed_len; }
    for (int k = 0; k < m->ncolumns; k++) {
        const int ptype = ((m->columns[k].type >> 8) & 0xFF) - 1;
        P.vfix[k] = m->columns[k].fixed_len & 0xFFFF;
        if (ptype >= 0) { if (!P.ncx) P.cx_first = k; P.ptype[P.ncx] = ptype; P.pfix[P.ncx] = (int32_t)((uint32_t)m->columns[k].fixed_len >> 16); P.ncx++; }
        if ((m->columns[k].type & 0xFF) == B200C_TYPE_COUNTER) P.ctr_mask |= 1ull << k;
    }
    if (!P.ncx) P.cx_first = m->ncolumns;
    for (int k = 0; k < m->nstatic_columns; k++) if ((m->static_columns[k].type & 0xFF) == B200C_TYPE_COUNTER) P.sctr_mask |= 1ull << k;
    P.nstat = m->nstatic_columns; P.mcols = std::max(m->ncolumns, m->nstatic_columns);
    for (int k = 0; k < m->nstatic_columns; k++) P.sfix[k] = m->static_columns[k].fixed_len;
    P.o_min_ts = m->out_stats.min_timestamp; P.o_min_ldt = m->out_stats.min_local_deletion_time; P.o_min_ttl = m->out_stats.min_ttl;
    P.now = m->now_in_sec; P.gc_before = m->gc_before; P.purge_max_ts = m->purge_max_timestamp;
    P.partitioner = m->partitioner;
    std::vector<InDesc> hin(K); memset(hin.data(), 0, sizeof(InDesc) * K); P.in = hin.data();
    for (int i = 0; i < K; i++) {
        const b200c_input& in = m->inputs[i]; InDesc& d = hin[i];
        d.ubase = ubase[i]; d.ulen = in.data_length; d.min_ts = in.header_stats.min_timestamp; d.min_ldt = in.header_stats.min_local_deletion_time; d.min_ttl = in.header_stats.min_ttl;
        d.ncols = in.ncolumns; for (int k = 0; k < in.ncolumns; k++) d.colmap[k] = in.column_map[k];
        d.nstat = in.nstatic_columns; for (int k = 0; k < in.nstatic_columns; k++) d.smap[k] = in.static_column_map[k];
    }
    // ---- K4, the code under test: size pass, host-side scans, emit pass ---------------------------------------------------------------
    std::vector<Cur32> cur(MAXK); /* the cursor type of k_partition_thr */ std::vector<DT> open_dt(MAXK); std::vector<MCell> merged(MAXCOLS);
    std::vector<PartOut> po(nparts); PartStats st{0, 0};
    for (uint64_t j = 0; j < nparts; j++) {
        int e = 0; PartOut out{0, 0, 0, 0, 0};
        if (P.ncx || P.ctr_mask || P.sctr_mask) process_partition<false, Cur32, XlateGlobal, MAXK, true>(P, XlateGlobal(), contrib.data(), op_first[j], (uint32_t)(op_first[j + 1] - op_first[j]), upos.data(), pbase.data(), kp.data(), klen.data(), tok.data(),
                                 nullptr, ~0ull, 0, nullptr, 0, 0, 0, cur.data(), open_dt.data(), merged.data(), out, st, e);      // (tables with multi-cell columns: the CX instantiation, as compact.cu launches it)
        else process_partition<false>(P, XlateGlobal(), contrib.data(), op_first[j], (uint32_t)(op_first[j + 1] - op_first[j]), upos.data(), pbase.data(), kp.data(), klen.data(), tok.data(),
                                 nullptr, ~0ull, 0, nullptr, 0, 0, 0, cur.data(), open_dt.data(), merged.data(), out, st, e);
        if (e) return fail(err, errcap, e == PERR_UNSUPPORTED ? "unsupported" : "corrupt data");
        po[j] = out;
    }
    std::vector<uint64_t> dpos(nparts + 1, 0), ipos(nparts + 1, 0);
    for (uint64_t j = 0; j < nparts; j++) {
        dpos[j + 1] = dpos[j] + po[j].dsize;
        uint64_t isz = po[j].dsize ? po[j].ihead + vint_size(dpos[j]) + vint_size(po[j].ipay) + po[j].ipay : 0;
        ipos[j + 1] = ipos[j] + isz;
    }
    *ulen = dpos[nparts]; *ilen = ipos[nparts];
    if (dpos[nparts] + 16 > ucap || ipos[nparts] + 16 > icap) return fail(err, errcap, "output buffers too small");
    PartStats st2{0, 0}; uint64_t written = 0;
    for (uint64_t j = 0; j < nparts; j++) {
        if (!po[j].dsize) continue;
        int e = 0; PartOut out{0, 0, 0, 0, 0};
        if (P.ncx || P.ctr_mask || P.sctr_mask) process_partition<true, Cur32, XlateGlobal, MAXK, true>(P, XlateGlobal(), contrib.data(), op_first[j], (uint32_t)(op_first[j + 1] - op_first[j]), upos.data(), pbase.data(), kp.data(), klen.data(), tok.data(),
                                uout + dpos[j], ~0ull, dpos[j], iout + ipos[j], po[j].nblk, po[j].ipay, 0, cur.data(), open_dt.data(), merged.data(), out, st2, e);
        else process_partition<true>(P, XlateGlobal(), contrib.data(), op_first[j], (uint32_t)(op_first[j + 1] - op_first[j]), upos.data(), pbase.data(), kp.data(), klen.data(), tok.data(),
                                uout + dpos[j], ~0ull, dpos[j], iout + ipos[j], po[j].nblk, po[j].ipay, 0, cur.data(), open_dt.data(), merged.data(), out, st2, e);
        if (e || out.dsize != po[j].dsize) return fail(err, errcap, "size pass and emit pass disagree at partition " + std::to_string(j));
        written++;
    }
    // ---- the staged path (k_partition_staged): token-contiguous tiles of output partitions whose input byte ranges (one contiguous
    //      range per source) are copied into a small buffer, P.U pointing at the copy, 32-byte cursors with 32/16-bit offsets -----------
    if (!P.ncx && !P.ctr_mask && !P.sctr_mask) {                               // (tables with multi-cell or counter columns run the thread kernels only)
        std::vector<uint8_t> u2(dpos[nparts] + 64, 0), i2(ipos[nparts] + 64, 0);
        std::vector<CurS> curs(MAXK); PartStats st3{0, 0};
        const uint64_t TILE_BYTES = 40000;
        uint64_t j0 = 0;
        while (j0 < nparts) {
            // grow the tile while it fits
            uint64_t j1 = j0, bytes = 0; std::vector<uint64_t> lo(K, ~0ull), hi(K, 0);
            while (j1 < nparts && j1 - j0 < 128) {
                uint64_t add = 0;
                for (uint64_t c = op_first[j1]; c < op_first[j1 + 1]; c++) { uint64_t e = contrib[c]; uint64_t g = pbase[(e >> 56) & 0x7F] + (e & 0xFFFFFFFFFFull); add += upos[g + 1] - upos[g]; }
                if (j1 > j0 && bytes + add > TILE_BYTES) break;
                bytes += add; j1++;
            }
            for (uint64_t c = op_first[j0]; c < op_first[j1]; c++) { uint64_t e = contrib[c]; int src = (int)((e >> 56) & 0x7F); uint64_t l = e & 0xFFFFFFFFFFull; lo[src] = std::min(lo[src], l); hi[src] = std::max(hi[src], l); }
            std::vector<uint32_t> sbase(K, 0); std::vector<uint64_t> g0(K, 0); uint64_t tot = 0;
            for (int i = 0; i < K; i++) if (lo[i] != ~0ull) {
                uint64_t a = upos[pbase[i] + lo[i]] & ~15ull, b = (upos[pbase[i] + hi[i] + 1] + 15) & ~15ull;
                sbase[i] = (uint32_t)tot; g0[i] = a; tot += b - a;
            }
            if (tot + 64 > 65535 || bytes > 60000) {          // too big to stage: the global path handles it (as the kernel does)
                for (uint64_t j = j0; j < j1; j++) if (po[j].dsize) {
                    int e = 0; PartOut out{0, 0, 0, 0, 0};
                    process_partition<true>(P, XlateGlobal(), contrib.data(), op_first[j], (uint32_t)(op_first[j + 1] - op_first[j]), upos.data(), pbase.data(), kp.data(), klen.data(), tok.data(),
                                            u2.data() + dpos[j], ~0ull, dpos[j], i2.data() + ipos[j], po[j].nblk, po[j].ipay, 0, cur.data(), open_dt.data(), merged.data(), out, st3, e);
                    if (e) return fail(err, errcap, "staged pass (global fallback) failed");
                }
                j0 = j1; continue;
            }
            std::vector<uint8_t> tile(tot + 64, 0xEE);
            for (int i = 0; i < K; i++) if (lo[i] != ~0ull) memcpy(tile.data() + sbase[i], U + g0[i], ((upos[pbase[i] + hi[i] + 1] + 15) & ~15ull) - g0[i]);
            CParams PS = P; PS.U = tile.data();
            XlateStaged xl{sbase.data(), g0.data()};
            for (uint64_t j = j0; j < j1; j++) if (po[j].dsize) {
                int e = 0; PartOut out{0, 0, 0, 0, 0};
                process_partition<true>(PS, xl, contrib.data(), op_first[j], (uint32_t)(op_first[j + 1] - op_first[j]), upos.data(), pbase.data(), kp.data(), klen.data(), tok.data(),
                                        u2.data() + dpos[j], ~0ull, dpos[j], i2.data() + ipos[j], po[j].nblk, po[j].ipay, 0, curs.data(), open_dt.data(), merged.data(), out, st3, e);
                if (e || out.dsize != po[j].dsize) return fail(err, errcap, "staged pass disagrees at partition " + std::to_string(j));
            }
            j0 = j1;
        }
        if (memcmp(u2.data(), uout, dpos[nparts]) || memcmp(i2.data(), iout, ipos[nparts])) return fail(err, errcap, "staged pass produced different bytes");
        if (st3.merged_unfiltereds != st2.merged_unfiltereds || st3.rows_out != st2.rows_out) return fail(err, errcap, "staged pass counters differ");
    }
    if (stats) { stats[0] = st.merged_unfiltereds + nparts; stats[1] = st.rows_out; stats[2] = written; }
    return 0;
}
