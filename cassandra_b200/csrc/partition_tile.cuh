// partition_tile.cuh — K4 as a cooperative kernel: a tile of G lanes owns one output partition, lane v (and v+G when S == 2)
// holds the cursor of contributing input partition v in REGISTERS. Every step is a small tournament:
//   * all lanes compare their head unfiltered's clustering with the current leader's (broadcast by shuffle) until no lane is
//     smaller; the lanes that compare equal form the reduce group (MergeIterator semantics: equal items reduce together),
//   * the group's row headers / cells are decoded in parallel, one lane per source, and folded in source order by broadcasting
//     them (Row.Merger / Cells.reconcile / RangeTombstoneMarker.Merger — see partition.cuh for the citations),
//   * the merged, purged row is serialised once (all lanes track the byte position, lane 0 stores).
// The thread-per-partition version (process_partition in partition.cuh) kept 6.6 KB of cursor state per thread in local memory
// and decoded the m sources one after the other; this version has no local-memory arrays and decodes them side by side.
#pragma once
#include "partition.cuh"
#include <cooperative_groups.h>

namespace b200c {
namespace cg = cooperative_groups;

template <int G> __device__ __forceinline__ int64_t tshfl64(const cg::thread_block_tile<G>& t, int64_t v, int src) { return (int64_t)t.shfl((long long)v, src); }
template <int G> __device__ __forceinline__ uint64_t tshflu64(const cg::thread_block_tile<G>& t, uint64_t v, int src) { return (uint64_t)t.shfl((unsigned long long)v, src); }
template <int G> __device__ __forceinline__ DT tshfl_dt(const cg::thread_block_tile<G>& t, const DT& d, int src) { DT r; r.mfda = tshfl64<G>(t, d.mfda, src); r.ldt = tshfl64<G>(t, d.ldt, src); return r; }
template <int G> __device__ __forceinline__ Cur tshfl_cur(const cg::thread_block_tile<G>& t, const Cur& c, int src) {
    Cur r; r.pos = tshflu64<G>(t, c.pos, src); r.next = 0; r.end = 0; r.k0 = tshflu64<G>(t, c.k0, src);
    r.ck_rel = (uint8_t)t.shfl((uint32_t)c.ck_rel, src); r.ckend_rel = t.shfl(c.ckend_rel, src); r.body_rel = 0;
    uint32_t packed = (uint32_t)c.flags | ((uint32_t)c.kind << 8) | ((uint32_t)c.n << 16) | ((uint32_t)c.fast << 24);
    packed = t.shfl(packed, src);
    r.flags = (uint8_t)packed; r.kind = (uint8_t)(packed >> 8); r.n = (uint8_t)(packed >> 16); r.fast = (uint8_t)(packed >> 24); r.ext = 0; r.src = 0; r.done = false;
    return r;
}

// AbstractCell.purge for one merged cell (pure function): S/db/rows/AbstractCell.java:78-99
__device__ __forceinline__ MCell purge_cell(const Purger& pg, MCell m) {
    if (!m.present) return m;
    bool is_live = m.ldt == I64_MAX || (m.ttl != 0 && pg.now < m.ldt);
    if (!is_live) {
        if (pg.ts_ldt(m.ts, m.ldt)) { m.present = false; return m; }
        if (m.ttl != 0) { m.ldt = m.ldt - m.ttl; m.ttl = 0; m.vlen = 0; if (pg.ts_ldt(m.ts, m.ldt)) m.present = false; }
    }
    return m;
}

// decodes the next cell of input column `icol` at reader r (Cell.Serializer.deserialize, S/db/rows/Cell.java:307-349)
__device__ __forceinline__ MCell read_cell(const InDesc& in, Rd& r, const Live& info, int oc, const int32_t* vfix) {
    uint32_t cf = r.u8();
    bool has_value = !(cf & 0x04), deleted = cf & 0x01, expiring = cf & 0x02, use_ts = cf & 0x08, use_ttl = cf & 0x10;
    MCell m; m.present = true;
    m.ts = use_ts ? info.ts : (int64_t)(r.vint() + (uint64_t)in.min_ts);
    m.ldt = use_ttl ? info.ldt : ((deleted || expiring) ? (int64_t)r.vint32() + in.min_ldt : I64_MAX);
    m.ttl = use_ttl ? info.ttl : (expiring ? r.vint32() + in.min_ttl : 0);
    m.voff = r.p; m.vlen = 0;
    if (has_value) {
        int64_t len = vfix[oc] > 0 ? vfix[oc] : (int64_t)r.vint32();
        if (len < 0) { r.err = PERR_CORRUPT; len = 0; }
        m.voff = r.p; m.vlen = (int32_t)len; r.skip((uint64_t)len);
    }
    if (m.ttl < 0) r.err = PERR_CORRUPT;
    if (m.ldt != I64_MAX) m.ldt = decode_ldt(m.ldt, m.ttl);
    return m;
}

// One merged row out of the versions the lanes in grp[] stand on (Row.Merger.merge S/db/rows/Row.java:730-781; ColumnDataReducer :838-849):
// row headers and cells are decoded side by side, one lane per source, and folded in source order by broadcast. s_cells[0..columns) receives
// the merged AND purged cells (lane 0 stores). stat: the static column set (mergeStaticRows) instead of the regular one.
template <int G, int S> __device__ __forceinline__ void tile_merge_rows(const cg::thread_block_tile<G>& tile, const CParams& P, const Purger& pg, Cur (&cur)[S], const unsigned (&grp)[S],
                                                                        bool as_is, DT active, bool stat, MCell* s_cells, Live& info, DT& del, int& npresent_merged, int& npresent, int& lerr) {
    const int lane = tile.thread_rank();
    const int nout = stat ? P.nstat : P.ncols; const int32_t* const vfix = stat ? P.sfix : P.vfix;
    bool ing[S];
#pragma unroll
    for (int s = 0; s < S; s++) ing[s] = (grp[s] >> lane) & 1;
    Live myinfo[S]; DT mydel[S]; Rd rd[S]; uint64_t missing[S]; int icol[S];
#pragma unroll
    for (int s = 0; s < S; s++) {
        myinfo[s] = live_empty(); mydel[s] = dt_live(); rd[s] = Rd{P.U, 0, 0, 0}; missing[s] = 0; icol[s] = 0;
        if (ing[s]) {
            rd[s] = row_header(P, cur[s], myinfo[s], mydel[s]);
            if (!(cur[s].flags & 0x20)) missing[s] = rd[s].vint();
            if (rd[s].err) lerr = rd[s].err;
        }
    }
    info = live_empty(); del = dt_live();
#pragma unroll
    for (int s = 0; s < S; s++) {
        for (unsigned bits = grp[s]; bits; bits &= bits - 1) {
            int l = __ffs(bits) - 1;
            Live vi; vi.ts = tshfl64<G>(tile, myinfo[s].ts, l); vi.ldt = tshfl64<G>(tile, myinfo[s].ldt, l); vi.ttl = tile.shfl(myinfo[s].ttl, l);
            DT vd = tshfl_dt<G>(tile, mydel[s], l);
            if (live_supersedes(vi, info)) info = vi;
            if (dt_supersedes(vd, del)) del = vd;
        }
    }
    if (!as_is) {
        if (dt_supersedes(del, active)) active = del; else del = dt_live();
        if (dt_deletes(active, info.ts)) info = live_empty();
    }
    npresent_merged = 0; npresent = 0;
    for (int c = 0; c < nout; c++) {
        MCell mine[S]; bool hasc[S];
#pragma unroll
        for (int s = 0; s < S; s++) {
            hasc[s] = false; mine[s].present = false; mine[s].ts = 0; mine[s].ldt = 0; mine[s].voff = 0; mine[s].ttl = 0; mine[s].vlen = 0;
            if (ing[s]) {
                const InDesc& in = P.in[cur[s].src];
                const int nin = stat ? in.nstat : in.ncols; const int32_t* const map = stat ? in.smap : in.colmap;
                while (icol[s] < nin && ((missing[s] >> icol[s]) & 1)) icol[s]++;
                if (icol[s] < nin && map[icol[s]] == c) {
                    mine[s] = read_cell(in, rd[s], myinfo[s], c, vfix); hasc[s] = true; icol[s]++;
                    if (rd[s].err) lerr = rd[s].err;
                }
            }
        }
        MCell mc; mc.present = false; mc.ts = 0; mc.ldt = I64_MAX; mc.voff = 0; mc.ttl = 0; mc.vlen = 0;
#pragma unroll
        for (int s = 0; s < S; s++) {
            for (unsigned bits = tile.ballot(hasc[s]); bits; bits &= bits - 1) {
                int l = __ffs(bits) - 1;
                MCell x; x.present = true;
                x.ts = tshfl64<G>(tile, mine[s].ts, l); x.ldt = tshfl64<G>(tile, mine[s].ldt, l); x.voff = tshflu64<G>(tile, mine[s].voff, l);
                x.ttl = tile.shfl(mine[s].ttl, l); x.vlen = tile.shfl(mine[s].vlen, l);
                if (!as_is && dt_deletes(active, x.ts)) continue;
                if (!mc.present || !reconcile_keep_left(P, mc, x)) mc = x;
            }
        }
        npresent_merged += mc.present;
        MCell pc = purge_cell(pg, mc);
        npresent += pc.present;
        if (lane == 0) s_cells[c] = pc;
    }
}

template <int G, int S, bool EMIT>
__device__ void process_partition_tile(const cg::thread_block_tile<G>& tile, const CParams& P, const uint64_t* __restrict__ contrib, uint64_t c0, uint32_t m,
                                       const uint64_t* __restrict__ part_upos, const uint64_t* __restrict__ pbase,
                                       const uint64_t* __restrict__ part_kp, const uint16_t* __restrict__ part_klen, const int64_t* __restrict__ part_tok,
                                       uint8_t* dout, uint64_t dcap, uint64_t dpos, uint8_t* iout, uint32_t nblocks_final, uint32_t ipay_final, uint32_t ixs_cap,
                                       MCell* s_cells, PartOut& out, PartStats& st, int& err, StatAcc* acc = nullptr) {
    const int lane = tile.thread_rank();
    const bool multi = m > 1;
    if (P.ncx || P.ctr_mask || P.sctr_mask) { err = PERR_UNSUPPORTED; return; }        // multi-cell columns: the thread kernels only (compact.cu routes every fan-in there)
    Purger pg{P.now, P.gc_before, purge_threshold(P, contrib, c0, pbase, part_tok)};
    Cur cur[S]; bool have[S]; DT my_pd[S];
    int lerr = 0;
    uint64_t my_key_off = 0; uint32_t my_klen = 0;
#pragma unroll
    for (int s = 0; s < S; s++) {
        uint32_t v = lane + G * s;
        have[s] = v < m; cur[s].done = true; cur[s].pos = cur[s].next = cur[s].end = 0; cur[s].src = 0; my_pd[s] = dt_live();
        cur[s].flags = cur[s].kind = cur[s].n = cur[s].ext = cur[s].fast = 0; cur[s].k0 = 0; cur[s].ck_rel = 0; cur[s].ckend_rel = cur[s].body_rel = 0;
        if (have[s]) {
            uint64_t e = contrib[c0 + v];
            int src = (int)((e >> 56) & 0x7F); uint64_t g = pbase[src] + (e & 0xFFFFFFFFFFull);
            uint64_t pos = part_upos[g], end = part_upos[g + 1];
            Rd r{P.U, pos, end, 0};
            uint32_t kl = r.be16();
            {   // Index.db <-> Data.db consistency for keys up to 8 bytes (longer keys are compared in full by K2)
                uint64_t pre = load_be64(P.U + pos + 2);
                if (kl < 8) pre &= kl ? (~0ull << (8 * (8 - kl))) : 0ull;
                if (kl != part_klen[g] || pre != part_kp[g]) lerr = PERR_CORRUPT;
            }
            r.skip(kl);
            my_pd[s] = read_partition_dt(r);
            if (r.err) lerr = r.err;
            else if (kl > 8 && !P.partitioner && murmur3_token(P.U + pos + 2, kl) != part_tok[g]) lerr = PERR_CORRUPT;      // see process_partition
            if (v == 0) { my_key_off = pos + 2; my_klen = kl; }
            cur[s].src = (uint8_t)src; cur[s].pos = r.p; cur[s].end = end; cur[s].next = r.p; cur[s].done = false;
        }
    }
    // partition deletion = max over the sources (collectPartitionLevelDeletion :465-482)
    DT pdel = dt_live();
#pragma unroll
    for (int s = 0; s < S; s++) if (dt_supersedes(my_pd[s], pdel)) pdel = my_pd[s];
#pragma unroll
    for (int d = G / 2; d; d >>= 1) { DT o; o.mfda = tshfl64<G>(tile, pdel.mfda, lane ^ d); o.ldt = tshfl64<G>(tile, pdel.ldt, lane ^ d); if (dt_supersedes(o, pdel)) pdel = o; }
    const uint64_t key_off = tshflu64<G>(tile, my_key_off, 0); const uint32_t klen = tile.shfl(my_klen, 0);
    if (tile.any(lerr != 0)) { err = PERR_CORRUPT; return; }
    const DT out_pdel = pg.dt(pdel) ? dt_live() : pdel;

    PWriter<EMIT> w;
    const bool ixs = EMIT && ixs_cap != 0;           // scratch pass with a promoted-index slot (see process_partition)
    w.d.base = dout; w.d.pos = 0; w.d.on = (lane == 0); w.d.cap = dcap; w.ix.on = (lane == 0) && EMIT && iout && (ixs || nblocks_final > 1); w.ix.cap = ixs ? (uint64_t)ixs_cap : ~0ull;
    w.start = 0; w.header_len = 0; w.prev_row_start = 0; w.block_start = 0;
    w.nblocks = 0; w.nblocks_final = nblocks_final; w.started = false; w.have_first = false; w.open_marker = dt_live(); w.rows_out = 0;
    w.first = CkRef{0, 0, 0, 0}; w.last = w.first; w.acc = (lane == 0) ? acc : nullptr;      // every lane runs the writer, lane 0 stores and counts
    {
        uint32_t fixed = 2 + klen + vint_size(dpos) + vint_size(ipay_final);
        w.ix.base = nullptr; w.ix.pos = 0;
        w.ix_entry = (EMIT && iout && !ixs) ? iout : nullptr; w.ix_fixed = fixed + (dt_is_live(out_pdel) ? 1 : 12) + vint_size(nblocks_final);
        w.ix_offs = (EMIT && iout) ? iout + fixed + ipay_final - 4 * nblocks_final : nullptr;
        if (ixs) { w.ix_offs = iout + IXS_HEAD; w.ix.base = iout + IXS_HEAD + 4 * (size_t)nblocks_final; }
    }
    // ---- static rows: merged and purged before the clustered rows (see process_partition) -----------------------------------------------
    if (P.nstat > 0) {
        bool sne[S]; unsigned sg[S]; unsigned anys = 0;
#pragma unroll
        for (int s = 0; s < S; s++) {
            sne[s] = false;
            if (have[s] && P.in[cur[s].src].nstat > 0) { int e = static_load(P, cur[s], &sne[s]); if (e) { lerr = e; sne[s] = false; } }
        }
        if (tile.any(lerr != 0)) { err = tile.any(lerr == PERR_UNSUPPORTED) ? PERR_UNSUPPORTED : PERR_CORRUPT; return; }
#pragma unroll
        for (int s = 0; s < S; s++) { sg[s] = tile.ballot(sne[s]); anys |= sg[s]; }
        if (anys) {
            Live sinfo; DT sdel; int npm, np;
            tile_merge_rows<G, S>(tile, P, pg, cur, sg, m == 1 && dt_is_live(pdel), pdel, true, s_cells, sinfo, sdel, npm, np, lerr);
            tile.sync();
            if (tile.any(lerr != 0)) { err = tile.any(lerr == PERR_UNSUPPORTED) ? PERR_UNSUPPORTED : PERR_CORRUPT; return; }
            if (!(live_is_empty(sinfo) && dt_is_live(sdel) && npm == 0)) {
                if (pg.live(sinfo)) sinfo = live_empty();
                if (pg.dt(sdel)) sdel = dt_live();
                if (!(live_is_empty(sinfo) && dt_is_live(sdel) && np == 0)) pw_start(w, P, key_off, klen, out_pdel, s_cells, sinfo, sdel, np);
            }
            tile.sync();
        }
#pragma unroll
        for (int s = 0; s < S; s++) if (have[s] && P.in[cur[s].src].nstat > 0) cur[s].pos = cur[s].next;      // step over the static rows
    }

#pragma unroll
    for (int s = 0; s < S; s++) if (have[s]) cur_load(P, cur[s], lerr);
    DT my_open[S]; bool my_has_open[S];
#pragma unroll
    for (int s = 0; s < S; s++) { my_open[s] = dt_live(); my_has_open[s] = false; }
    DT cur_open = dt_live();                       // open deletion in the merged stream (uniform)
    uint64_t merged_unf = 0;

    for (;;) {
        if (tile.any(lerr != 0)) { err = tile.any(lerr == PERR_UNSUPPORTED) ? PERR_UNSUPPORTED : PERR_CORRUPT; return; }
        unsigned act[S]; unsigned anyact = 0;
#pragma unroll
        for (int s = 0; s < S; s++) { act[s] = tile.ballot(have[s] && !cur[s].done); anyact |= act[s]; }
        if (!anyact) break;
        // ---- tournament: find the smallest head and everything equal to it ------------------------------------------------------
        int ls = 0, ll = 0;
#pragma unroll
        for (int s = S - 1; s >= 0; s--) if (act[s]) { ls = s; ll = __ffs(act[s]) - 1; }
        unsigned grp[S]; Cur L;
        for (;;) {
            L = tshfl_cur<G>(tile, (S == 2 && ls == 1) ? cur[S - 1] : cur[0], ll);
            int c[S]; unsigned less[S]; unsigned anyless = 0;
#pragma unroll
            for (int s = 0; s < S; s++) {
                c[s] = 1;
                if (have[s] && !cur[s].done) c[s] = (lane == ll && s == ls) ? 0 : cmp_heads(P, cur[s], L);
                less[s] = tile.ballot(c[s] < 0); anyless |= less[s];
            }
            if (anyless) {
#pragma unroll
                for (int s = S - 1; s >= 0; s--) if (less[s]) { ls = s; ll = __ffs(less[s]) - 1; }
                continue;
            }
#pragma unroll
            for (int s = 0; s < S; s++) grp[s] = tile.ballot(c[s] == 0);
            break;
        }
        int gcount = 0;
#pragma unroll
        for (int s = 0; s < S; s++) gcount += __popc(grp[s]);
        bool ing[S];
#pragma unroll
        for (int s = 0; s < S; s++) ing[s] = (grp[s] >> lane) & 1;

        if (!(L.flags & 0x02)) {
            // ---- rows: Row.Merger.merge (multi source) or pass-through (single source, TrivialOneToOne) -----------------------------
            DT active = multi ? (dt_is_live(cur_open) ? pdel : cur_open) : dt_live();
            const bool as_is = !multi || (gcount == 1 && dt_is_live(active));
            Live info; DT del; int npresent_merged, npresent;
            tile_merge_rows<G, S>(tile, P, pg, cur, grp, as_is, active, false, s_cells, info, del, npresent_merged, npresent, lerr);
            tile.sync();
            bool haverow = !(live_is_empty(info) && dt_is_live(del) && npresent_merged == 0);
            if (haverow && !tile.any(lerr != 0)) {
                merged_unf++;
                if (pg.live(info)) info = live_empty();
                if (pg.dt(del)) del = dt_live();
                if (!(live_is_empty(info) && dt_is_live(del) && npresent == 0)) {
                    CkRef ck{L.pos + L.ck_rel, L.ckend_rel - L.ck_rel, K_CLUSTERING, L.n};
                    if (!w.started) pw_start(w, P, key_off, klen, out_pdel);
                    write_row(w, P, ck, info, del, s_cells, npresent);
                }
            }
            tile.sync();
        } else {
            // ---- markers ------------------------------------------------------------------------------------------------------------
            DT mc_ = dt_live(), mo_ = dt_live();
            uint8_t kind = L.kind;
            CkRef ck{L.pos + L.ck_rel, L.ckend_rel - L.ck_rel, L.kind, L.n};
            bool emit = false;
            if (!multi) {
                DT a = dt_live(), b = dt_live();
                if (ing[0]) read_marker_dts(P, cur[0], a, b, lerr);
                mc_ = tshfl_dt<G>(tile, a, ll); mo_ = tshfl_dt<G>(tile, b, ll);
                emit = true;
            } else {
                // RangeTombstoneMarker.Merger.merge :94-153 — per-source open markers live in the owning lane's registers
                int last_l = 0, last_s = 0;
#pragma unroll
                for (int s = 0; s < S; s++) {
                    if (ing[s]) {
                        DT a, b; read_marker_dts(P, cur[s], a, b, lerr);
                        bool is_open = kind_is_boundary(cur[s].kind) || kind_is_start(cur[s].kind);
                        my_has_open[s] = is_open; if (is_open) my_open[s] = b;
                    }
                    if (grp[s]) { last_s = s; last_l = 31 - __clz(grp[s]); }
                }
                DT best = dt_live(); bool anyopen = false;
#pragma unroll
                for (int s = 0; s < S; s++) if (my_has_open[s] && (!anyopen || dt_supersedes(my_open[s], best))) { best = my_open[s]; anyopen = true; }
#pragma unroll
                for (int d = G / 2; d; d >>= 1) {
                    DT o = tshfl_dt<G>(tile, best, lane ^ d); bool oa = tile.shfl((int)anyopen, lane ^ d);
                    if (oa && (!anyopen || dt_supersedes(o, best))) { best = o; anyopen = true; }
                }
                DT now_open = (anyopen && dt_supersedes(best, pdel)) ? best : dt_live();
                if (!dt_eq(cur_open, now_open)) {
                    // `bound` = clustering of the last marker added (:99-103)
                    Cur B = tshfl_cur<G>(tile, (S == 2 && last_s == 1) ? cur[S - 1] : cur[0], last_l);
                    bool before = kind_vs_clustering(B.kind) < 0;
                    ck = CkRef{B.pos + B.ck_rel, B.ckend_rel - B.ck_rel, 0, B.n};
                    if (dt_is_live(cur_open)) { kind = before ? K_INCL_START : K_EXCL_START; mo_ = now_open; }
                    else if (dt_is_live(now_open)) { kind = before ? K_EXCL_END : K_INCL_END; mc_ = cur_open; }
                    else { kind = before ? K_EXCL_END_INCL_START : K_INCL_END_EXCL_START; mc_ = cur_open; mo_ = now_open; }
                    emit = true;
                }
                cur_open = now_open;
            }
            if (emit && !tile.any(lerr != 0)) {
                merged_unf++;
                ck.kind = kind;
                if (purge_marker(pg, ck.kind, mc_, mo_)) { if (!w.started) pw_start(w, P, key_off, klen, out_pdel); write_marker(w, P, ck, mc_, mo_); }
            }
        }
#pragma unroll
        for (int s = 0; s < S; s++) if (ing[s]) { cur[s].pos = cur[s].next; cur_load(P, cur[s], lerr); }
    }
    if (tile.any(lerr != 0)) { err = tile.any(lerr == PERR_UNSUPPORTED) ? PERR_UNSUPPORTED : PERR_CORRUPT; return; }
    if (!w.started && !dt_is_live(out_pdel)) pw_start(w, P, key_off, klen, out_pdel);
    out.dsize = 0; out.ipay = 0; out.nblk = 0; out.ihead = 2 + klen; out.ovf = 0; out.cells = 0;
    st.merged_unfiltereds += merged_unf;
    if (w.started) {
        if (w.acc) out.cells = w.acc->part_cells;
        w.d.u8(0x01);
        if (w.rows_out && w.have_first) pw_add_index_block(w, P);
        out.dsize = w.d.pos; out.nblk = w.nblocks;
        if (w.nblocks > 1) out.ipay = vint_size(w.header_len) + (dt_is_live(out_pdel) ? 1 : 12) + vint_size(w.nblocks) + (uint32_t)w.ix.pos + 4 * w.nblocks;
        st.rows_out += w.rows_out;
        out.ovf = (EMIT && w.d.pos > dcap) ? 1 : 0;
        if (ixs) {
            if (w.nblocks > 1) {
                if (w.ix.pos > w.ix.cap || w.nblocks > nblocks_final) out.ovf = 1;
                if (lane == 0) { ((int64_t*)iout)[0] = out_pdel.mfda; ((int64_t*)iout)[1] = out_pdel.ldt; ((int64_t*)iout)[2] = (int64_t)w.header_len; }
            }
        } else if (EMIT && !iout && w.nblocks > 1) out.ovf = 1;
        else if (EMIT && iout && lane == 0) {
            Sink<true> e{iout, 0, true, ~0ull};
            e.be16(klen); e.copy(P.U + key_off, klen); e.vint(dpos); e.vint(ipay_final);
            if (nblocks_final > 1) { e.vint(w.header_len); write_partition_dt(e, out_pdel); e.vint(nblocks_final); }
        }
    }
}

} // namespace b200c
