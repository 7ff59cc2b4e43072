// synth/synth.cc — seeded generator of valid big-format `oa` SSTable inputs for the benchmark configs.
//
// Self-contained (no oracle/ and no engine code): it writes the *uncompressed* Data stream and Index.db of one input
// SSTable; the caller compresses the stream (tests: CPU oracle; bench.py: the engine's own b200c_compress_chunks) so that
// bench input synthesis never links the checker. Multi-threaded over token-order slices of the key universe.
//
// Schema N (mirrors T/microbench CompactionBench.java:57): (userid bigint, picid bigint, commentid bigint, PK(userid, picid)),
//   1-8 candidate rows per partition (geometric, mean ~3), each input holds a partition with probability p and a candidate
//   row with probability 3/4 => overlapping partitions AND rows (cell reconciliation).
// Schema W (wide time series): (sensor bigint, ts timestamp, v1 double, v2 double, tag text, PK(sensor, ts)), rows_per_partition
//   rows (1000 in cfg5, ~70 KB => 2 column-index blocks), 10 % late arrivals land on the previous input's time grid.
// Value mix: 5 % cell tombstones, 1 % row deletions, 5 % TTL cells (half expired at nowInSec), 0.1 % partition deletions,
//   0.2 % range tombstones (W only). Timestamps 1.6e15 + 1e9*sstable + U[0,1e9) µs. nowInSec = 1 700 000 000.
// Layouts written: SortedTablePartitionWriter.java:97-166, UnfilteredSerializer.java:151-305, Cell.java:268-305,
//   ClusteringPrefix.java:455-477, BigFormatPartitionWriter.java:128-251, RowIndexEntry.java:625-642, IndexInfo.java:107-117.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include <thread>
#include <algorithm>
#include <mutex>
#include <map>
#include <chrono>
#include <cstdio>

namespace {

const int64_t NOW = 1700000000, GC_GRACE = 864000;
inline uint64_t mix(uint64_t z) { z += 0x9e3779b97f4a7c15ULL; z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31); }
inline uint64_t h3(uint64_t s, uint64_t a, uint64_t b, uint64_t c) { return mix(mix(mix(mix(s) ^ a) ^ b) ^ c); }

inline uint64_t rotl64(uint64_t v, int n) { return (v << n) | (v >> (64 - n)); }
inline uint64_t fmix(uint64_t k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k; }
// Murmur3Partitioner token of an 8-byte key (MurmurHash.hash3_x64_128 tail case 8, sign-extended bytes: MurmurHash.java:214-233)
int64_t token8(const uint8_t* key) {
    uint64_t h1 = 0, h2 = 0, k1 = 0;
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    for (int i = 7; i >= 0; i--) k1 ^= (uint64_t)(int64_t)(int8_t)key[i] << (8 * i);
    k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    h1 ^= 8; h2 ^= 8; h1 += h2; h2 += h1; h1 = fmix(h1); h2 = fmix(h2); h1 += h2;
    int64_t v = (int64_t)h1; return v == INT64_MIN ? INT64_MAX : v;
}

struct Buf {
    std::vector<uint8_t> b;
    void u8(uint8_t v) { b.push_back(v); }
    void put(const void* p, size_t n) { const uint8_t* q = (const uint8_t*)p; b.insert(b.end(), q, q + n); }
    void be16(uint16_t v) { u8(v >> 8); u8((uint8_t)v); }
    void be32(uint32_t v) { u8(v >> 24); u8(v >> 16); u8(v >> 8); u8((uint8_t)v); }
    void be64(uint64_t v) { be32((uint32_t)(v >> 32)); be32((uint32_t)v); }
    static int vsize(uint64_t v) { return (639 - __builtin_clzll(v | 1) * 9) >> 6; }
    void vint(uint64_t v) {
        int size = vsize(v);
        if (size == 1) { u8((uint8_t)v); return; }
        if (size < 9) { uint64_t reg = (v << ((8 - size) << 3)) | ((uint64_t)(uint8_t)(~(0xff >> (size - 1))) << 56); for (int i = 0; i < size; i++) u8((uint8_t)(reg >> (56 - 8 * i))); return; }
        u8(0xFF); be64(v);
    }
    void vint32s(int32_t v) { vint((uint64_t)(int64_t)v); }
    size_t size() const { return b.size(); }
};

struct Universe { std::vector<std::pair<int64_t, uint64_t>> keys; };   // (token, key) sorted by token
std::mutex g_mu; std::map<std::pair<uint64_t, uint64_t>, Universe*> g_universes;

Universe* get_universe(uint64_t seed, uint64_t n, int threads) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_universes.find({seed, n});
    if (it != g_universes.end()) return it->second;
    Universe* u = new Universe(); u->keys.resize(n);
    auto work = [&](int t) {
        for (uint64_t i = t; i < n; i += threads) {
            uint64_t k = mix(seed ^ (i * 0x2545F4914F6CDD1DULL + 1));
            uint8_t kb[8]; for (int b = 0; b < 8; b++) kb[b] = (uint8_t)(k >> (56 - 8 * b));
            u->keys[i] = {token8(kb), k};
        }
    };
    std::vector<std::thread> th; for (int t = 0; t < threads; t++) th.emplace_back(work, t); for (auto& x : th) x.join();
    // bucket by the top 8 token bits, sort buckets in parallel
    std::vector<std::vector<std::pair<int64_t, uint64_t>>> bk(256);
    for (auto& kv : u->keys) bk[(uint8_t)((uint64_t)(kv.first ^ INT64_MIN) >> 56)].push_back(kv);
    std::vector<std::thread> th2;
    for (int t = 0; t < threads; t++) th2.emplace_back([&, t]() { for (int b = t; b < 256; b += threads) std::sort(bk[b].begin(), bk[b].end()); });
    for (auto& x : th2) x.join();
    size_t o = 0; for (auto& v : bk) { for (auto& kv : v) u->keys[o++] = kv; }
    // drop duplicate keys (astronomically unlikely, but keys must be unique)
    u->keys.erase(std::unique(u->keys.begin(), u->keys.end(), [](auto& a, auto& b) { return a.second == b.second; }), u->keys.end());
    g_universes[{seed, n}] = u;
    return u;
}

struct Cfg {
    int schema; uint64_t seed; int sstable; int nsstables; uint64_t universe; double p; int rows_per_partition; int column_index_size;
    int band_count; int band_overlap_l0;   // LCS shape: >0 => input s >= l0 covers only token band (s - l0) % band_count
};

struct IndexEntry { uint64_t key; uint64_t rel_pos; std::vector<uint8_t> promoted; };
struct Slice { Buf data; std::vector<IndexEntry> idx; uint64_t rows = 0, parts = 0; };

const char* WORDS[64] = {"alpha","bravo","charlie","delta","echo","foxtrot","golf","hotel","india","juliet","kilo","lima","mike","november","oscar","papa",
 "quebec","romeo","sierra","tango","uniform","victor","whiskey","xray","yankee","zulu","anode","boiler","cathode","dynamo","engine","flange",
 "gasket","heater","impeller","jacket","kelvin","louver","manifold","nozzle","orifice","piston","quench","rotor","stator","turbine","upstream","valve",
 "washer","xenon","yoke","zener","ampere","baud","candela","decibel","erg","farad","gauss","henry","joule","kilowatt","lumen","mole"};

struct Gen {
    const Cfg& c; int64_t min_ts, min_ldt; int32_t min_ttl;
    Gen(const Cfg& cfg) : c(cfg) { min_ts = 1600000000000000LL + 1000000000LL * cfg.sstable; min_ldt = NOW - 2 * GC_GRACE - 31 * 86400; min_ttl = 86400; }

    void clustering8(Buf& o, int64_t v) { o.u8(0); o.be64((uint64_t)v); }                  // header vint 0 + fixed 8-byte value
    void delta_dt(Buf& o, int64_t mfda, int64_t ldt) { o.vint((uint64_t)(mfda - min_ts)); o.vint32s((int32_t)(ldt - min_ldt)); }
    void part_dt(Buf& o, bool live, int64_t mfda, int64_t ldt) { if (live) o.u8(0x80); else { o.be64((uint64_t)mfda); o.be32((uint32_t)ldt); } }

    struct CellSpec { bool present; int kind; /*0 live 1 tombstone 2 expiring*/ int64_t ts; int64_t ldt; int32_t ttl; const uint8_t* val; int vlen; bool fixed; };

    // serialises one row; returns bytes appended. prev = previous-unfiltered size
    Buf scratch;
    void row(Buf& out, int64_t ck, int64_t row_ts, bool has_liveness, bool row_deleted, int64_t del_ldt, CellSpec* cells, int ncells, int ncols, uint64_t prev) {
        Buf& body = scratch; body.b.clear();
        int flags = 0;
        if (has_liveness) flags |= 0x04;
        if (row_deleted) flags |= 0x10;
        int present = 0; for (int i = 0; i < ncells; i++) present += cells[i].present;
        if (present == ncols) flags |= 0x20;
        if (has_liveness) body.vint((uint64_t)(row_ts - min_ts));
        if (row_deleted) delta_dt(body, row_ts, del_ldt);
        if (!(flags & 0x20)) { uint64_t missing = 0; for (int i = 0; i < ncells; i++) if (!cells[i].present) missing |= 1ull << i; body.vint(missing); }
        for (int i = 0; i < ncells; i++) {
            CellSpec& c = cells[i]; if (!c.present) continue;
            bool use_ts = has_liveness && c.ts == row_ts;
            int cf = (c.vlen ? 0 : 0x04) | (c.kind == 1 ? 0x01 : (c.kind == 2 ? 0x02 : 0)) | (use_ts ? 0x08 : 0);
            body.u8((uint8_t)cf);
            if (!use_ts) body.vint((uint64_t)(c.ts - min_ts));
            if (c.kind) body.vint32s((int32_t)(c.ldt - min_ldt));
            if (c.kind == 2) body.vint32s(c.ttl - min_ttl);
            if (c.vlen) { if (!c.fixed) body.vint((uint64_t)c.vlen); body.put(c.val, c.vlen); }
        }
        out.u8((uint8_t)flags); clustering8(out, ck);
        out.vint(body.size() + Buf::vsize(prev)); out.vint(prev); out.put(body.b.data(), body.size());
    }
    void marker(Buf& out, int kind, int64_t ck, int64_t mfda, int64_t ldt, uint64_t prev) {
        Buf& body = scratch; body.b.clear(); delta_dt(body, mfda, ldt);
        out.u8(0x02); out.u8((uint8_t)kind); out.be16(1); clustering8(out, ck);
        out.vint(body.size() + Buf::vsize(prev)); out.vint(prev); out.put(body.b.data(), body.size());
    }

    struct Block { int fkind, lkind; int64_t fck, lck; uint64_t off, width; bool open; int64_t omf, oldt; };
    std::vector<Block> blocks;
    void prefix(Buf& o, int kind, int64_t ck) { o.u8((uint8_t)kind); if (kind != 4) o.be16(1); clustering8(o, ck); }

    void partition(Slice& s, uint64_t key) {
        const Cfg& cfg = c;
        uint64_t hk = mix(cfg.seed ^ key);
        Buf& d = s.data;
        uint64_t start = d.size();
        uint8_t kb[8]; for (int b = 0; b < 8; b++) kb[b] = (uint8_t)(key >> (56 - 8 * b));
        uint64_t hps = h3(cfg.seed, key, cfg.sstable, 0x50);
        bool pdel = (hps % 1000) == 0;                                         // 0.1 % partition deletions
        int64_t base_ts = min_ts;
        int64_t pdel_ts = base_ts + (int64_t)(hps >> 20) % 500000000, pdel_ldt = NOW - 2 * GC_GRACE + (int64_t)((hps >> 8) % (2 * GC_GRACE));
        d.be16(8); d.put(kb, 8); part_dt(d, !pdel, pdel_ts, pdel_ldt);
        uint64_t header_len = d.size() - start, prev_start = 0;
        blocks.clear(); bool have_first = false; Block cur{}; bool open = false; int64_t omf = 0, oldt = 0; uint64_t nunf = 0;
        auto begin_unf = [&](int kind, int64_t ck) { uint64_t pos = d.size() - start; if (!have_first) { cur.fkind = kind; cur.fck = ck; cur.off = pos; have_first = true; } return pos; };
        auto end_unf = [&](int kind, int64_t ck, uint64_t pos) {
            prev_start = pos; cur.lkind = kind; cur.lck = ck; nunf++; s.rows++;
            if (d.size() - start - cur.off >= (uint64_t)cfg.column_index_size) { cur.width = d.size() - start - cur.off; cur.open = open; cur.omf = omf; cur.oldt = oldt; blocks.push_back(cur); have_first = false; }
        };
        if (cfg.schema == 0) {
            int nrows = 1 + (int)((hk >> 11) % 3) + (int)(((hk >> 17) % 4 == 0) ? (hk >> 23) % 5 : 0);
            if (nrows > 8) nrows = 8;
            bool any = false;
            for (int j = 0; j < nrows; j++) {
                uint64_t hr = h3(cfg.seed, key, (uint64_t)cfg.sstable * 64 + j, 0x52);
                bool present = (hr & 3) != 0 || (!any && j == nrows - 1);
                if (!present) continue; any = true;
                int64_t ck = (int64_t)j * 1000003 + (int64_t)(hk % 997);
                int64_t ts = base_ts + (int64_t)((hr >> 8) % 1000000000ULL);
                int m = (int)((hr >> 40) % 1000);
                uint8_t vb[8]; uint64_t val = h3(cfg.seed, key, j, cfg.sstable) & 0xFFFFFFFFFFULL; for (int b = 0; b < 8; b++) vb[b] = (uint8_t)(val >> (56 - 8 * b));
                CellSpec cell{true, 0, ts, 0, 0, vb, 8, true};
                bool row_deleted = false; int64_t del_ldt = 0; bool live = true;
                if (m < 10) { row_deleted = true; live = false; cell.present = false; del_ldt = NOW - 2 * GC_GRACE + (int64_t)((hr >> 13) % (2 * GC_GRACE)); }       // 1 % row deletions
                else if (m < 60) { cell.kind = 1; cell.vlen = 0; cell.ldt = NOW - 2 * GC_GRACE + (int64_t)((hr >> 13) % (2 * GC_GRACE)); }                          // 5 % cell tombstones
                else if (m < 110) { cell.kind = 2; cell.ttl = 86400 * (1 + (int)((hr >> 5) % 30)); cell.ldt = NOW - 10 * 86400 + (int64_t)((hr >> 13) % (20 * 86400)); }  // 5 % TTL, half expired
                uint64_t pos = begin_unf(4, ck);
                row(d, ck, ts, live, row_deleted, del_ldt, &cell, 1, 1, pos - prev_start);
                end_unf(4, ck, pos);
            }
        } else {
            int R = cfg.rows_per_partition;
            int64_t window = (int64_t)R * 1000;                                 // ms per input window
            int skip_until = -1; int64_t close_ck = 0;
            // two ordered passes: 10 % late arrivals (they land on the previous input's time grid, ascending) then the own window
            for (int pass = 0; pass < 2; pass++) {
                if (pass == 0 && cfg.sstable == 0) continue;
                int w = pass == 0 ? cfg.sstable - 1 : cfg.sstable;
                for (int j = 0; j < R; j++) {
                    uint64_t hr = h3(cfg.seed, key, (uint64_t)cfg.sstable * 4096 + j, 0x57);
                    bool late = (hr % 10) == 0 && cfg.sstable > 0;
                    if (pass == 1 && j <= skip_until) { if (j == skip_until) { uint64_t pos = begin_unf(6, close_ck); marker(d, 6, close_ck, omf, oldt, pos - prev_start); open = false; end_unf(6, close_ck, pos); skip_until = -1; } continue; }
                    if ((pass == 0) != late) continue;
                    int64_t ck = 1600000000000LL + (int64_t)w * window + (int64_t)j * 1000;
                    int64_t ts = base_ts + (int64_t)((hr >> 8) % 1000000000ULL);
                    int m = (int)((hr >> 40) % 1000);
                    if (m < 2 && pass == 1 && j + 3 < R) {                        // 0.2 % range tombstones: [ck_j, ck_{j+2}] deleted at this row's timestamp
                        omf = ts; oldt = NOW - 2 * GC_GRACE + (int64_t)((hr >> 13) % (2 * GC_GRACE));
                        uint64_t pos = begin_unf(1, ck); marker(d, 1, ck, omf, oldt, pos - prev_start); open = true; end_unf(1, ck, pos);
                        skip_until = j + 2; close_ck = ck + 2000; continue;
                    }
                    uint8_t v1[8], v2[8]; uint64_t a = h3(cfg.seed, key, j, 1), b2 = h3(cfg.seed, key, j, 2);
                    a = 0x4040000000000000ULL | ((a >> 20) << 8); b2 = 0x4059000000000000ULL | (b2 >> 24 << 12);
                    for (int q = 0; q < 8; q++) { v1[q] = (uint8_t)(a >> (56 - 8 * q)); v2[q] = (uint8_t)(b2 >> (56 - 8 * q)); }
                    char tag[64]; int tl = 0; uint64_t hw = h3(cfg.seed, key, j / 8, 3);
                    while (tl < 32) { const char* wd = WORDS[hw & 63]; hw >>= 6; int l = (int)strlen(wd); if (tl + l + 1 > 48) break; memcpy(tag + tl, wd, l); tl += l; tag[tl++] = '-'; }
                    CellSpec cells[3] = { {true, 0, ts, 0, 0, (const uint8_t*)tag, tl, false}, {true, 0, ts, 0, 0, v1, 8, true}, {true, 0, ts, 0, 0, v2, 8, true} };
                    bool row_deleted = false; int64_t del_ldt = 0; bool live = true;
                    if (m >= 10 && m < 20) { row_deleted = true; live = false; for (auto& cc : cells) cc.present = false; del_ldt = NOW - 2 * GC_GRACE + (int64_t)((hr >> 13) % (2 * GC_GRACE)); }
                    else if (m < 70) { cells[1].kind = 1; cells[1].vlen = 0; cells[1].ldt = NOW - 2 * GC_GRACE + (int64_t)((hr >> 13) % (2 * GC_GRACE)); }
                    else if (m < 120) { cells[0].kind = 2; cells[0].ttl = 86400 * (1 + (int)((hr >> 5) % 30)); cells[0].ldt = NOW - 10 * 86400 + (int64_t)((hr >> 13) % (20 * 86400)); }
                    else if (m < 150) { cells[2].present = false; }
                    uint64_t pos = begin_unf(4, ck);
                    row(d, ck, ts, live, row_deleted, del_ldt, cells, 3, 3, pos - prev_start);
                    end_unf(4, ck, pos);
                }
            }
        }
        d.u8(0x01);
        if (nunf && have_first) { cur.width = d.size() - start - cur.off; cur.open = open; cur.omf = omf; cur.oldt = oldt; blocks.push_back(cur); }
        IndexEntry e; e.key = key; e.rel_pos = start;
        if (blocks.size() > 1) {
            Buf infos; std::vector<uint32_t> offs;
            for (auto& b : blocks) {
                offs.push_back((uint32_t)infos.size());
                prefix(infos, b.fkind, b.fck); prefix(infos, b.lkind, b.lck);
                infos.vint(b.off); int64_t w = (int64_t)b.width - 65536; infos.vint(((uint64_t)w << 1) ^ (uint64_t)(w >> 63));
                infos.u8(b.open ? 1 : 0); if (b.open) part_dt(infos, false, b.omf, b.oldt);
            }
            Buf p; uint64_t size = Buf::vsize(header_len) + (pdel ? 12 : 1) + Buf::vsize(blocks.size()) + infos.size() + 4 * blocks.size();
            p.vint(size); p.vint(header_len); part_dt(p, !pdel, pdel_ts, pdel_ldt); p.vint(blocks.size()); p.put(infos.b.data(), infos.size());
            for (uint32_t o : offs) p.be32(o);
            e.promoted = std::move(p.b);
        }
        s.idx.push_back(std::move(e)); s.parts++;
    }
};

} // namespace

extern "C" {

struct synth_config {
    int32_t schema;            // 0 = N (narrow), 1 = W (wide time series)
    int32_t sstable;           // index of this input
    int32_t nsstables;
    int32_t rows_per_partition;// W only
    uint64_t seed;
    uint64_t universe;         // number of keys in the shared key universe
    double  p;                 // probability that an input holds a given partition
    int32_t column_index_size; // 65536
    int32_t threads;
    int32_t band_count;        // LCS: number of disjoint token bands (0 = none)
    int32_t l0_count;          // LCS: the first l0_count inputs overlap everything
};
struct synth_result {
    uint8_t* data; uint64_t data_len; uint8_t* index; uint64_t index_len;
    uint64_t partitions; uint64_t rows;
    int64_t min_timestamp; int64_t min_local_deletion_time; int32_t min_ttl; int32_t _pad;
    uint64_t* summary; uint64_t nsummary;      // Index.db offset of every 128th entry (what Summary.db holds at min_index_interval 128)
};

uint64_t synth_universe_for(int schema, uint64_t target_bytes, double p, int rows_per_partition) {
    double per_part = schema == 0 ? 62.8 : (14.0 + 78.0 * rows_per_partition);
    return (uint64_t)(target_bytes / per_part / p) + 1;
}

int synth_generate(const synth_config* cfg, synth_result* out) {
    int T = cfg->threads > 0 ? cfg->threads : (int)std::max(1u, std::thread::hardware_concurrency());
    Universe* u = get_universe(cfg->seed, cfg->universe, T);
    Cfg c{cfg->schema, cfg->seed, cfg->sstable, cfg->nsstables, cfg->universe, cfg->p, cfg->rows_per_partition, cfg->column_index_size, cfg->band_count, cfg->l0_count};
    size_t n = u->keys.size();
    int nslices = T * 4;
    std::vector<Slice> slices(nslices);
    uint64_t thr = cfg->p >= 1.0 ? UINT64_MAX : (uint64_t)(cfg->p * 18446744073709551615.0);
    auto work = [&](int t) {
        Gen g(c);
        for (int sl = t; sl < nslices; sl += T) {
            size_t lo = n * sl / nslices, hi = n * (sl + 1) / nslices;
            Slice& s = slices[sl];
            s.data.b.reserve((size_t)((hi - lo) * c.p * (c.schema == 0 ? 70 : 80.0 * c.rows_per_partition)) + 4096);
            for (size_t i = lo; i < hi; i++) {
                uint64_t key = u->keys[i].second;
                if (c.band_count > 0 && c.sstable >= c.band_overlap_l0) {       // LCS L1: one token band per input
                    int band = (int)(((unsigned __int128)(uint64_t)(u->keys[i].first ^ INT64_MIN) * (unsigned)c.band_count) >> 64);
                    if (band != (c.sstable - c.band_overlap_l0) % c.band_count) continue;
                }
                if (h3(c.seed, key, c.sstable, 0x11) > thr) continue;
                g.partition(s, key);
            }
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th; for (int t = 0; t < T; t++) th.emplace_back(work, t); for (auto& x : th) x.join();
    auto t1 = std::chrono::steady_clock::now();
    uint64_t total = 0, parts = 0, rows = 0; for (auto& s : slices) { total += s.data.size(); parts += s.parts; rows += s.rows; }
    uint8_t* data = (uint8_t*)malloc(total + 64);
    if (!data) return -1;
    std::vector<uint64_t> base(nslices); uint64_t o = 0; for (int i = 0; i < nslices; i++) { base[i] = o; o += slices[i].data.size(); }
    std::vector<std::thread> th3; for (int t = 0; t < T; t++) th3.emplace_back([&, t]() { for (int i = t; i < nslices; i += T) memcpy(data + base[i], slices[i].data.b.data(), slices[i].data.size()); });
    for (auto& x : th3) x.join();
    Buf index; std::vector<uint64_t> summary; uint64_t nentry = 0;
    for (int i = 0; i < nslices; i++) for (auto& e : slices[i].idx) {
        if ((nentry++ & 127) == 0) summary.push_back(index.size());
        index.be16(8); index.be64(e.key); index.vint(base[i] + e.rel_pos);
        if (e.promoted.empty()) index.vint(0); else index.put(e.promoted.data(), e.promoted.size());
    }
    uint8_t* idx = (uint8_t*)malloc(index.size() + 64); if (!idx) { free(data); return -1; }
    memcpy(idx, index.b.data(), index.size());
    Gen g(c);
    if (getenv("SYNTH_DEBUG")) fprintf(stderr, "synth: gen %.3fs, stitch+index %.3fs\n", std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count());
    out->data = data; out->data_len = total; out->index = idx; out->index_len = index.size(); out->partitions = parts; out->rows = rows;
    out->summary = (uint64_t*)malloc(summary.size() * 8 + 8); out->nsummary = summary.size();
    if (out->summary) memcpy(out->summary, summary.data(), summary.size() * 8);
    out->min_timestamp = g.min_ts; out->min_local_deletion_time = g.min_ldt; out->min_ttl = g.min_ttl; out->_pad = 0;
    return 0;
}
void synth_free(synth_result* r) { free(r->data); free(r->index); free(r->summary); r->data = r->index = nullptr; r->summary = nullptr; }
void synth_drop_universes() { std::lock_guard<std::mutex> lk(g_mu); for (auto& kv : g_universes) delete kv.second; g_universes.clear(); }

}
