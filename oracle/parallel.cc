// oracle/parallel.cc — TEST INFRASTRUCTURE (CPU oracle). Not part of the product path.
//
// The CPU oracle with every host thread: ONE compaction cut into token ranges (what the reference itself supports through ranged
// scanners, S/db/compaction/AbstractCompactionStrategy.java:247-269, SSTableReader.getPositionsForRanges
// S/io/sstable/format/SSTableReader.java:724, and UCS's ShardedCompactionWriter S/db/compaction/unified/ShardedCompactionWriter.java:65-82),
// every range merged by the single-threaded oracle (compaction.cc) on its own thread, the merged uncompressed streams stitched in
// token order (partition records are position independent), and the stitched stream then pushed through CompressedSequentialWriter
// + ChecksumWriter chunk by chunk, chunks being independent. The result is byte for byte the output of the single-threaded oracle
// (tests/test_oracle_parallel.py). bench.py uses it (a) as the strongest CPU baseline for the benchmarked workload — the reference
// runs one such compaction on ONE thread — and (b) to verify the GPU output of the full-size benchmark run in seconds.
#include "codec.h"
#include "parallel.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cstring>
#include <thread>
#include <ctime>
#include <mutex>
#include <malloc.h>
#include <memory>

namespace oracle {

static int64_t token_of_entry_at(const b200c_input& in, uint64_t off, int partitioner) {
    if (off + 2 > in.index_len) return INT64_MAX;
    uint32_t kl = ((uint32_t)in.index[off] << 8) | in.index[off + 1];
    if (off + 2 + kl > in.index_len) return INT64_MAX;
    const uint8_t* key = in.index + off + 2;
    if (partitioner == B200C_PARTITIONER_MURMUR3) return murmur3_token(key, kl);
    uint64_t pre = 0; for (uint32_t q = 0; q < 8; q++) pre = (pre << 8) | (q < kl ? key[q] : 0);
    return (int64_t)(pre ^ 0x8000000000000000ull);
}

// interior splitters: tokens of evenly spaced Summary.db samples of the largest input (the host-side choice SURVEY §8e describes)
static std::vector<int64_t> pick_cuts(const b200c_manifest* m, int nranges) {
    std::vector<int64_t> cuts;
    int big = -1;
    for (int i = 0; i < m->ninputs; i++) if (m->inputs[i].summary_positions && m->inputs[i].nsummary && (big < 0 || m->inputs[i].nsummary > m->inputs[big].nsummary)) big = i;
    if (big < 0 || nranges < 2) return cuts;
    const b200c_input& in = m->inputs[big];
    for (int r = 1; r < nranges; r++) {
        uint64_t k = (uint64_t)((double)in.nsummary * r / nranges);
        if (k >= in.nsummary) break;
        int64_t t = token_of_entry_at(in, in.summary_positions[k], m->partitioner);
        if (t <= m->token_lo || t >= m->token_hi) continue;
        if (cuts.empty() || t > cuts.back()) cuts.push_back(t);
    }
    return cuts;
}

// Index.db entries of one range, data positions shifted by `base` (RowIndexEntry.serialize S/io/sstable/format/big/RowIndexEntry.java:468-473:
// u16 keyLen | key | vint position | vint32 payload size | payload; nothing in the payload is an absolute position)
static void rebase_index(const std::vector<uint8_t>& in, uint64_t base, std::vector<uint8_t>& out) {
    const uint8_t* p = in.data(); const uint8_t* end = p + in.size();
    out.clear(); out.reserve(in.size() + in.size() / 4 + 16);
    while (p < end) {
        uint32_t kl = ((uint32_t)p[0] << 8) | p[1];
        out.insert(out.end(), p, p + 2 + kl); p += 2 + kl;
        uint64_t pos, ps; int n = vint_read(p, end, &pos); p += n; n = vint_read(p, end, &ps); const uint8_t* sz = p; p += n;
        uint8_t t[9]; int w = vint_write(t, pos + base); out.insert(out.end(), t, t + w);
        out.insert(out.end(), sz, p + ps); p += ps;
    }
}

template <typename F> static void parallel_for(int nthreads, uint64_t n, F f) {
    std::atomic<uint64_t> next{0};
    auto work = [&]() { for (;;) { uint64_t i = next.fetch_add(1); if (i >= n) return; f(i); } };
    std::vector<std::thread> th;
    int t = (int)std::min<uint64_t>((uint64_t)std::max(1, nthreads), std::max<uint64_t>(n, 1));
    for (int k = 1; k < t; k++) th.emplace_back(work);
    work();
    for (auto& x : th) x.join();
}

struct ParallelTimes { double merge_ms, stitch_ms, compress_ms; double task_wall_ms, task_cpu_ms, task_max_ms; };      // task_*: summed over the range tasks of the merge phase

static int compact_parallel_impl(const b200c_manifest* m, b200c_result* res, int nthreads, int nranges, int max_ranges, ParallelTimes* tm, int64_t* sample_hi) {
    auto t0 = std::chrono::steady_clock::now();
    if (m->max_sstable_bytes) throw Unsupported{"parallel oracle: single output only"};
    if (res->noutputs_cap < 1 || !res->outputs) return B200C_EINVAL;
    std::vector<int64_t> T{m->token_lo};
    for (int64_t c : pick_cuts(m, nranges)) T.push_back(c);
    T.push_back(m->token_hi);
    int R = (int)T.size() - 1;
    if (max_ranges > 0 && max_ranges < R) R = max_ranges;                 // bounded sample: the first R ranges of the ring (a prefix of the output)
    if (sample_hi) *sample_hi = T[R];
    std::vector<RangeOut> ro(R); std::vector<b200c_result> rr(R);
    std::vector<int> rc(R, B200C_OK); std::vector<Corrupt> cerr(R); std::vector<std::string> uerr(R);
    std::vector<double> twall(R, 0), tcpu(R, 0);
    auto thread_cpu_ms = []() { timespec ts; clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    parallel_for(nthreads, (uint64_t)R, [&](uint64_t r) {
        auto w0 = std::chrono::steady_clock::now(); const double c0 = thread_cpu_ms();
        b200c_manifest mm = *m; mm.token_lo = T[r]; mm.token_hi = T[r + 1];
        memset(&rr[r], 0, sizeof(b200c_result)); rr[r].noutputs_cap = 1; rr[r].outputs = res->outputs;   // (outputs untouched in raw mode)
        try { rc[r] = compact_impl(&mm, &rr[r], &ro[r]); }
        catch (Corrupt& c) { rc[r] = B200C_ECORRUPT; cerr[r] = c; }
        catch (Unsupported& u) { rc[r] = B200C_EUNSUPPORTED; uerr[r] = u.what; }
        twall[r] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count(); tcpu[r] = thread_cpu_ms() - c0;
    });
    if (tm) { tm->task_wall_ms = tm->task_cpu_ms = tm->task_max_ms = 0; for (int r = 0; r < R; r++) { tm->task_wall_ms += twall[r]; tm->task_cpu_ms += tcpu[r]; tm->task_max_ms = std::max(tm->task_max_ms, twall[r]); } }
    for (int r = 0; r < R; r++) {
        if (rc[r] == B200C_ECORRUPT) throw cerr[r];
        if (rc[r] == B200C_EUNSUPPORTED) throw Unsupported{uerr[r]};
        if (rc[r] != B200C_OK) return rc[r];
    }
    auto t1 = std::chrono::steady_clock::now();
    // ---- stitch: stream offsets, counters ------------------------------------------------------------------------------------------------
    std::vector<uint64_t> ubase(R + 1, 0);
    for (int r = 0; r < R; r++) ubase[r + 1] = ubase[r] + ro[r].ustream.size();
    const uint64_t ulen = ubase[R]; const uint64_t L = (uint64_t)m->out_chunk_len; const uint64_t nch = (ulen + L - 1) / L;
    b200c_output& out = res->outputs[0];
    memset(res->merged_row_counts, 0, sizeof(res->merged_row_counts));
    res->bytes_read = 0; res->bytes_in_range = 0; res->total_source_rows = 0; res->input_partitions = 0; out.partitions = 0; out.rows = 0;
    for (int i = 0; i < m->ninputs; i++) res->bytes_read += m->inputs[i].data_length;
    for (int r = 0; r < R; r++) {
        res->bytes_in_range += rr[r].bytes_in_range; res->total_source_rows += rr[r].total_source_rows; res->input_partitions += rr[r].input_partitions;
        for (int k = 0; k < B200C_MAX_INPUTS; k++) res->merged_row_counts[k] += rr[r].merged_row_counts[k];
        out.partitions += ro[r].partitions; out.rows += ro[r].rows;
    }
    std::unique_ptr<uint8_t[]> whole_buf(new uint8_t[ulen + 64]);          // uninitialised: every byte is written by the copies below
    uint8_t* const whole = whole_buf.get();
    std::vector<std::vector<uint8_t>> idx(R);
    parallel_for(nthreads, (uint64_t)R, [&](uint64_t r) {
        if (!ro[r].ustream.empty()) memcpy(whole + ubase[r], ro[r].ustream.data(), ro[r].ustream.size());
        rebase_index(ro[r].index, ubase[r], idx[r]);
        std::vector<uint8_t>().swap(ro[r].ustream);
    });
    uint64_t ilen = 0; std::vector<uint64_t> ibase(R + 1, 0);
    for (int r = 0; r < R; r++) { ibase[r + 1] = ibase[r] + idx[r].size(); } ilen = ibase[R];
    auto t2 = std::chrono::steady_clock::now();
    // ---- CompressedSequentialWriter.flushData + ChecksumWriter over independent chunks (compaction.cc Writer::flush_chunk) -----------------
    const int comp = m->out_compressor; const int bound = chunk_max_compressed(comp, (int)L) + 64;
    const uint64_t G = 256;                                                // chunks per work item
    const uint64_t ngroups = (nch + G - 1) / G;
    std::vector<std::vector<uint8_t>> gdata(ngroups); std::vector<std::vector<uint32_t>> glen(ngroups);
    parallel_for(nthreads, ngroups, [&](uint64_t g) {
        std::vector<uint8_t> tmp(bound); auto& d = gdata[g]; auto& ln = glen[g];
        d.reserve(G * (L / 2));
        for (uint64_t ch = g * G; ch < std::min(nch, (g + 1) * G); ch++) {
            const uint8_t* src = whole + ch * L; int n = (int)std::min<uint64_t>(L, ulen - ch * L);
            int clen = chunk_compress(comp, src, n, tmp.data());
            const uint8_t* w = tmp.data(); int wlen = clen; std::vector<uint8_t> raw;
            if ((int64_t)clen >= (int64_t)m->out_max_compressed_len) {
                raw.assign(src, src + n);
                if ((int64_t)raw.size() < (int64_t)m->out_max_compressed_len) raw.resize(m->out_max_compressed_len, 0);
                w = raw.data(); wlen = (int)raw.size();
            }
            uint32_t crc = crc32_ieee(0, w, wlen);
            d.insert(d.end(), w, w + wlen);
            uint8_t cb[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
            d.insert(d.end(), cb, cb + 4);
            ln.push_back((uint32_t)wlen + 4);
        }
    });
    std::vector<uint64_t> gbase(ngroups + 1, 0);
    for (uint64_t g = 0; g < ngroups; g++) gbase[g + 1] = gbase[g] + gdata[g].size();
    const uint64_t dlen = gbase[ngroups];
    res->noutputs = 1; res->bytes_written = ulen;
    res->required_data_cap = dlen; res->required_index_cap = ilen; res->required_chunk_cap = nch;
    res->kernel_ms = 0; res->kernel_launches = 0; res->index_slow_path_inputs = 0;
    if (dlen > out.data_cap || ilen > out.index_cap || nch > out.chunk_cap) return B200C_ETOOSMALL;
    std::vector<uint32_t> gcrc(ngroups, 0);
    parallel_for(nthreads, ngroups, [&](uint64_t g) {
        if (!gdata[g].empty()) memcpy(out.data + gbase[g], gdata[g].data(), gdata[g].size());
        uint64_t o = gbase[g]; uint64_t ch = g * G;
        for (uint32_t l : glen[g]) { out.chunk_offsets[ch++] = o; o += l; }
        gcrc[g] = crc32_ieee(0, gdata[g].data(), gdata[g].size());
    });
    uint32_t digest = 0;
    for (uint64_t g = 0; g < ngroups; g++) digest = g ? crc32_combine(digest, gcrc[g], gdata[g].size()) : gcrc[g];
    parallel_for(nthreads, (uint64_t)R, [&](uint64_t r) { if (!idx[r].empty()) memcpy(out.index + ibase[r], idx[r].data(), idx[r].size()); });
    out.data_len = dlen; out.index_len = ilen; out.nchunks = nch; out.data_length = ulen; out.digest = digest;
    auto t3 = std::chrono::steady_clock::now();
    res->total_ms = std::chrono::duration<double, std::milli>(t3 - t0).count();
    if (tm) { tm->merge_ms = std::chrono::duration<double, std::milli>(t1 - t0).count(); tm->stitch_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
              tm->compress_ms = std::chrono::duration<double, std::milli>(t3 - t2).count(); }
    return B200C_OK;
}

} // namespace oracle

// nranges: token ranges the ring is cut into (>= threads for balance); max_ranges > 0: only the first max_ranges of them (bounded
// sample: the result is then the compaction of the token range (token_lo, cut[max_ranges]]). times_ms (optional, 6 doubles): merge / stitch / compress phase, then the range tasks' summed wall time, summed thread CPU time and the longest task.
// *sample_token_hi (optional) = upper token bound of what was compacted (token_hi unless max_ranges cut the ring short).
// Range tasks allocate and free a few MiB per source and per output piece. With glibc's defaults every such buffer is its own mmap: first-touch
// page faults for each task and an munmap (TLB shoot-down to every core of the process) when it ends — on the 128-thread GPU host that made 64
// threads slower than 16. Serving these sizes from the per-thread arenas and never trimming them lets a thread reuse its warm pages.
static void tune_allocator_once() {
    static std::once_flag once;
    std::call_once(once, [] { mallopt(M_MMAP_THRESHOLD, 32 << 20); mallopt(M_TRIM_THRESHOLD, 1 << 30); mallopt(M_TOP_PAD, 64 << 20); });
}

extern "C" int orc_compact_parallel(const b200c_manifest* m, b200c_result* res, int nthreads, int nranges, int max_ranges, double* times_ms, int64_t* sample_token_hi,
                                    char* errbuf, int errcap) {
    try {
        oracle::ParallelTimes tm{0, 0, 0, 0, 0, 0};
        tune_allocator_once();
        int rc = oracle::compact_parallel_impl(m, res, nthreads, nranges, max_ranges, &tm, sample_token_hi);
        if (times_ms) { times_ms[0] = tm.merge_ms; times_ms[1] = tm.stitch_ms; times_ms[2] = tm.compress_ms; times_ms[3] = tm.task_wall_ms; times_ms[4] = tm.task_cpu_ms; times_ms[5] = tm.task_max_ms; }
        return rc;
    }
    catch (oracle::Unsupported& u) { if (errbuf) snprintf(errbuf, errcap, "unsupported: %s", u.what.c_str()); return B200C_EUNSUPPORTED; }
    catch (oracle::Corrupt& c) {
        res->corruption.input = c.input; res->corruption.kind = c.kind; res->corruption.chunk = c.chunk; res->corruption.offset = c.offset;
        if (errbuf) snprintf(errbuf, errcap, "corrupt: %s", c.what.c_str());
        return B200C_ECORRUPT;
    }
}
