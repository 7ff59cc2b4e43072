"""CPU parity of the engine's K4 source. tests/native/k4_host.cc compiles cassandra_b200/csrc/partition.cuh — the code every GPU
thread of k_partition_thr runs: row merge, reconciliation, purge, serialisation, promoted index — with g++ behind a few intrinsic
shims and feeds it from host-side stand-ins for K1-K3. Its merged Data stream and Index.db must equal the oracle's byte for byte.
(The GPU tests prove the same for the CUDA build; this one runs without a GPU and lets K4 logic be developed and fuzzed on the CPU.)"""
import ctypes as C, os, shutil, struct, subprocess, random
import numpy as np, pytest
import oracle_lib as O
from sstable_builder import *
from synth_util import synth_tables, decompress_output
from cassandra_b200 import native
from cassandra_b200.db.compaction import CompactionTask, CompactionController

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NOW = 1700000000
I32 = lambda v: struct.pack(">i", v)

@pytest.fixture(scope="module")
def k4lib():
    if shutil.which("g++") is None or not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"): pytest.skip("needs g++ and the CUDA headers")
    out = os.path.join(ROOT, "tests", "native", "_build", "libk4host.so")
    srcs = [os.path.join(ROOT, "tests", "native", "k4_host.cc"), os.path.join(ROOT, "oracle", "codec.cc")]
    deps = srcs + [os.path.join(ROOT, "cassandra_b200", "csrc", f) for f in ("partition.cuh", "common.cuh")] + [os.path.join(ROOT, "include", "b200c.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        r = subprocess.run(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-fno-strict-aliasing", "-I/usr/local/cuda/include", "-Wno-attributes",
                            "-Wno-unknown-pragmas", "-o", out] + srcs, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
    L = C.CDLL(out)
    L.k4host_compact.restype = C.c_int
    L.k4host_compact.argtypes = [C.POINTER(native.Manifest), C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64),
                                 C.POINTER(C.c_uint64), C.c_char_p, C.c_int]
    return L

def host_k4(L, task):
    m = task.build_manifest()
    total = sum(t.compression.data_length for t in task.inputs)
    ucap = int(total * 1.5) + 4096; icap = sum(len(t.index) for t in task.inputs) * 2 + 4096
    u = np.zeros(ucap, dtype=np.uint8); ix = np.zeros(icap, dtype=np.uint8); ul = C.c_uint64(); il = C.c_uint64(); st = (C.c_uint64 * 3)(); err = C.create_string_buffer(256)
    rc = L.k4host_compact(C.byref(m), u.ctypes.data, ucap, C.byref(ul), ix.ctypes.data, icap, C.byref(il), st, err, 256)
    assert rc == 0, err.value
    return bytes(u[:ul.value]), bytes(ix[:il.value]), list(st)

def check(L, tables, controller, **kw):
    for g, t in enumerate(tables): t.generation = g
    want = CompactionTask(tables, controller, **kw).execute(O.OracleEngine())
    data, index, st = host_k4(L, CompactionTask(tables, controller, **kw))
    w = want.outputs[0]
    assert data == decompress_output(w), "Data stream of the K4 host build differs from the oracle"
    assert index == w.index
    assert st[0] == want.stats["total_source_rows"] and st[2] == w.partitions and st[1] == w.rows
    return want

@pytest.mark.parametrize("schema,n,universe,rpp,cis", [(0, 4, 6000, 0, 65536), (0, 16, 1500, 0, 65536), (1, 3, 40, 600, 65536), (1, 4, 50, 200, 4096)])
def test_k4_source_matches_oracle_on_synthetic_tables(k4lib, schema, n, universe, rpp, cis):
    tabs = synth_tables(schema, n, 0x4B34 + schema + n, universe, rows_per_partition=rpp, column_index_size=cis)
    check(k4lib, tabs, CompactionController(NOW), column_index_size=cis)
    check(k4lib, tabs, CompactionController(0, 0), column_index_size=cis)
    check(k4lib, tabs, CompactionController(NOW, overlapping_min_timestamp=1600000000000000 + 1500000000), column_index_size=cis)

def test_k4_source_on_golden_files(k4lib, golden_dir):
    from cassandra_b200.io.sstable import SSTable
    for name in ("legacy_oa_simple", "legacy_oa_clust", "legacy_oa_simple_counter", "legacy_oa_clust_counter"):
        t = SSTable.open(os.path.join(golden_dir, "oa", "legacy_tables", name, "oa-1-big-"))
        check(k4lib, [t], CompactionController(NOW, 0))
        if name.endswith("counter"):                       # three-way self merge: every context meets its equal
            base = os.path.join(golden_dir, "oa", "legacy_tables", name, "oa-1-big-")
            check(k4lib, [SSTable.open(base, 1), SSTable.open(base, 2), SSTable.open(base, 3)], CompactionController(NOW, 0))

def test_k4_source_mixed_types_and_token_range(k4lib):
    rng = random.Random(11)
    s = Schema(["LongType", "UTF8Type"], [("a", "LongType"), ("b", "UTF8Type"), ("c", "Int32Type")])
    keys = sorted({bytes(rng.getrandbits(8) for _ in range(rng.choice([1, 3, 8, 9, 17]))) for _ in range(120)} | {b""})
    tables = []
    for t in range(4):
        parts = []
        for k in keys:
            if rng.random() < 0.5: continue
            us = []
            for ck in sorted({(rng.randint(-3, 3), rng.choice([b"", b"x", b"yy", b"zzzz" * 40])) for _ in range(rng.randint(1, 10))}):
                cells = [Cell(ci, 1000 + rng.randint(0, 5), v) for ci, v in ((0, struct.pack(">q", rng.getrandbits(40))), (1, rng.choice([b"", b"hello", b"w" * 150])), (2, I32(rng.randint(-9, 9)))) if rng.random() < 0.7]
                us.append(Row((struct.pack(">q", ck[0]), ck[1]), cells, ts=1000 + rng.randint(0, 5) if (rng.random() < 0.8 or not cells) else NO_TS))
            parts.append(Partition(k, us, (1002, NOW) if rng.random() < 0.1 else None))
        tables.append(Builder(s, (1000 - t, 0, 0), column_index_size=1024).build(parts))
    check(k4lib, tables, CompactionController(NOW, 10**9), column_index_size=1024)
    check(k4lib, tables, CompactionController(NOW, 10**9), column_index_size=1024, token_range=(-(1 << 62), 1 << 61))

def _corruption_fuzz(lib_path, trials, counters=False):
    """child process body (runs under LD_PRELOAD=libasan): damaged Data.db contents must end in 'corrupt data' / 'unsupported' or in a
    normal result — never outside the buffers (ASan aborts the process on any out-of-bounds access of the K4 source)"""
    L = C.CDLL(lib_path)
    L.k4host_compact.restype = C.c_int
    L.k4host_compact.argtypes = [C.POINTER(native.Manifest), C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64),
                                 C.POINTER(C.c_uint64), C.c_char_p, C.c_int]
    rng = random.Random(77)
    s = Schema(["LongType", "UTF8Type"], [("a", "LongType"), ("b", "UTF8Type"), ("c", "Int32Type")])
    keys = sorted({bytes(rng.getrandbits(8) for _ in range(rng.choice([2, 8, 11]))) for _ in range(40)})
    def table(t):
        parts = []
        for k in keys:
            if rng.random() < 0.4: continue
            us = [Row((struct.pack(">q", a), b), [Cell(ci, 1000 + rng.randint(0, 5), v) for ci, v in ((0, struct.pack(">q", rng.getrandbits(40))), (1, rng.choice([b"", b"hello", b"w" * 90])), (2, I32(rng.randint(-9, 9)))) if rng.random() < 0.7],
                      ts=1000 + rng.randint(0, 5)) for a, b in sorted({(rng.randint(-3, 3), rng.choice([b"", b"x", b"zz" * 30])) for _ in range(rng.randint(1, 12))})]
            parts.append(Partition(k, us, (1002, NOW) if rng.random() < 0.1 else None))
        return Builder(s, (1000 - t, 0, 0), column_index_size=512).build(parts)
    tables = [table(t) for t in range(3)]
    if counters:                                     # counter tables: the damage lands in contexts, headers and shard ids (ctr_merge walks them)
        from counter_tables import counter_tables
        tables = counter_tables(5, ntables=3, nkeys=40, cis=512)
    for g, t in enumerate(tables): t.generation = g
    outcomes = {"ok": 0, "rejected": 0}
    for trial in range(trials):
        victim = tables[trial % len(tables)]
        good_image, good_offs = victim.data, list(victim.compression.chunk_offsets)
        raw = bytearray(victim.uncompressed)
        for _ in range(rng.randint(1, 4)):
            k = rng.randrange(len(raw)); raw[k] = rng.choice([0, 0xFF, raw[k] ^ (1 << rng.randrange(8)), rng.getrandbits(8)])
        image = bytearray(); offs = []
        for i in range(0, len(raw), 16384):
            c = O.chunk_compress(O.COMP_LZ4, bytes(raw[i:i + 16384])); offs.append(len(image)); image += c + struct.pack(">I", O.crc32(c))
        victim.data = bytes(image); victim.compression.chunk_offsets = offs
        try:
            task = CompactionTask(tables, CompactionController(NOW, 10**9), column_index_size=512)
            m = task.build_manifest()
            total = sum(t.compression.data_length for t in tables); ucap = total * 2 + 4096; icap = sum(len(t.index) for t in tables) * 2 + 4096
            u = np.zeros(ucap, dtype=np.uint8); ix = np.zeros(icap, dtype=np.uint8); ul = C.c_uint64(); il = C.c_uint64(); st = (C.c_uint64 * 3)(); err = C.create_string_buffer(256)
            rc = L.k4host_compact(C.byref(m), u.ctypes.data, ucap, C.byref(ul), ix.ctypes.data, icap, C.byref(il), st, err, 256)
            outcomes["ok" if rc == 0 else "rejected"] += 1
        finally:
            victim.data = good_image; victim.compression.chunk_offsets = good_offs
    print("k4 corruption fuzz done", outcomes)

def test_k4_source_survives_damaged_data_under_asan(k4lib, tmp_path):
    import subprocess, sys
    libasan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(libasan) or not os.path.exists(libasan): pytest.skip("no libasan")
    out = str(tmp_path / "libk4host_asan.so")
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-fsanitize=address", "-fno-omit-frame-pointer", "-fno-strict-aliasing", "-I/usr/local/cuda/include",
                        "-Wno-attributes", "-Wno-unknown-pragmas", "-o", out, os.path.join(ROOT, "tests", "native", "k4_host.cc"), os.path.join(ROOT, "oracle", "codec.cc")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    code = "import sys; sys.path.insert(0, 'tests'); sys.path.insert(0, '.'); import test_k4_host_build as T; T._corruption_fuzz(%r, 150); T._corruption_fuzz(%r, 150, True)" % (out, out)
    c = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, LD_PRELOAD=libasan, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1:abort_on_error=0"))
    assert c.returncode == 0 and "k4 corruption fuzz done" in c.stdout, (c.stdout + c.stderr)[-4000:]

@pytest.mark.parametrize("seed,wide,cis", [(1, False, 2048), (2, False, 65536), (3, True, 1024)])
def test_k4_source_static_rows(k4lib, seed, wide, cis):
    from static_tables import static_tables
    tabs = static_tables(seed, ntables=5, wide=wide, cis=cis)
    check(k4lib, tabs, CompactionController(NOW, 864000), column_index_size=cis)                 # old tombstones purge
    check(k4lib, tabs, CompactionController(NOW, 10**9), column_index_size=cis)                  # nothing purges
    check(k4lib, tabs, CompactionController(NOW, 864000, overlapping_min_timestamp=1015), column_index_size=cis)
    check(k4lib, tabs[:1], CompactionController(NOW, 864000), column_index_size=cis)             # single source: partition deletion still shadows the static row
    check(k4lib, tabs[2:3], CompactionController(NOW, 864000), column_index_size=cis)            # input without static columns

@pytest.mark.parametrize("seed,big,cis", [(1, False, 65536), (2, False, 65536), (3, True, 2048)])
def test_k4_source_complex_columns(k4lib, seed, big, cis):
    """multi-cell columns (map / set / list beside simple ones): complex deletions, cells merged in cell-path order (TimeUUID order for the list),
    purge, HAS_COMPLEX_DELETION / subset bitmap, wide partitions with a promoted index — the K4 source against the oracle"""
    from complex_tables import complex_tables
    tabs = complex_tables(seed, ntables=5 if not big else 3, nkeys=60 if not big else 6, cis=cis, big=big)
    check(k4lib, tabs, CompactionController(NOW, 864000), column_index_size=cis)
    check(k4lib, tabs, CompactionController(NOW, 10 ** 9), column_index_size=cis)          # nothing purgeable
    check(k4lib, tabs[:1], CompactionController(NOW, 864000), column_index_size=cis)       # single source: pass-through + purge
    check(k4lib, tabs, CompactionController(NOW, 0), column_index_size=cis)

@pytest.mark.parametrize("seed,big,cis,legacy,static", [(1, False, 65536, True, False), (2, False, 65536, False, False), (3, True, 2048, True, False),
                                                        (4, False, 65536, True, True), (5, True, 2048, True, True)])
def test_k4_source_counter_columns(k4lib, seed, big, cis, legacy, static):
    """counter columns: contexts merged shard by shard (global / local / remote rules), tombstones and empty values, cells under a deletion left out
    of the merge, the merged timestamp, wide partitions — the K4 source's K-way merge against the oracle's pairwise fold"""
    from counter_tables import counter_tables
    tabs = counter_tables(seed, ntables=5 if not big else 3, nkeys=60 if not big else 6, cis=cis, big=big, legacy=legacy, static=static)
    check(k4lib, tabs, CompactionController(NOW, 864000), column_index_size=cis)
    check(k4lib, tabs, CompactionController(NOW, 10 ** 9), column_index_size=cis)
    check(k4lib, tabs[:1], CompactionController(NOW, 864000), column_index_size=cis)
    check(k4lib, tabs, CompactionController(NOW, 0), column_index_size=cis)

def test_k4_source_refuses_contexts_the_reference_does_not_write(k4lib):
    """a header element that meets no shard / ids out of order: inside a merge the pairwise fold may hand such a context on unchanged where the K-way
    merge would rebuild it — refused (unsupported), never guessed; alone (no merge) the cell passes through as in the reference"""
    from counter_tables import SCTR, ctx as cctx, cid, G, T0
    good = cctx([(cid(1), 2, 2, G)])
    for odd in (struct.pack(">hh", 1, 5) + cid(1) + struct.pack(">qq", 1, 1),                         # element index 5, one shard
                struct.pack(">h", 0) + cid(2) + struct.pack(">qq", 1, 1) + cid(1) + struct.pack(">qq", 1, 1),      # ids descending
                struct.pack(">hh", -1, -32768) + cid(1) + struct.pack(">qq", 1, 1)):                  # "clear local shards" marker (negative count)
        t = [Builder(SCTR).build([Partition(b"k", [Row((I32(1),), [Cell(0, T0, v)])])]) for v in (odd, good)]
        for g_, tb in enumerate(t): tb.generation = g_
        with pytest.raises(AssertionError, match="unsupported"):
            host_k4(k4lib, CompactionTask(t, CompactionController(NOW)))
        data, _, _ = host_k4(k4lib, CompactionTask(t[:1], CompactionController(NOW)))
        assert odd in data

def test_k4_source_on_the_reference_counter_cell_answers(k4lib):
    """CounterCellTest.testReconcile's cases (pinned to their expected cells in tests/test_oracle_counter_kats.py) through the K4 source"""
    from counter_tables import SCTR, ctx as cctx, L as LOCAL
    lid = struct.pack(">QQ", 0x0123456789AB11EE, 0x8000000000000077)
    local = lambda count: cctx([(lid, 1, count, LOCAL)])
    dead = lambda ts, ldt: Cell.tombstone(0, ts, ldt)
    live = [Cell(0, 2, local(1)), Cell(0, 5, local(3)), Cell(0, 4, local(10)), Cell(0, 7, local(3))]
    for cells in ([dead(2, 5), dead(2, 10)], [dead(2, 5), Cell(0, 10, local(1))], [dead(6, 6), Cell(0, 5, local(1))], [dead(8, 8), Cell(0, 8, local(1))],
                  live[:2], live[:3], live, live + [dead(8, 8)]):
        tabs = [Builder(SCTR).build([Partition(b"k", [Row((I32(1),), [c])])]) for c in cells]
        check(k4lib, tabs, CompactionController(NOW, NOW + 10))
