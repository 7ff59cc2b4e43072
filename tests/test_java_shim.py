"""The JNI binding (java/b200c_jni.c) without a JDK: it must compile against include/b200c.h (behind a minimal stand-in for <jni.h>), export
one function per native method B200C.java declares, and its sizeof/offsetof table — which B200C.Layout consumes positionally — must agree
with the ctypes mirror of the same structs (cassandra_b200/native.py) and have exactly as many entries as Layout reads."""
import ctypes as C, os, re, subprocess, pytest
from cassandra_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

@pytest.fixture(scope="module")
def jni():
    out = os.path.join(ROOT, "tests", "native", "_build", "libb200c_jni_stub.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    r = subprocess.run(["gcc", "-shared", "-fPIC", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "tests", "native", "jni_stub"), "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "java", "b200c_jni.c"), "-L", os.path.join(ROOT, "cassandra_b200"), "-lb200compact", "-Wl,-rpath," + os.path.join(ROOT, "cassandra_b200"), "-o", out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    native.lib()
    return C.CDLL(out)

def java_natives():
    src = open(os.path.join(ROOT, "java", "org", "apache", "cassandra", "db", "compaction", "B200C.java")).read()
    return re.findall(r"public static native\s+[\w\[\]]+\s+(\w+)\(", src)

def test_every_native_method_has_its_c_function(jni):
    names = java_natives()
    assert len(names) >= 20
    for n in names:
        assert hasattr(jni, "Java_org_apache_cassandra_db_compaction_B200C_" + n), n

def test_layout_table_matches_the_ctypes_mirror(jni):
    jni.b200c_jni_layout.restype = C.POINTER(C.c_int32)
    n = C.c_int(); tab = jni.b200c_jni_layout(C.byref(n)); got = [tab[i] for i in range(n.value)]
    src = open(os.path.join(ROOT, "java", "org", "apache", "cassandra", "db", "compaction", "B200C.java")).read()
    layout = src[src.index("public static final class Layout"):]
    assert layout.count("next()") - 1 == len(got)            # one next() per table entry (+ its own definition)
    def offs(cls, fields): return [getattr(cls, f).offset for f in fields]
    want = [C.sizeof(x) for x in (native.Input, native.Manifest, native.Output, native.Result, native.Progress, native.SSTableStats, native.Corruption)]
    want += offs(native.Input, ["data", "data_len", "index", "index_len", "chunk_offsets", "nchunks", "data_length", "compressor", "chunk_len", "max_compressed_len",
                                "ncolumns", "column_map", "header_stats", "level", "summary_positions", "nsummary", "nstatic_columns", "static_column_map"])
    want += offs(native.Manifest, ["abi_version", "ninputs", "inputs", "nclustering", "clustering", "ncolumns", "columns", "nstatic_columns", "out_stats", "out_compressor",
                                   "out_chunk_len", "out_max_compressed_len", "column_index_size", "now_in_sec", "gc_before", "purge_max_timestamp", "tombstone_option",
                                   "enforce_strict_liveness", "token_lo", "token_hi", "max_sstable_bytes", "partitioner", "npurge_ranges", "purge_range_hi",
                                   "purge_range_max_ts", "bloom_hash_count", "min_index_interval", "bloom_words", "static_columns"])
    want += offs(native.Output, ["data", "data_cap", "data_len", "index", "index_cap", "index_len", "chunk_offsets", "chunk_cap", "nchunks", "data_length", "digest",
                                 "partitions", "rows", "key_buf", "key_cap", "first_key_len", "last_key_len", "filter", "filter_cap", "filter_len", "summary",
                                 "summary_cap", "summary_len", "stats"])
    want += offs(native.Result, ["noutputs_cap", "noutputs", "outputs", "bytes_read", "bytes_in_range", "bytes_written", "total_source_rows", "input_partitions",
                                 "merged_row_counts", "required_data_cap", "required_index_cap", "required_chunk_cap", "corruption", "kernel_ms", "total_ms"])
    want += offs(native.SSTableStats, ["min_timestamp", "max_timestamp", "min_local_deletion_time", "max_local_deletion_time", "min_ttl", "max_ttl", "total_rows",
                                       "total_columns_set", "total_cells", "total_tombstones", "has_partition_level_deletions", "tdrop_overflow", "partition_size_hist",
                                       "cells_per_partition_hist", "ntdrop", "has_legacy_counter_shards", "tdrop_point", "tdrop_count", "hll_registers"])
    assert got == want

def test_java_sources_are_present_and_name_the_reference_hooks():
    base = os.path.join(ROOT, "java", "org", "apache", "cassandra")
    for rel, needle in (("db/compaction/GpuCompactionTask.java", "extends CompactionTask"), ("db/compaction/GpuSizeTieredCompactionStrategy.java", "extends SizeTieredCompactionStrategy"),
                        ("io/compress/GpuLZ4Compressor.java", "implements ICompressor"), ("io/compress/GpuSnappyCompressor.java", "extends GpuLZ4Compressor"),
                        ("db/compaction/B200C.java", "System.loadLibrary")):
        assert needle in open(os.path.join(base, rel)).read(), rel
