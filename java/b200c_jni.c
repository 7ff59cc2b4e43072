/*
 * b200c_jni.c — JNI binding of include/b200c.h for org.apache.cassandra.db.compaction.B200C (java/org/apache/cassandra/db/compaction/B200C.java).
 * One C function per native method; every argument is a raw address or a scalar, nothing touches the Java heap on the hot path.
 * Build (needs a JDK, which the build image of this repository does not have):
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude java/b200c_jni.c -Lcassandra_b200 -lb200compact -o libb200compact_jni.so
 * tests/test_java_shim.py compiles this file against a minimal stand-in for <jni.h> (tests/native/jni_stub/jni.h) so that signature drift
 * against include/b200c.h is caught without a JDK, and checks layout() against the ctypes mirror.
 */
#include <jni.h>
#include <stddef.h>
#include <stdint.h>
#include "b200c.h"

#define CLS(name) Java_org_apache_cassandra_db_compaction_B200C_##name
#define CTX(x) ((b200c_ctx*)(intptr_t)(x))
#define PTR(t, x) ((t*)(intptr_t)(x))

JNIEXPORT jint JNICALL CLS(abiVersion)(JNIEnv* e, jclass c) { (void)e; (void)c; return b200c_abi_version(); }
JNIEXPORT jint JNICALL CLS(deviceCount)(JNIEnv* e, jclass c) { (void)e; (void)c; return b200c_device_count(); }
JNIEXPORT jlong JNICALL CLS(create)(JNIEnv* e, jclass c, jint device, jlong ws) { (void)e; (void)c; return (jlong)(intptr_t)b200c_create(device, (size_t)ws); }
JNIEXPORT void JNICALL CLS(destroy)(JNIEnv* e, jclass c, jlong ctx) { (void)e; (void)c; b200c_destroy(CTX(ctx)); }
JNIEXPORT jstring JNICALL CLS(lastError)(JNIEnv* e, jclass c, jlong ctx) { (void)c; return (*e)->NewStringUTF(e, b200c_last_error(CTX(ctx))); }
JNIEXPORT jint JNICALL CLS(hostRegister)(JNIEnv* e, jclass c, jlong p, jlong n) { (void)e; (void)c; return b200c_host_register(PTR(void, p), (size_t)n); }
JNIEXPORT jint JNICALL CLS(hostUnregister)(JNIEnv* e, jclass c, jlong p) { (void)e; (void)c; return b200c_host_unregister(PTR(void, p)); }

JNIEXPORT jint JNICALL CLS(compact)(JNIEnv* e, jclass c, jlong ctx, jlong man, jlong res, jint flags)
{ (void)e; (void)c; return b200c_compact(CTX(ctx), PTR(const b200c_manifest, man), PTR(b200c_result, res), flags); }
JNIEXPORT jint JNICALL CLS(poll)(JNIEnv* e, jclass c, jlong ctx, jlong prog) { (void)e; (void)c; return b200c_poll(CTX(ctx), PTR(b200c_progress, prog)); }
JNIEXPORT jint JNICALL CLS(pollInputs)(JNIEnv* e, jclass c, jlong ctx, jlong pos, jint n) { (void)e; (void)c; return b200c_poll_inputs(CTX(ctx), PTR(uint64_t, pos), n); }
JNIEXPORT void JNICALL CLS(cancel)(JNIEnv* e, jclass c, jlong ctx) { (void)e; (void)c; b200c_cancel(CTX(ctx)); }
JNIEXPORT void JNICALL CLS(cancelReset)(JNIEnv* e, jclass c, jlong ctx) { (void)e; (void)c; b200c_cancel_reset(CTX(ctx)); }
JNIEXPORT jlong JNICALL CLS(token)(JNIEnv* e, jclass c, jint partitioner, jlong key, jint len) { (void)e; (void)c; return b200c_token(partitioner, PTR(const uint8_t, key), (uint32_t)len); }

JNIEXPORT jlong JNICALL CLS(compressBound)(JNIEnv* e, jclass c, jint comp, jlong n, jint chunk) { (void)e; (void)c; return (jlong)b200c_compress_bound(comp, (uint64_t)n, chunk); }
JNIEXPORT jint JNICALL CLS(compressChunks)(JNIEnv* e, jclass c, jlong ctx, jint comp, jlong in, jlong n, jint chunk, jint maxc, jlong out, jlong cap, jlong outLen, jlong offs, jlong digest, jint flags)
{ (void)e; (void)c; return b200c_compress_chunks(CTX(ctx), comp, PTR(const uint8_t, in), (uint64_t)n, chunk, maxc, PTR(uint8_t, out), (uint64_t)cap, PTR(uint64_t, outLen), PTR(uint64_t, offs), PTR(uint32_t, digest), flags); }
JNIEXPORT jint JNICALL CLS(decompressChunks)(JNIEnv* e, jclass c, jlong ctx, jint comp, jlong data, jlong dlen, jlong offs, jlong nch, jint chunk, jint maxc, jlong ulen, jlong out, jint verify, jlong where, jint flags)
{ (void)e; (void)c; return b200c_decompress_chunks(CTX(ctx), comp, PTR(const uint8_t, data), (uint64_t)dlen, PTR(const uint64_t, offs), (uint64_t)nch, chunk, maxc, (uint64_t)ulen, PTR(uint8_t, out), verify, PTR(b200c_corruption, where), flags); }
JNIEXPORT jint JNICALL CLS(initialCompressedBufferLength)(JNIEnv* e, jclass c, jint comp, jint chunk) { (void)e; (void)c; return b200c_initial_compressed_buffer_length(comp, chunk); }
JNIEXPORT jint JNICALL CLS(compress)(JNIEnv* e, jclass c, jlong ctx, jint comp, jlong in, jint n, jlong out, jint cap)
{ (void)e; (void)c; return b200c_compress(CTX(ctx), comp, PTR(const uint8_t, in), n, PTR(uint8_t, out), cap); }
JNIEXPORT jint JNICALL CLS(uncompress)(JNIEnv* e, jclass c, jlong ctx, jint comp, jlong in, jint n, jlong out, jint cap)
{ (void)e; (void)c; return b200c_uncompress(CTX(ctx), comp, PTR(const uint8_t, in), n, PTR(uint8_t, out), cap); }
JNIEXPORT jlong JNICALL CLS(address)(JNIEnv* e, jclass c, jobject buf) { (void)c; return (jlong)(intptr_t)(*e)->GetDirectBufferAddress(e, buf); }

/* sizeof / offsetof table consumed by B200C.Layout in exactly this order */
#define OFF(t, f) (jint)offsetof(t, f)
static const jint LAYOUT[] = {
    (jint)sizeof(b200c_input), (jint)sizeof(b200c_manifest), (jint)sizeof(b200c_output), (jint)sizeof(b200c_result), (jint)sizeof(b200c_progress),
    (jint)sizeof(b200c_sstable_stats), (jint)sizeof(b200c_corruption),
    OFF(b200c_input, data), OFF(b200c_input, data_len), OFF(b200c_input, index), OFF(b200c_input, index_len), OFF(b200c_input, chunk_offsets), OFF(b200c_input, nchunks),
    OFF(b200c_input, data_length), OFF(b200c_input, compressor), OFF(b200c_input, chunk_len), OFF(b200c_input, max_compressed_len), OFF(b200c_input, ncolumns),
    OFF(b200c_input, column_map), OFF(b200c_input, header_stats), OFF(b200c_input, level), OFF(b200c_input, summary_positions), OFF(b200c_input, nsummary),
    OFF(b200c_input, nstatic_columns), OFF(b200c_input, static_column_map),
    OFF(b200c_manifest, abi_version), OFF(b200c_manifest, ninputs), OFF(b200c_manifest, inputs), OFF(b200c_manifest, nclustering), OFF(b200c_manifest, clustering),
    OFF(b200c_manifest, ncolumns), OFF(b200c_manifest, columns), OFF(b200c_manifest, nstatic_columns), OFF(b200c_manifest, out_stats), OFF(b200c_manifest, out_compressor),
    OFF(b200c_manifest, out_chunk_len), OFF(b200c_manifest, out_max_compressed_len), OFF(b200c_manifest, column_index_size), OFF(b200c_manifest, now_in_sec),
    OFF(b200c_manifest, gc_before), OFF(b200c_manifest, purge_max_timestamp), OFF(b200c_manifest, tombstone_option), OFF(b200c_manifest, enforce_strict_liveness),
    OFF(b200c_manifest, token_lo), OFF(b200c_manifest, token_hi), OFF(b200c_manifest, max_sstable_bytes), OFF(b200c_manifest, partitioner), OFF(b200c_manifest, npurge_ranges),
    OFF(b200c_manifest, purge_range_hi), OFF(b200c_manifest, purge_range_max_ts), OFF(b200c_manifest, bloom_hash_count), OFF(b200c_manifest, min_index_interval),
    OFF(b200c_manifest, bloom_words), OFF(b200c_manifest, static_columns),
    OFF(b200c_output, data), OFF(b200c_output, data_cap), OFF(b200c_output, data_len), OFF(b200c_output, index), OFF(b200c_output, index_cap), OFF(b200c_output, index_len),
    OFF(b200c_output, chunk_offsets), OFF(b200c_output, chunk_cap), OFF(b200c_output, nchunks), OFF(b200c_output, data_length), OFF(b200c_output, digest),
    OFF(b200c_output, partitions), OFF(b200c_output, rows), OFF(b200c_output, key_buf), OFF(b200c_output, key_cap), OFF(b200c_output, first_key_len),
    OFF(b200c_output, last_key_len), OFF(b200c_output, filter), OFF(b200c_output, filter_cap), OFF(b200c_output, filter_len), OFF(b200c_output, summary),
    OFF(b200c_output, summary_cap), OFF(b200c_output, summary_len), OFF(b200c_output, stats),
    OFF(b200c_result, noutputs_cap), OFF(b200c_result, noutputs), OFF(b200c_result, outputs), OFF(b200c_result, bytes_read), OFF(b200c_result, bytes_in_range),
    OFF(b200c_result, bytes_written), OFF(b200c_result, total_source_rows), OFF(b200c_result, input_partitions), OFF(b200c_result, merged_row_counts),
    OFF(b200c_result, required_data_cap), OFF(b200c_result, required_index_cap), OFF(b200c_result, required_chunk_cap), OFF(b200c_result, corruption),
    OFF(b200c_result, kernel_ms), OFF(b200c_result, total_ms),
    OFF(b200c_sstable_stats, min_timestamp), OFF(b200c_sstable_stats, max_timestamp), OFF(b200c_sstable_stats, min_local_deletion_time),
    OFF(b200c_sstable_stats, max_local_deletion_time), OFF(b200c_sstable_stats, min_ttl), OFF(b200c_sstable_stats, max_ttl), OFF(b200c_sstable_stats, total_rows),
    OFF(b200c_sstable_stats, total_columns_set), OFF(b200c_sstable_stats, total_cells), OFF(b200c_sstable_stats, total_tombstones),
    OFF(b200c_sstable_stats, has_partition_level_deletions), OFF(b200c_sstable_stats, tdrop_overflow), OFF(b200c_sstable_stats, partition_size_hist),
    OFF(b200c_sstable_stats, cells_per_partition_hist), OFF(b200c_sstable_stats, ntdrop), OFF(b200c_sstable_stats, has_legacy_counter_shards), OFF(b200c_sstable_stats, tdrop_point), OFF(b200c_sstable_stats, tdrop_count),
    OFF(b200c_sstable_stats, hll_registers),
};
JNIEXPORT jintArray JNICALL CLS(layout)(JNIEnv* e, jclass c)
{
    (void)c;
    const jsize n = (jsize)(sizeof(LAYOUT) / sizeof(LAYOUT[0]));
    jintArray a = (*e)->NewIntArray(e, n);
    if (a) (*e)->SetIntArrayRegion(e, a, 0, n, LAYOUT);
    return a;
}
/* for tests without a JVM: the same table through a plain C symbol */
const jint* b200c_jni_layout(int* n) { *n = (int)(sizeof(LAYOUT) / sizeof(LAYOUT[0])); return LAYOUT; }
