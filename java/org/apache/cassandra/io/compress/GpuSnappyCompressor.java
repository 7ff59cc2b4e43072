/*
 * Snappy flavour of GpuLZ4Compressor: raw snappy blocks (S/io/compress/SnappyCompressor.java:77-105), byte-identical to Google snappy
 * 1.1.10 (tests/golden/snappy, tests/test_snappy_golden.py). Table option: compression = {'class': '...GpuSnappyCompressor'}.
 * Not compiled in the build image (no JDK); see B200C.java for the compile line.
 */
package org.apache.cassandra.io.compress;

import java.util.Collections;
import java.util.Map;
import java.util.Set;

import org.apache.cassandra.db.compaction.B200C;

public class GpuSnappyCompressor extends GpuLZ4Compressor
{
    private static final GpuSnappyCompressor INSTANCE = new GpuSnappyCompressor();

    public static GpuSnappyCompressor create(Map<String, String> options) { return INSTANCE; }

    @Override
    protected int compressorId() { return B200C.COMP_SNAPPY; }

    @Override
    public Set<String> supportedOptions() { return Collections.emptySet(); }
}
