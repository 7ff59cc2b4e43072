/*
 * Table option:  ... WITH compression = {'class': 'org.apache.cassandra.io.compress.GpuLZ4Compressor'};
 * ICompressor (S/io/compress/ICompressor.java:28-86) over b200c_compress / b200c_uncompress. The bytes are those of LZ4Compressor
 * (4-byte little-endian uncompressed length + LZ4 block, S/io/compress/LZ4Compressor.java:113-190), bit-exact with liblz4's
 * LZ4_compress_default, so stock nodes read the files; CompressionInfo.db persists the class SIMPLE name
 * (CompressionMetadata.java:379), which this class reports as "LZ4Compressor" through the parameters it is registered with
 * (see GpuCompactionTask.compressionParamsForDisk).
 * The single-buffer calls are PCIe-latency bound (one 16 KiB chunk per round trip); the fast path is the batched writer used by
 * GpuCompactionTask (B200C.compressChunks: whole Data.db image + chunk offsets + digest in one call).
 * Not compiled in the build image (no JDK); see B200C.java for the compile line.
 */
package org.apache.cassandra.io.compress;

import java.io.IOException;
import java.nio.ByteBuffer;
import java.util.Collections;
import java.util.EnumSet;
import java.util.Map;
import java.util.Set;

import org.apache.cassandra.db.compaction.B200C;

public class GpuLZ4Compressor implements ICompressor
{
    private static final GpuLZ4Compressor INSTANCE = new GpuLZ4Compressor();

    /** CompressionParams instantiates compressors through this static factory by reflection (S/schema/CompressionParams.java:266-283) */
    public static GpuLZ4Compressor create(Map<String, String> options)
    {
        String type = options == null ? null : options.get(LZ4Compressor.LZ4_COMPRESSOR_TYPE);
        if (type != null && !LZ4Compressor.LZ4_FAST_COMPRESSOR.equals(type))
            throw new IllegalArgumentException("GpuLZ4Compressor implements lz4_compressor_type=fast only");
        return INSTANCE;
    }

    protected int compressorId() { return B200C.COMP_LZ4; }

    @Override
    public int initialCompressedBufferLength(int chunkLength)
    {
        return B200C.initialCompressedBufferLength(compressorId(), chunkLength);
    }

    /** consume input.position()..limit(), write at output.position(), advance both positions, leave the limits alone (ICompressor.java:45-51) */
    @Override
    public void compress(ByteBuffer input, ByteBuffer output) throws IOException
    {
        requireDirect(input, output);
        int n = B200C.compress(B200C.context(), compressorId(), B200C.address(input) + input.position(), input.remaining(),
                               B200C.address(output) + output.position(), output.remaining());
        if (n < 0)
            throw new IOException("b200c_compress: " + B200C.lastError(B200C.context()));
        input.position(input.limit());
        output.position(output.position() + n);
    }

    @Override
    public void uncompress(ByteBuffer input, ByteBuffer output) throws IOException
    {
        requireDirect(input, output);
        int n = B200C.uncompress(B200C.context(), compressorId(), B200C.address(input) + input.position(), input.remaining(),
                                 B200C.address(output) + output.position(), output.remaining());
        if (n < 0)
            throw new IOException("b200c_uncompress: " + B200C.lastError(B200C.context()));
        input.position(input.limit());
        output.position(output.position() + n);
    }

    /** byte[] variant used by a few legacy call sites: staged through direct buffers */
    @Override
    public int uncompress(byte[] input, int inputOffset, int inputLength, byte[] output, int outputOffset) throws IOException
    {
        ByteBuffer in = ByteBuffer.allocateDirect(inputLength);
        in.put(input, inputOffset, inputLength).flip();
        ByteBuffer out = ByteBuffer.allocateDirect(output.length - outputOffset);
        uncompress(in, out);
        out.flip();
        int n = out.remaining();
        out.get(output, outputOffset, n);
        return n;
    }

    @Override
    public BufferType preferredBufferType() { return BufferType.OFF_HEAP; }

    @Override
    public boolean supports(BufferType bufferType) { return bufferType == BufferType.OFF_HEAP; }

    @Override
    public Set<String> supportedOptions() { return Collections.singleton(LZ4Compressor.LZ4_COMPRESSOR_TYPE); }

    @Override
    public Set<Uses> recommendedUses() { return EnumSet.of(Uses.GENERAL); }   // not FAST_COMPRESSION: flushes compress chunk by chunk

    private static void requireDirect(ByteBuffer a, ByteBuffer b) throws IOException
    {
        if (!a.isDirect() || !b.isDirect())
            throw new IOException("GpuLZ4Compressor needs direct buffers (preferredBufferType() is OFF_HEAP)");
    }
}
