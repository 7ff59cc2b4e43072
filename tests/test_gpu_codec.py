"""GPU parity tests of the chunk codec (K1/K5) through the C ABI: bit-exact against the CPU oracle and against the
reference's golden SSTables (Data.db bytes, chunk offsets, Digest.crc32)."""
import os, random, struct, zlib, pytest
import oracle_lib as O
from sstable_files import read_compression_info, split_chunks, find_tables
from test_oracle_codec import _corpus

pytestmark = pytest.mark.gpu

@pytest.fixture(scope="module")
def ctx():
    from cassandra_b200 import native
    c = native.Context(0)
    yield c
    c.close()

def _oracle_image(comp, stream, chunk_len):
    image = bytearray(); offs = []
    for i in range(0, len(stream), chunk_len):
        c = O.chunk_compress(comp, stream[i:i + chunk_len])
        offs.append(len(image)); image += c + struct.pack(">I", O.crc32(c))
    return bytes(image), offs, O.crc32(bytes(image))

def test_golden_sstables_recompress_bit_exact(ctx, golden_dir):
    """decompress every golden Data.db on the GPU (CRC verified) and recompress it: the Data.db image, the chunk offsets and
    Digest.crc32 written by real Cassandra releases must come back byte for byte."""
    from cassandra_b200 import native
    from cassandra_b200.io.compress import CompressionMetadata, LZ4Compressor, write_compressed, read_compressed
    n = 0
    for base in find_tables(golden_dir):
        info = read_compression_info(base + "CompressionInfo.db")
        data = open(base + "Data.db", "rb").read()
        mcl = info["max_compressed_length"] or native.INT32_MAX
        raw = ctx.decompress_chunks(native.COMP_LZ4, data, info["offsets"], info["data_length"], info["chunk_length"], mcl, True)
        want = b"".join(O.chunk_decompress(O.COMP_LZ4, c, u) for c, _, u in split_chunks(data, info))
        assert raw == want
        image, offs, digest = ctx.compress_chunks(native.COMP_LZ4, raw, info["chunk_length"], mcl)
        assert image == data and offs == info["offsets"]
        assert digest == int(open(base + "Digest.crc32").read())
        n += 1
    assert n >= 60

def test_compressioninfo_roundtrip_oa(ctx, golden_dir):
    from cassandra_b200.io.compress import CompressionMetadata, LZ4Compressor, write_compressed, read_compressed
    for base in find_tables(golden_dir, versions=("oa", "nb", "nc")):
        ci = open(base + "CompressionInfo.db", "rb").read(); data = open(base + "Data.db", "rb").read()
        meta = CompressionMetadata.parse(ci)
        assert meta.serialize() == ci
        stream = read_compressed(ctx, data, ci)
        image, ci2, digest = write_compressed(ctx, LZ4Compressor.create(ctx), stream, meta.chunk_length, meta.max_compressed_length)
        assert (image, ci2, digest) == (data, ci, open(base + "Digest.crc32").read())

@pytest.mark.parametrize("comp", [O.COMP_LZ4, O.COMP_SNAPPY])
def test_random_streams_match_oracle(ctx, comp):
    rng = random.Random(0xB200 + comp)
    pieces = _corpus(rng)
    for chunk_len in (16384, 4096, 65536):
        stream = b"".join(pieces)
        stream = stream[: (len(stream) // 7) * 7 + 3]              # ragged tail
        image, offs, digest = ctx.compress_chunks(comp, stream, chunk_len)
        w_image, w_offs, w_digest = _oracle_image(comp, stream, chunk_len)
        assert offs == w_offs
        assert image == w_image
        assert digest == w_digest == zlib.crc32(image)
        assert ctx.decompress_chunks(comp, image, offs, len(stream), chunk_len) == stream

def test_each_corpus_piece_as_single_chunk(ctx):
    from cassandra_b200.io.compress import LZ4Compressor, SnappyCompressor
    rng = random.Random(77)
    lz4, snp = LZ4Compressor.create(ctx), SnappyCompressor.create(ctx)
    for s in _corpus(rng):
        if len(s) > 65536: continue
        for comp_obj, cid in ((lz4, O.COMP_LZ4), (snp, O.COMP_SNAPPY)):
            c = comp_obj.compress(s)
            assert c == O.chunk_compress(cid, s), (cid, len(s))
            assert comp_obj.uncompress(c, max(len(s), 1)) == s

def test_corruption_is_detected(ctx):
    from cassandra_b200 import native
    rng = random.Random(5)
    stream = bytes(rng.choice(b"abcdefgh") for _ in range(100000))
    image, offs, _ = ctx.compress_chunks(native.COMP_LZ4, stream, 16384)
    bad = bytearray(image); bad[offs[3] + 10] ^= 0x40
    with pytest.raises(native.CorruptSSTableError) as e:
        ctx.decompress_chunks(native.COMP_LZ4, bytes(bad), offs, len(stream), 16384)
    assert e.value.corruption.chunk == 3 and e.value.corruption.kind == 1
    # without CRC verification the malformed block itself must be rejected or decode to different bytes, never crash
    try:
        out = ctx.decompress_chunks(native.COMP_LZ4, bytes(bad), offs, len(stream), 16384, verify_crc=False)
        assert out != stream
    except native.CorruptSSTableError as e2:
        assert e2.corruption.kind == 2

def test_max_compressed_length_raw_fallback(ctx):
    """CompressedSequentialWriterTest 'uncompressed chunk' case: incompressible chunks are stored raw when
    compressedLength >= maxCompressedLength; a short final chunk is zero padded up to maxCompressedLength."""
    from cassandra_b200 import native
    rng = random.Random(9)
    stream = bytes(rng.getrandbits(8) for _ in range(16384 * 2 + 100)) + b"\x00" * 16384
    mcl = 16384  # min_compress_ratio 1.0
    image, offs, digest = ctx.compress_chunks(native.COMP_LZ4, stream, 16384, mcl)
    assert digest == zlib.crc32(image)
    assert offs[1] - offs[0] == 16384 + 4 and image[:16384] == stream[:16384]
    assert ctx.decompress_chunks(native.COMP_LZ4, image, offs, len(stream), 16384, mcl) == stream

def test_large_stream_properties(ctx):
    """size-independent properties at a size the oracle would not be asked to check chunk by chunk:
    round trip + digest == CRC32 of the image + offsets strictly increasing and consistent with chunk CRC trailers."""
    from cassandra_b200 import native
    import numpy as np
    rng = np.random.default_rng(1)
    words = rng.integers(0, 256, size=(512, 24), dtype=np.uint8)
    idx = rng.integers(0, 512, size=(64 << 20) // 24)
    stream = words[idx].reshape(-1).tobytes()
    image, offs, digest = ctx.compress_chunks(native.COMP_LZ4, stream, 16384)
    assert digest == zlib.crc32(image)
    assert all(b > a for a, b in zip(offs, offs[1:]))
    for i in (0, 1, len(offs) // 2, len(offs) - 1):
        end = offs[i + 1] if i + 1 < len(offs) else len(image)
        assert struct.unpack(">I", image[end - 4:end])[0] == zlib.crc32(image[offs[i]:end - 4])
        assert image[offs[i]:end - 4] == O.chunk_compress(O.COMP_LZ4, stream[i * 16384:(i + 1) * 16384])
    assert ctx.decompress_chunks(native.COMP_LZ4, image, offs, len(stream), 16384) == stream

def test_snappy_equals_googles_library_on_the_golden_vectors(ctx):
    """tests/golden/snappy/vectors.json holds what Google's snappy library (>= 1.2.0 generation) produced; the GPU compressor in that
    generation (B200C_COMP_SNAPPY15) must give the same bytes, and both generations must round-trip through the GPU decompressor"""
    from cassandra_b200 import native
    from test_snappy_golden import vectors
    for name, data, want in vectors():
        if not data: continue
        cl = 1
        while cl < len(data): cl <<= 1
        image, offs, digest = ctx.compress_chunks(native.COMP_SNAPPY15, data, cl)
        assert image[:-4] == want, name
        assert ctx.decompress_chunks(native.COMP_SNAPPY15, image, offs, len(data), cl) == data, name
        img14, offs14, _ = ctx.compress_chunks(native.COMP_SNAPPY, data, cl)
        assert img14[:-4] == O.chunk_compress(O.COMP_SNAPPY, data), name
        assert ctx.decompress_chunks(native.COMP_SNAPPY, img14, offs14, len(data), cl) == data, name
