"""Snappy byte parity, pinned to the real library: the oracle's raw-snappy compressor must reproduce, byte for byte, what Google's snappy
library produced for the vectors in tests/golden/snappy/vectors.json (SnappyCompressor.compress -> Snappy.compress, S/io/compress/
SnappyCompressor.java:77-88; snappy-java wraps that C++ library), and decode them back. Where pyarrow is importable (it is in this image, here
and on the GPU box) a live differential run adds random inputs on top of the committed vectors."""
import base64, json, os, random, pytest
import oracle_lib as O
from snappy_vectors import inputs

ROOT = os.path.dirname(os.path.abspath(__file__))

def vectors():
    j = json.load(open(os.path.join(ROOT, "golden", "snappy", "vectors.json")))
    data = dict(inputs())
    return [(v["name"], data[v["name"]], base64.b64decode(v["compressed_b64"])) for v in j["vectors"]]

def test_vector_file_matches_the_generator():
    v = vectors()
    assert len(v) >= 25 and all(len(d) == n for (_, d, _), n in zip(v, [x["n"] for x in json.load(open(os.path.join(ROOT, "golden", "snappy", "vectors.json")))["vectors"]]))

@pytest.mark.parametrize("name,data,want", vectors(), ids=[v[0] for v in vectors()])
def test_oracle_snappy_equals_the_library(name, data, want):
    if len(data) == 0: pytest.skip("empty input: varint 0 only (checked in the codec tests)")
    got = O.chunk_compress(O.COMP_SNAPPY15, data)
    assert got == want
    assert O.chunk_decompress(O.COMP_SNAPPY15, want, len(data)) == data
    # the 1.1.x generation (what snappy-java 1.1.10.4 writes) differs only by its smaller hash table: it must decode to the same bytes
    assert O.chunk_decompress(O.COMP_SNAPPY, O.chunk_compress(O.COMP_SNAPPY, data), len(data)) == data

def test_live_differential_against_the_library():
    pa = pytest.importorskip("pyarrow")
    if not pa.Codec.is_available("snappy"): pytest.skip("pyarrow built without snappy")
    codec = pa.Codec("snappy"); rng = random.Random(77)
    alphabet = [bytes(rng.getrandbits(8) for _ in range(rng.randint(1, 9))) for _ in range(40)]
    for it in range(300):
        n = rng.choice([1, 7, 100, 1000, 4096, 16384, 16385, 40000, 65536])
        kind = rng.random()
        if kind < 0.3: d = bytes(rng.getrandbits(8) for _ in range(n))
        elif kind < 0.6: d = b"".join(rng.choice(alphabet) for _ in range(n))[:n]
        else: d = bytes(rng.choice(b"ab\x00\xff") for _ in range(n))
        assert O.chunk_compress(O.COMP_SNAPPY15, d) == codec.compress(d, asbytes=True), (it, n)
