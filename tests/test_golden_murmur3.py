"""Reference-held vectors for the partition order and the key hash: tables written by real Cassandra releases under
Murmur3Partitioner (tests/golden/murmur3-tables, imported by tests/golden/tools/import_murmur3_tables.sh).
  * Index.db lists the keys in DecoratedKey order => the oracle's Murmur3 token of consecutive keys must increase
    (S/dht/Murmur3Partitioner.java:256-296, S/db/DecoratedKey.java:79-91);
  * Filter.db is BloomFilter.add of every key (S/utils/BloomFilter.java:79-122: indexes = |(h2 + i*h1) % capacity| over
    MurmurHash.hash3_x64_128(key, seed 0), S/utils/MurmurHash.java:178-260) => rebuilding the bit set from the keys with the oracle's
    hash must give the stored bytes exactly. The same holds for the ByteOrderedPartitioner `oa` tables (the filter hash does not
    depend on the partitioner)."""
import ctypes as C, glob, os, struct, pytest
import oracle_lib as O
from cassandra_b200.io.sstable import _vint

ROOT = os.path.dirname(os.path.abspath(__file__))

def index_keys(ix):
    o = 0; keys = []
    while o < len(ix):
        kl = (ix[o] << 8) | ix[o + 1]; keys.append(bytes(ix[o + 2:o + 2 + kl])); p = o + 2 + kl
        _, p = _vint(ix, p); ps, p = _vint(ix, p); o = p + ps
    return keys

def hash3(key):
    out = (C.c_uint64 * 2)(); O.lib().orc_murmur3_x64_128(key, len(key), 0, out); return out[0], out[1]

def s64(v): return v - (1 << 64) if v >= 1 << 63 else v

def bloom_bytes(keys, hash_count, nwords):
    """BloomFilter.add over an OffHeapBitSet of nwords 64-bit words; Java long arithmetic (wrapping add, truncating remainder)."""
    cap = nwords * 64; bits = bytearray(nwords * 8)
    for k in keys:
        h0, h1 = hash3(k)
        base, inc = s64(h1), s64(h0)                       # setIndexes(indexes[1], indexes[0], ...)
        for _ in range(hash_count):
            r = abs(base) % cap; r = -r if base < 0 else r  # Java % truncates toward zero
            idx = abs(r)
            bits[idx >> 3] |= 1 << (idx & 7)
            base = s64((base + inc) & ((1 << 64) - 1))
    return bytes(bits)

def parse_filter(b, old_format):
    hc, nw = struct.unpack_from(">ii", b, 0); raw = b[8:8 + 8 * nw]
    assert len(b) == 8 + 8 * nw
    if old_format:                                           # serializeOldBfFormat: each word written as a big-endian long
        raw = b"".join(raw[i:i + 8][::-1] for i in range(0, len(raw), 8))
    return hc, nw, raw

def murmur3_tables():
    return sorted(glob.glob(os.path.join(ROOT, "golden", "murmur3-tables", "*", "*", "*-big-Index.db")))

def test_fixtures_present():
    assert len(murmur3_tables()) >= 6

@pytest.mark.parametrize("path", murmur3_tables(), ids=lambda p: "/".join(p.split(os.sep)[-3:-1]))
def test_murmur3_tables_are_in_oracle_token_order(path):
    keys = index_keys(open(path, "rb").read())
    toks = [O.token(k) for k in keys]
    assert toks == sorted(toks) and len(set(toks)) == len(toks), (keys, toks)

def test_some_murmur3_table_is_not_byte_ordered():
    """the vectors must be able to tell Murmur3 order from byte order"""
    assert any(index_keys(open(p, "rb").read()) != sorted(index_keys(open(p, "rb").read())) for p in murmur3_tables())

@pytest.mark.parametrize("path", murmur3_tables() + sorted(glob.glob(os.path.join(ROOT, "golden", "legacy-sstables", "oa", "legacy_tables", "*", "*-big-Index.db"))),
                         ids=lambda p: "/".join(p.split(os.sep)[-3:-1]))
def test_filter_db_is_rebuilt_exactly_from_the_keys(path):
    version = os.path.basename(path)[:2]
    hc, nw, raw = parse_filter(open(path.replace("Index.db", "Filter.db"), "rb").read(), old_format=version < "na")
    keys = index_keys(open(path, "rb").read())
    assert bloom_bytes(keys, hc, nw) == raw
