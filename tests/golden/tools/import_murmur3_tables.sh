#!/bin/bash
# Copies the reference's Murmur3Partitioner-ordered fixture tables (test/data/negative-ldts-invalid-deletions-test and
# negative-local-expiration-test: written by real Cassandra releases under the default partitioner) into tests/golden/murmur3-tables.
# DATA fixtures, not source. They are pre-`oa` (mc/nb/nc) so they cannot be identity-compacted, but their Index.db key ORDER pins
# Murmur3Partitioner.getToken / DecoratedKey.compareTo, their Filter.db pins MurmurHash.hash3_x64_128 + BloomFilter.add, and their
# Summary.db pins the index-summary layout. Run in the build container only (needs /root/reference); the result is committed.
set -e
DST="$(dirname "$0")/../murmur3-tables"
for SRC in /root/reference/test/data/negative-ldts-invalid-deletions-test /root/reference/test/data/negative-local-expiration-test; do
  for d in "$SRC"/*/; do
    name=$(basename "$SRC")/$(basename "$d")
    mkdir -p "$DST/$name"
    cp "$d"/*-big-{Index.db,Filter.db,Summary.db,Statistics.db,Data.db,CompressionInfo.db,Digest.crc32} "$DST/$name/" 2>/dev/null || true
  done
done
chmod -R u+w "$DST"
