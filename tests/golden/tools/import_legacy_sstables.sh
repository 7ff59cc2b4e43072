#!/bin/bash
# Copies the reference's golden SSTables (written by real Cassandra releases) into tests/golden/legacy-sstables.
# These are DATA fixtures (not source): for every big-format table the Data/CompressionInfo/Digest components
# (they pin LZ4 bytes + CRC32), and the complete component set for the current `oa` version (identity-compaction gate).
# Run in the build container only (needs /root/reference); the result is committed.
set -e
SRC=/root/reference/test/data/legacy-sstables
DST="$(dirname "$0")/../legacy-sstables"
for v in ma mb mc md me na nb nc oa; do
  find "$SRC/$v" -name '*-big-Data.db' | while read -r f; do
    d=$(dirname "$f"); rel=${d#$SRC/}; base=$(basename "$f" Data.db)
    mkdir -p "$DST/$rel"
    for c in Data.db CompressionInfo.db Digest.crc32; do [ -f "$d/$base$c" ] && cp "$d/$base$c" "$DST/$rel/"; done
    if [ "$v" = oa ]; then cp "$d"/"$base"* "$DST/$rel/"; fi
  done
done
chmod -R u+w "$DST"
