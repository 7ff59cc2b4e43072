"""cassandra_b200 — host-side mirror of the Cassandra plugin surfaces for the B200-native compaction engine.

The product is libb200compact.so (hand-written sm_100a CUDA behind the C ABI in include/b200c.h). This package is the
thin host layer above it: ctypes bindings (`native`), the ICompressor / CompressionMetadata mirror (`io.compress`),
SSTable component readers/writers (`io.sstable`) and the CompactionTask mirror (`db.compaction`).
There is no CPU fallback: importing works anywhere, but every compute call needs the CUDA library and a device.
"""
__version__ = "0.1.0"
