"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/b200c.h declares,
and the product path fails loudly (no CPU fallback) when there is no CUDA device. No compute calls here."""
import ctypes as C, os, re, pytest
from cassandra_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def _declared_symbols():
    h = open(os.path.join(ROOT, "include", "b200c.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return set(re.findall(r"\b(b200c_[a-z0-9_]+)\s*\(", h))

def test_library_exports_every_declared_symbol():
    L = native.lib()
    declared = _declared_symbols()
    assert declared, "no prototypes parsed"
    for s in declared:
        assert hasattr(L, s), "missing export: " + s
    assert declared == set(native.SYMBOLS), (declared ^ set(native.SYMBOLS))
    assert L.b200c_abi_version() == native.ABI_VERSION == 2

def test_struct_sizes_match_header():
    # compile a tiny C program against the header and compare sizeof with the ctypes mirror
    import subprocess, tempfile
    src = r'''
#include <stdio.h>
#include "b200c.h"
int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(b200c_corruption), sizeof(b200c_column), sizeof(b200c_encoding_stats),
 sizeof(b200c_input), sizeof(b200c_manifest), sizeof(b200c_output), sizeof(b200c_result), sizeof(b200c_progress), sizeof(b200c_sstable_stats));return 0;}'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        got = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    want = [C.sizeof(x) for x in (native.Corruption, native.Column, native.EncodingStats, native.Input, native.Manifest,
                                  native.Output, native.Result, native.Progress, native.SSTableStats)]
    assert got == want

def test_no_cpu_fallback_without_device():
    L = native.lib()
    if L.b200c_device_count() > 0:
        pytest.skip("a CUDA device is present")
    assert not L.b200c_create(0, 0)
    with pytest.raises(native.B200CError):
        native.Context(0)

def test_static_helpers():
    L = native.lib()
    assert L.b200c_chunk_count(0, 16384) == 0 and L.b200c_chunk_count(16385, 16384) == 2
    assert L.b200c_initial_compressed_buffer_length(native.COMP_LZ4, 16384) == 4 + 16384 + 16384 // 255 + 16
    assert L.b200c_initial_compressed_buffer_length(native.COMP_SNAPPY, 16384) == 32 + 16384 + 16384 // 6
