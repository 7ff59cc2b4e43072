"""The per-thread LZ4 decoder that K1's thread-per-chunk kernels run (cassandra_b200/csrc/lz4_thread.cuh) is plain C++ compiled for
host and device. Here the same source is built with g++ -fsanitize=address,undefined and fuzzed against the oracle on the CPU: valid
blocks at every source alignment must decode exactly; flipped and truncated blocks must fail or stay inside the buffers (+16 bytes of
slack, as every engine buffer has). The walk of the two-pass decoder (lz4_batch.cuh) rides along: it must accept exactly the blocks the
decoder decodes to the expected size, with a record buffer of exactly (n - 1) / 3 + 1 entries."""
import os, shutil, subprocess, pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_thread_decoder_fuzz_under_asan(tmp_path):
    exe = str(tmp_path / "lz4_thread_fuzz")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=address", "-fno-omit-frame-pointer", "-fno-strict-aliasing",
           "-o", exe, os.path.join(ROOT, "tests", "native", "lz4_thread_fuzz.cc"), os.path.join(ROOT, "oracle", "codec.cc")]
    b = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert b.returncode == 0, b.stderr[-2000:]
    r = subprocess.run([exe, "250"], capture_output=True, text=True, timeout=600, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1"))
    assert r.returncode == 0 and "lz4_thread_fuzz ok" in r.stdout, (r.stdout + r.stderr)[-3000:]
    assert "runtime error" not in r.stderr, r.stderr[-3000:]
