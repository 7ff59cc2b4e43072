#!/usr/bin/env python
"""Generates tests/golden/snappy/vectors.json: raw-snappy compressions produced by GOOGLE'S SNAPPY LIBRARY (the C++ library snappy-java 1.1.10.4
wraps, .build/parent-pom-template.xml:292-296; here the copy bundled in pyarrow, pyarrow.Codec('snappy')) of deterministic inputs that
tests/snappy_vectors.py regenerates from seeds. The reference holds no Snappy-compressed fixture, so these vectors are what pins
the oracle's and the GPU's Snappy bytes to the real library. Run in the build container; the JSON is committed."""
import base64, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import pyarrow as pa
from snappy_vectors import inputs

codec = pa.Codec("snappy")
out = {"library": "Google snappy as bundled in pyarrow %s (pyarrow.Codec('snappy'))" % pa.__version__, "vectors": []}
for name, data in inputs():
    c = codec.compress(data, asbytes=True) if len(data) else b"\x00"
    out["vectors"].append({"name": name, "n": len(data), "compressed_b64": base64.b64encode(c).decode()})
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "snappy", "vectors.json")
json.dump(out, open(dst, "w"), indent=0)
print("wrote %d vectors, %d bytes" % (len(out["vectors"]), os.path.getsize(dst)))
