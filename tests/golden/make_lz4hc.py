"""Generates the LZ4HC fixtures: three 64 KiB inputs and their LZ4_compress_HC (level 9) streams from the
UNMODIFIED reference compiled under oracle/_ref (oracle/Makefile builds algos/lz4/lz4hc.c too).  Run here, where
/root/reference exists; the GPU box only reads the committed files."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "aocl-compression_b200", "python"))
import oracle_lib as ol  # noqa: E402
from llc_b200 import gen  # noqa: E402

ref = ol.ref_lib()
assert ref is not None, "oracle/_ref could not be built"
inputs = {"text": gen.text_like(65536, seed=101), "mixed": gen.mixed_entropy(65536), "pages": gen.pages(1).reshape(-1)[:65536]}
for name, data in inputs.items():
    d = ref.new_desc(ol.LZ4HC)
    d.level = 9
    r, stream = ref.compress(data, ol.LZ4HC, desc=d)
    assert r > 0
    r2, back = ref.decompress(stream, ol.LZ4HC, len(data))
    assert r2 == len(data) and back == data.tobytes()
    open(os.path.join(HERE, f"lz4hc_{name}.src"), "wb").write(data.tobytes())
    open(os.path.join(HERE, f"lz4hc_{name}.lz4hc"), "wb").write(stream)
    print(name, len(data), "->", len(stream))
