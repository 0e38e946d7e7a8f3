#!/usr/bin/env python
"""Regenerates tests/golden/golden.json from the UNMODIFIED reference compiled by oracle/Makefile
(oracle/_ref/libaocl_ref.so; needs /root/reference, so this only runs in the build container).

For every (codec, generator, size) of tests/kat.py the reference compresses the input through
aocl_llc_compress with OMP threads = saturated partition count; we store the SHA-256 and length of
its output (and the output itself, hex, when it is at most 512 bytes) plus the SHA-256 of the input.
Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "aocl-compression_b200", "python"))

import kat  # noqa: E402
import oracle_lib as ol  # noqa: E402


def main():
    ref = ol.ref_lib()
    assert ref is not None, "reference library not available"
    orc = ol.Oracle()
    out = {"reference": ref.L.aocl_llc_version().decode(), "cases": []}
    for codec in (kat.LZ4, kat.SNAPPY):
        for name in kat.GOLDEN_GENS:
            for size in kat.GOLDEN_SIZES:
                data = kat.make_input(name, size)
                T = orc.partition_count(size, codec)
                ref.set_threads(max(T, 1))
                r, comp = ref.compress(data, codec)
                assert r > 0 and len(comp) == r
                r2, back = ref.decompress(comp, codec, max(size, 1))
                assert r2 == size and back == data.tobytes()
                case = {"codec": codec, "gen": name, "size": size, "partitions": T, "in_sha256": kat.sha(data.tobytes()),
                        "out_len": r, "out_sha256": kat.sha(comp)}
                if r <= 512:
                    case["out_hex"] = comp.hex()
                out["cases"].append(case)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
