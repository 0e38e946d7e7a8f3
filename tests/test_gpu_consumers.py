"""Acceptance test with the reference's own consumers: test/codec_bench.c (with its
native-API front end) and the program of docs/EXAMPLE_README.md, compiled UNCHANGED from /root/reference and linked
against the GPU drop-in library (oracle/Makefile target `consumers`; the binaries travel under oracle/_ref).  They
must run their own verification (-t) green on a 64 MiB file."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "oracle", "_ref", "codec_bench_gpu")
EXAMPLE = os.path.join(ROOT, "oracle", "_ref", "example_gpu")


def test_consumer_binaries_link_against_the_gpu_library():
    """Built here (where /root/reference exists): the reference's bench resolves every aocl_llc_* / native symbol it
    uses from the GPU library and starts (its -h path needs no device)."""
    if not os.path.exists(BENCH):
        if not os.path.isdir("/root/reference"):
            pytest.skip("no prebuilt consumer binaries and no reference tree")
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "consumers"])
    out = subprocess.run([BENCH, "-h"], capture_output=True, text=True)
    assert "B200 LZ4/Snappy RAP path" in out.stdout and "aocl_compression_bench <options> input" in out.stdout
    assert os.path.exists(EXAMPLE)


@pytest.fixture(scope="module")
def input_file(tmp_path_factory):
    from llc_b200 import gen
    p = tmp_path_factory.mktemp("consumers") / "mixed64.bin"
    data = np.concatenate([gen.mixed_entropy(32 << 20), gen.text_like(32 << 20, seed=81)])
    data.tofile(p)
    return str(p)


@pytest.mark.gpu
@pytest.mark.parametrize("args", [["-elz4", "-t"], ["-esnappy", "-t"], ["-n", "-elz4", "-t"], ["-n", "-esnappy", "-t"],
                                  ["-elz4", "-p", "-i3"], ["-esnappy", "-p", "-i3"]])
def test_reference_codec_bench_runs_on_the_gpu_library(input_file, args):
    if not os.path.exists(BENCH):
        pytest.skip("oracle/_ref/codec_bench_gpu not built")
    r = subprocess.run([BENCH] + args + [input_file], capture_output=True, text=True, timeout=600)
    text = r.stdout + r.stderr
    print(text[-1500:])
    assert r.returncode == 0, text[-2000:]
    if "-t" in args:
        assert "verification: passed" in text and "failed" not in text, text[-2000:]
    else:
        assert "Compression" in text and "Decompression" in text, text[-2000:]


@pytest.mark.gpu
def test_reference_docs_example_runs_on_the_gpu_library(input_file):
    if not os.path.exists(EXAMPLE):
        pytest.skip("oracle/_ref/example_gpu not built")
    r = subprocess.run([EXAMPLE, input_file], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Compression: done" in r.stdout and "Decompression: done" in r.stdout, r.stdout + r.stderr
