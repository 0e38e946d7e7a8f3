"""GPU parity: the CUDA path, called through the reference's own C ABI (aocl_llc_*), against
the CPU oracle on the same seeded inputs.  Bit-exact for every byte (integer/byte domain)."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

SIZES = [0, 1, 12, 13, 100, 4096, 30000, 65535, 65536, 65546, 65547, 131072, 262143, 262144, 262267, 262268,
         393215, 393216, 393401, 393402, 1 << 20, 1500001, 2097229]
CODECS = [ol.LZ4, ol.SNAPPY]


def first_diff(a: bytes, b: bytes) -> str:
    x, y = np.frombuffer(a, dtype=np.uint8), np.frombuffer(b, dtype=np.uint8)
    m = min(len(x), len(y))
    d = np.nonzero(x[:m] != y[:m])[0]
    return f"len {len(a)} vs {len(b)}, first diff at {int(d[0]) if len(d) else m}"


@pytest.mark.parametrize("codec", CODECS)
@pytest.mark.parametrize("name", ["mixed", "text", "random", "zeros", "period7"])
def test_compress_matches_oracle(gpu_lib, oracle, corpus, codec, name):
    data = corpus[name]
    for n in SIZES:
        if n > len(data):
            continue
        d = data[:n]
        want = oracle.compress(d, codec)
        r, got = gpu_lib.compress(d, codec)
        assert r == len(want), (name, n, r, len(want))
        assert got == want, (name, n, first_diff(got, want))


@pytest.mark.parametrize("codec", CODECS)
@pytest.mark.parametrize("name", ["mixed", "text", "log", "random", "zeros", "period7", "pages"])
def test_decompress_oracle_streams(gpu_lib, oracle, corpus, codec, name):
    data = corpus[name]
    for n in SIZES:
        if n > len(data):
            continue
        d = data[:n]
        stream = oracle.compress(d, codec)
        r, got = gpu_lib.decompress(stream, codec, max(n, 1))
        assert r == n, (name, n, r)
        assert got == d.tobytes(), (name, n, first_diff(got, d.tobytes()))


@pytest.mark.parametrize("codec", CODECS)
def test_round_trip_large(gpu_lib, oracle, corpus, codec):
    data = np.concatenate([corpus["text"], corpus["mixed"], corpus["log"], corpus["random"], corpus["zeros"]])
    r, stream = gpu_lib.compress(data, codec)
    assert r > 0
    assert stream == oracle.compress(data, codec)
    r2, back = gpu_lib.decompress(stream, codec, len(data))
    assert r2 == len(data) and back == data.tobytes()
    # the oracle decoder accepts the GPU stream as well
    assert oracle.decompress(stream, codec, len(data)) == data.tobytes()
