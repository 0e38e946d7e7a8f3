import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aocl-compression_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The compiled unmodified reference (oracle/_ref); None when it cannot be built/found."""
    import oracle_lib
    return oracle_lib.ref_lib()


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library through the same generic aocl_llc_* binding used for the reference."""
    import llc_b200
    import oracle_lib
    if not os.path.exists(llc_b200.LIB_PATH):
        llc_b200.build()
    return oracle_lib.LlcLib(llc_b200.LIB_PATH)


@pytest.fixture(scope="session")
def corpus():
    """Small deterministic inputs shared by the parity tests."""
    import numpy as np
    from llc_b200 import gen
    rng = np.random.default_rng(7)
    mixed = gen.mixed_entropy(4 << 20)
    return {
        "mixed": mixed,
        "text": gen.text_like(3 << 20, seed=11),
        "log": gen.log_like(3 << 20, seed=12),
        "random": rng.integers(0, 256, size=1 << 20, dtype=np.uint8),
        "zeros": np.zeros(1 << 20, dtype=np.uint8),
        "period7": np.resize(np.frombuffer(b"abcdefg", dtype=np.uint8), 1 << 20).copy(),
        "pages": gen.pages(6).reshape(-1),
    }
