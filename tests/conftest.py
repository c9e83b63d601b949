import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))


@pytest.fixture(scope="session")
def golden_seqs(golden):
    text, off = golden["text"].tobytes(), golden["text_off"]
    return [text[off[i]:off[i + 1]] for i in range(len(off) - 1)]


@pytest.fixture(scope="session")
def built_lib():
    """Make sure the in-tree CUDA library and the C oracle exist (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    from meshclust2_b200 import capi
    return capi


@pytest.fixture(scope="session")
def ctx(built_lib):
    c = built_lib.Context(0)
    yield c
    c.close()


def weights_path(name):
    return os.path.join(ROOT, "tests", "golden", name + ".txt")


def weights_text(name):
    return open(weights_path(name)).read()
