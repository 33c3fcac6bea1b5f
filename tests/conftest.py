"""pytest configuration: `gpu` marker, shared fixtures.

`-m "not gpu"`: oracle vs the reference-generated golden vectors, host logic, ABI symbol checks (no compute calls).
`-m gpu`:       parity tests proper — the CUDA path, called through the C ABI, against the oracle.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_sessionstart(session):
    """A fresh checkout has no built library (the .so files are git-ignored): build them once, like
    __graft_entry__.build() does. An existing library is used as it is."""
    from hpfw_b200 import build as _build
    if not os.path.exists(_build.LIB):
        _build.build()


def _load_cases(path):
    z = np.load(path)
    cases = {}
    for key in z.files:
        name, field = key.split("__")
        cases.setdefault(name, {})[field] = z[key]
    return cases


@pytest.fixture(scope="session")
def matcher_golden():
    return _load_cases(os.path.join(GOLDEN, "matcher.npz"))


@pytest.fixture(scope="session")
def hashprint_golden():
    return dict(np.load(os.path.join(GOLDEN, "hashprint.npz")))


@pytest.fixture(scope="session")
def kat_golden():
    return dict(np.load(os.path.join(GOLDEN, "kat.npz")))


@pytest.fixture(scope="session")
def collector_golden():
    return dict(np.load(os.path.join(GOLDEN, "collector.npz")))


@pytest.fixture(scope="session")
def ctx():
    """One hpfw_ctx on cuda:0 for the gpu tests. Fails (does not skip) if the CUDA library is missing."""
    import hpfw_b200
    c = hpfw_b200.Context(0)
    yield c
    c.close()
