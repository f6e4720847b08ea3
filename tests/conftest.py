"""Shared fixtures.  `-m "not gpu"`: oracle vs golden vectors, host logic, C-ABI symbol export (CPU box).
`-m gpu`: parity of the CUDA path against the oracle / golden vectors through the C ABI (B200)."""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (ctypes over libsspyr.so).  Import fails loudly if the library is not built."""
    return entry.load_package()


@pytest.fixture(scope="session")
def O():
    """The oracle (test infrastructure): C restatement + the compiled reference header when present."""
    o = entry.load_oracle()
    if not os.path.exists(o.port_path()):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True)
    o.load_port()
    return o


@pytest.fixture(scope="session")
def synth(pkg):
    return pkg.synth


@pytest.fixture(scope="session")
def golden_small():
    return np.load(os.path.join(GOLDEN, "header_small.npz"))


@pytest.fixture(scope="session")
def golden_hashes():
    with open(os.path.join(GOLDEN, "header_hashes.json")) as f:
        return json.load(f)


def split_flat(flat: np.ndarray, n: int, S: int) -> list[np.ndarray]:
    """Flat golden array of a square side-n all-octave pyramid -> list over octaves of [S+3, n_o, n_o]."""
    out, off, lo = [], 0, n
    while lo:
        cnt = (S + 3) * lo * lo
        out.append(flat[off:off + cnt].reshape(S + 3, lo, lo))
        off += cnt
        lo //= 2
    assert off == flat.size
    return out


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


SMALL_CASES = [(1, 3), (2, 3), (3, 3), (5, 2), (8, 3), (16, 2), (16, 3), (37, 3), (64, 0), (67, 3), (100, 5)]
KINDS = ("ones", "pattern", "noise")
