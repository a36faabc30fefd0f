import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(autouse=True)
def _seed():
    np.random.seed(0)


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


def geodesic(Ra, Rb):
    """Rotation angle of Ra^T Rb; atan2 form stays accurate near 0 (arccos alone bottoms out at 1.5e-8)."""
    E = np.asarray(Ra, dtype=np.float64).T @ np.asarray(Rb, dtype=np.float64)
    c = 0.5 * (np.trace(E) - 1.0)
    s = 0.5 * np.linalg.norm([E[2, 1] - E[1, 2], E[0, 2] - E[2, 0], E[1, 0] - E[0, 1]])
    return float(np.arctan2(s, c))


def check_compact(g, key, arr, exact=False, tol=None):
    """
    Compare a full per-point array with a compact golden entry written by tests/golden/make_golden.py for the
    full-size cases: `key_rows` = every full_stride-th row, `key_sha256` = digest of the whole array (bit-exact
    arrays), `key_sum` / `key_abssum` = column sums (floating arrays: relative tolerance `tol`).
    """
    import hashlib
    a = np.ascontiguousarray(arr)
    stride = int(g["full_stride"])
    rows = g[key + "_rows"]
    if exact:
        assert a.dtype == rows.dtype, (key, a.dtype, rows.dtype)
        assert np.array_equal(a[::stride], rows), key
        digest = np.frombuffer(hashlib.sha256(a.tobytes()).digest(), np.uint8)
        assert np.array_equal(digest, g[key + "_sha256"]), key + ": digest of the whole array differs"
        return
    scale = float(np.max(np.abs(rows))) + 1e-300
    assert float(np.max(np.abs(a[::stride].astype(np.float64) - rows))) <= tol * scale, key
    s_abs = g[key + "_abssum"]
    assert np.all(np.abs(a.astype(np.float64).sum(axis=0) - g[key + "_sum"]) <= tol * s_abs + 1e-300), key + " (column sums)"
    assert np.all(np.abs(np.abs(a.astype(np.float64)).sum(axis=0) - s_abs) <= tol * s_abs + 1e-300), key + " (abs sums)"


class Gold:
    """A golden file whose large arrays may be stored in compact form (see check_compact): eq() / close() pick the form."""

    def __init__(self, name):
        self.g = golden(name)
        self.files = set(self.g.files)

    def __getitem__(self, k):
        return self.g[k]

    def has_full(self, k):
        return k in self.files

    def eq(self, key, arr):
        arr = np.asarray(arr)
        if key in self.files:
            assert np.array_equal(arr, self.g[key]), key
        else:
            rows = self.g[key + "_rows"]
            check_compact(self.g, key, arr.astype(rows.dtype) if arr.dtype != rows.dtype and arr.dtype.kind == rows.dtype.kind else arr,
                          exact=True)

    def close(self, key, arr, tol):
        if key in self.files:
            assert rel_err(arr, self.g[key]) < tol, key
        else:
            check_compact(self.g, key, np.asarray(arr), tol=tol)
