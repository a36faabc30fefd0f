"""oracle/prim_path.py's sparse_cost and sinkhorn_unbalanced against the vectors produced by the reference's own
_compute_sparse_cost_matrix_jax / _sinkhorn_unbalanced_fixed_k_jax (tests/golden/make_golden_assoc_inner.py).  CPU only."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden

COST_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "associnner_cost_*.npz")))
SINKHORN_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "associnner_sinkhorn_*.npz")))


def close(x, ref, rtol):
    x, ref = np.asarray(x), np.asarray(ref)
    return np.all(np.abs(x - ref) <= rtol * np.maximum(np.abs(ref), 1e-300) + 1e-300)


def test_cases_present():
    assert len(COST_CASES) >= 2 and len(SINKHORN_CASES) >= 3


@pytest.mark.parametrize("case", COST_CASES)
def test_oracle_cost_matches_reference(case):
    from oracle import prim_path as op
    g = golden(case)
    C = op.sparse_cost(g["mp"], g["md"], g["mk"], g["vp"], g["vd"], g["vk"], g["cand"], beta=float(g["beta"]))
    assert C.shape == g["out_cost"].shape and close(C, g["out_cost"], 1e-13)


@pytest.mark.parametrize("case", SINKHORN_CASES)
def test_oracle_sinkhorn_matches_reference(case):
    from oracle import prim_path as op
    g = golden(case)
    pi = op.sinkhorn_unbalanced(g["C"], g["a"], g["b"], float(g["epsilon"]), float(g["tau_a"]), float(g["tau_b"]), int(g["iters"]))
    assert pi.shape == g["out_pi"].shape and close(pi, g["out_pi"], 1e-12)
