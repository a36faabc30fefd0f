"""
The association's inner functions on the device (the reference's warm-up calls them directly, backend_node.py:884-905)
against the reference's own outputs (tests/golden/associnner_*.npz, made by tests/golden/make_golden_assoc_inner.py).
Through the C-ABI entries gcs_sparse_cost_matrix / gcs_sinkhorn_unbalanced_fixed_k.  Element-wise relative 1e-10 (costs: the
Hellinger term cancels to ~1e-12 absolute; plan entries after 50 power iterations).
"""
import numpy as np
import pytest

from conftest import golden
from test_oracle_assoc_inner_vs_golden import COST_CASES, SINKHORN_CASES

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def P():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from gc_slam_b200 import primitives
    return primitives


@pytest.mark.parametrize("case", COST_CASES)
def test_cost_matrix_vs_reference_golden(P, case):
    g = golden(case)
    C = P.compute_sparse_cost_matrix(g["mp"], g["md"], g["mk"], g["vp"], g["vd"], g["vk"], g["cand"], beta=float(g["beta"])).cpu().numpy()
    ref = g["out_cost"]
    assert C.shape == ref.shape
    assert np.all(np.abs(C - ref) <= 1e-12 * np.abs(ref) + 1e-11)


@pytest.mark.parametrize("case", SINKHORN_CASES)
def test_sinkhorn_vs_reference_golden(P, case):
    g = golden(case)
    args = (g["C"], g["a"], g["b"], float(g["epsilon"]), float(g["tau_a"]), float(g["tau_b"]), int(g["iters"]))
    pi = P.sinkhorn_unbalanced_fixed_k(*args)
    ref = g["out_pi"]
    out = pi.cpu().numpy()
    assert out.shape == ref.shape
    assert np.all(np.abs(out - ref) <= 1e-10 * np.abs(ref) + 1e-300)
    assert np.array_equal(out == 0.0, ref == 0.0)            # rows with a = 0 carry no mass, exactly
    assert torch.equal(pi, P.sinkhorn_unbalanced_fixed_k(*args))   # fixed-order sums: bit-identical rerun


def test_warmup_block_runs_unchanged(P):
    """The shapes and constants of the reference's warm-up (backend_node.py:868-905) through the reference's private names."""
    n_total, k_assoc = 1536, 8
    C = P._compute_sparse_cost_matrix_jax(np.zeros((n_total, 3)), np.tile([1.0, 0.0, 0.0], (n_total, 1)), np.ones(n_total),
                                          np.zeros((1, 3)), np.array([[1.0, 0.0, 0.0]]), np.array([1.0]),
                                          np.zeros((n_total, k_assoc), dtype=np.int32))
    assert tuple(C.shape) == (n_total, k_assoc) and float(C.abs().max()) < 1e-12
    a = np.ones(n_total) / n_total
    pi = P._sinkhorn_unbalanced_fixed_k_jax(C, a, np.ones(k_assoc) / k_assoc, 0.1, 0.5, 0.5, 50)
    assert tuple(pi.shape) == (n_total, k_assoc) and bool(torch.isfinite(pi).all()) and float(pi.sum()) > 0.0
