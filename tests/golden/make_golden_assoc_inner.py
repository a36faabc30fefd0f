#!/usr/bin/env python
"""
Golden vectors for the two inner functions of the OT association, produced by the REFERENCE's own
_compute_sparse_cost_matrix_jax and _sinkhorn_unbalanced_fixed_k_jax
(fl/backend/operators/primitive_association.py:105-197) on top of oracle/jax_shim -- the functions its start-up warm-up
calls directly (fl/backend/backend_node.py:884-905).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_assoc_inner.py
Outputs tests/golden/associnner_*.npz.  Nothing under /root/reference is written or copied.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "jax_shim"))
sys.path.insert(0, os.path.join(REF, "fl_ws", "src", "fl_slam_poc"))
sys.path.insert(0, ROOT)


def main():
    import jax.numpy as jnp
    from fl_slam_poc.backend.operators import primitive_association as pa

    def unit(v):
        return v / np.linalg.norm(v, axis=-1, keepdims=True)

    for name, N, M, K, seed, beta in (("associnner_cost_300x8", 300, 500, 8, 71, 0.5), ("associnner_cost_64x20", 64, 90, 20, 72, 2.0)):
        rng = np.random.default_rng(seed)
        mp, vp = rng.normal(0, 2.0, (N, 3)), rng.normal(0, 2.0, (M, 3))
        md, vd = unit(rng.normal(size=(N, 3))), unit(rng.normal(size=(M, 3)))
        # concentrations across all three branches of A_vmf (k < 1e-2, 1e-2 <= k <= 20, k > 20) and exact zeros
        mk = 10.0 ** rng.uniform(-4.0, 3.0, N) * (rng.random(N) < 0.9)
        vk = 10.0 ** rng.uniform(-4.0, 3.0, M) * (rng.random(M) < 0.9)
        cand = rng.integers(0, M, (N, K)).astype(np.int32)
        C = pa._compute_sparse_cost_matrix_jax(jnp.asarray(mp), jnp.asarray(md), jnp.asarray(mk), jnp.asarray(vp), jnp.asarray(vd),
                                               jnp.asarray(vk), jnp.asarray(cand), beta=beta)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), mp=mp, md=md, mk=mk, vp=vp, vd=vd, vk=vk, cand=cand, beta=beta,
                            out_cost=np.asarray(C))
        print(name, np.asarray(C).shape, float(np.asarray(C).mean()))

    for name, N, M, seed, eps, ta, tb, iters in (("associnner_sinkhorn_1536x8", 1536, 8, 81, 0.1, 0.5, 0.5, 50),
                                                 ("associnner_sinkhorn_200x5", 200, 5, 82, 0.05, 2.0, 0.3, 7),
                                                 ("associnner_sinkhorn_40x32_zero_iters", 40, 32, 83, 0.2, 1.0, 1.0, 0)):
        rng = np.random.default_rng(seed)
        C = rng.random((N, M)) * 0.6 + (rng.random((N, M)) < 0.1) * 1e12 * 0.0 + (rng.random((N, M)) < 0.05) * 50.0
        a = (rng.random(N) < 0.8).astype(np.float64)
        a = a / max(a.sum(), 1e-12)                       # the pipeline's uniform marginal over valid rows (zeros for invalid)
        b = np.ones(M) / M if M == 8 else rng.dirichlet(np.ones(M))
        pi = pa._sinkhorn_unbalanced_fixed_k_jax(jnp.asarray(C), jnp.asarray(a), jnp.asarray(b), eps, ta, tb, iters)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), C=C, a=a, b=b, epsilon=eps, tau_a=ta, tau_b=tb, iters=iters,
                            out_pi=np.asarray(pi))
        print(name, float(np.asarray(pi).sum()))


if __name__ == "__main__":
    main()
