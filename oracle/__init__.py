"""
oracle/ -- CPU (NumPy float64) restatement of GC-SLAM's per-scan LiDAR evidence path.

TEST INFRASTRUCTURE.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
package.  The product (``gc_slam_b200``) never does: it fails loudly when its
CUDA library is missing.

Parity status
-------------
* The reference (whabacivch/GC-SLAM) ships **no golden vectors** for this path
   and needs JAX, which is absent here, so it cannot run
  unmodified in this container.
* Pin used instead: ``oracle/jax_shim`` substitutes only the array runtime
  (XLA -> NumPy) so that the reference's *own operator source files* execute
  here; ``tests/golden/make_golden.py`` ran them on seeded inputs and committed
  the outputs under ``tests/golden/*.npz``.  ``tests/test_oracle_vs_golden.py``
  checks every oracle function against those vectors, and
  ``tests/test_oracle_properties.py`` re-states the reference's own property
  tests (kappa batch==scalar, SO(3)/SE(3) round trips, PSD eigen >= eps,
  softmax rows sum to 1, mass preservation).
* What stays unpinned: XLA's own float64 kernels for exp/log/sin/cos/eigh/svd
  (third-party ``jax[cuda13]==0.9.0``, not vendored) -- differences are at the
  1e-15 level -- and ``PoseCovInflationPushforward``, whose source was deleted
  from the reference (prose only).

Every function cites the reference file:line it follows (paths relative to
``/root/reference``; ``fl/`` = ``fl_ws/src/fl_slam_poc/fl_slam_poc/``).
"""
