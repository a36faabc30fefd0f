#!/usr/bin/env python
"""
Golden vectors for the hypothesis combine (HypothesisBarycenterProjection core), produced by executing the REFERENCE's
own `_hypothesis_barycenter_core` (fl/backend/operators/hypothesis.py:51-115) on top of oracle/jax_shim.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_hyp.py
Outputs tests/golden/hyp_*.npz (small, committed).  Nothing under /root/reference is written or copied.
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
    from fl_slam_poc.backend.operators.hypothesis import _hypothesis_barycenter_core
    from gc_slam_b200 import synth

    cases = {"hyp_k4": (4, 22, 11, False), "hyp_k64": (64, 22, 12, False), "hyp_k5_indefinite_floor": (5, 22, 13, True)}
    for name, (K, D, seed, nasty) in cases.items():
        Ls, hs, zs, w = synth.hypothesis_evidence_stack(K, D, seed, indefinite=nasty)
        out = _hypothesis_barycenter_core(Ls, hs, zs, w, 0.0025, 1e-12, 1e-9)
        keys = ("L", "h", "z_lin", "floor_adjustment", "weights_normalized", "psd_cert", "spread_proxy")
        d = {k: np.asarray(v, dtype=np.float64) for k, v in zip(keys, out)}
        np.savez_compressed(os.path.join(HERE, name + ".npz"), L_stack=Ls, h_stack=hs, z_lin_stack=zs, weights=w,
                            weight_floor=0.0025, eps_psd=1e-12, eps_lift=1e-9, **d)
        print(name, "spread", float(d["spread_proxy"]), "psd", d["psd_cert"])


if __name__ == "__main__":
    main()
