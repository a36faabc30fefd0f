"""How much does the bin path running on another stream slow a pinned host -> device copy down?  (explains why the
e2e step is longer than the bare copy: tools/e2e_probe.py)"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from gc_slam_b200 import _lib as L, operators as ops, synth  # noqa: E402

S, P = 128, 65536
bins = synth.fibonacci_atlas(48)
plan = ops.BinPathPlan(S, P, P, n_hyp=1, n_bins=48, tau=0.1, origin=synth.lidar_origin_base(), precision=L.PREC_TC,
                       want_evidence=True, materialize_deskewed=True, own_context=True)
plan.set_bins(bins, 0.1)
plan.set_map(synth.random_map_bin_stats(48, 7, bins))
sc = [synth.vlp16_scan(P, 1000 + k, t0=synth.EPOCH_T0) for k in range(4)]
arr = [np.stack([sc[k % 4][i] for k in range(S)]) for i in range(5)]
t0 = np.full(S, synth.EPOCH_T0)
plan.upload(*arr, t0, t0 + 0.1, np.stack([synth.scan_twist(5 + k) for k in range(S)]), synth.hypothesis_poses(S, 3), non_blocking=False)
n = 184563712
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
sk, sc_ = torch.cuda.Stream(), torch.cuda.Stream()
for mode in ("copy alone", "copy + bin path on another stream", "copy + float64 bin path on another stream"):
    if mode.startswith("copy + float64"):
        plan.args.precision = L.PREC_F64
    torch.cuda.synchronize()
    a = time.perf_counter()
    if mode != "copy alone":
        with torch.cuda.stream(sk):
            for _ in range(60 if plan.args.precision == L.PREC_TC else 18):
                plan.run()
    with torch.cuda.stream(sc_):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            d.copy_(h, non_blocking=True)
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{mode:45s}: 10 x 184.6 MB in {ms:6.2f} ms = {10 * n / ms / 1e6:5.1f} GB/s (kernels busy for {(time.perf_counter() - a) * 1e3:.1f} ms)")
