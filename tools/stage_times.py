"""Per-stage device times of the bin path (mass / accumulate / finalize) for a batch shape: python tools/stage_times.py S P [prec]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from gc_slam_b200 import _lib as L, operators as ops, synth  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1
P = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
prec = {"f64": L.PREC_F64, "mixed": L.PREC_MIXED, "tc": L.PREC_TC}[sys.argv[3] if len(sys.argv) > 3 else "tc"]
bins = synth.fibonacci_atlas(48)
plan = ops.BinPathPlan(S, P, P, n_hyp=1, n_bins=48, tau=0.1, origin=synth.lidar_origin_base(), precision=prec,
                       want_evidence=True, materialize_deskewed=True)
plan.set_bins(bins, 0.1); plan.set_map(synth.random_map_bin_stats(48, 7, bins))
sc = [synth.vlp16_scan(P, 1000 + k, t0=synth.EPOCH_T0) for k in range(min(S, 4))]
sel = [sc[k % len(sc)] for k in range(S)]
arr = [np.stack([s[i] for s in sel]) for i in range(5)]
t0 = np.full(S, synth.EPOCH_T0)
plan.upload(*arr, t0, t0 + 0.1, np.stack([synth.scan_twist(5 + k) for k in range(S)]), synth.hypothesis_poses(S, 3), non_blocking=False)
dev = plan.io.dev
mass = torch.zeros((S, 4), dtype=torch.float64, device=dev)
raw = torch.zeros((plan.U, plan.raw_len), dtype=torch.float64, device=dev)
mx = torch.zeros((plan.U, 2), dtype=torch.float64, device=dev)
stages = [("mass", lambda: plan.run_mass(mass)), ("accumulate", lambda: plan.run_accumulate(mass, raw, mx)),
          ("finalize", lambda: plan.run_finalize(mass, raw, mx)), ("fused run()", plan.run)]
for name, fn in stages:
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    print(f"S={S} P={P} {name:12s} p50 {np.median(ts):8.1f} us   min {min(ts):8.1f} us")
