"""Dev check of the tensor-core precision of the bin path against the float64 path on the same device buffers.
Usage (GPU box):  python tools/tc_check.py [n_scans] [points] [n_hyp]
Prints worst relative errors per output and the kernel times; exits non-zero when an error exceeds 1e-5."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from gc_slam_b200 import _lib as L, operators as ops, synth  # noqa: E402


def rel(a, b):
    a = a.double().cpu().numpy(); b = b.double().cpu().numpy()
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


def run(S, P, H, prec, cap=None):
    cap = cap or P
    bins = synth.fibonacci_atlas(48)
    plan = ops.BinPathPlan(S, P, cap, n_hyp=H, n_bins=48, tau=0.1, origin=synth.lidar_origin_base(), precision=prec,
                           want_evidence=True, materialize_deskewed=True)
    plan.set_bins(bins, 0.1)
    plan.set_map(synth.random_map_bin_stats(48, 7, bins))
    scans = [synth.vlp16_scan(P, 1000 + k, t0=synth.EPOCH_T0) for k in range(min(S, 4))]
    sel = [scans[k % len(scans)] for k in range(S)]
    pts = np.stack([s[0] for s in sel]); t = np.stack([s[1] for s in sel]); w = np.stack([s[2] for s in sel])
    ring = np.stack([s[3] for s in sel]); tag = np.stack([s[4] for s in sel])
    t0 = np.full(S, synth.EPOCH_T0); t1 = t0 + 0.1
    xi = np.stack([synth.scan_twist(5 + k) for k in range(S * H)])
    poses = synth.hypothesis_poses(S * H, 3)
    plan.upload(pts, t, w, ring, tag, t0, t1, xi, poses, non_blocking=False)
    ctx = plan.io.ctx
    plan.run(); torch.cuda.synchronize()
    ctx.timing_enable(True)
    for _ in range(3):
        plan.run()
    torch.cuda.synchronize()
    ms, n = ctx.timing_collect(); ctx.timing_enable(False)
    return plan, plan.outputs(), ms / max(n, 1)


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    cap = int(sys.argv[4]) if len(sys.argv) > 4 else None
    pa, a, ms_a = run(S, P, H, L.PREC_F64, cap)
    pm, m, ms_m = run(S, P, H, L.PREC_MIXED, cap)
    pb, b, ms_b = run(S, P, H, L.PREC_TC, cap)
    pb2, b2, _ = run(S, P, H, L.PREC_TC, cap)
    print(f"S={S} P={P} H={H} cap={cap}: scan kernel f64 {ms_a:.3f} ms  mixed {ms_m:.3f} ms  tc {ms_b:.3f} ms")
    worst = 0.0
    for k in a.stats:
        e = rel(b.stats[k], a.stats[k]); em = rel(m.stats[k], a.stats[k])
        worst = max(worst, e)
        print(f"  stats[{k:14s}] tc rel {e:.3e}   mixed rel {em:.3e}")
    for name, x, y, z in (("L22", b.L22, a.L22, m.L22), ("h22", b.h22, a.h22, m.h22), ("cert", b.cert, a.cert, m.cert)):
        e = rel(x, y)
        print(f"  {name:21s} tc rel {e:.3e}   mixed rel {rel(z, y):.3e}")
    if a.deskewed is not None:
        print("  deskewed points equal:", torch.equal(a.deskewed["points"], b.deskewed["points"]),
              " weights equal:", torch.equal(a.deskewed["weights"], b.deskewed["weights"]))
    same = all(torch.equal(b.stats[k], b2.stats[k]) for k in b.stats) and torch.equal(b.L22, b2.L22)
    print("  tc rerun bit-identical:", same)
    sys.exit(0 if worst < 1e-5 and same else 1)


if __name__ == "__main__":
    main()
