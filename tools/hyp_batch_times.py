"""Wall / device time of lidar_evidence_primitives_batched at several hypothesis counts (config 3 shapes)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from gc_slam_b200 import hypothesis_batch as HB, primitives as PR, synth  # noqa: E402

n = 65536
Hs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 4, 16, 64]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n_map = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
atlas_np = synth.synthetic_atlas(n_map, 50000, 7, scan_seq=20)
amap = PR.AtlasMap.from_numpy(atlas_np, n_tiles_cap=len(atlas_np["tiles"]) + 16)
pts, t, w, _, _ = synth.vlp16_scan(n, 4242, t0=synth.EPOCH_T0)
cam = synth.camera_splats(512, 99)
base = PR.measurement_batch_from_camera_splats(cam["positions"], cam["covariances"], cam["directions"], cam["kappas"],
                                               cam["weights"], cam["timestamps"], cam["colors"])
pts_d, t_d, w_d = [torch.from_numpy(a).cuda() for a in (pts, t, w)]
t0, t1 = synth.EPOCH_T0, synth.EPOCH_T0 + 0.1
for H in Hs:
    xis = torch.from_numpy(np.stack([synth.scan_twist(4242 + h) for h in range(H)])).cuda()
    poses = synth.hypothesis_poses(H, 42) * 0.2
    poses[:, :3] += np.array([0.1, -0.2, 0.5])   # away from a tile boundary: one stencil for all hypotheses
    for upd in (False, True):
        ts, ds = [], []
        for r in range(reps + 2):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a = time.perf_counter()
            e0.record()
            out = HB.lidar_evidence_primitives_batched(pts_d, t_d, w_d, t0, t1, xis, amap, poses, 30 + r, base_batch=base, update_map=upd)
            e1.record()
            torch.cuda.synchronize()
            if r >= 2:
                ts.append(time.perf_counter() - a)
                ds.append(e0.elapsed_time(e1))
        ms = 1e3 * float(np.median(ts))
        print(f"H={H:3d} update_map={upd!s:5}  groups={len(out.groups)}  {ms:8.3f} ms/scan wall  {float(np.median(ds)):8.3f} ms device span  "
              f"{H / ms * 1e3:10.1f} hypothesis-scans/s", flush=True)

# steady state: scans enqueued back to back (defer), the wait for scan k-1 issued after scan k has been enqueued
import os
for H in Hs:
    xis = torch.from_numpy(np.stack([synth.scan_twist(4242 + h) for h in range(H)])).cuda()
    poses = synth.hypothesis_poses(H, 42) * 0.2
    poses[:, :3] += np.array([0.1, -0.2, 0.5])
    K = 20
    for rep in range(2):
        torch.cuda.synchronize()
        a = time.perf_counter()
        prev, t_enq = None, 0.0
        for k in range(K):
            b = time.perf_counter()
            cur = HB.lidar_evidence_primitives_batched(pts_d, t_d, w_d, t0, t1, xis, amap, poses, 60 + k, base_batch=base, update_map=False, defer=True)
            t_enq += time.perf_counter() - b
            if prev is not None:
                prev.wait()
            prev = cur
        prev.wait()
        torch.cuda.synchronize()
        dt = time.perf_counter() - a
    print(f"H={H:3d} pipelined x{K}: {1e3 * dt / K:8.3f} ms/scan ({1e3 * t_enq / K:6.3f} ms host enqueue)  {H * K / dt:10.1f} hypothesis-scans/s", flush=True)
if os.environ.get("PROFILE"):
    import cProfile, pstats
    H = Hs[0]
    pr = cProfile.Profile()
    pr.enable()
    for k in range(20):
        HB.lidar_evidence_primitives_batched(pts_d, t_d, w_d, t0, t1, xis[:H], amap, poses[:H], 90 + k, base_batch=base, update_map=False)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
