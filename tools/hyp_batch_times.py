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
    poses[:, :3] += np.array([0.1, -0.2, 0.0])
    for upd in (False, True):
        ts = []
        for r in range(reps + 2):
            torch.cuda.synchronize()
            a = time.perf_counter()
            out = HB.lidar_evidence_primitives_batched(pts_d, t_d, w_d, t0, t1, xis, amap, poses, 30 + r, base_batch=base, update_map=upd)
            torch.cuda.synchronize()
            if r >= 2:
                ts.append(time.perf_counter() - a)
        ms = 1e3 * float(np.median(ts))
        print(f"H={H:3d} update_map={upd!s:5}  groups={len(out.groups)}  {ms:8.3f} ms/scan  {H / ms * 1e3:10.1f} hypothesis-scans/s", flush=True)
