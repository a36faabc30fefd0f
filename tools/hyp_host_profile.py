"""Developer tool: cProfile of the host side of lidar_evidence_primitives_batched (H from argv, with map update)."""
import cProfile, pstats, sys, io as _io
import numpy as np, torch
sys.path.insert(0, ".")
from gc_slam_b200 import hypothesis_batch as HB, primitives as PR, synth
H = int(sys.argv[1]) if len(sys.argv) > 1 else 64
upd = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
atlas_np = synth.synthetic_atlas(1_000_000, 50000, 7, scan_seq=20)
amap = PR.AtlasMap.from_numpy(atlas_np, n_tiles_cap=len(atlas_np["tiles"]) + 16)
pts, t, w, _, _ = synth.vlp16_scan(65536, 4242, t0=synth.EPOCH_T0)
cam = synth.camera_splats(512, 99)
base = PR.measurement_batch_from_camera_splats(cam["positions"], cam["covariances"], cam["directions"], cam["kappas"], cam["weights"], cam["timestamps"], cam["colors"])
pts_d, t_d, w_d = [torch.from_numpy(a).cuda() for a in (pts, t, w)]
t0, t1 = synth.EPOCH_T0, synth.EPOCH_T0 + 0.1
xis = torch.from_numpy(np.stack([synth.scan_twist(4242 + h) for h in range(H)])).cuda()
poses = synth.hypothesis_poses(H, 42) * 0.2; poses[:, :3] += np.array([0.1, -0.2, 0.5])
run = lambda k: HB.lidar_evidence_primitives_batched(pts_d, t_d, w_d, t0, t1, xis, amap, poses, 30 + k, base_batch=base, update_map=upd)
for k in range(8): run(k)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for k in range(40): run(10 + k)
torch.cuda.synchronize()
pr.disable()
s = _io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue())
