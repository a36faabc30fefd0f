"""Developer probe: config 3 timings inside the full default bench process, then the pipelined H=64 loop again."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import bench
orig = bench.config3_block
def pipelined(reps=60, timing=False, H=64):
    from gc_slam_b200 import hypothesis_batch as HB, primitives as PR, synth, _lib as L
    atlas_np = synth.synthetic_atlas(1_000_000, 50000, 7, scan_seq=20)
    amap = PR.AtlasMap.from_numpy(atlas_np, n_tiles_cap=len(atlas_np["tiles"]) + 16)
    pts, t, w, _, _ = synth.vlp16_scan(65536, 4242, t0=synth.EPOCH_T0)
    cam = synth.camera_splats(512, 99)
    base = PR.measurement_batch_from_camera_splats(cam["positions"], cam["covariances"], cam["directions"], cam["kappas"], cam["weights"], cam["timestamps"], cam["colors"])
    pts_d, t_d, w_d = [torch.from_numpy(a).cuda() for a in (pts, t, w)]
    t0, t1 = synth.EPOCH_T0, synth.EPOCH_T0 + 0.1
    xis = torch.from_numpy(np.stack([synth.scan_twist(4242 + h) for h in range(H)])).cuda()
    poses = synth.hypothesis_poses(H, 42) * 0.2; poses[:, :3] += np.array([0.1, -0.2, 0.5])
    ctx = L.context()
    out = []
    for rnd in range(3):
        for k in range(6): HB.lidar_evidence_primitives_batched(pts_d, t_d, w_d, t0, t1, xis, amap, poses, 31 + k, base_batch=base, update_map=False)
        torch.cuda.synchronize()
        if timing: ctx.timing_enable(True, only="topk")
        a0 = time.perf_counter(); prev = None; enq = 0.0; wt = 0.0
        for k in range(reps):
            b = time.perf_counter()
            cur = HB.lidar_evidence_primitives_batched(pts_d, t_d, w_d, t0, t1, xis, amap, poses, 41 + k, base_batch=base, update_map=False, defer=True)
            enq += time.perf_counter() - b
            b = time.perf_counter()
            if prev is not None: prev.wait()
            wt += time.perf_counter() - b
            prev = cur
        prev.wait(); torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - a0) / reps
        if timing:
            ctx.timing_collect(); ctx.timing_enable(False)
        out.append((round(ms, 3), round(1e3 * enq / reps, 3), round(1e3 * wt / reps, 3)))
    return out
def wrapped(*a, **k):
    sys.stderr.write(f"before config3 (in-bench), pipelined H=64 (ms, enqueue, wait): {pipelined()}\n")
    r = orig(*a, **k)
    for H in ("1", "4", "64"):
        c = r["per_hypotheses"][H]
        sys.stderr.write(f"in-bench {H} {c['ms_per_scan_with_map_update']:.3f} {c['ms_per_scan_evidence_only']:.3f}\n")
    sys.stderr.write(f"after config3, pipelined H=64: {pipelined()}  with topk timing: {pipelined(timing=True)}\n")
    return r
bench.config3_block = wrapped
sys.argv = ["bench.py"]
bench.main()
