"""Developer tool: A/B timing of the primitive-family hypothesis batch (config 3 shapes) across library variants on ONE box.
Usage (GPU box):  python tools/prim_ab.py main variantB ...   Prints per variant: device span of an evidence-only batch at
H = 1 and 64, and the summed device time of the tagged kernels (topk, sinkhorn) of one batch."""
import os
import subprocess
import sys

CHILD = r'''
import sys, numpy as np, torch
sys.path.insert(0, ".")
from gc_slam_b200 import _lib as L, hypothesis_batch as HB, primitives as PR, synth
n = 65536
atlas_np = synth.synthetic_atlas(1_000_000, 50000, 7, scan_seq=20)
amap = PR.AtlasMap.from_numpy(atlas_np, n_tiles_cap=len(atlas_np["tiles"]) + 16)
pts, t, w, _, _ = synth.vlp16_scan(n, 4242, t0=synth.EPOCH_T0)
cam = synth.camera_splats(512, 99)
base = PR.measurement_batch_from_camera_splats(cam["positions"], cam["covariances"], cam["directions"], cam["kappas"],
                                               cam["weights"], cam["timestamps"], cam["colors"])
pts_d, t_d, w_d = [torch.from_numpy(a).cuda() for a in (pts, t, w)]
t0, t1 = synth.EPOCH_T0, synth.EPOCH_T0 + 0.1
ctx = L.context()
out = []
for H in (1, 64):
    xis = torch.from_numpy(np.stack([synth.scan_twist(4242 + h) for h in range(H)])).cuda()
    poses = synth.hypothesis_poses(H, 42) * 0.2
    poses[:, :3] += np.array([0.1, -0.2, 0.5])
    run = lambda k: HB.lidar_evidence_primitives_batched(pts_d, t_d, w_d, t0, t1, xis, amap, poses, 30 + k, base_batch=base, update_map=False)
    for k in range(4): run(k)
    ds = []
    for k in range(12):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(k); e1.record(); torch.cuda.synchronize()
        ds.append(e0.elapsed_time(e1))
    tags = {}
    for tag in ("topk", "sinkhorn"):
        ctx.timing_enable(True, only=tag)
        for k in range(6): run(k)
        torch.cuda.synchronize()
        ms, cnt = ctx.timing_collect(); ctx.timing_enable(False)
        tags[tag] = 1e3 * ms / max(cnt, 1)
    import time
    pl = []
    for rep in range(3):
        torch.cuda.synchronize(); a0 = time.perf_counter(); prev = None
        for k in range(30):
            cur = HB.lidar_evidence_primitives_batched(pts_d, t_d, w_d, t0, t1, xis, amap, poses, 60 + k, base_batch=base, update_map=False, defer=True)
            if prev is not None: prev.wait()
            prev = cur
        prev.wait(); torch.cuda.synchronize()
        pl.append(1e3 * (time.perf_counter() - a0) / 30)
    out.append(f"H={H}: span {float(np.median(ds)):.3f} ms, pipelined {min(pl):.3f} ms, topk {tags['topk']:.0f} us, sinkhorn {tags['sinkhorn']:.0f} us")
print(" | ".join(out))
'''

def main():
    names = sys.argv[1:] or ["main"]
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for rnd in range(2):
        for n in names:
            env = dict(os.environ)
            label = n
            n, *sets = n.split(":")          # name[:ENV=VALUE...]
            for kv in sets:
                k, v = kv.split("=")
                env[k] = v
            if n != "main":
                env["GCS_B200_LIB"] = os.path.join(here, "gc-slam_b200", "lib", "variants", f"libgcs_b200.{n}.so")
            else:
                env.pop("GCS_B200_LIB", None)
            r = subprocess.run([sys.executable, "-c", CHILD], cwd=here, env=env, capture_output=True, text=True, timeout=600)
            print(f"{label:22s} {r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else 'ERR ' + r.stderr[-400:]}", flush=True)

if __name__ == "__main__":
    main()
