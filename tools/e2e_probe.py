"""Where the end-to-end step time of bench.py goes: the same two-plan / two-stream loop with stages switched off."""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from gc_slam_b200 import _lib as L, operators as ops, synth  # noqa: E402

S, P = 128, 65536
bins = synth.fibonacci_atlas(48)
R_bl, t_bl = synth.base_lidar_extrinsics()
msgs = [synth.vlp16_pointcloud2(P, 1000 + k, time_unit="s") for k in range(4)]
payload = torch.from_numpy(np.concatenate([msgs[k % 4][0] for k in range(S)])).pin_memory()
plans = []
for i in range(2):
    pl = ops.BinPathPlan(S, P, P, n_hyp=1, n_bins=48, tau=0.1, origin=synth.lidar_origin_base(), precision=L.PREC_TC,
                         want_evidence=True, materialize_deskewed=True, own_context=True)
    pl.set_bins(bins, 0.1)
    pl.set_map(synth.random_map_bin_stats(48, 7, bins))
    pl.enable_pointcloud2(msgs[0][1], msgs[0][2], R_bl, t_bl)
    plans.append(pl)
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
t0 = torch.zeros(S, dtype=torch.float64).pin_memory()
t1 = torch.full((S,), 0.1, dtype=torch.float64).pin_memory()
xi = torch.from_numpy(np.stack([synth.scan_twist(5 + k) for k in range(S)])).pin_memory()
poses = torch.from_numpy(synth.hypothesis_poses(S, 3)).pin_memory()
oh = [torch.empty(S * (L.BC_NCERT + 22 * 22 + 22), dtype=torch.float64).pin_memory() for _ in range(2)]
n = S * P * msgs[0][2]


def step(k, mode):
    pl, st = plans[k % 2], streams[k % 2]
    with torch.cuda.stream(st):
        if mode == "copy":
            pl._pc2_dev[:n].copy_(payload[:n], non_blocking=True)
            return
        if mode in ("copy+small", "copy+small+run", "copy+run(no small)", "small+copy+run"):
            if mode == "small+copy+run":
                for dst, src in ((pl.t0, t0), (pl.t1, t1), (pl.xi, xi), (pl.poses, poses)):
                    dst.copy_(src.reshape(dst.shape), non_blocking=True)
            pl._pc2_dev[:n].copy_(payload[:n], non_blocking=True)
            if "small" in mode and "no small" not in mode and mode != "small+copy+run":
                for dst, src in ((pl.t0, t0), (pl.t1, t1), (pl.xi, xi), (pl.poses, poses)):
                    dst.copy_(src.reshape(dst.shape), non_blocking=True)
            if "run" in mode:
                io = pl.io
                io.ctx.check(io.ctx.lib.gcs_parse_pointcloud2_vlp16(
                    io.ctx.handle, io.stream(), L.ptr(pl._pc2_dev), pl.S, pl.n_raw, C.byref(pl._pc2_lay), L.ptr(pl._pc2_stamp),
                    pl._pc2_R, pl._pc2_t, L.ptr(pl.pts), L.ptr(pl.t), L.ptr(pl.w), L.ptr(pl.ring), L.ptr(pl.tag), L.ptr(pl._pc2_cert)))
                pl.run()
            return
        pl.upload_pointcloud2(payload, None, t0, t1, xi, poses)
        if mode == "copy+parse":
            return
        pl.run()
        if mode == "copy+parse+run":
            return
        o = pl.outputs()
        oh[k % 2].copy_(torch.cat([o.cert.reshape(-1), o.L22.reshape(-1), o.h22.reshape(-1)]), non_blocking=True)


for mode in ("copy", "copy+small", "copy+run(no small)", "copy+small+run", "small+copy+run", "copy+parse", "copy+parse+run", "full"):
    for k in range(2):
        step(k, mode)
    torch.cuda.synchronize()
    a = time.perf_counter()
    for k in range(10):
        step(k, mode)
    cpu = time.perf_counter() - a
    torch.cuda.synchronize()
    dt = time.perf_counter() - a
    print(f"{mode:16s}: {dt * 100:.3f} ms/step  ({S * 10 / dt:.0f} scans/s, {10 * n / dt / 1e9:.1f} GB/s H2D)  cpu enqueue {cpu * 100:.3f} ms/step")

