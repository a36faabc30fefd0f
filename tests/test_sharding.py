"""
Multi-rank logic of the path.

CPU (gloo, world_size 2): shard geometry, the all-reduce of additive per-bin statistics and the host-side
hypothesis combine, with the NumPy oracle standing in for the per-rank kernel.
GPU (nccl, needs >= 2 devices; skipped otherwise): the real point-sharded bin path vs the single-GPU path.
"""
import os
import socket

import numpy as np
import pytest

from conftest import rel_err


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_ranges_cover_and_balance():
    from gc_slam_b200.sharding import point_shard_rows, shard_range
    for n, w in ((64, 8), (10, 4), (7, 8), (10000, 3)):
        spans = [shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    for n_raw, cap, w in ((4194304, 4194304, 8), (30000, 8192, 2), (5000, 8192, 4), (4194304, 1048576, 4)):
        sh = [point_shard_rows(n_raw, cap, w, r) for r in range(w)]
        assert sh[0]["row0"] == 0 and sum(s["n_raw"] for s in sh) == n_raw
        assert all(s["row0"] % s["stride"] == 0 for s in sh)
        assert sum(s["cap_local"] for s in sh) == cap
        for s in sh:
            assert -(-s["n_raw"] // s["stride"]) <= s["cap_local"]


def _oracle_additive(points, t, w, t0, t1, xi, origin, bins, tau, mass_scale):
    """Per-rank stand-in for gcs_bins_accumulate: additive raw sums of the rows this rank owns (oracle arithmetic)."""
    from oracle import bin_path as ob
    dk, _ = ob.deskew_constant_twist(points, t, w * mass_scale, t0, t1, xi)
    d = ob.ray_directions(dk["points"], origin)
    sa, _ = ob.bin_soft_assign(d, bins, tau)
    r = sa["responsibilities"]
    wr = dk["weights"][:, None] * r
    p = dk["points"]
    ent = float(np.sum(-np.sum(r * np.log(r + 1e-12), axis=1)))
    return dict(N=wr.sum(0), s_dir=wr.T @ d, S=np.einsum("nb,ni,nj->bij", wr, d, d), sum_p=wr.T @ p,
                sum_pp=np.einsum("nb,ni,nj->bij", wr, p, p), ent=np.array([ent, np.sum(dk["weights"])]),
                mx=np.array([r.max()]))


def _worker(rank, world, port, n_raw, out_q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gc_slam_b200 import synth
    from types import SimpleNamespace
    from gc_slam_b200.sharding import PointShardExchange, gather_evidence, point_shard_rows, shard_range
    pts, t, w, _, _ = synth.vlp16_scan(n_raw, 5, t0=0.0)
    sh = point_shard_rows(n_raw, n_raw, world, rank)
    sl = slice(sh["row0"], sh["row0"] + sh["n_raw"])
    # the exchange of the point-sharded path (one packed buffer, one all-gather, rank-ordered reduction) on host tensors
    raw_len = 48 * (1 + 3 + 9 + 3 + 9) + 2
    x = PointShardExchange(SimpleNamespace(S=1, U=1, raw_len=raw_len, io=SimpleNamespace(dev=torch.device("cpu"))))
    x.mass[0] = torch.tensor([w[sl].sum(), w[sl].sum(), (w[sl] ** 2).sum(), float(sh["n_raw"])], dtype=torch.float64)
    x.exchange_mass()
    mass = x.mass.clone()
    scale = float(mass[0, 0] / (mass[0, 1] + 1e-12))
    bins = synth.fibonacci_atlas(48)
    a = _oracle_additive(pts[sl], t[sl], w[sl], 0.0, 0.1, synth.scan_twist(5), synth.lidar_origin_base(), bins, 0.1, scale)
    x.raw_sums[0] = torch.from_numpy(np.concatenate([a["N"], a["s_dir"].ravel(), a["S"].ravel(), a["sum_p"].ravel(), a["sum_pp"].ravel(), a["ent"]]))
    x.raw_max[0, 0] = float(a["mx"][0])
    x.exchange_sums()
    raw, mx = x.raw_sums[0].clone(), x.raw_max[0, :1].clone()
    # five hypotheses on two ranks: 3 + 2 (uneven stacks through the gather)
    lo, hi = shard_range(5, world, rank)
    L = torch.stack([torch.full((22, 22), float(k + 1), dtype=torch.float64) for k in range(lo, hi)])
    h = torch.stack([torch.full((22,), float(k + 1), dtype=torch.float64) for k in range(lo, hi)])
    Lg, hg = gather_evidence(L, h)
    if rank == 0:
        out_q.put((mass.numpy(), raw.numpy(), mx.numpy(), Lg.numpy()[:, 0, 0].copy(), hg.shape))
    dist.barrier()
    dist.destroy_process_group()


def test_point_sharded_allreduce_gloo_world2():
    import torch.multiprocessing as mp
    from gc_slam_b200 import synth
    n_raw, world = 6000, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_raw, q)) for r in range(world)]
    for p in procs:
        p.start()
    mass, raw, mx, Ldiag, hshape = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pts, t, w, _, _ = synth.vlp16_scan(n_raw, 5, t0=0.0)
    assert abs(mass[0, 0] - w.sum()) < 1e-9 * w.sum() and mass[0, 3] == n_raw
    scale = float(mass[0, 0] / (mass[0, 1] + 1e-12))
    full = _oracle_additive(pts, t, w, 0.0, 0.1, synth.scan_twist(5), synth.lidar_origin_base(), synth.fibonacci_atlas(48), 0.1, scale)
    ref = np.concatenate([full["N"], full["s_dir"].ravel(), full["S"].ravel(), full["sum_p"].ravel(), full["sum_pp"].ravel(), full["ent"]])
    assert rel_err(raw, ref) < 1e-12 and abs(mx[0] - full["mx"][0]) < 1e-15
    assert list(Ldiag) == [1.0, 2.0, 3.0, 4.0, 5.0] and tuple(hshape) == (5, 22)
    from gc_slam_b200.sharding import point_shard_rows
    with pytest.raises(ValueError):
        point_shard_rows(3, 8192, 4, 0)      # fewer selected rows than ranks: rejected, not an empty shard


def test_hypothesis_barycenter_host_combine():
    from gc_slam_b200.sharding import hypothesis_barycenter
    rng = np.random.default_rng(0)
    A = rng.normal(size=(4, 22, 22))
    Ls = A @ A.transpose(0, 2, 1)
    hs = rng.normal(size=(4, 22))
    w = np.array([0.7, 0.2, 0.1, 0.0])
    L, h, wn, adj = hypothesis_barycenter(Ls, hs, w)
    assert abs(wn.sum() - 1.0) < 1e-15 and wn.min() >= 0.0025 / 1.0025 - 1e-12 and abs(adj - 0.0025) < 1e-15
    assert rel_err(L, np.einsum("k,kij->ij", wn, Ls)) < 1e-12 and np.linalg.eigvalsh(L).min() >= 1e-12 - 1e-18
    assert rel_err(h, wn @ hs) < 1e-14


def _gpu_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from gc_slam_b200 import operators as ops, synth
    from gc_slam_b200.sharding import point_shard_rows, run_point_sharded
    n_raw, cap = 262144, 262144
    pts, t, w, ring, tag = synth.vlp16_scan(n_raw, 9, t0=synth.EPOCH_T0)
    sh = point_shard_rows(n_raw, cap, world, rank)
    sl = slice(sh["row0"], sh["row0"] + sh["n_raw"])
    bins = synth.fibonacci_atlas(48)
    plan = ops.BinPathPlan(1, sh["n_raw"], sh["cap_local"], n_hyp=1, origin=synth.lidar_origin_base(), shard_row0=sh["row0"],
                           n_raw_total=n_raw, cap_total=cap, materialize_deskewed=False)
    plan.set_bins(bins, 0.1)
    plan.set_map(synth.random_map_bin_stats(48, 7, bins))
    plan.upload(pts[sl], t[sl], w[sl], ring[sl], tag[sl], np.array([synth.EPOCH_T0]), np.array([synth.EPOCH_T0 + 0.1]),
                synth.scan_twist(9)[None], synth.hypothesis_poses(1, 3), non_blocking=False)
    from gc_slam_b200.sharding import PointShardExchange
    x = PointShardExchange(plan)                      # peer windows (CUDA IPC over NVLink) when every rank can map them
    run_point_sharded(plan, exchange=x)
    torch.cuda.synchronize()
    out = plan.outputs()
    res = (rank, out.L22.cpu().numpy(), out.stats["Sigma_p"].cpu().numpy(), out.cert.cpu().numpy())
    assert x.peer_status() == 0
    how = x.peer_note
    x.close()
    # the library path (one all-gather of the packed buffer + the rank-ordered reduction kernel) adds the same blocks in the
    # same order: bit-identical to the peer-window kernel
    x2 = PointShardExchange(plan, use_peer=False)
    run_point_sharded(plan, exchange=x2)
    torch.cuda.synchronize()
    out2 = plan.outputs()
    assert np.array_equal(out2.L22.cpu().numpy(), res[1]) and np.array_equal(out2.cert.cpu().numpy(), res[3]), how
    x2.close()
    q.put(res + (how,))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_point_sharded_nccl_matches_single_gpu():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    from gc_slam_b200 import operators as ops, synth
    world = 4 if torch.cuda.device_count() >= 4 else 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda x: x[0])
    print("exchange:", res[0][4])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # every rank ends with bit-identical results (replicated epilogue on all-reduced sums)
    for r in range(1, world):
        assert np.array_equal(res[0][1], res[r][1]) and np.array_equal(res[0][2], res[r][2]) and np.array_equal(res[0][3], res[r][3])
    n_raw = 262144
    pts, t, w, ring, tag = synth.vlp16_scan(n_raw, 9, t0=synth.EPOCH_T0)
    bins = synth.fibonacci_atlas(48)
    plan = ops.BinPathPlan(1, n_raw, n_raw, n_hyp=1, origin=synth.lidar_origin_base(), materialize_deskewed=False)
    plan.set_bins(bins, 0.1)
    plan.set_map(synth.random_map_bin_stats(48, 7, bins))
    plan.upload(pts, t, w, ring, tag, np.array([synth.EPOCH_T0]), np.array([synth.EPOCH_T0 + 0.1]), synth.scan_twist(9)[None],
                synth.hypothesis_poses(1, 3), non_blocking=False)
    plan.run()
    torch.cuda.synchronize()
    o = plan.outputs()
    assert rel_err(res[0][1], o.L22.cpu().numpy()) < 1e-9
    assert rel_err(res[0][2], o.stats["Sigma_p"].cpu().numpy()) < 1e-10
    assert rel_err(res[0][3], o.cert.cpu().numpy()) < 1e-10
