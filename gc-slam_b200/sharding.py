"""
Multi-GPU partitioning of the LiDAR evidence path.  One process per GPU, torch.distributed
for the plumbing (NCCL over NVLink on the GPU box, gloo in the CPU tests).

  * pose hypotheses (BASELINE config 4) and independent scans (config 5a): units are independent -> contiguous unit
    ranges per rank, NO data-path collective; per-hypothesis 22-D evidence (4.3 KB each) is gathered and combined on
    the host exactly where the reference combines it (hypothesis_barycenter_projection, ops/hypothesis.py:51-118).
  * one very large cloud, point-sharded (config 5b): every rank accumulates the additive per-bin raw sums of its
    rows (gcs_bins_accumulate); the only exchange is an all-reduce of ~9.7 KB per unit (SUM of the 48 x 25 + 8 raw
    sums, MAX of the 2 running maxima) plus the 4 resample masses per scan; the epilogue (gcs_bins_finalize) is then
    replicated and bit-identical on every rank.
  * surfel extraction / association / map update do not shard ("replicas only").
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def shard_range(n_units: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) of n_units for `rank` (first n % world ranks get one extra)."""
    base, rem = divmod(int(n_units), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def point_shard_rows(n_raw_total: int, n_points_cap: int, world: int, rank: int):
    """
    Row range of a point-sharded cloud.  Shard boundaries are multiples of the resample stride so that every rank
    selects exactly the rows the single-GPU path would (stride = max(1, ceil(n_raw / cap)), point_budget.py:160).
    Returns dict(row0, n_raw, cap_local, stride, cap_total).
    """
    stride = 1 if n_raw_total <= n_points_cap else -(-int(n_raw_total) // int(n_points_cap))
    n_sel_total = -(-int(n_raw_total) // stride)
    lo, hi = shard_range(n_sel_total, world, rank)          # shard the SELECTED rows evenly
    row0, row1 = lo * stride, min(hi * stride, int(n_raw_total))
    # padded output rows (cap - n_sel) are dealt to the last rank so that their count is preserved globally
    pad = int(n_points_cap) - n_sel_total
    cap_local = (hi - lo) + (pad if rank == world - 1 else 0)
    return dict(row0=row0, n_raw=row1 - row0, cap_local=cap_local, stride=stride, cap_total=int(n_points_cap))


def allreduce_bin_sums(mass=None, raw_sums=None, raw_max=None, group=None):
    """In-place all-reduce of the additive statistics of the bin path (SUM; MAX for the running maxima)."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    if mass is not None:
        dist.all_reduce(mass, op=dist.ReduceOp.SUM, group=group)
    if raw_sums is not None:
        dist.all_reduce(raw_sums, op=dist.ReduceOp.SUM, group=group)
    if raw_max is not None:
        dist.all_reduce(raw_max, op=dist.ReduceOp.MAX, group=group)


def run_point_sharded(plan, group=None):
    """
    Bin path of clouds whose rows are split across ranks.  `plan` is a BinPathPlan built with this rank's
    (shard_row0, n_raw, cap) and the global (n_raw_total, cap_total).  Three phases with two tiny collectives.
    """
    import torch

    dev = plan.io.dev
    mass = torch.zeros((plan.S, 4), dtype=torch.float64, device=dev)
    raw_sums = torch.zeros((plan.U, plan.raw_len), dtype=torch.float64, device=dev)
    raw_max = torch.zeros((plan.U, 2), dtype=torch.float64, device=dev)
    plan.run_mass(mass)
    allreduce_bin_sums(mass=mass, group=group)
    plan.run_accumulate(mass, raw_sums, raw_max)
    allreduce_bin_sums(raw_sums=raw_sums, raw_max=raw_max, group=group)
    plan.run_finalize(mass, raw_sums, raw_max)
    return mass, raw_sums, raw_max


def gather_evidence(L22, h22, group=None):
    """All-gather per-hypothesis (L, h) from every rank to every rank (4.3 KB per hypothesis)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return L22, h22
    world = dist.get_world_size(group)
    Ls = [torch.empty_like(L22) for _ in range(world)]
    hs = [torch.empty_like(h22) for _ in range(world)]
    dist.all_gather(Ls, L22, group=group)
    dist.all_gather(hs, h22, group=group)
    return torch.cat(Ls, 0), torch.cat(hs, 0)


def hypothesis_barycenter(L_stack, h_stack, weights, weight_floor: float = 0.0025, eps_psd: float = 1e-12):
    """
    Host-side hypothesis combine (HypothesisBarycenterProjection core, ops/hypothesis.py:51-118): floor and
    renormalise the weights, weighted sum of (L, h), PSD projection of L.  22 x 22 arithmetic on the host, as in the
    reference (backend_node.py:2093).
    """
    L_stack = np.asarray(L_stack, np.float64)
    h_stack = np.asarray(h_stack, np.float64)
    w = np.maximum(np.asarray(weights, np.float64), weight_floor)
    floor_adjustment = float(np.sum(np.abs(w - np.asarray(weights, np.float64))))
    w = w / np.sum(w)
    L_raw = np.einsum("k,kij->ij", w, L_stack)
    h = np.einsum("k,ki->i", w, h_stack)
    Ls = 0.5 * (L_raw + L_raw.T)
    vals, vecs = np.linalg.eigh(Ls)
    L = vecs @ np.diag(np.maximum(vals, eps_psd)) @ vecs.T
    return L, h, w, floor_adjustment
