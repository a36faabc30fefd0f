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


def bind_to_gpu_numa_node(device_index: int):
    """
    Keep the calling process's host memory on the NUMA node its GPU hangs off, so that pinned staging buffers are
    first-touched there.  With one rank per GPU and eight GPUs on a two-socket host, ranks that stage through the other
    socket's memory lower the aggregate host -> device rate.  Two means, in this order: pin the process to the node's CPUs
    (sysfs `local_cpulist` of the PCI device) when the process is allowed to run on any of them; otherwise (a container
    whose CPU set lies on one socket) ask the kernel to prefer the node for this process's allocations
    (set_mempolicy(MPOL_PREFERRED)).  Returns the CPU list it bound to (a list, as before), or a dict
    {"node": n, "mempolicy": True} for the second means, or None when the topology cannot be read or neither is possible
    (nothing is changed then).  `numa_binding_note()` describes the outcome for logs.
    """
    import os

    global _NUMA_NOTE
    try:
        import torch

        pr = torch.cuda.get_device_properties(int(device_index))
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            cpus = parse_cpulist(f.read())
        node = -1
        try:
            with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
                node = int(f.read().strip())
        except Exception:
            pass
        allowed = os.sched_getaffinity(0)
        cpus = sorted(set(cpus) & set(allowed))
        if cpus:
            os.sched_setaffinity(0, cpus)
            _NUMA_NOTE = f"gpu {device_index} ({bdf}) node {node}: bound to {len(cpus)} of its cpus"
            return cpus
        if node >= 0 and _prefer_numa_node(node):
            _NUMA_NOTE = f"gpu {device_index} ({bdf}) node {node}: none of its cpus allowed here, memory policy PREFERRED node {node}"
            return {"node": node, "mempolicy": True}
        _NUMA_NOTE = f"gpu {device_index} ({bdf}) node {node}: none of its cpus allowed here and set_mempolicy refused; unbound"
        return None
    except Exception as e:
        _NUMA_NOTE = f"gpu {device_index}: topology unreadable ({type(e).__name__}); unbound"
        return None


_NUMA_NOTE = "not attempted"


def numa_binding_note() -> str:
    return _NUMA_NOTE


def _prefer_numa_node(node: int) -> bool:
    """set_mempolicy(MPOL_PREFERRED, {node}) for the calling thread (x86-64 / aarch64 syscall numbers); False when refused."""
    import ctypes
    import platform

    nr = {"x86_64": 238, "aarch64": 237}.get(platform.machine())
    if nr is None or node < 0 or node >= 1024:
        return False
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        words = (ctypes.c_ulong * 16)()
        words[node // 64] = 1 << (node % 64)
        MPOL_PREFERRED = 1
        rc = libc.syscall(ctypes.c_long(nr), ctypes.c_long(MPOL_PREFERRED), ctypes.byref(words), ctypes.c_ulong(16 * 64 + 1))
        return rc == 0
    except Exception:
        return False


def parse_cpulist(text: str):
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11] (the kernel's cpulist format)."""
    out = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            out.extend(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return out


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
    if n_sel_total < int(world):
        raise ValueError(f"point_shard_rows: {n_sel_total} selected rows cannot be split over {world} ranks "
                         "(every rank needs at least one row: use fewer ranks for clouds this small)")
    lo, hi = shard_range(n_sel_total, world, rank)          # shard the SELECTED rows evenly
    row0, row1 = lo * stride, min(hi * stride, int(n_raw_total))
    # padded output rows (cap - n_sel) are dealt to the last rank so that their count is preserved globally
    pad = int(n_points_cap) - n_sel_total
    cap_local = (hi - lo) + (pad if rank == world - 1 else 0)
    return dict(row0=row0, n_raw=row1 - row0, cap_local=cap_local, stride=stride, cap_total=int(n_points_cap))


def allreduce_bin_sums(mass=None, raw_sums=None, raw_max=None, group=None):
    """In-place all-reduce of the additive statistics of the bin path (SUM; MAX for the running maxima): three library
    all-reduces.  Kept for callers that hold the three arrays separately; run_point_sharded uses PointShardExchange."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    if mass is not None:
        dist.all_reduce(mass, op=dist.ReduceOp.SUM, group=group)
    if raw_sums is not None:
        dist.all_reduce(raw_sums, op=dist.ReduceOp.SUM, group=group)
    if raw_max is not None:
        dist.all_reduce(raw_max, op=dist.ReduceOp.MAX, group=group)


class PointShardExchange:
    """
    Buffers of the two exchanges of a point-sharded bin path, allocated once per plan.  Each exchange moves ONE packed
    buffer -- this rank's partial block [additive | maxima] -- and reduces the `world` blocks in rank order on every rank
    (same data, same order: bit-identical results everywhere):
      * peer windows (default on CUDA): gcs_peer_xchg_reduce, one kernel per rank that pushes the block into every rank's
        IPC-mapped receive window over NVLink, signals, waits and reduces -- no library collective;
      * all-gather path (host tensors, IPC not available, use_peer=False / GCS_EXCHANGE=nccl): one
        all_gather_into_tensor (ncclAllGather) + gcs_bins_reduce_gathered.
    `mass`, `raw_sums`, `raw_max` are views into the packed buffers, so the bin kernels write their partials straight into
    the exchange buffer and read the reduced values from the same place: no copy on either side of the exchange.
    """

    def __init__(self, plan, group=None, use_peer=None):
        self._use_peer = use_peer
        import torch
        import torch.distributed as dist

        self.plan, self.group = plan, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        dev = plan.io.dev
        self.n_mass = plan.S * 4
        self.n_sum, self.n_max = plan.U * plan.raw_len, plan.U * 2
        self.pack1 = torch.zeros(self.n_mass, dtype=torch.float64, device=dev)
        self.pack2 = torch.zeros(self.n_sum + self.n_max, dtype=torch.float64, device=dev)
        self.mass = self.pack1.view(plan.S, 4)
        self.raw_sums = self.pack2[:self.n_sum].view(plan.U, plan.raw_len)
        self.raw_max = self.pack2[self.n_sum:].view(plan.U, 2)
        self.gath1 = torch.empty(self.world * self.n_mass, dtype=torch.float64, device=dev)
        self.gath2 = torch.empty(self.world * (self.n_sum + self.n_max), dtype=torch.float64, device=dev)
        self.ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if dev.type == "cuda" else None
        # Peer-memory exchange (one kernel per rank over NVLink peer windows, no library collective) when every rank can
        # map every other rank's window; otherwise -- IPC not permitted, host tensors, GCS_EXCHANGE=nccl -- all ranks use
        # the all-gather path.  The decision is collective: either every rank uses peer windows or none does.
        self.peer, self.peer_note = None, "single rank"
        if self.world > 1:
            self._setup_peer(dev)

    def _setup_peer(self, dev):
        import ctypes as C
        import os

        import torch.distributed as dist

        from . import _lib as L

        want = dev.type == "cuda" and (os.environ.get("GCS_EXCHANGE", "peer") == "peer" if self._use_peer is None else bool(self._use_peer))
        handle, buf, err = C.c_void_p(), C.create_string_buffer(64), ""
        ok = False
        if want:
            io = self.plan.io
            rank = dist.get_rank(self.group)
            rc = io.ctx.lib.gcs_peer_xchg_create(io.ctx.handle, rank, self.world, 8 * max(self.n_mass, self.n_sum + self.n_max),
                                                 C.byref(handle), buf)
            ok = rc == 0
            if not ok:
                err = io.ctx.lib.gcs_last_error(io.ctx.handle).decode()
        infos = [None] * self.world
        dist.all_gather_object(infos, (ok, buf.raw if ok else b"", err), group=self.group)
        if all(i[0] for i in infos):
            rc = io.ctx.lib.gcs_peer_xchg_connect(io.ctx.handle, handle, b"".join(i[1] for i in infos))
            ok2 = rc == 0
            err = "" if ok2 else io.ctx.lib.gcs_last_error(io.ctx.handle).decode()
            oks = [None] * self.world
            dist.all_gather_object(oks, (ok2, err), group=self.group)
            if all(o[0] for o in oks):
                self.peer, self.peer_note = handle, "peer windows (CUDA IPC over NVLink), one kernel per exchange"
                return
            self.peer_note = "all-gather path: " + "; ".join(o[1] for o in oks if o[1])
        else:
            self.peer_note = "all-gather path" + ("" if not want else ": " + "; ".join(i[2] for i in infos if i[2]))
        if ok:
            self.plan.io.ctx.lib.gcs_peer_xchg_destroy(self.plan.io.ctx.handle, handle)

    def close(self):
        if self.peer is not None:
            self.plan.io.ctx.lib.gcs_peer_xchg_destroy(self.plan.io.ctx.handle, self.peer)
            self.peer = None

    def peer_status(self) -> int:
        """0, or the number of the exchange whose wait for a peer timed out (synchronises)."""
        import ctypes as C
        if self.peer is None:
            return 0
        e = C.c_uint32(0)
        io = self.plan.io
        io.ctx.check(io.ctx.lib.gcs_peer_xchg_status(io.ctx.handle, self.peer, C.byref(e)))
        return int(e.value)

    def _exchange(self, pack, gath, n_sum, n_max, e0=None, e1=None):
        import torch
        import torch.distributed as dist

        from . import _lib as L

        if self.world == 1:
            return
        if e0 is not None:
            e0.record()
        if self.peer is not None:
            io = self.plan.io
            io.ctx.check(io.ctx.lib.gcs_peer_xchg_reduce(io.ctx.handle, self.peer, io.stream(), L.ptr(pack), n_sum, n_max))
            if e1 is not None:
                e1.record()
            return
        dist.all_gather_into_tensor(gath, pack, group=self.group)
        if pack.is_cuda:
            io = self.plan.io
            io.ctx.check(io.ctx.lib.gcs_bins_reduce_gathered(io.ctx.handle, io.stream(), L.ptr(gath), self.world, n_sum, n_max,
                                                             L.ptr(pack), L.ptr(pack[n_sum:]) if n_max else None))
        else:   # host tensors (gloo tests of the exchange logic): the same rank-ordered arithmetic
            g = gath.view(self.world, n_sum + n_max)
            acc = g[0, :n_sum].clone()
            for r in range(1, self.world):
                acc += g[r, :n_sum]
            pack[:n_sum] = acc
            if n_max:
                pack[n_sum:] = torch.amax(g[:, n_sum:], dim=0)
        if e1 is not None:
            e1.record()

    def exchange_mass(self, timed=False):
        self._exchange(self.pack1, self.gath1, self.n_mass, 0, *(self.ev[0:2] if timed and self.ev else (None, None)))

    def exchange_sums(self, timed=False):
        self._exchange(self.pack2, self.gath2, self.n_sum, self.n_max, *(self.ev[2:4] if timed and self.ev else (None, None)))

    def exchange_ms(self):
        """Device time of the two exchanges of the last timed run (collective + rank-ordered reduction), in ms."""
        return self.ev[0].elapsed_time(self.ev[1]), self.ev[2].elapsed_time(self.ev[3])


def run_point_sharded(plan, group=None, exchange: "PointShardExchange" = None, timed: bool = False):
    """
    Bin path of clouds whose rows are split across ranks.  `plan` is a BinPathPlan built with this rank's
    (shard_row0, n_raw, cap) and the global (n_raw_total, cap_total).  Three phases with two tiny exchanges, each one
    collective on one packed buffer (PointShardExchange).  Returns (mass, raw_sums, raw_max) -- views into the exchange.
    """
    x = exchange if exchange is not None else PointShardExchange(plan, group)
    x.pack1.zero_()
    x.pack2.zero_()
    plan.run_mass(x.mass)
    x.exchange_mass(timed)
    plan.run_accumulate(x.mass, x.raw_sums, x.raw_max)
    x.exchange_sums(timed)
    plan.run_finalize(x.mass, x.raw_sums, x.raw_max)
    if exchange is None:
        x.close()          # a one-off exchange: waits for the enqueued work and unmaps the peer windows
    return x.mass, x.raw_sums, x.raw_max


def gather_evidence(L22, h22, group=None):
    """
    All-gather per-hypothesis (L, h) from every rank to every rank (4.3 KB per hypothesis).  Ranks may hold different
    numbers of hypotheses (shard_range gives 3 + 2 for five hypotheses on two GPUs): the counts are gathered first, every
    stack is padded to the largest count for the collective and cut back afterwards.
    """
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return L22, h22
    world = dist.get_world_size(group)
    k_local = int(L22.shape[0])
    D = int(h22.shape[1])
    cnt = torch.tensor([k_local], dtype=torch.int64, device=L22.device)
    cnts = torch.empty(world, dtype=torch.int64, device=L22.device)
    dist.all_gather_into_tensor(cnts, cnt, group=group)
    counts = [int(c) for c in cnts.tolist()]
    k_max = max(counts)
    pack = torch.zeros(k_max, D * D + D, dtype=L22.dtype, device=L22.device)   # (L | h) of one hypothesis per row
    if k_local:
        pack[:k_local, :D * D] = L22.reshape(k_local, D * D)
        pack[:k_local, D * D:] = h22
    gath = torch.empty(world * k_max, D * D + D, dtype=L22.dtype, device=L22.device)
    dist.all_gather_into_tensor(gath, pack, group=group)
    rows = torch.cat([gath[r * k_max:r * k_max + counts[r]] for r in range(world)], 0)
    return rows[:, :D * D].reshape(-1, D, D).contiguous(), rows[:, D * D:].contiguous()


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


# --------------------------------------------------------------------------------------------------
# device-side combine: no per-hypothesis host round trip
# --------------------------------------------------------------------------------------------------
class HypothesisProjectionResult:
    """Result of hypothesis_barycenter_projection: the fused information pair on the device (fields of the reference's
    belief_out that this path produces) and the weight-floor adjustment."""

    def __init__(self, L, h, z_lin, weights_normalized, means, floor_adjustment):
        self.L, self.h, self.z_lin = L, h, z_lin
        self.weights_normalized, self.means = weights_normalized, means
        self.floor_adjustment = float(floor_adjustment)


def hypothesis_barycenter_projection(L_stack, h_stack, weights, z_lin_stack=None, K_HYP: int = None,
                                     HYP_WEIGHT_FLOOR: float = 0.0025, eps_psd: float = 1e-12, eps_lift: float = 1e-9,
                                     anchor_id: str = "hypothesis_barycenter"):
    """
    HypothesisBarycenterProjection (fl/backend/operators/hypothesis.py:123-236) on stacked arrays that may already live on
    the device -- e.g. BinEvidenceBatch.L22 / h22 of a 64-hypothesis plan, or the all-gathered stacks of gather_evidence():
    weight floor + renormalisation, barycenter in information form, PSD projection, spread proxy; one kernel
    (gcs_hypothesis_barycenter), one certificate read-back.  Returns (result, CertBundle, ExpectedEffect) with the
    reference's triggers and certificate fields.  A list of belief-like objects (attributes L, h, z_lin) is accepted in
    place of L_stack, as the reference passes it.
    """
    import ctypes as C

    import torch

    from . import _lib as L
    from .certs import CertBundle, ConditioningCert, ExpectedEffect, InfluenceCert, SupportCert
    from .operators import _IO

    if isinstance(L_stack, (list, tuple)) and hasattr(L_stack[0], "L"):
        hyps = L_stack
        anchor_id = getattr(hyps[0], "anchor_id", anchor_id)
        as_t = lambda v: v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v, dtype=np.float64))
        z_lin_stack = torch.stack([as_t(b.z_lin) for b in hyps]) if hasattr(hyps[0], "z_lin") else None
        h_stack = torch.stack([as_t(b.h) for b in hyps])
        L_stack = torch.stack([as_t(b.L) for b in hyps])
    io = _IO()
    Ls = io.dev_in(L_stack)
    if Ls.dim() != 3 or Ls.shape[1] != Ls.shape[2]:
        raise ValueError(f"L_stack must be (K, D, D), got {tuple(Ls.shape)}")
    K, D = int(Ls.shape[0]), int(Ls.shape[1])
    if K_HYP is not None and K != int(K_HYP):
        raise ValueError(f"Expected {K_HYP} hypotheses, got {K}")
    hs = io.dev_in(h_stack)
    w = io.dev_in(weights).reshape(-1)
    if tuple(hs.shape) != (K, D):
        raise ValueError(f"h_stack must be ({K}, {D}), got {tuple(hs.shape)}")
    if tuple(w.shape) != (K,):
        raise ValueError(f"Expected weights shape ({K},), got {tuple(w.shape)}")
    zs = io.dev_in(z_lin_stack) if z_lin_stack is not None else None
    if zs is not None and tuple(zs.shape) != (K, D):
        raise ValueError(f"z_lin_stack must be ({K}, {D}), got {tuple(zs.shape)}")
    L_out, h_out, wn, means, cert_d = io.empty(D, D), io.empty(D), io.empty(K), io.empty(K, D), io.zeros(L.HB["NCERT"])
    z_out = io.empty(D) if zs is not None else None
    io.ctx.check(io.ctx.lib.gcs_hypothesis_barycenter(io.ctx.handle, io.stream(), L.ptr(Ls), L.ptr(hs), L.ptr(zs), L.ptr(w), K, D,
                                                      float(HYP_WEIGHT_FLOOR), float(eps_psd), float(eps_lift), L.ptr(L_out),
                                                      L.ptr(h_out), L.ptr(z_out), L.ptr(wn), L.ptr(means), L.ptr(cert_d)))
    c = io.host(torch.cat([cert_d, wn]))
    cs, wn_h = c[:L.HB["NCERT"]], c[L.HB["NCERT"]:]
    cert = CertBundle.create_approx(
        chart_id="GC-RIGHT-01", anchor_id=anchor_id, triggers=["HypothesisProjection", "I-projection-info-barycenter"],
        conditioning=ConditioningCert(eig_min=float(cs[L.HB["PSD_EIG_MIN"]]), eig_max=float(cs[L.HB["PSD_EIG_MAX"]]),
                                      cond=float(cs[L.HB["PSD_COND"]]), near_null_count=int(cs[L.HB["PSD_NEAR_NULL"]])),
        support=SupportCert(ess_total=float(1.0 / np.sum(wn_h ** 2)),
                            support_frac=float(np.sum(wn_h > HYP_WEIGHT_FLOOR) / K)),
        influence=InfluenceCert.identity().with_overrides(psd_projection_delta=float(cs[L.HB["PSD_PROJECTION_DELTA"]]),
                                                          mass_epsilon_ratio=float(cs[L.HB["FLOOR_ADJUSTMENT"]]) / K),
        compute=io.compute())
    result = HypothesisProjectionResult(L_out, h_out, z_out, wn, means, cs[L.HB["FLOOR_ADJUSTMENT"]])
    _ = C
    return result, cert, ExpectedEffect("predicted_projection_spread_proxy", float(cs[L.HB["SPREAD_PROXY"]]), None)
