"""
Certificate contract types returned by every operator: ``(result, CertBundle, ExpectedEffect)``.

The field names, defaults and ``to_dict`` keys are the reference's audit schema
(fl_ws/src/fl_slam_poc/fl_slam_poc/common/certificates.py:21-503, asserted by its
test/test_cert_schema.py) because downstream consumers (pipeline aggregation, /gc/certificate JSON)
read them by name.  The implementation is our own: one small mixin instead of per-class boilerplate.
"""

from __future__ import annotations

import dataclasses as _dc
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional


class _Rec:
    """Shallow, order-preserving ``to_dict`` for flat certificate records."""

    def to_dict(self) -> Dict[str, Any]:
        out = {}
        for f in _dc.fields(self):
            v = getattr(self, f.name)
            if isinstance(v, _Rec):
                v = v.to_dict()
            elif isinstance(v, list):
                v = list(v)
            elif isinstance(v, dict):
                v = dict(v)
            out[f.name] = v
        return out


@dataclass
class ConditioningCert(_Rec):
    eig_min: float = 1.0
    eig_max: float = 1.0
    cond: float = 1.0
    near_null_count: int = 0


@dataclass
class SupportCert(_Rec):
    ess_total: float = 0.0
    support_frac: float = 1.0


@dataclass
class MismatchCert(_Rec):
    nll_per_ess: float = 0.0
    directional_score: float = 1.0


@dataclass
class ExcitationCert(_Rec):
    dt_effect: float = 0.0
    extrinsic_effect: float = 0.0


@dataclass
class InfluenceCert(_Rec):
    lift_strength: float = 0.0
    psd_projection_delta: float = 0.0
    nu_projection_delta: float = 0.0
    mass_epsilon_ratio: float = 0.0
    anchor_drift_rho: float = 0.0
    dt_scale: float = 1.0
    extrinsic_scale: float = 1.0
    trust_alpha: float = 1.0
    power_beta: float = 1.0

    @classmethod
    def identity(cls) -> "InfluenceCert":
        return cls()

    def with_overrides(self, **kw: Any) -> "InfluenceCert":
        return _dc.replace(self, **kw)


@dataclass
class OverconfidenceCert(_Rec):
    excitation_total: float = 0.0
    ess_to_excitation: float = 0.0
    cond_to_support: float = 0.0
    dt_asymmetry: float = 0.0
    z_to_xy_ratio: float = 0.0
    ess_growth_rate: float = 0.0
    excitation_growth_rate: float = 0.0
    nullspace_energy_ratio: float = 0.0


@dataclass
class OTCert(_Rec):
    marginal_defect_a: float = 0.0
    marginal_defect_b: float = 0.0
    transport_mass_total: float = 0.0
    dual_gap_proxy: float = 0.0
    sum_a: float = 0.0
    sum_b: float = 0.0
    sum_m: float = 0.0
    sum_novel: float = 0.0
    p95_a: float = 0.0
    p95_b: float = 0.0
    nonzero_a: int = 0
    nonzero_b: int = 0
    epsilon: float = 0.0
    tau_a: float = 0.0
    tau_b: float = 0.0
    n_iters: int = 0
    b_policy: str = ""
    b_recency_decay_lambda: float = 0.0
    b_recency_p95: float = 0.0


@dataclass
class MapUpdateCert(_Rec):
    n_active_tiles: int = 0
    tile_ids_active: List[int] = field(default_factory=list)
    n_inactive_tiles: int = 0
    tile_ids_inactive: List[int] = field(default_factory=list)
    tile_cache_hits: int = 0
    tile_cache_misses: int = 0
    candidate_tiles_per_meas_mean: float = 0.0
    candidate_primitives_per_meas_mean: float = 0.0
    candidate_primitives_per_meas_p95: float = 0.0
    insert_count_total: int = 0
    insert_mass_total: float = 0.0
    insert_mass_p95: float = 0.0
    evicted_count: int = 0
    evicted_mass_total: float = 0.0
    fused_count: int = 0
    fused_mass_total: float = 0.0
    merged_count: int = 0
    staleness_inflation_strength: float = 0.0
    staleness_cov_inflation_trace: float = 0.0
    stale_precision_downscale_total: float = 0.0


@dataclass
class ScanIOCert(_Rec):
    scan_seq: int = 0
    scan_stamp_sec: float = 0.0
    scan_window_start_sec: float = 0.0
    scan_window_end_sec: float = 0.0
    streams: Dict[str, Dict[str, float]] = field(default_factory=dict)


@dataclass
class DeviceRuntimeCert(_Rec):
    """Bytes actually copied and host syncs actually performed by the ctypes layer (not estimates here)."""

    host_sync_count_est: int = 0
    device_to_host_bytes_est: int = 0
    host_to_device_bytes_est: int = 0
    jit_recompile_count: int = 0


@dataclass
class ComputeCert(_Rec):
    alloc_bytes_est: int = 0
    largest_tensor_shape: tuple = (0, 0)
    segment_sum_k: int = 0
    psd_projection_count: int = 0
    chol_solve_count: int = 0
    scan_io: ScanIOCert = field(default_factory=ScanIOCert)
    device_runtime: DeviceRuntimeCert = field(default_factory=DeviceRuntimeCert)


_SUBCERTS = (
    ("conditioning", ConditioningCert), ("support", SupportCert), ("mismatch", MismatchCert),
    ("excitation", ExcitationCert), ("influence", InfluenceCert), ("overconfidence", OverconfidenceCert),
    ("compute", ComputeCert),
)


@dataclass
class CertBundle:
    chart_id: str
    anchor_id: str
    exact: bool
    approximation_triggers: List[str] = field(default_factory=list)
    frobenius_applied: bool = False
    conditioning: ConditioningCert = field(default_factory=ConditioningCert)
    support: SupportCert = field(default_factory=SupportCert)
    mismatch: MismatchCert = field(default_factory=MismatchCert)
    excitation: ExcitationCert = field(default_factory=ExcitationCert)
    influence: InfluenceCert = field(default_factory=InfluenceCert)
    overconfidence: OverconfidenceCert = field(default_factory=OverconfidenceCert)
    compute: ComputeCert = field(default_factory=ComputeCert)
    ot: Optional[OTCert] = None
    map_update: Optional[MapUpdateCert] = None

    @classmethod
    def _make(cls, chart_id, anchor_id, exact, triggers, frobenius_applied, kw):
        parts = {name: (kw.get(name) or typ()) for name, typ in _SUBCERTS}
        return cls(chart_id=chart_id, anchor_id=anchor_id, exact=exact, approximation_triggers=list(triggers),
                   frobenius_applied=frobenius_applied, ot=kw.get("ot"), map_update=kw.get("map_update"), **parts)

    @classmethod
    def create_exact(cls, chart_id: str, anchor_id: str, **kw) -> "CertBundle":
        return cls._make(chart_id, anchor_id, True, [], False, kw)

    @classmethod
    def create_approx(cls, chart_id: str, anchor_id: str, triggers: List[str], frobenius_applied: bool = False,
                      **kw) -> "CertBundle":
        return cls._make(chart_id, anchor_id, False, triggers, frobenius_applied, kw)

    def total_trigger_magnitude(self) -> float:
        i = self.influence
        return (i.lift_strength + i.psd_projection_delta + i.nu_projection_delta + i.mass_epsilon_ratio
                + i.anchor_drift_rho + abs(1.0 - i.dt_scale) + abs(1.0 - i.extrinsic_scale)
                + abs(1.0 - i.trust_alpha) + abs(1.0 - i.power_beta))

    def to_dict(self) -> Dict[str, Any]:
        d: Dict[str, Any] = {
            "chart_id": self.chart_id, "anchor_id": self.anchor_id, "exact": self.exact,
            "approximation_triggers": self.approximation_triggers, "frobenius_applied": self.frobenius_applied,
        }
        for name, _ in _SUBCERTS:
            d[name] = getattr(self, name).to_dict()
        d["total_trigger_magnitude"] = self.total_trigger_magnitude()
        if self.ot is not None:
            d["ot"] = self.ot.to_dict()
        if self.map_update is not None:
            d["map_update"] = self.map_update.to_dict()
        return d


@dataclass
class ExpectedEffect(_Rec):
    objective_name: str
    predicted: float
    realized: Optional[float] = None


def aggregate_certificates(certs: List[CertBundle]) -> CertBundle:
    """Pipeline-level summary (certificates.py:515-708): worst-case conditioning, mean support, summed
    mismatch / influence magnitudes, max excitation / overconfidence / compute, merged OT and map-update."""
    if not certs:
        return CertBundle.create_exact(chart_id="unknown", anchor_id="unknown")
    n = len(certs)

    def col(path):
        a, b = path.split(".")
        return [getattr(getattr(c, a), b) for c in certs]

    cond = ConditioningCert(min(col("conditioning.eig_min")), max(col("conditioning.eig_max")),
                            max(col("conditioning.cond")), sum(col("conditioning.near_null_count")))
    sup = SupportCert(sum(col("support.ess_total")) / n, sum(col("support.support_frac")) / n)
    mis = MismatchCert(sum(col("mismatch.nll_per_ess")), sum(col("mismatch.directional_score")) / n)
    exc = ExcitationCert(max(col("excitation.dt_effect")), max(col("excitation.extrinsic_effect")))
    inf = InfluenceCert(
        lift_strength=sum(col("influence.lift_strength")),
        psd_projection_delta=sum(col("influence.psd_projection_delta")),
        nu_projection_delta=sum(col("influence.nu_projection_delta")),
        mass_epsilon_ratio=max(col("influence.mass_epsilon_ratio")),
        anchor_drift_rho=max(col("influence.anchor_drift_rho")),
        dt_scale=min(col("influence.dt_scale")), extrinsic_scale=min(col("influence.extrinsic_scale")),
        trust_alpha=min(col("influence.trust_alpha")), power_beta=min(col("influence.power_beta")))
    over = OverconfidenceCert(**{f.name: max(col(f"overconfidence.{f.name}")) for f in _dc.fields(OverconfidenceCert)})

    def shape_score(shape):
        if isinstance(shape, (tuple, list)) and shape:
            p = 1
            for v in shape:
                try:
                    p *= int(v)
                except Exception:
                    return 0
            return p
        return 0

    comp = ComputeCert(
        alloc_bytes_est=max(c.compute.alloc_bytes_est for c in certs),
        largest_tensor_shape=max((c.compute.largest_tensor_shape for c in certs), key=shape_score, default=(0, 0)),
        segment_sum_k=max(c.compute.segment_sum_k for c in certs),
        psd_projection_count=max(c.compute.psd_projection_count for c in certs),
        chol_solve_count=max(c.compute.chol_solve_count for c in certs),
        scan_io=max((c.compute.scan_io for c in certs), key=lambda s: s.scan_seq),
        device_runtime=DeviceRuntimeCert(
            host_sync_count_est=max(c.compute.device_runtime.host_sync_count_est for c in certs),
            device_to_host_bytes_est=max(c.compute.device_runtime.device_to_host_bytes_est for c in certs),
            host_to_device_bytes_est=max(c.compute.device_runtime.host_to_device_bytes_est for c in certs),
            jit_recompile_count=max(c.compute.device_runtime.jit_recompile_count for c in certs)))

    ots = [c.ot for c in certs if c.ot is not None]
    ot = None
    if ots:
        mx = lambda k: max(getattr(o, k) for o in ots)  # noqa: E731
        sm = lambda k: sum(getattr(o, k) for o in ots)  # noqa: E731
        f0 = ots[0]
        ot = OTCert(marginal_defect_a=mx("marginal_defect_a"), marginal_defect_b=mx("marginal_defect_b"),
                    transport_mass_total=sm("transport_mass_total"), dual_gap_proxy=mx("dual_gap_proxy"),
                    sum_a=sm("sum_a"), sum_b=sm("sum_b"), sum_m=sm("sum_m"), sum_novel=sm("sum_novel"),
                    p95_a=mx("p95_a"), p95_b=mx("p95_b"), nonzero_a=sm("nonzero_a"), nonzero_b=sm("nonzero_b"),
                    epsilon=f0.epsilon, tau_a=f0.tau_a, tau_b=f0.tau_b, n_iters=f0.n_iters, b_policy=f0.b_policy,
                    b_recency_decay_lambda=f0.b_recency_decay_lambda, b_recency_p95=mx("b_recency_p95"))

    mcs = [c.map_update for c in certs if c.map_update is not None]
    mu = None
    if mcs:
        act, inact = [], []
        for m in mcs:
            act.extend(m.tile_ids_active)
            inact.extend(m.tile_ids_inactive)
        mx = lambda k: max(getattr(o, k) for o in mcs)  # noqa: E731
        sm = lambda k: sum(getattr(o, k) for o in mcs)  # noqa: E731
        mu = MapUpdateCert(
            n_active_tiles=len(set(act)), tile_ids_active=list(set(act)), n_inactive_tiles=len(set(inact)),
            tile_ids_inactive=list(set(inact)), tile_cache_hits=sm("tile_cache_hits"),
            tile_cache_misses=sm("tile_cache_misses"),
            candidate_tiles_per_meas_mean=mx("candidate_tiles_per_meas_mean"),
            candidate_primitives_per_meas_mean=mx("candidate_primitives_per_meas_mean"),
            candidate_primitives_per_meas_p95=mx("candidate_primitives_per_meas_p95"),
            insert_count_total=sm("insert_count_total"), insert_mass_total=sm("insert_mass_total"),
            insert_mass_p95=mx("insert_mass_p95"), evicted_count=sm("evicted_count"),
            evicted_mass_total=sm("evicted_mass_total"), fused_count=sm("fused_count"),
            fused_mass_total=sm("fused_mass_total"), merged_count=sm("merged_count"),
            staleness_inflation_strength=mx("staleness_inflation_strength"),
            staleness_cov_inflation_trace=mx("staleness_cov_inflation_trace"),
            stale_precision_downscale_total=sm("stale_precision_downscale_total"))

    trig: List[str] = []
    for c in certs:
        trig.extend(c.approximation_triggers)
    return CertBundle(chart_id=certs[0].chart_id, anchor_id=certs[0].anchor_id,
                      exact=not any(not c.exact for c in certs), approximation_triggers=trig,
                      frobenius_applied=any(c.frobenius_applied for c in certs), conditioning=cond, support=sup,
                      mismatch=mis, excitation=exc, influence=inf, overconfidence=over, compute=comp, ot=ot,
                      map_update=mu)
