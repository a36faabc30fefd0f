#!/usr/bin/env python
"""
Generate golden vectors by executing the REFERENCE's own operator sources.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

How: ``oracle/jax_shim`` (a NumPy-backed stand-in for the JAX array runtime, which is not installed
here) is put first on sys.path, then the reference package ``fl_slam_poc`` and the archived bin-family
modules are imported unmodified from /root/reference and called on seeded synthetic inputs.  The two
constants the archived modules need but the tree no longer defines (GC_B_BINS, GC_TAU_SOFT_ASSIGN;
fl/common/constants.py:4-5) are injected as attributes on the loaded constants module: 48 (prose) and
0.1 (stated harness value).  Nothing under /root/reference is written or copied.

Outputs: tests/golden/bin_*.npz, tests/golden/prim_*.npz (small, committed).
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "jax_shim"))
sys.path.insert(0, os.path.join(REF, "fl_ws", "src", "fl_slam_poc"))
sys.path.insert(0, ROOT)


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def A(x):
    return np.asarray(x)


class _Belief:
    """Duck-typed belief: the bin-family evidence operators only read the pose and the ids."""

    def __init__(self, pose6):
        self._p = np.asarray(pose6, dtype=np.float64)
        self.chart_id = "GC-RIGHT-01"
        self.anchor_id = "golden"

    def mean_world_pose(self, eps_lift=1e-9):
        import jax.numpy as jnp

        return jnp.asarray(self._p)


def cert_scalars(c):
    return np.array([float(c.exact), c.support.ess_total, c.support.support_frac,
                     c.influence.psd_projection_delta, c.influence.mass_epsilon_ratio,
                     c.conditioning.eig_min, c.conditioning.eig_max, c.conditioning.cond,
                     float(c.conditioning.near_null_count), c.mismatch.nll_per_ess,
                     c.mismatch.directional_score])


FULL_STRIDE = 97


def make_bin_family():
    from fl_slam_poc.common import constants

    constants.GC_B_BINS = 48
    constants.GC_TAU_SOFT_ASSIGN = 0.1
    from fl_slam_poc.backend.operators.point_budget import point_budget_resample
    from fl_slam_poc.backend.operators.deskew_constant_twist import deskew_constant_twist
    from fl_slam_poc.backend.operators.kappa import kappa_from_resultant_batch, kappa_from_resultant_v2
    from fl_slam_poc.common.geometry import se3_jax

    binning = _load("ref_binning", os.path.join(REF, "archive/legacy_operators/binning.py"))
    atlas = _load("ref_bin_atlas", os.path.join(REF, "archive/bin_atlas.py"))
    mfe = _load("ref_mf", os.path.join(REF, "archive/legacy_operators/matrix_fisher_evidence.py"))

    from gc_slam_b200 import synth

    bin_dirs = A(atlas.create_fibonacci_atlas(48).dirs)
    origin = synth.lidar_origin_base()

    # case name -> (n_raw, cap, t0, seed)
    cases = {
        "c1_8192": (8192, 8192, 0.0, 1000),          # BASELINE config 1
        "c2_30000_cap8192": (30000, 8192, synth.EPOCH_T0, 1001),  # config 2, stride 4, epoch stamps (quirk Q1)
        "ragged_5000_cap8192": (5000, 8192, 0.0, 1002),  # padded rows, relative stamps
        "ragged_777_cap1024_epoch": (777, 1024, synth.EPOCH_T0, 1003),  # padded rows + epoch stamps: garbage pads
    }
    def run_case(n_raw, cap, t0, seed):
        pts, t, w, ring, tag = synth.vlp16_scan(n_raw, seed, t0=t0)
        xi = synth.scan_twist(seed)
        t1 = t0 + synth.SCAN_PERIOD
        rs, c_rs, e_rs = point_budget_resample(pts, t, w, ring, tag, n_points_cap=cap)
        dk, c_dk, e_dk = deskew_constant_twist(rs.points, rs.timestamps, rs.weights, t0, t1, xi, 1.0,
                                               "GC-RIGHT-01", "golden")
        rays = A(dk.points) - origin[None, :]
        dirs = rays / (np.linalg.norm(rays, axis=1, keepdims=True) + 1e-12)  # pipeline.py:589-593
        sa, c_sa, e_sa = binning.bin_soft_assign(dirs, bin_dirs, 0.1)
        covs = np.zeros((cap, 3, 3))
        st, c_st, e_st = binning.scan_bin_moment_match(dk.points, covs, dk.weights, sa.responsibilities,
                                                       direction_origin=origin)
        # map side: additive stats of a perturbed copy of the same scan statistics
        ms = synth.random_map_bin_stats(48, seed + 7, bin_dirs)
        map_stats = atlas.MapBinStats(**{k: A(v) for k, v in ms.items()})
        mu_dir, kappa_map, centroid, Sigma_c = atlas.compute_map_derived_stats(map_stats)
        pose = synth.hypothesis_poses(1, seed)[0]
        bel = _Belief(pose)
        mf, c_mf, e_mf = mfe.matrix_fisher_rotation_evidence(bel, st.s_dir, st.S_dir_scatter, st.N,
                                                             map_stats.S_dir, map_stats.S_dir_scatter,
                                                             map_stats.N_dir)
        pt, c_pt, e_pt = mfe.planar_translation_evidence(bel, st.p_bar, st.Sigma_p, st.N, centroid, Sigma_c,
                                                         map_stats.N_pos, map_stats.S_dir_scatter,
                                                         map_stats.N_dir, mf.R_mf)
        L22, h22 = mfe.build_combined_lidar_evidence_22d(mf, pt)
        forgot = atlas.apply_forgetting(map_stats, 0.99)
        resp = A(sa.responsibilities)
        # update_map_stats (archive/bin_atlas.py:137): the map statistics above plus the scan's own additive increments
        inc_sum_p = A(st.p_bar) * A(st.N)[:, None]
        inc_sum_ppT = (A(st.Sigma_p) + np.einsum("bi,bj->bij", A(st.p_bar), A(st.p_bar))) * A(st.N)[:, None, None]
        upd = atlas.update_map_stats(map_stats, st.s_dir, st.S_dir_scatter, st.N, 0.5 * A(st.N), inc_sum_p, inc_sum_ppT)
        return dict(
            n_raw=n_raw, cap=cap, t0=t0, t1=t1, seed=seed, xi=xi, tau=0.1, origin=origin, bin_dirs=bin_dirs,
            pose=pose,
            rs_points=A(rs.points), rs_t=A(rs.timestamps), rs_w=A(rs.weights), rs_ring=A(rs.ring),
            rs_tag=A(rs.tag), rs_n_output=rs.n_output, rs_mass_in=rs.total_mass_in, rs_cert=cert_scalars(c_rs),
            rs_effect=e_rs.predicted,
            dk_points=A(dk.points), dk_w=A(dk.weights), dk_cert=cert_scalars(c_dk),
            dirs=dirs,
            # responsibilities are (N,48): keep row sums of a strided sample + three reductions instead of 3 MB
            resp_rows=resp[:: max(1, cap // 64)], resp_colsum=resp.sum(0), resp_max=resp.max(),
            sa_cert=cert_scalars(c_sa), sa_effect=e_sa.predicted,
            st_N=A(st.N), st_s_dir=A(st.s_dir), st_S=A(st.S_dir_scatter), st_p_bar=A(st.p_bar),
            st_Sigma_p=A(st.Sigma_p), st_kappa=A(st.kappa_scan), st_cert=cert_scalars(c_st),
            st_effect=e_st.predicted,
            **{f"map_{k}": v for k, v in ms.items()},
            map_mu_dir=A(mu_dir), map_kappa=A(kappa_map), map_centroid=A(centroid), map_Sigma_c=A(Sigma_c),
            map_forgot_N_dir=A(forgot.N_dir), map_forgot_sum_ppT=A(forgot.sum_ppT),
            mf_R=A(mf.R_mf), mf_L=A(mf.L_rot), mf_h=A(mf.h_rot), mf_delta=A(mf.delta_rot),
            mf_s=A(mf.svd_singular_values), mf_cert=cert_scalars(c_mf), mf_effect=e_mf.predicted,
            mf_scan_metrics=np.array([mf.scan_scatter_metrics.linearity, mf.scan_scatter_metrics.planarity,
                                      mf.scan_scatter_metrics.sphericity, mf.scan_scatter_metrics.anisotropy,
                                      mf.scan_scatter_metrics.effective_rank]),
            mf_map_metrics=np.array([mf.map_scatter_metrics.linearity, mf.map_scatter_metrics.planarity,
                                     mf.map_scatter_metrics.sphericity, mf.map_scatter_metrics.anisotropy,
                                     mf.map_scatter_metrics.effective_rank]),
            mf_scan_eigs=A(mf.scan_scatter_metrics.eigenvalues), mf_map_eigs=A(mf.map_scatter_metrics.eigenvalues),
            pt_t=A(pt.t_wls), pt_L=A(pt.L_trans), pt_h=A(pt.h_trans), pt_delta=A(pt.delta_trans),
            pt_scales=np.array([pt.xy_info_scale, pt.z_info_scale]), pt_cert=cert_scalars(c_pt),
            pt_effect=e_pt.predicted, L22=A(L22), h22=A(h22),
            upd_inc_N_pos=0.5 * A(st.N), upd_inc_sum_p=inc_sum_p, upd_inc_sum_ppT=inc_sum_ppT,
            upd_S_dir=A(upd.S_dir), upd_S_dir_scatter=A(upd.S_dir_scatter), upd_N_dir=A(upd.N_dir),
            upd_N_pos=A(upd.N_pos), upd_sum_p=A(upd.sum_p), upd_sum_ppT=A(upd.sum_ppT),
        )

    which = set(os.environ.get("GCS_GOLDEN_BIN", "small,full").split(","))
    if "small" in which:
        for name, (n_raw, cap, t0, seed) in cases.items():
            d = run_case(n_raw, cap, t0, seed)
            np.savez_compressed(os.path.join(HERE, f"bin_{name}.npz"), **d)
            print("wrote", name, "n_sel", d["rs_n_output"], "kappa[:3]", d["st_kappa"][:3])

    # BASELINE config 3 / the bench shape: 65,536 points, stride 1, epoch stamps.  Per-point arrays are kept as every
    # FULL_STRIDE-th row + SHA-256 digests (bit-exact arrays) + column sums, so that the fixture stays small.
    if "full" in which:
        import hashlib
        for name, (n_raw, cap, t0, seed) in {"c3_65536": (65536, 65536, synth.EPOCH_T0, 1004)}.items():
            d = run_case(n_raw, cap, t0, seed)
            big = ("rs_points", "rs_t", "rs_w", "rs_ring", "rs_tag", "dk_points", "dk_w", "dirs")
            out = {k: v for k, v in d.items() if k not in big}
            out["full_stride"] = FULL_STRIDE
            for k in big:
                a = np.ascontiguousarray(d[k])
                out[k + "_rows"] = a[::FULL_STRIDE]
                out[k + "_sha256"] = np.frombuffer(hashlib.sha256(a.tobytes()).digest(), np.uint8)
                out[k + "_sum"] = a.astype(np.float64).sum(axis=0)
                out[k + "_abssum"] = np.abs(a.astype(np.float64)).sum(axis=0)
            np.savez_compressed(os.path.join(HERE, f"binfull_{name}.npz"), **out)
            print("wrote full", name, "n_sel", d["rs_n_output"], "kappa[:3]", d["st_kappa"][:3])

    # scalar known-answer table: kappa batch + scalar variant, so3 exp/log
    Rb = np.concatenate([np.linspace(0, 1, 41), [0.799, 0.8, 0.801, 0.999999, 1.0, 1.5, -0.2]])
    kb = A(kappa_from_resultant_batch(Rb))
    ks = np.array([kappa_from_resultant_v2(float(r))[0].kappa for r in Rb])
    rng = np.random.default_rng(5)
    rv = np.concatenate([rng.normal(size=(16, 3)), 1e-9 * rng.normal(size=(4, 3)),
                         (np.pi - 1e-9) * np.eye(3), np.zeros((1, 3))])
    Rm = np.stack([A(se3_jax.so3_exp(v)) for v in rv])
    lg = np.stack([A(se3_jax.so3_log(R)) for R in Rm])
    xi6 = np.concatenate([rng.normal(size=(16, 6)), 1e-9 * rng.normal(size=(4, 6))])
    ex = np.stack([A(se3_jax.se3_exp(x)) for x in xi6])
    np.savez_compressed(os.path.join(HERE, "bin_scalars.npz"), R_bar=Rb, kappa_batch=kb, kappa_scalar=ks,
                        rotvec=rv, so3_exp=Rm, so3_log=lg, xi6=xi6, se3_exp=ex)
    print("wrote scalars")


if __name__ == "__main__":
    which = sys.argv[1:] or ["bin", "prim"]
    if "bin" in which:
        make_bin_family()
    if "prim" in which:
        try:
            from make_golden_prim import make_primitive_family
        except ImportError:
            make_primitive_family = None
        if make_primitive_family is not None:
            make_primitive_family(_load, A, cert_scalars)
