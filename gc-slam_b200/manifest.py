"""
RuntimeManifest entries for the drop-in (fl/backend/pipeline.py:1629-1793, `backends` dict :1712-1732).

The reference audits "which implementation ran" through RuntimeManifest.backends; its tests only assert presence
of the keys (test/test_visual_lidar_plan.py:33-67).  `patch_backends()` overwrites the entries this library
replaces with "gcs_sm100a:<entry point>@<version>" and adds keys for the bin family.
"""
from __future__ import annotations

from typing import Dict

from . import _lib

ENTRY_POINTS = {
    "core_array": "torch.cuda+libgcs_b200",
    "se3": "gcs_common.cuh:so3_exp/so3_log/deskew_point",
    "domain_projection_psd": "gcs_common.cuh:psd_project3",
    "pointcloud2_ingest": "gcs_parse_pointcloud2_vlp16",
    "imu_preintegration": "gcs_imu_scan_twist",
    "hypothesis_combine": "gcs_hypothesis_barycenter",
    "point_budget": "gcs_point_budget_resample",
    "deskew": "gcs_deskew_constant_twist",
    "bin_soft_assign": "gcs_bin_soft_assign",
    "scan_bin_moment_match": "gcs_scan_bin_moment_match",
    "kappa": "gcs_kappa_from_resultant_batch",
    "matrix_fisher": "gcs_matrix_fisher_rotation",
    "planar_translation": "gcs_planar_translation",
    "lidar_evidence": "gcs_lidar_evidence_bins|gcs_visual_pose_evidence",
    "surfel_extraction": "gcs_extract_lidar_surfels",
    "map_view": "gcs_extract_atlas_map_view",
    "association": "gcs_associate_primitives_ot",
    "sinkhorn_backend": "gcs_associate_primitives_ot:assoc_sinkhorn_kernel(fixed 50 iter, unbalanced KL)",
    "pose_evidence": "gcs_visual_pose_evidence",
    "map_update": "gcs_map_update",
    "map_recency_inflate": "gcs_map_recency_inflate",
    "map_export": "gcs_export_map_points",
    "map_merge_reduce": "gcs_map_merge_reduce",
    "evidence_fusion": "gcs_evidence_fusion",
    "sparse_cost_matrix": "gcs_sparse_cost_matrix",
    "sinkhorn_unbalanced_fixed_k": "gcs_sinkhorn_unbalanced_fixed_k",
    "map_fuse": "gcs_map_fuse",
    "map_insert_masked": "gcs_map_insert_masked",
    "map_cull": "gcs_map_cull",
    "map_forget": "gcs_map_forget",
}


def backend_ids() -> Dict[str, str]:
    ver = _lib.version_string().split()[-1]
    return {k: f"gcs_sm100a:{v}@{ver}" for k, v in ENTRY_POINTS.items()}


def patch_backends(manifest_backends: Dict[str, str]) -> Dict[str, str]:
    """Return a copy of RuntimeManifest.backends with the entries this library implements replaced."""
    out = dict(manifest_backends)
    out.update(backend_ids())
    return out
