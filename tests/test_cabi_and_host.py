"""
CPU: the C-ABI library loads without a GPU and exports every symbol include/gcs_b200.h declares; the ctypes
prototypes cover all of them; contract types behave like the reference's (schema pins from the reference's
test/test_cert_schema.py and test/test_budget_assertions.py); no product module imports the oracle.
"""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "gcs_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gcs_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from gc_slam_b200 import _lib, fusion, hypothesis_batch, primitives  # noqa: F401  (these modules register their prototypes)
    lib = _lib.load()
    syms = _declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/gcs_b200.h but not exported"
        assert s in _lib.PROTOTYPES, f"{s} has no ctypes prototype"
    assert lib.gcs_version() >= 100 and b"gcs_sm100a" in lib.gcs_version_string()
    assert lib.gcs_bins_raw_sums_len(48) == 48 * 25 + 8


def test_struct_layouts_match_header_sizes(tmp_path):
    """Every ctypes mirror has the size the C compiler gives the header's struct (gcc on include/gcs_b200.h)."""
    import shutil
    import subprocess
    from gc_slam_b200 import _lib, fusion, hypothesis_batch as HB, primitives as P
    pairs = [("gcs_bin_stats", _lib.BinStats), ("gcs_map_bin_stats", _lib.MapBinStats), ("gcs_bins_args", _lib.BinsArgs),
             ("gcs_pc2_layout", _lib.Pc2Layout), ("gcs_meas_batch", P.CMeasBatch), ("gcs_atlas", P.CAtlas), ("gcs_map_view", P.CMapView),
             ("gcs_assoc_result", P.CAssocResult), ("gcs_surfel_cfg", P.CSurfelCfg), ("gcs_assoc_cfg", P.CAssocCfg),
             ("gcs_map_update_cfg", P.CMapUpdateCfg), ("gcs_map_export", P.CMapExport), ("gcs_prim_batch_args", HB.CPrimBatchArgs)]
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "gcs_b200.h"\nint main(void){\n' +
                   "".join(f'printf("%zu\\n", sizeof({c}));\n' for c, _ in pairs) + "return 0;}\n")
    exe = tmp_path / "sizes"
    subprocess.run([cc, "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    for (c, t), n in zip(pairs, sizes):
        assert ctypes.sizeof(t) == n, f"{c}: header {n} bytes, ctypes mirror {ctypes.sizeof(t)}"
    _ = fusion


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gc_slam_b200 import operators
    import numpy as np
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        operators.point_budget_resample(np.zeros((4, 3)), np.zeros(4), np.ones(4))
    from gc_slam_b200 import _lib
    h = ctypes.c_void_p()
    assert _lib.load().gcs_create(ctypes.byref(h), 0) != 0
    assert b"no CPU fallback" in _lib.load().gcs_last_error(None) or b"CUDA" in _lib.load().gcs_last_error(None)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gc-slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle/"


def test_cert_schema_and_aggregation():
    from gc_slam_b200.certs import (CertBundle, ComputeCert, DeviceRuntimeCert, ExpectedEffect, InfluenceCert, MapUpdateCert,
                                    OTCert, ScanIOCert, SupportCert, aggregate_certificates)
    c = CertBundle.create_approx("GC-RIGHT-01", "a", ["X"], support=SupportCert(3.0, 0.5),
                                 influence=InfluenceCert.identity().with_overrides(psd_projection_delta=0.25))
    d = c.to_dict()
    for k in ("chart_id", "anchor_id", "exact", "approximation_triggers", "frobenius_applied", "conditioning", "support",
              "mismatch", "excitation", "influence", "overconfidence", "compute", "total_trigger_magnitude"):
        assert k in d
    assert d["exact"] is False and d["total_trigger_magnitude"] == 0.25 and "ot" not in d
    cc = ComputeCert()
    assert isinstance(cc.alloc_bytes_est, int) and isinstance(cc.largest_tensor_shape, tuple) and len(cc.largest_tensor_shape) == 2
    assert isinstance(cc.scan_io, ScanIOCert) and isinstance(cc.device_runtime, DeviceRuntimeCert)
    assert set(cc.to_dict()["device_runtime"]) == {"host_sync_count_est", "device_to_host_bytes_est", "host_to_device_bytes_est",
                                                   "jit_recompile_count"}
    e = CertBundle.create_exact("GC-RIGHT-01", "b", ot=OTCert(transport_mass_total=2.0), map_update=MapUpdateCert(fused_count=3))
    agg = aggregate_certificates([c, e])
    assert agg.exact is False and agg.approximation_triggers == ["X"] and agg.ot.transport_mass_total == 2.0
    assert agg.map_update.fused_count == 3 and agg.support.ess_total == 1.5
    assert ExpectedEffect("x", 1.0).to_dict() == {"objective_name": "x", "predicted": 1.0, "realized": None}
    assert aggregate_certificates([]).chart_id == "unknown"


def test_constants_and_manifest():
    from gc_slam_b200 import constants as K, manifest
    assert (K.GC_D_Z, K.GC_K_HYP, K.GC_N_POINTS_CAP, K.GC_K_ASSOC, K.GC_K_SINKHORN) == (22, 4, 8192, 8, 50)
    assert K.GC_N_ACTIVE_TILES == 7 == K.GC_N_STENCIL_TILES and K.GC_M_TILE == 50000 and K.GC_M_TILE_VIEW == 1024
    b = manifest.patch_backends({"core_array": "jax", "imu": "jax"})
    for k in ("core_array", "se3", "domain_projection_psd", "deskew", "lidar_evidence", "map_update", "sinkhorn_backend",
              "point_budget", "bin_soft_assign", "scan_bin_moment_match", "matrix_fisher", "surfel_extraction", "association"):
        assert b[k].startswith("gcs_sm100a:")
    assert b["imu"] == "jax"


def test_tiling_helpers_match_oracle():
    import numpy as np
    from gc_slam_b200 import primitives as P
    from oracle import prim_path as op
    rng = np.random.default_rng(0)
    for _ in range(50):
        c = rng.uniform(-30, 30, 3)
        assert P.ma_hex_stencil_tile_ids(c, 2.0, 1, 0) == op.stencil_tile_ids(c, 2.0, 1, 0)
    assert len(P.hex_disk_axial(1)) == 7 and len(P.hex_disk_axial(2)) == 19


def test_cpulist_parser_and_numa_binding_is_safe_without_gpu():
    from gc_slam_b200.sharding import bind_to_gpu_numa_node, parse_cpulist
    assert parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11] and parse_cpulist("") == [] and parse_cpulist("5") == [5]
    import os
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa_node(0) is None or isinstance(bind_to_gpu_numa_node(0), list)
    os.sched_setaffinity(0, before)


def test_stencil_cells_batch_equals_scalar_and_oracle():
    """Vectorised cell computation of many hypothesis poses == the scalar function (also on and next to tile boundaries)
    == the oracle's stencil (fl/common/tiling.py:171-186)."""
    from gc_slam_b200 import primitives as PR
    from oracle import prim_path as op
    rng = np.random.default_rng(3)
    P = rng.uniform(-60.0, 60.0, (4000, 3))
    P[:40, 0] = 2.0 * np.arange(40) - 40.0          # on x boundaries
    P[40:80, 2] = 2.0 * np.arange(40) - 40.0        # on z boundaries
    P[80:120, :2] = 0.0
    P[120:160, 1] = (2.0 * np.arange(40) - 40.0 - 0.5 * P[120:160, 0]) / (0.5 * np.sqrt(3.0))   # on the second hex axis
    P[160:200] = np.nextafter(P[:40], np.inf)       # one ulp beside a boundary
    cells = PR.ma_hex_cells_3d_from_xyz_batch(P, 2.0)
    assert cells == [PR.ma_hex_cell_3d_from_xyz(p, 2.0) for p in P]
    for p, c in zip(P[:300], cells[:300]):
        assert list(PR.stencil_of_cell(c, 1, 0)) == PR.ma_hex_stencil_tile_ids(p, 2.0, 1, 0) == op.stencil_tile_ids(p)


def test_batch_args_pointer_slots_cover_every_arena_pointer():
    """The vectorised pointer store of the hypothesis batch writes exactly what field-by-field assignment would."""
    import torch
    from gc_slam_b200 import hypothesis_batch as HB
    A = HB._Arena("cpu")
    names = (["dk_cert", "n_valid", "ot_cert", "rec", "view_n_valid", "inflate_stats", "dk_pts", "dk_w"] +
             ["b_" + f for f, _, _ in HB._BATCH_FIELDS] + ["v_" + f for f, _, _ in HB._VIEW_FIELDS] +
             ["a_" + f for f, _, _ in HB._ASSOC_FIELDS] + ["L22", "h22"])
    for i, nm in enumerate(names):
        A.add(nm, (3 + i,), torch.float64)
    a = HB.CPrimBatchArgs()
    words, offs = HB._arena_pointer_slots(("test-layout",), A.spec)
    base = 0x7F0000001000
    np.frombuffer(a, dtype=np.uint64)[words] = offs + np.uint64(base)
    want = lambda name: base + A.spec[name][0]
    assert (a.dk_pts, a.dk_w, a.dk_cert, a.n_lidar_valid) == (want("dk_pts"), want("dk_w"), want("dk_cert"), want("n_valid"))
    assert (a.view_n_valid, a.inflate_stats, a.ot_cert) == (want("view_n_valid"), want("inflate_stats"), want("ot_cert"))
    assert (a.L22, a.h22, a.rec) == (want("L22"), want("h22"), want("rec"))
    for f, _, _ in HB._BATCH_FIELDS:
        assert getattr(a.batch, HB._C_BATCH_NAME.get(f, f)) == want("b_" + f), f
    for f, _, _ in HB._VIEW_FIELDS:
        assert getattr(a.view, HB._C_VIEW_NAME.get(f, f)) == want("v_" + f), f
    for f, _, _ in HB._ASSOC_FIELDS:
        assert getattr(a.assoc, f) == want("a_" + f), f
    assert len(words) == len(names) and a.base.Lambdas is None and a.atlas.__bool__() is False
