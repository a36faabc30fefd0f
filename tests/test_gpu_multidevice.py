"""One process driving two devices (one gcs_ctx per device, include/gcs_b200.h threading note): kernels that need more
than 48 KB of dynamic shared memory must have their attribute set on EACH device's context."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def test_two_devices_in_one_process():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs in one process")
    from gc_slam_b200 import _lib as L, operators as ops, synth
    bins = synth.fibonacci_atlas(48)
    pts, t, w, ring, tag = synth.vlp16_scan(8192, 3, t0=synth.EPOCH_T0)
    outs = []
    for dev in (0, 1):
        with torch.cuda.device(dev):
            for prec in (L.PREC_F64, L.PREC_TC):
                plan = ops.BinPathPlan(1, 8192, 8192, n_hyp=1, n_bins=48, tau=0.1, origin=synth.lidar_origin_base(),
                                       precision=prec, want_evidence=True, device=dev)
                plan.set_bins(bins, 0.1)
                plan.set_map(synth.random_map_bin_stats(48, 7, bins))
                plan.upload(pts[None], t[None], w[None], ring[None], tag[None], np.array([synth.EPOCH_T0]),
                            np.array([synth.EPOCH_T0 + 0.1]), synth.scan_twist(3)[None], synth.hypothesis_poses(1, 3),
                            non_blocking=False)
                plan.run()
                torch.cuda.synchronize(dev)
                outs.append(plan.outputs().L22.cpu())
    assert torch.equal(outs[0], outs[2]) and torch.equal(outs[1], outs[3])      # same results on both devices
