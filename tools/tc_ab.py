"""Developer tool: A/B timing of bin_scan_tc_kernel across library variants on ONE box.
Usage (GPU box):  python tools/tc_ab.py variantA variantB ...   ('main' = the in-tree library; others from lib/variants)
Each variant runs in a fresh process (the library is chosen at import), rounds interleaved; prints median kernel ms."""
import os
import subprocess
import sys

CHILD = r'''
import sys, numpy as np, torch
sys.path.insert(0, ".")
from gc_slam_b200 import _lib as L, operators as ops, synth
S, P = 128, 65536
bins = synth.fibonacci_atlas(48)
plan = ops.BinPathPlan(S, P, P, n_hyp=1, n_bins=48, tau=0.1, origin=synth.lidar_origin_base(), precision=L.PREC_TC,
                       want_evidence=True, materialize_deskewed=True)
plan.set_bins(bins, 0.1); plan.set_map(synth.random_map_bin_stats(48, 7, bins))
sc = [synth.vlp16_scan(P, 1000 + k, t0=synth.EPOCH_T0) for k in range(4)]
sel = [sc[k % 4] for k in range(S)]
t0 = np.full(S, synth.EPOCH_T0)
plan.upload(np.stack([s[0] for s in sel]), np.stack([s[1] for s in sel]), np.stack([s[2] for s in sel]),
            np.stack([s[3] for s in sel]), np.stack([s[4] for s in sel]), t0, t0 + 0.1,
            np.stack([synth.scan_twist(5 + k) for k in range(S)]), synth.hypothesis_poses(S, 3), non_blocking=False)
ctx = plan.io.ctx
for _ in range(5): plan.run()
torch.cuda.synchronize()
ts = []
for _ in range(40):
    ctx.timing_enable(True, only="bin_scan"); plan.run(); torch.cuda.synchronize()
    ms, n = ctx.timing_collect(); ts.append(ms / max(n, 1))
ctx.timing_enable(False)
ts.sort()
print(f"{ts[len(ts)//2]:.4f} {ts[0]:.4f}")
'''

def main():
    names = sys.argv[1:] or ["main"]
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {n: [] for n in names}
    for rnd in range(2):
        for n in names:
            env = dict(os.environ)
            n, *sets = n.split(":")          # name[:ENV=VALUE...]
            for kv in sets:
                k, v = kv.split("=")
                env[k] = v
            if n != "main":
                env["GCS_B200_LIB"] = os.path.join(here, "gc-slam_b200", "lib", "variants", f"libgcs_b200.{n}.so")
            else:
                env.pop("GCS_B200_LIB", None)
            r = subprocess.run([sys.executable, "-c", CHILD], cwd=here, env=env, capture_output=True, text=True, timeout=300)
            n = ":".join([n, *sets])
            res[n].append(r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else "ERR " + r.stderr[-300:])
    for n in names:
        print(f"{n:12s} median/min ms per launch, two rounds: {res[n]}")

if __name__ == "__main__":
    main()
