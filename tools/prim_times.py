"""Run the primitive-family path (config 3) a few times; meant to sit under `ncu --metrics gpu__time_duration.sum`."""
import json
import sys
sys.path.insert(0, ".")
import bench  # noqa: E402

r = bench.primitive_path_extra(65536, n_map=int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, reps=int(sys.argv[2]) if len(sys.argv) > 2 else 4)
print(json.dumps(r["p50_ms_per_stage"]))
