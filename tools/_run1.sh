timeout 600 python -m pytest tests/test_gpu_prims.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/prim_times.py 1000000 10
