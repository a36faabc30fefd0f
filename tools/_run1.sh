timeout 120 python tools/stage_times.py 128 65536 tc 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_bins.py -m gpu -x -q 2>&1 | tail -2
