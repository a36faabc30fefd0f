timeout 60 python tools/tc_check.py 128 65536 1 2>&1 | grep -E "scan kernel|N  |rerun|deskewed"
timeout 600 python -m pytest tests/test_gpu_bins.py -m gpu -x -q 2>&1 | tail -3
