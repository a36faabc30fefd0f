timeout 60 python tools/tc_check.py 128 65536 1 2>&1 | grep -vE "WARN|^$" > gpurun_out/ab4.log
GCS_TC_OPERANDS=tf32 timeout 60 python tools/tc_check.py 128 65536 1 2>&1 | grep -E "scan kernel|rerun|N  " >> gpurun_out/ab4.log
timeout 60 python tools/tc_check.py 1 65536 1 2>&1 | grep -E "scan kernel|rerun|N  " >> gpurun_out/ab4.log
timeout 60 python tools/tc_check.py 3 30000 4 8192 2>&1 | grep -vE "WARN|^$" >> gpurun_out/ab4.log
timeout 300 python -m pytest tests/test_gpu_bins.py -m gpu -x -q 2>&1 | tail -15 >> gpurun_out/ab4.log
cat gpurun_out/ab4.log
