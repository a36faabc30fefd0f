timeout 600 python -m pytest tests/test_gpu_imu.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/t8.log
cat gpurun_out/t8.log
