timeout 600 python -m pytest tests/test_gpu_prims.py -m gpu -x -q 2>&1 | tail -12
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/prim_l5.csv python tools/prim_times.py 1000000 1 > gpurun_out/prim_l5.log 2>&1
timeout 300 python tools/prim_times.py 1000000 10
