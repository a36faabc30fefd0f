timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/t7.log
timeout 900 python bench.py > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err
cat gpurun_out/t7.log; cut -c1-1500 gpurun_out/bench_r1b.json
