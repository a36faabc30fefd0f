timeout 200 python -m pytest tests/test_gpu_multidevice.py tests/test_sharding.py -x -q 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-prim > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 400 gpurun_out/bench_2gpu.json
