nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_sharding.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/t_shard4.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 3 --no-prim > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err
cat gpurun_out/t_shard4.log; cut -c1-400 gpurun_out/bench_4gpu.json; tail -3 gpurun_out/bench_4gpu.err
