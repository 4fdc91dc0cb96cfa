mkdir -p gpurun_out/r5c
timeout 600 python -m pytest tests/test_gpu_dp.py tests/test_gpu_slab.py -m gpu -q > gpurun_out/r5c/gpu_tests_2gpu.log 2>&1
tail -3 gpurun_out/r5c/gpu_tests_2gpu.log | cut -c1-300
N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r5c/bench_n$N.json 2> gpurun_out/r5c/bench_n$N.err
tail -1 gpurun_out/r5c/bench_n$N.json | cut -c1-260
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/slab_bench.py 7 graph peer > gpurun_out/r5c/slab_n$N.json 2> gpurun_out/r5c/slab_n$N.err
tail -1 gpurun_out/r5c/slab_n$N.json | cut -c1-300
