mkdir -p gpurun_out/r3k
timeout 600 python -m pytest tests/test_gpu_dp.py tests/test_gpu_slab.py -m gpu -q -x > gpurun_out/r3k/gpu_tests.log 2>&1
tail -3 gpurun_out/r3k/gpu_tests.log | cut -c1-300
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r3k/bench_n$N.json 2> gpurun_out/r3k/bench_n$N.err
echo "rc=$?"; tail -1 gpurun_out/r3k/bench_n$N.json | cut -c1-300; grep -v "OMP_NUM\|\*\*\*" gpurun_out/r3k/bench_n$N.err | tail -3 | cut -c1-300
