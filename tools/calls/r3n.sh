mkdir -p gpurun_out/r3n
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q -x -s > gpurun_out/r3n/gpu_tests.log 2>&1
tail -3 gpurun_out/r3n/gpu_tests.log | cut -c1-300; grep -h "grad_rel" gpurun_out/r3n/gpu_tests.log | cut -c1-400 | head -3
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r3n/bench_n$N.json 2> gpurun_out/r3n/bench_n$N.err
echo "rc=$?"; tail -1 gpurun_out/r3n/bench_n$N.json | cut -c1-300; grep -v "OMP_NUM\|\*\*\*" gpurun_out/r3n/bench_n$N.err | tail -3 | cut -c1-300
