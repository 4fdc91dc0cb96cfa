mkdir -p gpurun_out/r2l
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_dp.py tests/test_gpu_slab.py::test_peer_memory_comm_two_gpus -q -s -x > gpurun_out/r2l/dp_tests.log 2>&1
tail -12 gpurun_out/r2l/dp_tests.log | cut -c1-1200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2l/bench2.json 2> gpurun_out/r2l/bench2.err
echo "bench rc=$?"
cut -c1-1500 gpurun_out/r2l/bench2.json; tail -5 gpurun_out/r2l/bench2.err
timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py -q -x -k "128cube" > gpurun_out/r2l/base.log 2>&1; tail -3 gpurun_out/r2l/base.log
