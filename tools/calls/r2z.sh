mkdir -p gpurun_out/r2z
timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_gpu_baseline_shapes.py -m gpu -q -x > gpurun_out/r2z/gpu_tests.log 2>&1
tail -4 gpurun_out/r2z/gpu_tests.log | cut -c1-300
N=$(nvidia-smi -L | wc -l)
B3D_SLAB_P16=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/slab_bench.py 5 graph peer > gpurun_out/r2z/slab_n$N.json 2> gpurun_out/r2z/slab_n$N.err
echo "N=$N: $(tail -1 gpurun_out/r2z/slab_n$N.json | cut -c1-400)"; grep -v "OMP_NUM\|\*\*\*" gpurun_out/r2z/slab_n$N.err | tail -5 | cut -c1-300
