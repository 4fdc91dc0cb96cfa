mkdir -p gpurun_out/r2x
timeout 600 python -m pytest tests/test_gpu_slab.py -m gpu -q -x -k "peer_memory" > gpurun_out/r2x/gpu_tests.log 2>&1
tail -5 gpurun_out/r2x/gpu_tests.log | cut -c1-300
N=$(nvidia-smi -L | wc -l)
for p16 in 1 0; do
B3D_SLAB_P16=$p16 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/slab_bench.py 5 graph peer > gpurun_out/r2x/slab_p16_${p16}_n$N.json 2> gpurun_out/r2x/slab_p16_${p16}_n$N.err
echo "N=$N p16=$p16: $(tail -1 gpurun_out/r2x/slab_p16_${p16}_n$N.json | cut -c1-400)"; tail -3 gpurun_out/r2x/slab_p16_${p16}_n$N.err | cut -c1-300
done
