mkdir -p gpurun_out/r3b
N=$(nvidia-smi -L | wc -l)
for p16 in 1 0; do
B3D_SLAB_P16=$p16 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/slab_bench.py 7 graph peer > gpurun_out/r3b/slab_p16_${p16}_n$N.json 2> gpurun_out/r3b/slab_p16_${p16}_n$N.err
echo "N=$N p16=$p16: $(tail -1 gpurun_out/r3b/slab_p16_${p16}_n$N.json | cut -c1-400)"; grep -v "OMP_NUM\|\*\*\*" gpurun_out/r3b/slab_p16_${p16}_n$N.err | tail -3 | cut -c1-300
done
