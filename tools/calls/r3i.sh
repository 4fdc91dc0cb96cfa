mkdir -p gpurun_out/r3i
for t in 1 0; do
B3D_WGRAD_TSF=$t timeout 300 python tools/conv_bench.py fwd16 5 fp16 > gpurun_out/r3i/fwd16_tsf$t.txt 2>&1
echo "TSF=$t"; grep -o "^ *[0-9]*^3 *[0-9]*-> *[0-9]*\|wgrad(P16[^|]*" gpurun_out/r3i/fwd16_tsf$t.txt | paste - - | cut -c1-120
B3D_WGRAD_TSF=$t timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/r3i/bench_$t.json 2> gpurun_out/r3i/bench_$t.err
echo "tsf=$t: $(grep -o '"ms_per_step[^,]*' gpurun_out/r3i/bench_$t.json | head -1)"; tail -1 gpurun_out/r3i/bench_$t.err | cut -c1-200
done
