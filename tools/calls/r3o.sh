mkdir -p gpurun_out/r3o
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r3o/gpu_tests.log 2>&1
tail -3 gpurun_out/r3o/gpu_tests.log | cut -c1-300
for f in inkernel kernel; do
  B3D_SPLITK_FINISH=$f timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/r3o/bench_$f.json 2> gpurun_out/r3o/bench_$f.err
  echo "finish=$f: $(grep -o '"ms_per_step[^,]*' gpurun_out/r3o/bench_$f.json | head -1) $(grep -o '"inference": {[^}]*}' gpurun_out/r3o/bench_$f.json | grep -o 'ms_per_forward": [0-9.]*' | head -1)"; tail -1 gpurun_out/r3o/bench_$f.err | cut -c1-200
done
timeout 300 python tools/conv_bench.py fwd16 5 fp16 > gpurun_out/r3o/fwd16.txt 2>&1; grep "32^3\|16^3" gpurun_out/r3o/fwd16.txt | cut -c1-150
