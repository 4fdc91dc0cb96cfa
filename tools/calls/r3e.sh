mkdir -p gpurun_out/r3e
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_p16.py -m gpu -q -x > gpurun_out/r3e/gpu_tests.log 2>&1
tail -3 gpurun_out/r3e/gpu_tests.log | cut -c1-300
for sd in 0 1 0 1; do
  B3D_SHARE_DGRAD=$sd timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/r3e/bench_$sd.json 2> gpurun_out/r3e/bench_$sd.err
  echo "share_dgrad=$sd: $(grep -o '"ms_per_step[^,]*' gpurun_out/r3e/bench_$sd.json | head -1)"; tail -1 gpurun_out/r3e/bench_$sd.err | cut -c1-200
done
