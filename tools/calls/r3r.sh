mkdir -p gpurun_out/r3r
timeout 600 python -m pytest tests/test_gpu_p16.py -m gpu -q -x -k "block_input or backward_p16 or with_and_without" > gpurun_out/r3r/p16.log 2>&1
tail -3 gpurun_out/r3r/p16.log | cut -c1-300; grep "^E " gpurun_out/r3r/p16.log | head -5 | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r3r/gpu_tests.log 2>&1
tail -3 gpurun_out/r3r/gpu_tests.log | cut -c1-300
for f in 1 0; do
  B3D_FUSE_BLOCK_DGRAD=$f timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/r3r/bench_$f.json 2> gpurun_out/r3r/bench_$f.err
  echo "fuse=$f: $(grep -o '"ms_per_step[^,]*' gpurun_out/r3r/bench_$f.json | head -1)"; tail -1 gpurun_out/r3r/bench_$f.err | cut -c1-200
done
