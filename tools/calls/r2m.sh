mkdir -p gpurun_out/r2m
timeout 600 python tools/hbm_bench.py 10 gpurun_out/r2m/hbm.json > gpurun_out/r2m/hbm.txt 2>&1
grep -E "reduce" gpurun_out/r2m/hbm.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2m/gpu_tests.log 2>&1
tail -6 gpurun_out/r2m/gpu_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r2m/bench.json 2> gpurun_out/r2m/bench.err
cut -c1-260 gpurun_out/r2m/bench.json; tail -3 gpurun_out/r2m/bench.err
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_target.py > gpurun_out/r2m/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/r2m/sanitizer_$tool.log
done
