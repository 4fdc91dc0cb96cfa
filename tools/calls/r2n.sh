mkdir -p gpurun_out/r2o
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2o/gpu_tests.log 2>&1
tail -8 gpurun_out/r2o/gpu_tests.log | cut -c1-400
for cfg in "1 1" "0 1" "1 0"; do set -- $cfg
  B3D_SHARE_DGRAD=$1 B3D_DEDUP=$2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r2o/bench_$1$2.json 2> gpurun_out/r2o/bench_$1$2.err
  echo "share_dgrad=$1 dedup=$2: $(cut -c1-200 gpurun_out/r2o/bench_$1$2.json | grep -o 'ms_per_step[^,]*')"; tail -2 gpurun_out/r2o/bench_$1$2.err
done
