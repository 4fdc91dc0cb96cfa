set -x
mkdir -p gpurun_out/r2c
timeout 600 python -m pytest tests/test_gpu_p16.py -x -q -s > gpurun_out/r2c/p16_tests.log 2>&1
tail -12 gpurun_out/r2c/p16_tests.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c/gpu_tests.log 2>&1
tail -12 gpurun_out/r2c/gpu_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r2c/bench.json 2> gpurun_out/r2c/bench.err
cat gpurun_out/r2c/bench.json | cut -c1-400; tail -3 gpurun_out/r2c/bench.err
