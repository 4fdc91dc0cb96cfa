mkdir -p gpurun_out/r2e
timeout 600 python tools/debug/p16_debug.py > gpurun_out/r2e/debug.log 2>&1
grep -A 12 "whole model" gpurun_out/r2e/debug.log | cut -c1-400
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e/gpu_tests.log 2>&1
tail -12 gpurun_out/r2e/gpu_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r2e/bench.json 2> gpurun_out/r2e/bench.err
cat gpurun_out/r2e/bench.json | cut -c1-300; tail -3 gpurun_out/r2e/bench.err
