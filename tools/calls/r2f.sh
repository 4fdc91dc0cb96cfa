mkdir -p gpurun_out/r2f
timeout 600 python tools/debug/slab_debug.py > gpurun_out/r2f/slab.log 2>&1
tail -12 gpurun_out/r2f/slab.log
timeout 600 python tools/debug/p16_debug.py > gpurun_out/r2f/debug.log 2>&1
grep -A 6 "whole model" gpurun_out/r2f/debug.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2f/gpu_tests.log 2>&1
tail -15 gpurun_out/r2f/gpu_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r2f/bench.json 2> gpurun_out/r2f/bench.err
cat gpurun_out/r2f/bench.json | cut -c1-300; tail -3 gpurun_out/r2f/bench.err
