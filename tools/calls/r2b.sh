set -x
mkdir -p gpurun_out/r2b
timeout 600 python -m pytest tests/test_gpu_p16.py -x -q -s > gpurun_out/r2b/p16_tests.log 2>&1
tail -15 gpurun_out/r2b/p16_tests.log
timeout 300 python tools/kdfold_check.py > gpurun_out/r2b/kdfold.txt 2>&1
tail -6 gpurun_out/r2b/kdfold.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b/gpu_tests.log 2>&1
tail -15 gpurun_out/r2b/gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b/bench.json 2> gpurun_out/r2b/bench.err
cat gpurun_out/r2b/bench.json | cut -c1-600
