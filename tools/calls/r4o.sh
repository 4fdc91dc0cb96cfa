mkdir -p gpurun_out/r4o
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r4o/gpu_tests.log 2>&1
tail -3 gpurun_out/r4o/gpu_tests.log | cut -c1-300
