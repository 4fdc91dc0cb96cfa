mkdir -p gpurun_out/r2w
timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/r2w/gpu_tests.log 2>&1
tail -25 gpurun_out/r2w/gpu_tests.log | cut -c1-300
