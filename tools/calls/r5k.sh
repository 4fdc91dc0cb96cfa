mkdir -p gpurun_out/r5k
timeout 80 python -m pytest tests/test_gpu_p16.py tests/test_gpu_kernels.py -m gpu -q -x -k "with_and_without_twins or resnet_block or wgrad_tensor_core" > gpurun_out/r5k/t.log 2>&1
grep -E "passed|failed|^E " gpurun_out/r5k/t.log | cut -c1-200 | tail -6
