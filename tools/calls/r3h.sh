mkdir -p gpurun_out/r3h
timeout 600 python -m pytest tests/test_gpu_p16.py -m gpu -q -x -s -k "backward_p16" > gpurun_out/r3h/p16_tests.log 2>&1
grep -E "^wgrad|passed|failed|Error|error" gpurun_out/r3h/p16_tests.log | cut -c1-220 | tail -30
