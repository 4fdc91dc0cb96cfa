mkdir -p gpurun_out/r5i
timeout 600 python -m pytest tests/test_gpu_p16.py -m gpu -q -x -k "block_weight or backward_p16" > gpurun_out/r5i/p16.log 2>&1
grep -E "passed|failed|^E " gpurun_out/r5i/p16.log | cut -c1-200 | tail -8
timeout 300 python tools/conv_bench.py fwd16 5 fp16 2>&1 | grep -o "^ *[0-9]*^3 *[0-9]*-> *[0-9]*\|wgrad(P16[^|]*" | paste - - | sed -n '2,7p' | cut -c1-100
timeout 300 python tools/wgrad_block_bench.py 10 2>&1 | tail -6
