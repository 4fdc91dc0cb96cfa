mkdir -p gpurun_out/r3j
timeout 600 python -m pytest tests/test_gpu_p16.py -m gpu -q -x -k "backward_p16" 2>&1 | tail -2
B3D_WGRAD_TSF=1 timeout 300 python tools/conv_bench.py fwd16 5 fp16 > gpurun_out/r3j/fwd16_tsf1.txt 2>&1
grep -o "^ *[0-9]*^3 *[0-9]*-> *[0-9]*\|wgrad(P16[^|]*" gpurun_out/r3j/fwd16_tsf1.txt | paste - - | cut -c1-120
