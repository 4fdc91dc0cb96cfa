mkdir -p gpurun_out/r4b
timeout 600 python -m pytest tests/test_gpu_p16.py tests/test_gpu_kernels.py -m gpu -q -x -k "wgrad or backward" 2>&1 | tail -1
echo "baseline"; timeout 300 python tools/conv_bench.py fwd16 5 fp16 2>&1 | grep -o "^ *[0-9]*^3 *[0-9]*-> *[0-9]*\|wgrad(P16[^|]*" | paste - - | head -8 | cut -c1-100
for n in 2 3 4 6; do echo "CIN32 issuers=$n"; B3D_WGRAD_TS_CIN32=1 B3D_WGRAD_TS_ISSUERS=$n timeout 300 python tools/conv_bench.py fwd16 5 fp16 2>&1 | grep -o "^ *[0-9]*^3 *[0-9]*-> *[0-9]*\|wgrad(P16[^|]*" | paste - - | sed -n '3,8p' | cut -c1-100; done
B3D_WGRAD_TS_CIN32=1 timeout 600 python -m pytest tests/test_gpu_p16.py -m gpu -q -x -k "backward" 2>&1 | tail -1
