mkdir -p gpurun_out/r4d
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r4d/gpu_tests.log 2>&1
tail -1 gpurun_out/r4d/gpu_tests.log | cut -c1-200; grep "^FAILED" gpurun_out/r4d/gpu_tests.log | cut -c1-200
timeout 300 python tools/conv_bench.py fwd16 5 fp16 2>&1 | grep -o "^ *[0-9]*^3 *[0-9]*-> *[0-9]*\|wgrad(P16[^|]*" | paste - - | sed -n '2,4p' | cut -c1-100
timeout 300 python tools/conv_bench.py k1p16 5 fp16 2>&1 | grep "32^3\|16^3" | cut -c1-110
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/r4d/bench.json 2> gpurun_out/r4d/bench.err
echo "$(grep -o '"ms_per_step[^,]*' gpurun_out/r4d/bench.json | head -1) $(grep -o '"inference": {[^}]*}' gpurun_out/r4d/bench.json | grep -o 'ms_per_forward": [0-9.]*' | head -1)"
