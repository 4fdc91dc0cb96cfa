mkdir -p gpurun_out/r2r
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2r/gpu_tests.log 2>&1
tail -5 gpurun_out/r2r/gpu_tests.log | cut -c1-400
timeout 300 python tools/conv_bench.py k1p16 5 fp16 > gpurun_out/r2r/k1.txt 2>&1; cut -c1-200 gpurun_out/r2r/k1.txt
timeout 300 python tools/conv_bench.py fwd16 5 fp16 > gpurun_out/r2r/fwd16.txt 2>&1; cut -c1-260 gpurun_out/r2r/fwd16.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r2r/bench.json 2> gpurun_out/r2r/bench.err
grep -o '"ms_per_step[^,]*' gpurun_out/r2r/bench.json | head -3; grep -o '"inference": {[^}]*}' gpurun_out/r2r/bench.json | cut -c1-200; tail -2 gpurun_out/r2r/bench.err
