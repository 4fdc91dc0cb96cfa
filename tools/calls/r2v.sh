mkdir -p gpurun_out/r2v
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/r2v/gpu_tests.log 2>&1
tail -3 gpurun_out/r2v/gpu_tests.log | cut -c1-400
timeout 300 python tools/hbm_bench.py 10 gpurun_out/r2v/hbm.json > gpurun_out/r2v/hbm.txt 2>&1; grep -v "32^3\|64^3" gpurun_out/r2v/hbm.txt | cut -c1-150
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r2v/bench.json 2> gpurun_out/r2v/bench.err
grep -o '"ms_per_step[^,]*' gpurun_out/r2v/bench.json | head -3; tail -2 gpurun_out/r2v/bench.err
