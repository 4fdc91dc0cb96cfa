mkdir -p gpurun_out/r3z
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_p16.py -m gpu -q -x > gpurun_out/r3z/gpu_tests.log 2>&1
tail -2 gpurun_out/r3z/gpu_tests.log | cut -c1-300
timeout 300 python tools/hbm_bench.py 10 gpurun_out/r3z/hbm.json > gpurun_out/r3z/hbm.txt 2>&1; grep "gn_bwd_apply_p16" gpurun_out/r3z/hbm.txt | cut -c1-150
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/r3z/bench.json 2> gpurun_out/r3z/bench.err
echo "$(grep -o '"ms_per_step[^,]*' gpurun_out/r3z/bench.json | head -1)"
