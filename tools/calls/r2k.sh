mkdir -p gpurun_out/r2k
timeout 600 python -m pytest tests/test_gpu_p16.py -q > gpurun_out/r2k/p16_tests.log 2>&1
tail -5 gpurun_out/r2k/p16_tests.log
timeout 600 python tools/hbm_bench.py 10 gpurun_out/r2k/hbm.json > gpurun_out/r2k/hbm.txt 2>&1
grep -E "128\^3x16|64\^3x32" gpurun_out/r2k/hbm.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r2k/bench.json 2> gpurun_out/r2k/bench.err
cat gpurun_out/r2k/bench.json | cut -c1-300; tail -3 gpurun_out/r2k/bench.err
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2k/gpu_tests.log 2>&1
tail -5 gpurun_out/r2k/gpu_tests.log
