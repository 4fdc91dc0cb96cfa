mkdir -p gpurun_out/r3d
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r3d/gpu_tests.log 2>&1
tail -3 gpurun_out/r3d/gpu_tests.log | cut -c1-300
timeout 300 python tools/hbm_bench.py 10 gpurun_out/r3d/hbm.json > gpurun_out/r3d/hbm.txt 2>&1; grep "block_epi" gpurun_out/r3d/hbm.txt | cut -c1-150
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r3d/bench.json 2> gpurun_out/r3d/bench.err
grep -o '"ms_per_step[^,]*' gpurun_out/r3d/bench.json | head -3; grep -o '"inference": {[^}]*}' gpurun_out/r3d/bench.json | cut -c1-200; tail -2 gpurun_out/r3d/bench.err
