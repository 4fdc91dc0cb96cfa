mkdir -p gpurun_out/r2u
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2u/gpu_tests.log 2>&1
tail -3 gpurun_out/r2u/gpu_tests.log | cut -c1-400
timeout 300 python tools/hbm_bench.py 10 gpurun_out/r2u/hbm.json > gpurun_out/r2u/hbm.txt 2>&1; cut -c1-150 gpurun_out/r2u/hbm.txt | head -70
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r2u/bench.json 2> gpurun_out/r2u/bench.err
grep -o '"ms_per_step[^,]*' gpurun_out/r2u/bench.json | head -3; grep -o '"inference": {[^}]*}' gpurun_out/r2u/bench.json | cut -c1-200; tail -2 gpurun_out/r2u/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2u/launches_train.csv python tools/one_step.py 2 > gpurun_out/r2u/one_step.log 2>&1
