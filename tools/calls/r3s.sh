mkdir -p gpurun_out/r3s
for i in 1 2; do
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3s/gpu_tests_$i.log 2>&1
tail -1 gpurun_out/r3s/gpu_tests_$i.log | cut -c1-200; grep "^FAILED" gpurun_out/r3s/gpu_tests_$i.log | cut -c1-200
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/r3s/bench.json 2> gpurun_out/r3s/bench.err
echo "$(grep -o '"ms_per_step[^,]*' gpurun_out/r3s/bench.json | head -1)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3s/launches_train.csv python tools/one_step.py 2 > gpurun_out/r3s/one_step.log 2>&1
