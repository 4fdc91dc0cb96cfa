mkdir -p gpurun_out/r3q
for i in 1 2; do
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3q/gpu_tests_$i.log 2>&1
tail -1 gpurun_out/r3q/gpu_tests_$i.log | cut -c1-200; grep "^FAILED" gpurun_out/r3q/gpu_tests_$i.log | cut -c1-200
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/r3q/bench.json 2> gpurun_out/r3q/bench.err
echo "$(grep -o '"ms_per_step[^,]*' gpurun_out/r3q/bench.json | head -1) $(grep -o '"inference": {[^}]*}' gpurun_out/r3q/bench.json | grep -o 'ms_per_forward": [0-9.]*' | head -1)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3q/launches_train.csv python tools/one_step.py 2 > gpurun_out/r3q/one_step.log 2>&1
