mkdir -p gpurun_out/r2p
timeout 1500 python bench.py > gpurun_out/r2p/bench_full.json 2> gpurun_out/r2p/bench_full.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/r2p/bench_full.json; tail -3 gpurun_out/r2p/bench_full.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2p/bench_full.json'))
    print('extra:', json.dumps(d.get('extra_configs'))[:1500])
    print('cpu:', d.get('cpu_baseline'))
except Exception as e: print('ERR', e)
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2p/bench_ref.json 2> gpurun_out/r2p/bench_ref.err
cut -c1-400 gpurun_out/r2p/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2p/launches_train.csv python tools/one_step.py 2 > gpurun_out/r2p/one_step.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2p/launches_infer.csv python tools/one_infer.py 2 > gpurun_out/r2p/one_infer.log 2>&1
tail -2 gpurun_out/r2p/one_infer.log
