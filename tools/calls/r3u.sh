mkdir -p gpurun_out/r3u
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r3u/bench_n$N.json 2> gpurun_out/r3u/bench_n$N.err
echo "rc=$?"; tail -1 gpurun_out/r3u/bench_n$N.json | cut -c1-260; grep -v "OMP_NUM\|\*\*\*" gpurun_out/r3u/bench_n$N.err | tail -3 | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/r3u/bench_n$N.json').read().strip().splitlines()[-1])
print('e2e', d.get('e2e')); print('inference', json.dumps(d.get('inference'))[:500]); print('tta', d.get('inference_tta'))
PY
