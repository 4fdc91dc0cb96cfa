mkdir -p gpurun_out/r5g
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r5g/gpu_tests.log 2>&1
tail -3 gpurun_out/r5g/gpu_tests.log | cut -c1-300
for k in 1 0; do echo "BIGBOX=$k"; B3D_KDF_BIGBOX=$k timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'])"; done
