mkdir -p gpurun_out/r4y
B3D_PDL=1 timeout 900 python -m pytest tests/test_gpu_p16.py tests/test_gpu_model.py -m gpu -q -x > gpurun_out/r4y/p16.log 2>&1
grep -E "passed|failed|^E " gpurun_out/r4y/p16.log | cut -c1-200 | tail -6
for k in 1 0 1 0; do echo "PDL=$k"; B3D_PDL=$k timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'])"; done
