mkdir -p gpurun_out/r4x
timeout 600 python -m pytest tests/test_gpu_p16.py -m gpu -q -x -k "block_weight or backward_p16" > gpurun_out/r4x/p16.log 2>&1
grep -E "passed|failed|^E " gpurun_out/r4x/p16.log | cut -c1-200 | tail -12
timeout 300 python tools/wgrad_block_bench.py 10 2>&1 | tail -8
for k in 1 0 1 0; do echo "FUSE_W=$k"; B3D_FUSE_BLOCK_WGRAD=$k timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'])"; done
