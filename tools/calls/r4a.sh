mkdir -p gpurun_out/r4a
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r4a/gpu_tests.log 2>&1
tail -1 gpurun_out/r4a/gpu_tests.log | cut -c1-200; grep "^FAILED" gpurun_out/r4a/gpu_tests.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r4a/bench_ref.json 2> gpurun_out/r4a/bench_ref.err
cut -c1-300 gpurun_out/r4a/bench_ref.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/r4a/bench.json 2> gpurun_out/r4a/bench.err
echo "$(grep -o '"ms_per_step[^,]*' gpurun_out/r4a/bench.json | head -2 | tr '\n' ' ') $(grep -o '"clocks": {[^}]*}' gpurun_out/r4a/bench.json)"
