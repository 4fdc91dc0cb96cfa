mkdir -p gpurun_out/r5j
O=gpurun_out/r5j
timeout 900 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1
tail -2 $O/gpu_tests.log | cut -c1-300
timeout 600 python bench.py > $O/bench_full.json 2> $O/bench_full.err
echo "bench rc=$?"; cut -c1-200 $O/bench_full.json
timeout 120 python tools/conv_bench.py fwd16 5 fp16 > $O/conv_bench.txt 2>&1
