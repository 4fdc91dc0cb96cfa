mkdir -p gpurun_out/r4z
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r4z/gpu_tests.log 2>&1
tail -3 gpurun_out/r4z/gpu_tests.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python tools/one_infer.py 2 1 2>&1 | tail -1
