mkdir -p gpurun_out/r3p
for i in 1 2 3; do
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3p/gpu_tests_$i.log 2>&1
tail -1 gpurun_out/r3p/gpu_tests_$i.log | cut -c1-200; grep "^FAILED" gpurun_out/r3p/gpu_tests_$i.log | cut -c1-200
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
