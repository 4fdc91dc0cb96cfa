mkdir -p gpurun_out/r3t
for i in 1 2 3 4; do
timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_slab.py -m gpu -q -s > gpurun_out/r3t/slab_$i.log 2>&1
tail -1 gpurun_out/r3t/slab_$i.log | cut -c1-200; grep "^FAILED" gpurun_out/r3t/slab_$i.log | cut -c1-200; grep -h "8 slabs" gpurun_out/r3t/slab_$i.log | cut -c1-140
done
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3t/gpu_tests.log 2>&1
tail -1 gpurun_out/r3t/gpu_tests.log | cut -c1-200; grep "^FAILED" gpurun_out/r3t/gpu_tests.log | cut -c1-200
