mkdir -p gpurun_out/r2y
timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_gpu_baseline_shapes.py -m gpu -q -x > gpurun_out/r2y/gpu_tests.log 2>&1
tail -3 gpurun_out/r2y/gpu_tests.log | cut -c1-300
for w in 2 8; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2y/launches_infer_w$w.csv python tools/one_infer.py 2 $w > gpurun_out/r2y/one_infer_w$w.log 2>&1
tail -1 gpurun_out/r2y/one_infer_w$w.log
done
