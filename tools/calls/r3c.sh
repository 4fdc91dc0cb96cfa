mkdir -p gpurun_out/r3c
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3c/gpu_tests.log 2>&1
tail -4 gpurun_out/r3c/gpu_tests.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3c/launches_infer_w8.csv python tools/one_infer.py 2 8 > gpurun_out/r3c/one_infer_w8.log 2>&1
tail -1 gpurun_out/r3c/one_infer_w8.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3c/launches_infer_w1.csv python tools/one_infer.py 2 1 > gpurun_out/r3c/one_infer_w1.log 2>&1
tail -1 gpurun_out/r3c/one_infer_w1.log
