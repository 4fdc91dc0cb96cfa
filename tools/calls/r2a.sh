set -x
mkdir -p gpurun_out/r2a
# 1. ncu full capture of the kd-folded kernels (dominant kernel) + TS wgrad
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -o gpurun_out/r2a/prof_fold16_32to16 python tools/conv_bench.py fwd 2 fp16 128,32,16 > gpurun_out/r2a/ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -o gpurun_out/r2a/prof_fold32_32to32 python tools/conv_bench.py fwd 2 fp16 64,32,32 > gpurun_out/r2a/ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3_wgrad_ts_kernel -s 1 -c 1 -o gpurun_out/r2a/prof_wgrad_ts_16to16 python tools/conv_bench.py wgrad 2 fp16 128,16,16 > gpurun_out/r2a/ncu3.log 2>&1
# 2. launch list of the current build (one eager step)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a/launches_train.csv python tools/one_step.py 2 > gpurun_out/r2a/one_step.log 2>&1
# 3. per-layer bench
timeout 300 python tools/conv_bench.py all 5 fp16 > gpurun_out/r2a/conv_bench.txt 2>&1
nvidia-smi > gpurun_out/r2a/smi.txt
ls -la gpurun_out/r2a
# 4. the new BASELINE-shape parity tests on the round-1 kernels (calibration of the derived bounds)
timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py -x -q -s > gpurun_out/r2a/baseline_tests.log 2>&1
tail -5 gpurun_out/r2a/baseline_tests.log
