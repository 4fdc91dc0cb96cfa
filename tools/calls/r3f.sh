mkdir -p gpurun_out/r3f
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3f/gpu_tests.log 2>&1
tail -3 gpurun_out/r3f/gpu_tests.log | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -o gpurun_out/r3f/prof_fwd16_128_32to16 python tools/conv_bench.py fwd16 2 fp16 128,32,16 > gpurun_out/r3f/ncu_fwd.log 2>&1
tail -2 gpurun_out/r3f/ncu_fwd.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -o gpurun_out/r3f/prof_k1_128_32to16 python tools/conv_bench.py k1p16 2 fp16 128,32,16 > gpurun_out/r3f/ncu_k1.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3f/launches_train.csv python tools/one_step.py 2 > gpurun_out/r3f/one_step.log 2>&1
timeout 300 python tools/conv_bench.py fwd16 5 fp16 > gpurun_out/r3f/fwd16.txt 2>&1
timeout 300 python tools/conv_bench.py k1p16 5 fp16 > gpurun_out/r3f/k1.txt 2>&1
timeout 300 python tools/hbm_bench.py 10 gpurun_out/r3f/hbm.json > gpurun_out/r3f/hbm.txt 2>&1
timeout 1500 python bench.py > gpurun_out/r3f/bench_full.json 2> gpurun_out/r3f/bench_full.err
echo "bench rc=$?"; cut -c1-250 gpurun_out/r3f/bench_full.json; tail -2 gpurun_out/r3f/bench_full.err
