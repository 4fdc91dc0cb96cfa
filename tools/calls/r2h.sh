mkdir -p gpurun_out/r2h
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2h/gpu_tests.log 2>&1
tail -8 gpurun_out/r2h/gpu_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r2h/bench.json 2> gpurun_out/r2h/bench.err
cat gpurun_out/r2h/bench.json | cut -c1-300; tail -3 gpurun_out/r2h/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h/launches_train.csv python tools/one_step.py 2 > gpurun_out/r2h/one_step.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -o gpurun_out/r2h/prof_tma_fold16_32to16 python tools/conv_bench.py fwd16 2 fp16 128,32,16 > gpurun_out/r2h/ncu1.log 2>&1
tail -3 gpurun_out/r2h/ncu1.log
timeout 300 python tools/conv_bench.py all 5 fp16 > gpurun_out/r2h/conv_bench.txt 2>&1
cat gpurun_out/r2h/conv_bench.txt | cut -c1-400
