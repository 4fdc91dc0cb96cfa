mkdir -p gpurun_out/r4e
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -o /tmp/prof_fwd python tools/conv_bench.py fwd16 2 fp16 128,32,16 > gpurun_out/r4e/ncu_fwd.log 2>&1
ncu -i /tmp/prof_fwd.ncu-rep --page raw --csv > gpurun_out/r4e/ncu_fwd16_128_32to16.csv 2>/dev/null
ncu -i /tmp/prof_fwd.ncu-rep --page details > gpurun_out/r4e/ncu_fwd16_128_32to16_details.txt 2>/dev/null
ncu -i /tmp/prof_fwd.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r4e/ncu_fwd16_128_32to16_source.csv.gz
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -o /tmp/prof_k1 python tools/conv_bench.py k1p16 2 fp16 128,32,16 > gpurun_out/r4e/ncu_k1.log 2>&1
ncu -i /tmp/prof_k1.ncu-rep --page raw --csv > gpurun_out/r4e/ncu_k1_128_32to16.csv 2>/dev/null
ncu -i /tmp/prof_k1.ncu-rep --page details > gpurun_out/r4e/ncu_k1_128_32to16_details.txt 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r4e/launches_train.csv python tools/one_step.py 2 > gpurun_out/r4e/one_step.log 2>&1
timeout 300 python tools/conv_bench.py fwd16 5 fp16 > gpurun_out/r4e/fwd16.txt 2>&1
timeout 300 python tools/conv_bench.py k1p16 5 fp16 > gpurun_out/r4e/k1.txt 2>&1
timeout 300 python tools/hbm_bench.py 10 gpurun_out/r4e/hbm.json > gpurun_out/r4e/hbm.txt 2>&1
timeout 1500 python bench.py > gpurun_out/r4e/bench_full.json 2> gpurun_out/r4e/bench_full.err
echo "bench rc=$?"; cut -c1-250 gpurun_out/r4e/bench_full.json; tail -2 gpurun_out/r4e/bench_full.err
du -sh gpurun_out
