mkdir -p gpurun_out/r5a
O=gpurun_out/r5a
timeout 300 ncu --set full --clock-control none -k regex:conv_tc_kernel -s 1 -c 1 -o /tmp/prof_fwd python tools/conv_bench.py fwd16 2 fp16 128,32,16 > $O/ncu_fwd.log 2>&1
ncu -i /tmp/prof_fwd.ncu-rep --page raw --csv > $O/ncu_conv_tc_tma_kdfold_n16_128cube_32to16.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:conv3_wgrad_kdf -s 1 -c 1 -o /tmp/prof_kdf python tools/conv_bench.py fwd16 2 fp16 128,32,16 > $O/ncu_kdf.log 2>&1
ncu -i /tmp/prof_kdf.ncu-rep --page raw --csv > $O/ncu_wgrad_kdf_128cube_32to16.csv 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_train.csv python tools/one_step.py 2 > $O/one_step.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_infer.csv python tools/one_infer.py 2 1 > $O/one_infer.log 2>&1
timeout 300 python tools/conv_bench.py fwd16 5 fp16 > $O/conv_bench.txt 2>&1
timeout 300 python tools/wgrad_block_bench.py 10 > $O/wgrad_block_bench.txt 2>&1
timeout 1500 python bench.py > $O/bench_full.json 2> $O/bench_full.err
echo "bench rc=$?"; cut -c1-200 $O/bench_full.json; tail -2 $O/bench_full.err
du -sh gpurun_out
