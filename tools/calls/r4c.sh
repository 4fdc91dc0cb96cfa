mkdir -p gpurun_out/r4c
timeout 300 python tools/conv_bench.py fwd16 5 fp16 2>&1 | grep -o "^ *[0-9]*^3 *[0-9]*-> *[0-9]*\|wgrad(P16[^|]*" | paste - - | sed -n '2,5p' | cut -c1-100
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r4c/launches_infer.csv python tools/one_infer.py 2 1 > gpurun_out/r4c/one_infer.log 2>&1
tail -1 gpurun_out/r4c/one_infer.log
timeout 300 ncu --set full --clock-control none -k regex:conv3_wgrad_ts_kernel -s 1 -c 1 -o /tmp/prof_ts python tools/conv_bench.py fwd16 2 fp16 128,16,16 > gpurun_out/r4c/ncu_ts.log 2>&1
ncu -i /tmp/prof_ts.ncu-rep --page raw --csv > gpurun_out/r4c/ncu_wgrad_ts_128_16to16.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:conv3_wgrad_tc_kernel -s 1 -c 1 -o /tmp/prof_wg python tools/conv_bench.py fwd16 2 fp16 128,32,16 > gpurun_out/r4c/ncu_wg.log 2>&1
ncu -i /tmp/prof_wg.ncu-rep --page raw --csv > gpurun_out/r4c/ncu_wgrad_tc_128_32to16.csv 2>/dev/null
ls -la gpurun_out/r4c | head; du -sh gpurun_out
