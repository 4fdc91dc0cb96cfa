mkdir -p gpurun_out/r2q
timeout 300 python tools/conv_bench.py k1p16 5 fp16 > gpurun_out/r2q/k1.txt 2>&1; cat gpurun_out/r2q/k1.txt | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -o gpurun_out/r2q/prof_k1_128_32to16 python tools/conv_bench.py k1p16 2 fp16 128,32,16 > gpurun_out/r2q/ncu1.log 2>&1
tail -2 gpurun_out/r2q/ncu1.log
