mkdir -p gpurun_out/r4k
timeout 600 ncu --set full --clock-control none -k regex:conv3_wgrad_kdf -c 4 -o /tmp/kdf python tools/conv_bench.py fwd16 1 fp16 > gpurun_out/r4k/run.log 2>&1
ncu -i /tmp/kdf.ncu-rep --page raw --csv > gpurun_out/r4k/kdf_raw.csv 2>/dev/null
ls -la gpurun_out/r4k; tail -3 gpurun_out/r4k/run.log
