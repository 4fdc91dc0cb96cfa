mkdir -p gpurun_out/r3x
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3x/hbm_ncu.csv python tools/hbm_bench.py 1 > gpurun_out/r3x/hbm_under_ncu.txt 2>&1
wc -l gpurun_out/r3x/hbm_ncu.csv; du -sh gpurun_out/r3x
