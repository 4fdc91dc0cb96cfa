mkdir -p gpurun_out/r2i
timeout 600 python tools/hbm_bench.py 10 gpurun_out/r2i/hbm.json > gpurun_out/r2i/hbm.txt 2>&1
cat gpurun_out/r2i/hbm.txt | grep -v "^$"
