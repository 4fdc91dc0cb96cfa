mkdir -p gpurun_out/r2d
timeout 600 python tools/debug/p16_debug.py > gpurun_out/r2d/debug.log 2>&1
tail -80 gpurun_out/r2d/debug.log
