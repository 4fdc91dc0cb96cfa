mkdir -p gpurun_out/r2g
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g/launches_train.csv python tools/one_step.py 2 > gpurun_out/r2g/one_step.log 2>&1
tail -3 gpurun_out/r2g/one_step.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_apply_kernel -s 30 -c 1 -o gpurun_out/r2g/prof_gn_apply python tools/one_step.py 1 > gpurun_out/r2g/ncu_gn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:block_epilogue_bwd_apply -s 10 -c 1 -o gpurun_out/r2g/prof_bwd_apply python tools/one_step.py 1 > gpurun_out/r2g/ncu_bwd.log 2>&1
ls -la gpurun_out/r2g
