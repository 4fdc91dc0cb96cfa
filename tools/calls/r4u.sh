mkdir -p gpurun_out/r4u
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r4u/launches_train.csv python tools/one_step.py 2 > gpurun_out/r4u/one_step.log 2>&1
tail -2 gpurun_out/r4u/one_step.log
