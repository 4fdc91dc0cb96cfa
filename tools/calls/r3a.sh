mkdir -p gpurun_out/r3a
for i in 1 2 3; do
timeout 600 python -m pytest tests/test_gpu_baseline_shapes.py -m gpu -q -x -s -k "train_step_128cube" > gpurun_out/r3a/run$i.log 2>&1
grep -E "^  gradients|^128|passed|failed" gpurun_out/r3a/run$i.log | cut -c1-200
done
