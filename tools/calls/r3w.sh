for i in 1 2 3; do timeout 300 python tools/conv_bench.py fwd16 7 fp16 128,32,16 2>&1 | grep "128^3" | cut -c1-150; done
for i in 1 2; do timeout 300 python tools/conv_bench.py fwd16 7 fp16 128,16,16 2>&1 | grep "128^3" | cut -c1-150; done
