mkdir -p gpurun_out/r3l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/dp_debug2.py > gpurun_out/r3l/dbg.log 2>&1
grep -v "OMP_NUM\|\*\*\*" gpurun_out/r3l/dbg.log | cut -c1-400 | tail -40
