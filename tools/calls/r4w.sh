timeout 300 python tools/wgrad_block_bench.py 10 2>&1 | tail -8
for k in 1 0; do B3D_FUSE_BLOCK_WGRAD=$k timeout 300 python tools/one_step.py 2 2>&1 | tail -1; done
