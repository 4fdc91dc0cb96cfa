"""Times the fused weight gradients of a ResnetBlock's input convs (b3d_conv3d_wgrad_p16_block) against the two separate
launches.  Usage: wgrad_block_bench.py [iters]"""
import sys
from importlib import import_module
import torch

sys.path.insert(0, ".")
b3d = import_module("3d-brain-tumor-segmentation_b200")
ops = b3d.ops
dev = torch.device("cuda:0")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for S, pieces, cb in [(128, [32], 16), (128, [16], 16), (64, [32, 32], 32), (64, [32, 32, 32], 32), (64, [32], 32), (64, [16], 32)]:
    cin = sum(pieces)
    xs16 = [ops.p16_empty((1, S, S, S, c), flush, torch.bfloat16).normal_() for c in pieces]
    dy16 = ops.p16_empty((1, S, S, S, cb), flush, torch.bfloat16).normal_()
    dres16 = ops.p16_empty((1, S, S, S, cb), flush, torch.bfloat16).normal_()
    pad = xs16 + [None] * (4 - len(xs16))
    dw3, dw1 = torch.empty(3, 3, 3, cin, cb, device=dev), torch.empty(1, 1, 1, cin, cb, device=dev)
    plan = b3d._lib.lib.b3d_conv3d_wgrad_p16_plan(3, 1, 0, cin, cb, S)
    scratch = torch.empty(dy16.numel(), device=dev, dtype=torch.bfloat16) if plan == 3 else None
    t3 = timed(lambda: ops._call("b3d_conv3d_wgrad_p16", *pad, dy16, dw3, 1, 0, scratch))
    t1 = timed(lambda: ops._call("b3d_conv3d_wgrad_p16", *pad, dres16, dw1, 1, 0, None))
    tf = timed(lambda: ops._call("b3d_conv3d_wgrad_p16_block", *pad, dy16, dres16, dw3, dw1))
    print(f"{S}^3 {cin:3d}->{cb:2d}  3x3x3 {t3:6.1f} us  1x1x1 {t1:6.1f} us  fused {tf:6.1f} us  (L2 flushed between calls)")
