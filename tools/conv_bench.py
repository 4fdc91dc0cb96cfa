"""Per-layer timing harness for the tcgen05 conv kernels (fwd = dgrad kernel, and wgrad), CUDA events,
L2 flushed between launches.  Usage: python tools/conv_bench.py [fwd|wgrad|all] [reps]"""
import importlib
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b3d = importlib.import_module("3d-brain-tumor-segmentation_b200")
ops = b3d.ops
dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
ops.set_conv_precision(prec, prec)
print("conv precision:", ops.get_conv_precision())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

# (spatial, Cin, Cout) of the default model's 3x3x3 convs, largest first
SHAPES = [(128, 2, 16), (128, 16, 2), (128, 32, 16), (128, 16, 16), (64, 96, 32), (64, 64, 32), (64, 32, 32), (64, 16, 32),
          (32, 256, 64), (32, 192, 64), (32, 64, 64), (16, 512, 128), (16, 128, 128)]


def timeit(fn):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


if len(sys.argv) > 4:      # e.g. "128,32,16": only this shape (ncu captures)
    SHAPES = [tuple(int(v) for v in sys.argv[4].split(","))]
for n, cin, cout in SHAPES:
    x = torch.randn(1, n, n, n, cin, device=dev)
    dy = torch.randn(1, n, n, n, cout, device=dev)
    w = torch.randn(3, 3, 3, cin, cout, device=dev) * 0.05
    bias = torch.zeros(cout, device=dev)
    y = torch.empty(1, n, n, n, cout, device=dev)
    stats = torch.empty(1, 8, 2, dtype=torch.float64, device=dev)
    flops = 2.0 * n ** 3 * 27 * cin * cout
    line = f"{n:4d}^3 {cin:4d}->{cout:4d}  {flops / 1e9:7.1f} GF "
    if which in ("fwd", "all"):
        wp = ops.pack_weights(w, False)
        st = stats if cout % 8 == 0 else None
        ms = timeit(lambda: ops._call("b3d_conv3d_fwd", x, w, bias, y, 1, 0, 0, st, 8, None, 0, wp))
        ms2 = timeit(lambda: ops._call("b3d_conv3d_fwd", x, w, bias, y, 1, 0, 0, None, 1, None, 0, wp))
        io = 4.0 * n ** 3 * (cin + cout)
        line += (f"| fwd {ms * 1e3:8.1f} us {flops / ms / 1e9:7.1f} TF/s (io {io / ms / 1e6:6.0f} GB/s) "
                 f"| no-stats {ms2 * 1e3:8.1f} us {flops / ms2 / 1e9:7.1f} TF/s ")
    if which in ("fwd16", "all") and cin % 16 == 0 and cout % 16 == 0:
        # the way the model feeds it: fp16 / bf16 P16 twins, TMA-fetched
        dt = torch.float16 if prec == "fp16" else torch.bfloat16
        wp = ops.pack_weights(w, False)
        tw = ops.to_p16(x, dt)
        st = stats if cout % 8 == 0 else None
        ms = timeit(lambda: ops._call("b3d_conv3d_fwd_p16", tw, None, None, None, w, bias, y, 1, 0, 0, st, 8, None, 0, wp))
        line += f"| fwd(P16/TMA) {ms * 1e3:8.1f} us {flops / ms / 1e9:7.1f} TF/s "
        dyb = ops.to_p16(dy, torch.bfloat16)
        xb16 = ops.to_p16(x, torch.bfloat16)
        prev = ops.get_conv_precision()
        ops.set_conv_precision(prev[0], "bf16")
        wpd = ops.pack_weights(w, True)
        dx = torch.empty_like(x)
        ms = timeit(lambda: ops._call("b3d_conv3d_dgrad_p16", dyb, w, dx, 1, 0, 0, wpd))
        line += f"| dgrad(P16/TMA) {ms * 1e3:8.1f} us {flops / ms / 1e9:7.1f} TF/s "
        plan = b3d._lib.lib.b3d_conv3d_wgrad_p16_plan(3, 1, 0, cin, cout, n)
        scratch = torch.empty(dyb.numel(), device=dev, dtype=torch.bfloat16) if plan == 3 else None
        dw = torch.empty_like(w)
        ms = timeit(lambda: ops._call("b3d_conv3d_wgrad_p16", xb16, None, None, None, dyb, dw, 1, 0, scratch))
        line += f"| wgrad(P16, plan {plan}) {ms * 1e3:8.1f} us {flops / ms / 1e9:7.1f} TF/s "
        ops.set_conv_precision(*prev)
    if which in ("wgrad", "all"):
        dw = torch.empty_like(w)
        import ctypes
        xc, yc = ctypes.c_longlong(), ctypes.c_longlong()
        assert b3d._lib.lib.b3d_conv3d_wgrad_plan(3, 1, 0, cin, cout, ctypes.byref(xc), ctypes.byref(yc))
        xb = torch.empty(n ** 3 * xc.value, device=dev, dtype=torch.bfloat16)
        yb = torch.empty(n ** 3 * yc.value, device=dev, dtype=torch.bfloat16)
        ms = timeit(lambda: ops._call("b3d_conv3d_wgrad", x, dy, dw, None, 1, 0, xb, yb, 0))
        line += f"| wgrad(+casts) {ms * 1e3:8.1f} us {flops / ms / 1e9:7.1f} TF/s "
    if which in ("k1p16", "all") and cin % 16 == 0 and cout % 16 == 0:
        w1 = torch.randn(1, 1, 1, cin, cout, device=dev) * 0.05
        wp1 = ops.pack_weights(w1, False)
        gap = torch.empty(1, cout, device=dev)
        tw = ops.to_p16(x, torch.float16 if prec == "fp16" else torch.bfloat16)
        ms = timeit(lambda: ops._call("b3d_conv3d_fwd_p16", tw, None, None, None, w1, bias, y, 1, 0, 0, None, 1, gap, 0, wp1))
        io = n ** 3 * (2.0 * cin + 4.0 * cout)
        line += f"| 1x1x1+gap(P16) {ms * 1e3:8.1f} us (io {io / ms / 1e6:6.0f} GB/s) "
    if which in ("k1", "all") and cin >= 8 and cout >= 16:
        w1 = torch.randn(1, 1, 1, cin, cout, device=dev) * 0.05
        wp1 = ops.pack_weights(w1, False)
        gap = torch.empty(1, cout, device=dev)
        ms = timeit(lambda: ops._call("b3d_conv3d_fwd", x, w1, bias, y, 1, 0, 0, None, 1, gap, 0, wp1))
        io = 4.0 * n ** 3 * (cin + cout)
        line += f"| 1x1x1+gap {ms * 1e3:8.1f} us (io {io / ms / 1e6:6.0f} GB/s) "
    print(line, flush=True)
