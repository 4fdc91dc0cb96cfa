"""Timing of the transposed / strided convs (stride-2 family on tcgen05) at inference and training sizes, with and
without the fused GroupNorm statistics.  Usage: python tools/up_bench.py [reps]"""
import importlib, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b3d = importlib.import_module("3d-brain-tumor-segmentation_b200")
ops = b3d.ops
dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn):
    for _ in range(2): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)

for sp, cin, cout, tr in [((40, 48, 40), 64, 32, True), ((80, 96, 80), 32, 16, True), ((20, 24, 20), 512, 64, True),
                          ((32, 32, 32), 64, 32, True), ((64, 64, 64), 32, 16, True),
                          ((160, 192, 160), 16, 16, False), ((80, 96, 80), 64, 32, False)]:
    x = torch.randn((1,) + sp + (cin,), device=dev)
    w = torch.randn((3, 3, 3) + ((cout, cin) if tr else (cin, cout)), device=dev) * 0.05
    bias = torch.zeros(cout, device=dev)
    od = tuple(2 * s for s in sp) if tr else tuple(s // 2 for s in sp)
    y = torch.empty((1,) + od + (cout,), device=dev)
    stats = torch.empty(1, 8, 2, dtype=torch.float64, device=dev)
    wp = ops.pack_weights(w, False, 2, tr)
    a = timeit(lambda: ops._call("b3d_conv3d_fwd", x, w, bias, y, 2, int(tr), 0, stats, 8, None, 0, wp))
    b = timeit(lambda: ops._call("b3d_conv3d_fwd", x, w, bias, y, 2, int(tr), 0, None, 1, None, 0, wp))
    mb = y.numel() * 4 / 1e6
    print(f"{'convT' if tr else 'conv s2'} {sp} {cin}->{cout}: stats {a*1e3:7.1f} us | no stats {b*1e3:7.1f} us | out {mb:6.1f} MB "
          f"({mb / b / 1e3:5.2f} TB/s)", flush=True)
