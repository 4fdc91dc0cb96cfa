"""One-shot hardware check of the kd-folded 3x3x3 kernel (csrc/conv_tc.cu, TcCfg::FOLD): results with the fold on vs
off (same operand rounding; only the summation order differs) for forward (+bias, GroupNorm statistics) and data
gradient over full / partial tiles, narrow inputs and outputs, then timings of the two 128^3 layers it targets."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b3d = importlib.import_module("3d-brain-tumor-segmentation_b200")
ops = b3d.ops
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(0)


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, generator=g) * scale).to(dev)


def run(x, w, bias, dy, act, groups):
    y, stats, _ = ops.conv3d(x, w, bias, 1, False, act, groups, False)
    dx = torch.empty_like(x)
    ops._call("b3d_conv3d_dgrad", dy, w, dx, 1, 0, 0, ops.pack_weights(w, True, 1, False))
    torch.cuda.synchronize()
    return y, stats, dx


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


CASES = [((1, 16, 32, 16), 16, 16, 0), ((1, 5, 24, 20), 32, 16, 0), ((2, 8, 16, 16), 16, 32, 0),
         ((1, 12, 20, 9), 2, 16, 0), ((1, 9, 17, 11), 16, 3, 1), ((1, 8, 16, 8), 64, 32, 0), ((1, 4, 16, 16), 32, 96, 0)]
ok = True
for (B, D, H, W), cin, cout, act in CASES:
    x, w = rnd(B, D, H, W, cin), rnd(3, 3, 3, cin, cout, scale=(2.0 / (27 * cin)) ** 0.5)
    bias, dy = rnd(cout), rnd(B, D, H, W, cout)
    S = D * H * W
    groups = 8 if (S % 8 == 0 and cout % 8 == 0) else 0
    ops.set_kd_fold(False)
    ref = run(x, w, bias, dy, act, groups)
    ops.set_kd_fold(True)
    got = run(x, w, bias, dy, act, groups)
    ops.set_kd_fold(False)
    e = [rel(got[0], ref[0]), rel(got[2], ref[2])] + ([rel(got[1], ref[1])] if groups else [])
    good = all(v < 1e-5 for v in e)
    ok &= good
    print(f"{(B, D, H, W)} {cin}->{cout} act={act}: y {e[0]:.1e} dx {e[1]:.1e}" + (f" stats {e[2]:.1e}" if groups else "") +
          ("  OK" if good else "  MISMATCH"), flush=True)

for cin, cout in ((16, 16), (32, 16), (16, 32)):
    n = 128 if cout == 16 else 64
    x, w, bias = rnd(1, n, n, n, cin), rnd(3, 3, 3, cin, cout, scale=0.05), rnd(cout)
    line = f"{n}^3 {cin}->{cout} fwd:"
    for fold in (False, True):
        ops.set_kd_fold(fold)
        wp = ops.pack_weights(w, False, 1, False)
        y = torch.empty(1, n, n, n, cout, device=dev)
        st = torch.zeros(1, 8, 2, device=dev, dtype=torch.float64)
        for _ in range(2):
            ops._call("b3d_conv3d_fwd", x, w, bias, y, 1, 0, 0, st, 8, None, 0, wp)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops._call("b3d_conv3d_fwd", x, w, bias, y, 1, 0, 0, st, 8, None, 0, wp)
        e1.record()
        torch.cuda.synchronize()
        line += f"  fold={int(fold)} {e0.elapsed_time(e1) / 5 * 1e3:.1f} us"
    ops.set_kd_fold(False)
    print(line, flush=True)
print("ALL OK" if ok else "FAILED")
