"""Small single-GPU workload for compute-sanitizer (memcheck / racecheck / synccheck): every warp-specialised mbarrier
kernel of the path once, at tiny shapes — the tcgen05 conv (TMA-fed P16 form: kd-folded, unfolded, pointwise, transposed;
thread-loader form: fp32 input, space-to-depth), its split-K finish, both weight-gradient kernels, the P16 producers.
Usage: compute-sanitizer --tool racecheck python tools/sanitize_target.py"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b3d = importlib.import_module("3d-brain-tumor-segmentation_b200")
ops = b3d.ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g).to(dev)

CASES = [((8, 16, 8), [16], 16, 3, 1, False), ((4, 16, 16), [16, 16], 32, 3, 1, False), ((4, 16, 16), [32, 64], 64, 3, 1, False),
         ((4, 8, 8), [64, 64], 128, 3, 1, False), ((4, 8, 8), [32], 16, 1, 1, False), ((8, 8, 16), [32], 32, 3, 2, False),
         ((4, 4, 8), [64], 32, 3, 2, True)]
for sp, cs, cout, k, stride, tr in CASES:
    xs = [rnd(1, *sp, c) for c in cs]
    x = torch.cat(xs, -1).contiguous()
    cin = sum(cs)
    w = rnd(k, k, k, *((cout, cin) if tr else (cin, cout))) * 0.1
    bias = rnd(cout)
    od = tuple(2 * n for n in sp) if tr else tuple(n // stride for n in sp)
    y = torch.empty(1, *od, cout, device=dev)
    stats = torch.empty(1, 8, 2, dtype=torch.float64, device=dev)
    wp = ops.pack_weights(w, False, stride, tr)
    ops._call("b3d_conv3d_fwd", x, w, bias, y, stride, int(tr), 0, stats, 8, None, 0, wp)          # thread loader
    tw = [ops.to_p16(t, torch.float16) for t in xs]
    ops._call("b3d_conv3d_fwd_p16", *(tw + [None] * (4 - len(tw))), w, bias, y, stride, int(tr), 0, stats, 8, None, 0, wp)
    dy = rnd(1, *od, cout)
    dy16 = ops.to_p16(dy, torch.bfloat16)
    dx = torch.empty_like(x)
    ops._call("b3d_conv3d_dgrad_p16", dy16, w, dx, stride, int(tr), 0, ops.pack_weights(w, True, stride, tr))
    plan = b3d._lib.lib.b3d_conv3d_wgrad_p16_plan(k, stride, int(tr), cin, cout, od[2])
    twb = [ops.to_p16(t, torch.bfloat16) for t in xs]
    if tr:
        twb = [ops._p16_cat(twb)]
    scratch = None
    if plan == 2:
        scratch = torch.empty(dy16.numel() if tr else sum(t.numel() for t in twb), device=dev, dtype=torch.bfloat16)
    elif plan == 3:
        scratch = torch.empty(dy16.numel(), device=dev, dtype=torch.bfloat16)
    dw = torch.empty_like(w)
    ops._call("b3d_conv3d_wgrad_p16", *(twb + [None] * (4 - len(twb))), dy16, dw, stride, int(tr), scratch)
    if k == 3 and stride == 1 and not tr and b3d._lib.lib.b3d_conv3d_wgrad_p16_block_ok(cin, cout, od[1], od[2]):
        dres16 = ops.to_p16(rnd(1, *od, cout), torch.bfloat16)        # kd-in-M kernel with the pointwise layer's dw
        dwp = torch.empty(1, 1, 1, cin, cout, device=dev)
        ops._call("b3d_conv3d_wgrad_p16_block", *(twb + [None] * (4 - len(twb))), dy16, dres16, dw, dwp)
    torch.cuda.synchronize()
    print("ok", sp, cs, cout, k, stride, tr, "wgrad plan", plan, flush=True)

# whole training step at 16^3 (every other kernel of the path, graph-free)
import synthdata as R
crop = (16, 16, 16)
p = R.init_params(R.param_shapes(crop=crop), dtype=torch.float32)
x, y, eps, mask = R.synth_batch((1,) + crop, dtype=torch.float32)
model = b3d.Model()
with torch.no_grad():
    model(x.to(dev), training=False, inference=False)
model.load_named_weights(p)
opt = b3d.ScheduledOptim(learning_rate=1e-4)
opt(epoch=0)
out = b3d.train_step(model, opt, b3d.DiceVAELoss(), b3d.DiceCoefficient(), x.to(dev), y.to(dev))
torch.cuda.synchronize()
print("train step ok, loss", float(out[0]))
