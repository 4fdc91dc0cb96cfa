import importlib, os, sys, collections, traceback
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
b3d = importlib.import_module("3d-brain-tumor-segmentation_b200")
from oracle import ref_model as R
ops = b3d.ops
dev = torch.device("cuda:0")
rel = lambda a, b: float((a.double().cpu() - b.double().cpu()).norm() / (b.double().cpu().norm() + 1e-30))

# count materialize call sites
sites = collections.Counter()
_orig = ops.materialize
def mat(t):
    if ops.is_virtual(t):
        fr = traceback.extract_stack(limit=3)[0]
        sites[(fr.name, fr.lineno)] += 1
    return _orig(t)
ops.materialize = mat

print("=== single ResnetBlock, standalone (not fused), B=2, Cin=16 -> 16, fp16 / bf16 forward")
for prec in ("fp16", "bf16"):
    ops.set_conv_precision(prec, "bf16")
    torch.manual_seed(0)
    x = torch.randn(2, 8, 8, 8, 16, device=dev)
    outs = {}
    for on in (False, True):
        ops.P16["on"] = on
        b3d.keras_compat.set_seed(5)
        blk = b3d.ResnetBlock(16)
        xd = x.clone().requires_grad_(True)
        y = blk(xd)
        g = torch.randn(y.shape, generator=torch.Generator().manual_seed(1)).to(dev)
        (y * g).sum().backward()
        torch.cuda.synchronize()
        outs[on] = (y.detach().clone(), xd.grad.clone(), {n: v.tensor.grad.clone() for n, v in zip(range(99), blk.variables())},
                    getattr(y, "_p16", None))
    ops.P16["on"] = True
    y0, dx0, g0, _ = outs[False]; y1, dx1, g1, tw = outs[True]
    print(prec, "y", rel(y1, y0), "dx", rel(dx1, dx0), "twin-vs-y", None if tw is None else rel(ops._materialize_from(y1.shape, [tw]), y1))
    names = [v.name for v in blk.variables()]
    print("   grads", {f"{i}:{names[i]}": f"{rel(g1[i], g0[i]):.1e}" for i in g0})
ops.set_conv_precision("fp16", "bf16")

print("=== whole model 32^3: per-parameter gradient difference P16 on vs off")
crop = (32, 32, 32)
p = R.init_params(R.param_shapes(crop=crop), dtype=torch.float32)
x, y, eps, mask = R.synth_batch((1,) + crop, dtype=torch.float32)
f = lambda t: t.to(dev)
res = {}
for on in (False, True):
    ops.P16["on"] = on
    sites.clear()
    model = b3d.Model()
    with torch.no_grad():
        model(torch.zeros((1,) + crop + (2,), device=dev), training=False, inference=False)
    model.load_named_weights(p)
    with b3d.GradientTape() as tape:
        outs = model(f(x), training=True, inference=False, dropout_mask=f(mask), eps=f(eps))
        loss = b3d.DiceVAELoss()(f(x), f(y), *outs) + b3d.reduce_sum(model.losses)
    tape.gradient(loss, model.trainable_variables, direct=True)
    torch.cuda.synchronize()
    res[on] = {k: v.grad.clone() for k, v in model.named_variables().items()}
    print("P16", on, "loss", float(loss), "materialize sites:", dict(sites))
ops.P16["on"] = True
bad = sorted(((rel(res[True][k], res[False][k]), k) for k in res[True]), reverse=True)
for e, k in bad[:40]:
    print(f"  {e:.2e} {k}")
print("median", bad[len(bad)//2])
