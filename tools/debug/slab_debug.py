import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
b3d = importlib.import_module("3d-brain-tumor-segmentation_b200")
import synthdata as R
ops = b3d.ops
dev = torch.device("cuda:0")
rel = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
shape = (160, 192, 160)
p = R.init_params(R.param_shapes(crop=(16, 16, 16)), dtype=torch.float32)
g = torch.Generator().manual_seed(123)
x = torch.randn((1,) + shape + (2,), generator=g)
x[:, 155:], x[:, :, 190:], x[:, :, :, 147:] = 0, 0, 0
xd = x.to(dev)
model = b3d.Model()
with torch.no_grad():
    model(torch.zeros(1, 16, 16, 16, 2, device=dev), training=False, inference=False)
model.load_named_weights(p)
for fold in (True, False):
    ops.set_kd_fold(fold)
    for p16 in (True, False):
        ops.P16["on"] = p16
        with torch.no_grad():
            whole = model(xd, training=False, inference=True)[0]
        for world in (4, 8):
            got, _ = b3d.slab.run_virtual_ranks(model, xd, world)
            bounds = b3d.slab_bounds(shape[0], world)
            per = [f"{rel(got[:, a:b], whole[:, a:b]):.1e}" for a, b in bounds]
            print(f"fold {fold} p16(whole) {p16} world {world}: total {rel(got, whole):.2e} per slab {per}", flush=True)
ops.set_kd_fold(True); ops.P16["on"] = True
