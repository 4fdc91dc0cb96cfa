"""2-GPU debug: DP (replica_mean) gradient vs the mean of per-crop single-GPU gradients, per bucket and per tensor."""
import importlib, os, sys, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
b3d = importlib.import_module("3d-brain-tumor-segmentation_b200")
import synthdata as R
crop, depth = (64, 64, 64), 3
p = R.init_params(R.param_shapes(crop=crop, depth=depth), dtype=torch.float32)
data = [R.synth_batch((1,) + crop, seed=100 * r, latent=32, dtype=torch.float32) for r in range(world)]
f = lambda t: t.to(dev)
rel = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
def fresh():
    m = b3d.Model(depth=depth)
    with torch.no_grad():
        m(torch.zeros((1,) + crop + (2,), device=dev), training=False, inference=False)
    m.load_named_weights(p)
    o = b3d.ScheduledOptim(learning_rate=1e-4); o(epoch=0)
    return m, o
def grads_of(model, x, y, eps, mask):
    with b3d.GradientTape() as tape:
        outs = model(f(x), training=True, inference=False, dropout_mask=f(mask), eps=f(eps))
        loss = b3d.DiceVAELoss()(f(x), f(y), *outs)
    tape.gradient(loss, model.trainable_variables, direct=True)
    return float(loss), model.flat.grad.clone()
model, _ = fresh()
per = [grads_of(model, *d) for d in data]
per2 = [grads_of(model, *d) for d in data]
print(f"[r{rank}] repeatability of the single-GPU gradient: {[rel(a[1], b[1]) for a, b in zip(per, per2)]}", flush=True)
g_mean = sum(g for _, g in per) / world
x, y, eps, mask = data[rank]
for overlap in (True, False):
    model, opt = fresh()
    loss_fn = b3d.DiceVAELoss()
    dp = b3d.DataParallel(model, opt, world, objective="replica_mean", loss_fn=loss_fn, overlap=overlap)
    b3d.train_step(model, opt, loss_fn, b3d.DiceCoefficient(), f(x), f(y), dropout_mask=f(mask), eps=f(eps), dp=dp)
    torch.cuda.synchronize()
    g = model.flat.grad * opt.grad_scale
    print(f"[r{rank}] overlap={overlap}: total rel {rel(g, g_mean):.3e}; vs own crop/2 {rel(g, per[rank][1] / world):.3e}; |g| {float(g.norm()):.3e} |ref| {float(g_mean.norm()):.3e}", flush=True)
    if rank == 0:
        for bi, (lo, hi, m) in enumerate(dp.buckets):
            print(f"   bucket {bi} [{lo},{hi}) {len(m)} tensors: rel {rel(g[lo:hi], g_mean[lo:hi]):.3e} |g| {float(g[lo:hi].norm()):.3e} |ref| {float(g_mean[lo:hi].norm()):.3e}", flush=True)
        nv = model.named_variables()
        worst = sorted(((rel(g[model.flat.spans[id(t)][0]:model.flat.spans[id(t)][0] + t.numel()], g_mean[model.flat.spans[id(t)][0]:model.flat.spans[id(t)][0] + t.numel()]), k) for k, t in nv.items()), reverse=True)
        print("   worst tensors:", [(k, f"{e:.2e}") for e, k in worst[:8]], flush=True)
        print("   best tensors:", [(k, f"{e:.2e}") for e, k in worst[-5:]], flush=True)
dist.barrier()
dist.destroy_process_group()
