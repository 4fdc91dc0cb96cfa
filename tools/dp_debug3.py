"""1-GPU debug: which parameters notify the bucket scheduler more than once per backward?"""
import importlib, os, sys, collections, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29555")
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
b3d = importlib.import_module("3d-brain-tumor-segmentation_b200")
import synthdata as R
crop, depth = (64, 64, 64), 3
p = R.init_params(R.param_shapes(crop=crop, depth=depth), dtype=torch.float32)
x, y, eps, mask = R.synth_batch((1,) + crop, seed=0, latent=32, dtype=torch.float32)
f = lambda t: t.to(dev)
m = b3d.Model(depth=depth)
with torch.no_grad():
    m(torch.zeros((1,) + crop + (2,), device=dev), training=False, inference=False)
m.load_named_weights(p)
o = b3d.ScheduledOptim(learning_rate=1e-4); o(epoch=0)
loss_fn = b3d.DiceVAELoss()
dp = b3d.DataParallel(m, o, 1, objective="replica_mean", loss_fn=loss_fn)
cnt = collections.Counter()
orig = dp._on_grad
names = {id(t): k for k, t in m.named_variables().items()}
def spy(param):
    if dp._active:
        cnt[names.get(id(param), "?")] += 1
    return orig(param)
dp._on_grad = spy
m.flat.grad_ready_cb = spy
rc = collections.Counter()
orig_reduce = dp._reduce
def spy_reduce(bi):
    rc[bi] += 1
    print("reduce bucket", bi, "pending", list(dp._pending), "active", dp._active, flush=True)
    return orig_reduce(bi)
dp._reduce = spy_reduce
b3d.train_step(m, o, loss_fn, b3d.DiceCoefficient(), f(x), f(y), dropout_mask=f(mask), eps=f(eps), dp=dp)
torch.cuda.synchronize()
print("reduce counts:", dict(rc), "bucket sizes", [len(mm) for _, _, mm in dp.buckets])
print("params:", len(names), "notified:", len(cnt), "multiple:", {k: v for k, v in cnt.items() if v != 1})
print("never:", [k for k in names.values() if k not in cnt][:20])
dist.destroy_process_group()
