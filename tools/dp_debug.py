import importlib, os, sys, time, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
def log(*a):
    print(f"[r{rank} {time.time() % 1000:7.2f}]", *a, flush=True)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
log("init pg")
dist.init_process_group("nccl", device_id=dev)
log("pg ok")
b3d = importlib.import_module("3d-brain-tumor-segmentation_b200")
from oracle import ref_model as R
crop = (64, 64, 64)
p = R.init_params(R.param_shapes(crop=crop), dtype=torch.float32)
x, y, _, _ = R.synth_batch((1,) + crop, seed=100 * rank, dtype=torch.float32)
x, y = x.to(dev), y.to(dev)
model = b3d.Model()
model(x, training=False, inference=False)
model.load_named_weights(p)
opt = b3d.ScheduledOptim(learning_rate=1e-4); opt(epoch=0)
mode = sys.argv[1] if len(sys.argv) > 1 else "eager"
overlap = "nooverlap" not in sys.argv
dp = b3d.DataParallel(model, opt, world, overlap=overlap)
log("dp ok; buckets", [(lo, hi, len(m)) for lo, hi, m in dp.buckets])
args = (model, opt, b3d.DiceVAELoss(), b3d.DiceCoefficient())
for i in range(2):
    out = b3d.train_step(*args, x, y, dp=dp)
    torch.cuda.synchronize()
    log("eager step", i, float(out[0]))
th = model.flat.theta.clone()
dist.all_reduce(th, op=dist.ReduceOp.MAX)
log("replicas identical:", bool(torch.equal(th, model.flat.theta)))
if mode == "graph":
    log("capturing")
    step = b3d.GraphedTrainStep(*args, x, y, warmup=1, dp=dp)
    log("captured")
    for i in range(3):
        out = step()
        torch.cuda.synchronize()
        log("graph step", i, float(out[0]))
    th = model.flat.theta.clone()
    dist.all_reduce(th, op=dist.ReduceOp.MAX)
    log("replicas identical:", bool(torch.equal(th, model.flat.theta)))
dist.barrier()
log("done")
dist.destroy_process_group()
