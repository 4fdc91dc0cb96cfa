"""Profiling helper: N eager training steps of the default model at 128^3 (no CUDA graph), for
`ncu --metrics gpu__time_duration.sum` launch lists.  NVTX-free; the launch order is deterministic."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b3d = importlib.import_module("3d-brain-tumor-segmentation_b200")
import synthdata as R  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
crop = tuple(int(v) for v in sys.argv[2].split("x")) if len(sys.argv) > 2 else (128, 128, 128)
dev = torch.device("cuda:0")
p = R.init_params(R.param_shapes(crop=crop), dtype=torch.float32)
x, y, _, _ = R.synth_batch((1,) + crop, dtype=torch.float32)
x, y = x.to(dev), y.to(dev)
model = b3d.Model()
model(x, training=False, inference=False)
model.load_named_weights(p)
opt = b3d.ScheduledOptim(learning_rate=1e-4)
opt(epoch=0)
args = (model, opt, b3d.DiceVAELoss(), b3d.DiceCoefficient())
for i in range(steps):
    torch.cuda.synchronize()
    n0 = b3d.ops.LAUNCHES["n"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = b3d.train_step(*args, x, y)
    e1.record()
    torch.cuda.synchronize()
    print(f"step {i}: loss {float(out[0]):.5f}  {e0.elapsed_time(e1):.2f} ms  abi calls {b3d.ops.LAUNCHES['n'] - n0}",
          flush=True)
