"""Profiling helper: eager whole-volume inference forwards (160x192x160, VAE off) for ncu launch lists.
Usage: one_infer.py [forwards] [virtual ranks]  — with virtual ranks > 1 the volume is cut into depth slabs that run as
threads on this one GPU (slab.run_virtual_ranks): the launch list then holds every rank's kernels (exchanges as copies)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b3d = importlib.import_module("3d-brain-tumor-segmentation_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
world = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
model = b3d.Model()
x = torch.randn(1, 160, 192, 160, 2, device=dev)
with torch.no_grad():
    model(torch.zeros(1, 16, 16, 16, 2, device=dev), training=False, inference=True)      # build
    for i in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if world > 1:
            y, _ = b3d.slab.run_virtual_ranks(model, x, world)
        else:
            y = model(x, training=False, inference=True)[0]
        e1.record(); torch.cuda.synchronize()
        print(f"forward {i}: {e0.elapsed_time(e1):.2f} ms", flush=True)
