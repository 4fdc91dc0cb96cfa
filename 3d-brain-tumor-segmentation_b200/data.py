"""GPU-side training-example pipeline — what the reference's tf.data map function does per example
(/root/reference/train.py:12-47, `parse_example` inside `prepare_dataset`), on device-resident preprocessed volumes:

    var   = moments(x, axes=(0,1,2))                      x += U(-0.1, 0.1)[c] * sqrt(var);  x *= U(0.9, 1.1)[c]
    xy    = random_crop(concat(x, y), crop_size)          for axis in 0,1,2: flip with probability 1/2
    y     = one_hot(int(y), out_ch + 1)[..., 1:]

as one reduction pass and one fused gather kernel (csrc/augment.cu).  TFRecord parsing itself needs TF protos and is
out of scope: volumes come as tensors / .npy arrays.  Random draws are made here with a torch Generator (or injected,
for parity tests), never inside the kernels.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import ops


def parse_example(x: torch.Tensor, y: torch.Tensor, crop_size: Sequence[int], out_ch: int,
                  generator: Optional[torch.Generator] = None, shift=None, scale=None, offset=None, flips=None,
                  moments: Optional[torch.Tensor] = None):
    """x [H,W,D,C] fp32, y [H,W,D,1] fp32 labels (CUDA) -> (x_crop [*crop_size, C], y_onehot [*crop_size, out_ch]).
    shift/scale [C], offset (3 ints), flips (3 bools) override the random draws; `moments` (fp64 [C,2]) may be
    cached per volume, it only depends on x."""
    ops._check(x, "x")
    x, y = x.contiguous(), y.contiguous().to(torch.float32)
    C = x.shape[-1]
    g = generator
    u = lambda lo, hi: (torch.rand(C, generator=g) * (hi - lo) + lo)
    shift = torch.as_tensor(u(-0.1, 0.1) if shift is None else shift, dtype=torch.float32).to(x.device).contiguous()
    scale = torch.as_tensor(u(0.9, 1.1) if scale is None else scale, dtype=torch.float32).to(x.device).contiguous()
    if offset is None:
        offset = [int(torch.randint(0, s - c + 1, (1,), generator=g)) for s, c in zip(x.shape[:3], crop_size)]
    if flips is None:
        flips = [bool(torch.rand((), generator=g) > 0.5) for _ in range(3)]
    if moments is None:
        moments = torch.empty(C, 2, dtype=torch.float64, device=x.device)
        ops._call("b3d_channel_moments", x, moments)
    xo = torch.empty(tuple(crop_size) + (C,), dtype=torch.float32, device=x.device)
    yo = torch.empty(tuple(crop_size) + (out_ch,), dtype=torch.float32, device=x.device)
    bits = sum(1 << a for a in range(3) if flips[a])
    ops._call("b3d_augment_crop", x, y, moments, shift, scale, xo, yo, int(offset[0]), int(offset[1]), int(offset[2]),
              bits)
    return xo, yo


class VolumeDataset:
    """Device-resident stand-in for `prepare_dataset` (train.py:12-68): a list of preprocessed (x, y) volumes,
    optionally shuffled per epoch, each turned into a random crop by `parse_example` and stacked into batches."""

    def __init__(self, volumes, batch_size, crop_size, out_ch, shuffle=True, seed=0):
        self.volumes = [(x.contiguous(), y.contiguous()) for x, y in volumes]
        self.batch_size, self.crop_size, self.out_ch, self.shuffle = batch_size, list(crop_size), out_ch, shuffle
        self.gen = torch.Generator().manual_seed(seed)
        self._moments = [None] * len(self.volumes)

    def __len__(self):
        return len(self.volumes)

    def __iter__(self):
        order = torch.randperm(len(self.volumes), generator=self.gen).tolist() if self.shuffle \
            else list(range(len(self.volumes)))
        for i in range(0, len(order), self.batch_size):
            xs, ys = [], []
            for j in order[i:i + self.batch_size]:
                x, y = self.volumes[j]
                if self._moments[j] is None:
                    m = torch.empty(x.shape[-1], 2, dtype=torch.float64, device=x.device)
                    ops._call("b3d_channel_moments", x, m)
                    self._moments[j] = m
                xo, yo = parse_example(x, y, self.crop_size, self.out_ch, self.gen, moments=self._moments[j])
                xs.append(xo)
                ys.append(yo)
            yield torch.stack(xs), torch.stack(ys)
