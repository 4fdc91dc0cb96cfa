"""Whole-volume inference helpers with the reference's names and semantics
(/root/reference/test.py:75-161 TestTimeAugmentor, :164-178 pad_to_spatial_res).

The hot call is `model(aug, training=False, inference=True)` (test.py:133): the fully-convolutional forward
without the VAE branch.  Flips, normalisation, un-flip + mean and the brain mask run as two small fused
kernels around it instead of tf.reverse / concat / reduce_mean.
"""
from __future__ import annotations

import torch

from . import ops


def pad_to_spatial_res(res, x, mask):
    """test.py:164-178.  Trailing zero pad of a channels_last [D,H,W,C] volume (and its mask) to a multiple of
    `res`.  Like the reference, a dimension that is already aligned gets a FULL extra `res` (App. C)."""
    shape = list(x.shape[:-1])
    pad = [res - (s % res) for s in shape]
    orig_shape = list(shape)

    def _pad(t):
        out = torch.zeros([s + p for s, p in zip(shape, pad)] + [t.shape[-1]], dtype=t.dtype, device=t.device)
        out[:shape[0], :shape[1], :shape[2]] = t
        return out

    return _pad(x), _pad(mask), orig_shape


class TestTimeAugmentor(object):
    """Handles full inference on input with test-time augmentation (test.py:75-161)."""
    __test__ = False   # not a pytest class

    def __init__(self, mean, std, model, model_data_format, spatial_tta=True, channel_tta=0, threshold=0.5,
                 group=None):
        """`group` (B200-side addition, not in the reference): a torch.distributed process group — the flips are
        dealt round-robin to its ranks (replica mode: every GPU runs whole-volume forwards, no halo exchange) and
        the partial means are summed with one all-reduce; every rank returns the full result."""
        self.group = group
        if model_data_format != 'channels_last':
            raise NotImplementedError("b3d TestTimeAugmentor: feed channels_last volumes (the flips / mask kernels are "
                                      "NDHWC); a channels_first Model can be wrapped with ops.to_channels_first")
        if channel_tta:
            raise NotImplementedError("b3d: channel_tta (test.py:137-145 is itself broken in the reference, App. C)")
        self.mean, self.std, self.model = mean, std, model
        self.model_data_format = model_data_format
        self.channel_tta, self.threshold = channel_tta, threshold
        self.channel_axis = -1
        self.spatial_axes = [1, 2, 3]
        if spatial_tta:
            # test.py:96-101: [all three], [], then for each axis: [axis], [the other two]
            self.augment_axes = [self.spatial_axes, []]
            for axis in self.spatial_axes:
                pairs = self.spatial_axes.copy()
                pairs.remove(axis)
                self.augment_axes.append([axis])
                self.augment_axes.append(pairs)
        else:
            self.augment_axes = [[]]

    @staticmethod
    def _bits(axes):
        return sum(1 << (a - 1) for a in axes)        # axis 1 (D) -> 1, 2 (H) -> 2, 3 (W) -> 4

    def __call__(self, x, bmask):
        """x: [D,H,W,C] channels_last volume, bmask: [D,H,W,1] brain mask -> [D,H,W,out_ch]."""
        ops._check(x, "x")
        x = x.contiguous()
        mean = torch.as_tensor(self.mean, dtype=torch.float32, device=x.device).reshape(-1).contiguous()
        std = torch.as_tensor(self.std, dtype=torch.float32, device=x.device).reshape(-1).contiguous()
        if mean.numel() == 1:
            mean, std = mean.expand(x.shape[-1]).contiguous(), std.expand(x.shape[-1]).contiguous()
        bmask = bmask.to(torch.float32).contiguous()
        n = len(self.augment_axes)
        rank, world = 0, 1
        if self.group is not None:
            import torch.distributed as dist
            rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
        mine = list(range(rank, n, world))              # this rank's flips
        aug = torch.empty_like(x)
        acc = None
        with torch.no_grad():
            for j, i in enumerate(mine):
                bits = self._bits(self.augment_axes[i])
                ops._call("b3d_flip_normalize", x, mean, std, aug, bits)           # test.py:107,128
                y, *_ = self.model(aug.unsqueeze(0), training=False, inference=True)   # test.py:133
                y = y[0]
                if acc is None:
                    acc = torch.empty_like(y)
                last = world == 1 and j == len(mine) - 1     # single rank: the brain mask rides on the last pass
                ops._call("b3d_flip_accumulate", y, acc, bmask if last else None, bits, 1.0 / n, int(j == 0))
            if world > 1:
                if acc is None:                              # more ranks than flips
                    y0 = self.model(aug.unsqueeze(0) * 0, training=False, inference=True)[0][0]
                    acc = torch.zeros_like(y0)
                dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=self.group)
                ops._call("b3d_mul_scale", acc, bmask.expand_as(acc).contiguous(), acc, 1.0)   # test.py:147-151
        return acc
