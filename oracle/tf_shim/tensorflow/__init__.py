"""ORACLE / test infrastructure — NOT a TensorFlow re-implementation and never on the product path.

A torch-CPU backed stand-in for the ~40 TF/Keras symbols that the reference's hot-path
files (model.py, layers/*.py, util.py) touch, so that those files can be imported and
EXECUTED UNMODIFIED from /root/reference to generate golden vectors
(oracle/run_reference.py).  Tensors are plain torch tensors (fp64 by default so that the
result is a high-precision statement of the reference's math).  Built-in op semantics follow
SURVEY.md App. B (TF SAME padding, Conv3DTranspose = adjoint of the strided SAME conv, ...).
"""
import math as _math
import types as _types

import numpy as _np
import torch as _torch
import torch.nn.functional as _F

DTYPE = _torch.float64

float32 = "float32"
int32 = "int32"


def _t(x):
    if isinstance(x, Variable):
        return x.value_
    if isinstance(x, _torch.Tensor):
        return x
    return _torch.as_tensor(x, dtype=DTYPE)


class Variable:
    def __init__(self, initial_value, name=None, trainable=True, dtype=None):
        v = _torch.as_tensor(initial_value)
        if v.dtype.is_floating_point:
            v = v.to(DTYPE)
        self.value_ = v
        self.name = name
        self.trainable = trainable

    def value(self):
        return self.value_

    def numpy(self):
        return self.value_.numpy()

    def assign(self, v):
        self.value_ = _torch.as_tensor(v)


def constant(v, dtype=None):
    return _torch.as_tensor(v, dtype=DTYPE)


def convert_to_tensor(v, dtype=None):
    return _torch.as_tensor(v)


def reshape(x, shape):
    return _t(x).reshape([int(s) for s in shape])


def stack(vals, axis=0):
    return [int(v) for v in vals]


def cast(x, dtype):
    return x.to(DTYPE) if dtype == float32 else x.to(_torch.int64)


def reduce_mean(x, axis=None, keepdims=False):
    x = _t(x)
    return x.mean() if axis is None else x.mean(dim=axis, keepdim=keepdims)


def reduce_sum(x, axis=None, keepdims=False):
    if isinstance(x, (list, tuple)):
        x = _torch.stack([_t(v) for v in x]) if len(x) else _torch.zeros((), dtype=DTYPE)
    return x.sum() if axis is None else x.sum(dim=axis, keepdim=keepdims)


def reduce_max(x, axis=None, keepdims=False):
    return x.max() if axis is None else x.max(dim=axis, keepdim=keepdims).values


def argmax(x, axis=None, output_type=None):
    return x.argmax(dim=axis)


def one_hot(idx, depth, axis=-1, dtype=None):
    oh = _F.one_hot(idx, depth).to(DTYPE)
    if axis not in (-1, oh.dim() - 1):
        oh = oh.movedim(-1, axis)
    return oh


def reverse(x, axis):
    return _torch.flip(x, dims=list(axis)) if len(axis) else x


def concat(xs, axis=0):
    return _torch.cat(list(xs), dim=axis)


def expand_dims(x, axis):
    return x.unsqueeze(axis)


def squeeze(x, axis=None):
    return x.squeeze(axis)


def zeros(shape, dtype=None):
    return _torch.zeros(list(shape), dtype=DTYPE)


def transpose(x, perm):
    return _t(x).permute(*[int(a) for a in perm])


def pad(x, paddings, mode='CONSTANT', constant_values=0.0):
    """tf.pad, CONSTANT mode: paddings[d] = [before, after] per dimension (used by the reference's test.py:164-178)."""
    assert mode == 'CONSTANT'
    x = _t(x)
    flat = []
    for before, after in reversed([[int(a), int(b)] for a, b in paddings]):     # F.pad lists the LAST dim first
        flat += [before, after]
    return _F.pad(x, flat, mode='constant', value=float(constant_values))


class _Math:
    sqrt = staticmethod(_torch.sqrt)
    exp = staticmethod(_torch.exp)


math = _Math()
sqrt = _torch.sqrt


class _NN:
    @staticmethod
    def moments(x, axes, keepdims=False):
        mean = x.mean(dim=tuple(axes), keepdim=True)
        var = ((x - mean) ** 2).mean(dim=tuple(axes), keepdim=True)   # population variance
        if not keepdims:
            mean = mean.squeeze(tuple(axes))
            var = var.squeeze(tuple(axes))
        return mean, var


nn = _NN()


class _Random:
    _gen = _torch.Generator().manual_seed(1234)
    # test hook: when set, normal() returns this tensor instead of drawing (deterministic parity)
    injected_normal = None

    def normal(self, shape, dtype=None):
        if self.injected_normal is not None:
            return self.injected_normal
        return _torch.randn(list(shape), generator=self._gen, dtype=DTYPE)

    # test hook: a list of tensors handed out in call order by uniform() (the data pipeline's draws)
    injected_uniform = None

    def uniform(self, shape, lo=0.0, hi=1.0):
        if self.injected_uniform is not None:
            v = _torch.as_tensor(self.injected_uniform.pop(0), dtype=DTYPE)
            assert list(v.shape) == list(shape), (v.shape, shape)
            return v
        return lo + (hi - lo) * _torch.rand(list(shape), generator=self._gen, dtype=DTYPE)


random = _Random()


def cond(pred, true_fn, false_fn):
    return true_fn() if bool(pred) else false_fn()


def split(x, sizes, axis=0):
    return list(_torch.split(_t(x), [int(v) for v in sizes], dim=axis))


class _Image:
    # test hook: crop offset used instead of a random one
    injected_offset = None

    def random_crop(self, x, size):
        size = [int(v) for v in size]
        off = self.injected_offset
        if off is None:
            off = [int(_torch.randint(0, x.shape[i] - size[i] + 1, (1,), generator=random._gen)) for i in range(len(size))]
        off = list(off) + [0] * (len(size) - len(off))
        return x[tuple(slice(o, o + n) for o, n in zip(off, size))]


image = _Image()


class _IO:
    """Stand-in for the TFRecord proto parsing of train.py:15,49-51: a 'serialized example' is already a dict of
    flat tensors, parse_single_example hands its entries back."""

    class FixedLenFeature:
        def __init__(self, shape, dtype):
            self.shape, self.dtype = shape, dtype

    @staticmethod
    def parse_single_example(proto, desc):
        out = {}
        for k, f in desc.items():
            v = _t(proto[k]).reshape(-1)
            assert v.numel() == int(f.shape[0]), (k, v.numel(), f.shape)
            out[k] = v
        return out


io = _IO()


class _Dataset:
    def __init__(self, records, fn=None, batch=None):
        self.records, self.fn, self.nbatch = records, fn, batch

    def shuffle(self, buffer_size):
        return self

    def map(self, map_func, num_parallel_calls=None):
        return _Dataset(self.records, map_func, self.nbatch)

    def batch(self, batch_size):
        return _Dataset(self.records, self.fn, int(batch_size))

    def prefetch(self, buffer_size):
        return self

    def __iter__(self):
        items = [self.fn(r) if self.fn else r for r in self.records]
        n = self.nbatch or 1
        for i in range(0, len(items), n):
            chunk = items[i:i + n]
            yield tuple(_torch.stack([c[j] for c in chunk]) for j in range(len(chunk[0])))


class _Data:
    # test hook: the "file contents" every TFRecordDataset yields, one dict per file
    injected_records = None

    class experimental:
        AUTOTUNE = -1

    def TFRecordDataset(self, files):
        assert self.injected_records is not None and len(self.injected_records) == len(files)
        return _Dataset(list(self.injected_records))


data = _Data()

from . import keras  # noqa: E402,F401
