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

    def uniform(self, shape, lo=0.0, hi=1.0):
        return lo + (hi - lo) * _torch.rand(list(shape), generator=self._gen, dtype=DTYPE)


random = _Random()

from . import keras  # noqa: E402,F401
