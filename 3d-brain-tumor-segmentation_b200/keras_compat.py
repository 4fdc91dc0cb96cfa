"""Host-side stand-in for the small part of the Keras `Layer` protocol the reference's hot path relies on:
lazy `build` on first call, `trainable_variables` / `losses` collected in attribute order,
`get_config()`, Keras initialisers and `l2` regularisers.
Weights are torch CUDA tensors in Keras layouts; nothing here does arithmetic on the hot path.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch

_GEN = torch.Generator().manual_seed(20260117)


def set_seed(seed: int):
    _GEN.manual_seed(seed)


def _trunc_normal(shape, std):
    # Keras he_normal / glorot_normal: truncated normal (|z| <= 2), stddev / 0.87962566 (App. B)
    t = torch.empty(shape, dtype=torch.float32)
    torch.nn.init.trunc_normal_(t, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=_GEN)
    return t * (std / 0.87962566103423978)


def _fans(shape):
    if len(shape) == 2:
        return shape[0], shape[1]
    rf = 1
    for s in shape[:-2]:
        rf *= s
    return shape[-2] * rf, shape[-1] * rf


def make_initial_value(shape, initializer):
    shape = tuple(int(s) for s in shape)
    if initializer in (None, "zeros"):
        return torch.zeros(shape)
    if initializer == "ones":
        return torch.ones(shape)
    fan_in, fan_out = _fans(shape)
    if initializer == "he_normal":
        return _trunc_normal(shape, math.sqrt(2.0 / fan_in))
    if initializer == "glorot_normal":
        return _trunc_normal(shape, math.sqrt(2.0 / (fan_in + fan_out)))
    if initializer == "glorot_uniform":
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        return (torch.rand(shape, generator=_GEN) * 2 - 1) * lim
    raise ValueError(f"unknown initializer {initializer!r}")


class L2:
    """tf.keras.regularizers.l2(l): l * sum(w^2)."""

    def __init__(self, l=0.01):
        self.l = float(l)


class Variable:
    """A named trainable tensor + its regulariser (what Keras' add_weight returns)."""
    __slots__ = ("name", "tensor", "regularizer")

    def __init__(self, name, tensor, regularizer):
        self.name, self.tensor, self.regularizer = name, tensor, regularizer


import threading

_LAYOUT = threading.local()


def _internal_layout() -> bool:
    return getattr(_LAYOUT, "depth", 0) > 0


class internal_layout:
    """Inside: tensors are in the internal NDHWC storage even when the layers were built with
    data_format='channels_first' (only the outermost public call converts NCDHW <-> NDHWC)."""

    def __enter__(self):
        _LAYOUT.depth = getattr(_LAYOUT, "depth", 0) + 1

    def __exit__(self, *a):
        _LAYOUT.depth -= 1
        return False


def map5d(v, fn):
    if isinstance(v, torch.Tensor):
        return fn(v) if v.dim() == 5 else v
    if isinstance(v, (list, tuple)):
        return type(v)(map5d(u, fn) for u in v)
    return v


def check_data_format(data_format):
    if data_format not in ('channels_last', 'channels_first'):
        raise ValueError(f"data_format must be 'channels_last' or 'channels_first', got {data_format!r}")
    return data_format


# Keras names layers `snake_case(class)` + `_<n>` with one process-wide counter per class, starting without a suffix
# (tf.keras.backend.unique_object_name, zero-based): 'conv3d', 'conv3d_1', ... — what the weight names of a Keras
# checkpoint are made of.  `reset_uids()` is tf.keras.backend.clear_session()'s effect on them.
import collections as _collections
import re as _re

_UIDS = _collections.defaultdict(int)
# the Keras class each host class stands for (names are derived from the reference's class names)
_KERAS_CLASS = {"LinearUpsample": "LinearUpsample", "MaxDownsample": "MaxDownsample"}


def to_snake_case(name: str) -> str:
    """tf.python.keras.utils.generic_utils.to_snake_case."""
    intermediate = _re.sub("(.)([A-Z][a-z0-9]+)", r"\1_\2", name)
    insecure = _re.sub("([a-z])([A-Z])", r"\1_\2", intermediate).lower()
    return insecure if insecure[0] != "_" else "private" + insecure


def unique_layer_name(cls_name: str) -> str:
    base = to_snake_case(cls_name)
    n = _UIDS[base]
    _UIDS[base] += 1
    return base if n == 0 else f"{base}_{n}"


def reset_uids():
    _UIDS.clear()


class Layer:
    def __init__(self, name: Optional[str] = None, **kwargs):
        self.name = name or unique_layer_name(_KERAS_CLASS.get(type(self).__name__, type(self).__name__))
        self.built = False
        self._vars: List[Variable] = []

    # ---- Keras protocol
    def __call__(self, inputs, *args, **kwargs):
        # data_format='channels_first': the public surface takes / returns NCDHW tensors; storage inside is NDHWC
        # (the semantic difference — true channel GroupNorm — lives in the kernels, not in the layout)
        convert = getattr(self, "data_format", "channels_last") == "channels_first" and not _internal_layout()
        if convert:
            inputs = map5d(inputs, ops.to_channels_last)
            kwargs = {k: map5d(v, ops.to_channels_last) for k, v in kwargs.items()}
        with internal_layout():
            if not self.built:
                self.build(_shape_of(inputs), _device_of(inputs))
                self.built = True
            out = self.call(inputs, *args, **kwargs)
        return map5d(out, ops.to_channels_first) if convert else out

    def build(self, input_shape, device):
        pass

    def call(self, inputs, training=None):
        raise NotImplementedError

    def add_weight(self, name, shape, initializer, device, regularizer: Optional[L2] = None):
        t = make_initial_value(shape, initializer).to(device).requires_grad_(True)
        self._vars.append(Variable(name, t, regularizer))
        return t

    def get_config(self):
        return {"name": self.name, "trainable": True, "dtype": "float32"}

    # ---- tracking (attribute order; recurses into nested lists like Keras' ListWrapper)
    def _sublayers(self) -> List["Layer"]:
        out: List[Layer] = []

        def walk(v):
            if isinstance(v, Layer):
                out.append(v)
            elif isinstance(v, (list, tuple)):
                for u in v:
                    walk(u)

        for k, v in self.__dict__.items():
            if not k.startswith("_"):
                walk(v)
        return out

    def _all_layers(self) -> List["Layer"]:
        res = [self]
        for l in self._sublayers():
            res += l._all_layers()
        return res

    def variables(self) -> List[Variable]:
        return [v for l in self._all_layers() for v in l._vars]

    @property
    def trainable_variables(self) -> List[torch.Tensor]:
        return [v.tensor for v in self.variables()]


def _shape_of(x):
    if isinstance(x, torch.Tensor):
        return list(x.shape)
    if isinstance(x, (list, tuple)):
        return [_shape_of(v) for v in x]
    return None


def _device_of(x):
    if isinstance(x, torch.Tensor):
        return x.device
    if isinstance(x, (list, tuple)):
        for v in x:
            d = _device_of(v)
            if d is not None:
                return d
    return None


# ----------------------------------------------------------------------------------------------
# The Keras built-ins the reference composes, as thin weight holders over b3d kernels.
# ----------------------------------------------------------------------------------------------
from . import ops  # noqa: E402


def _require_channels_last(data_format):      # validates only: both layouts are served (Layer.__call__)
    check_data_format(data_format)


class Conv3D(Layer):
    """tf.keras.layers.Conv3D(padding='same'); kernel (k,k,k,Cin,Cout)."""

    def __init__(self, filters, kernel_size, strides=1, padding="same", data_format="channels_last",
                 activation=None, use_bias=True, kernel_initializer="glorot_uniform", kernel_regularizer=None,
                 **kw):
        super().__init__(**kw)
        self.data_format = check_data_format(data_format)
        if padding != "same":
            raise NotImplementedError("b3d Conv3D: only padding='same'")
        if activation not in (None, "sigmoid"):
            raise NotImplementedError("b3d Conv3D: activation must be None or 'sigmoid'")
        self.filters, self.kernel_size, self.strides = filters, kernel_size, strides
        self.activation, self.use_bias = activation, use_bias
        self.kernel_initializer, self.kernel_regularizer = kernel_initializer, kernel_regularizer
        self.kernel = self.bias = None

    def build(self, input_shape, device):
        k = self.kernel_size
        self.kernel = self.add_weight("kernel", (k, k, k, input_shape[-1], self.filters), self.kernel_initializer,
                                      device, self.kernel_regularizer)
        if self.use_bias:
            self.bias = self.add_weight("bias", (self.filters,), "zeros", device)
        self.built = True

    def call(self, x, training=None, gn_groups=0, want_gap=False, aux=False, share_x=False, grad_box=None,
             kernel=None):
        """kernel: a tensor derived from self.kernel to convolve with instead (the folded kernel of a dense-connection
        input, ops.fold_dup)."""
        y, stats, gap = ops.conv3d(x, self.kernel if kernel is None else kernel, self.bias, self.strides, False,
                                   1 if self.activation == "sigmoid" else 0, gn_groups, want_gap, share_x, grad_box)
        return (y, stats, gap) if aux else y


class Conv3DTranspose(Layer):
    """tf.keras.layers.Conv3DTranspose(kernel 3, strides 2, padding 'same'); kernel (3,3,3,Cout,Cin)."""

    def __init__(self, filters, kernel_size=3, strides=2, padding="same", data_format="channels_last",
                 kernel_initializer="glorot_uniform", **kw):
        super().__init__(**kw)
        self.data_format = check_data_format(data_format)
        if (kernel_size, strides, padding) != (3, 2, "same"):
            raise NotImplementedError("b3d Conv3DTranspose: only kernel_size=3, strides=2, padding='same'")
        self.filters, self.kernel_initializer = filters, kernel_initializer
        self.kernel = self.bias = None

    def build(self, input_shape, device):
        self.kernel = self.add_weight("kernel", (3, 3, 3, self.filters, input_shape[-1]), self.kernel_initializer,
                                      device)
        self.bias = self.add_weight("bias", (self.filters,), "zeros", device)
        self.built = True

    def call(self, x, training=None, gn_groups=0, aux=False):
        y, stats, _ = ops.conv3d(x, self.kernel, self.bias, 2, True, 0, gn_groups, False)
        return (y, stats) if aux else y


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer="glorot_uniform",
                 kernel_regularizer=None, **kw):
        super().__init__(**kw)
        if activation not in (None, "relu"):
            raise NotImplementedError("b3d Dense.call: activation must be None or 'relu' "
                                      "(the SE denses are fused into the ResnetBlock epilogue)")
        self.units, self.activation, self.use_bias = units, activation, use_bias
        self.kernel_initializer, self.kernel_regularizer = kernel_initializer, kernel_regularizer
        self.kernel = self.bias = None

    def build(self, input_shape, device):
        self.kernel = self.add_weight("kernel", (input_shape[-1], self.units), self.kernel_initializer, device,
                                      self.kernel_regularizer)
        if self.use_bias:
            self.bias = self.add_weight("bias", (self.units,), "zeros", device)
        self.built = True

    def call(self, x, training=None):
        return ops.dense(x, self.kernel, self.bias, 1 if self.activation == "relu" else 0)
