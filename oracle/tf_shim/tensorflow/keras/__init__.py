"""ORACLE / test infrastructure — torch-CPU stand-in for the Keras symbols used by the reference's
hot-path files.  See ../__init__.py.  Semantics restated from SURVEY.md App. B."""
import math
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

import tensorflow as tf


def _mod(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


layers = _mod("tensorflow.keras.layers")
models = _mod("tensorflow.keras.models")
initializers = _mod("tensorflow.keras.initializers")
regularizers = _mod("tensorflow.keras.regularizers")
constraints = _mod("tensorflow.keras.constraints")
optimizers = _mod("tensorflow.keras.optimizers")

_rng = np.random.default_rng(7)


# ------------------------------------------------------------------ initializers etc.
def _init_get(x):
    return x


initializers.get = _init_get
initializers.serialize = lambda x: x
constraints.get = lambda x: x
constraints.serialize = lambda x: x


class _L2:
    def __init__(self, l=0.01):
        self.l = l

    def __call__(self, w):
        return self.l * (tf._t(w) ** 2).sum()        # l * sum(w^2), no 1/2


regularizers.l2 = _L2
regularizers.get = lambda x: x
regularizers.serialize = lambda x: None if x is None else {"l2": x.l}


def _make_weight(shape, initializer, fan_in=None, fan_out=None):
    shape = tuple(int(s) for s in shape)
    if initializer in ("zeros", None):
        a = np.zeros(shape)
    elif initializer == "ones":
        a = np.ones(shape)
    elif initializer == "he_normal":
        a = np.clip(_rng.standard_normal(shape), -2, 2) * math.sqrt(2.0 / fan_in) / 0.87962566
    elif initializer == "glorot_normal":
        a = np.clip(_rng.standard_normal(shape), -2, 2) * math.sqrt(2.0 / (fan_in + fan_out)) / 0.87962566
    elif initializer == "glorot_uniform":
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        a = _rng.uniform(-lim, lim, shape)
    else:
        raise ValueError(initializer)
    t = torch.from_numpy(a).to(tf.DTYPE)
    t.requires_grad_(True)
    return t


# ------------------------------------------------------------------ Layer base
class InputSpec:
    def __init__(self, ndim=None, axes=None):
        self.ndim, self.axes = ndim, axes


def _shape_of(x):
    if isinstance(x, torch.Tensor):
        return list(x.shape)
    if isinstance(x, (list, tuple)):
        return [_shape_of(v) for v in x]
    return None


class Layer:
    def __init__(self, name=None, trainable=True, dtype=None, **kwargs):
        self.name = name or type(self).__name__.lower()
        self.built = False
        self._weights = []      # (tensor_holder_name, regularizer)

    # Keras: build(input_shape) on first __call__, then call()
    def __call__(self, inputs, *args, **kwargs):
        if not self.built:
            self.build(_shape_of(inputs))
            self.built = True
        return self.call(inputs, *args, **kwargs)

    def build(self, input_shape):
        pass

    def add_weight(self, shape=None, name=None, initializer=None, regularizer=None,
                   constraint=None, fan_in=None, fan_out=None, trainable=True):
        w = _make_weight(shape, initializer, fan_in, fan_out)
        self._weights.append([name, w, regularizer])
        return w

    def get_config(self):
        return {"name": self.name, "trainable": True, "dtype": "float32"}

    # --- tracking (attribute order, recursing into nested lists like Keras' ListWrapper)
    def _sublayers(self):
        out = []

        def walk(v):
            if isinstance(v, Layer):
                out.append(v)
            elif isinstance(v, (list, tuple)):
                for u in v:
                    walk(u)

        for k, v in self.__dict__.items():
            if k.startswith("_"):
                continue
            walk(v)
        return out

    def _all_layers(self):
        res = [self]
        for l in self._sublayers():
            res += l._all_layers()
        return res

    @property
    def trainable_variables(self):
        return [w[1] for l in self._all_layers() for w in l._weights]

    @property
    def losses(self):
        return [w[2](w[1]) for l in self._all_layers() for w in l._weights if w[2] is not None]


layers.Layer = Layer
layers.InputSpec = InputSpec


class Model(Layer):
    pass


models.Model = Model


def _activation(name):
    return {None: (lambda x: x), "relu": torch.relu, "sigmoid": torch.sigmoid,
            "linear": (lambda x: x)}[name]


def _same_pads(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def _to_cf(x, data_format):
    return x.permute(0, 4, 1, 2, 3) if data_format == "channels_last" else x


def _from_cf(x, data_format):
    return x.permute(0, 2, 3, 4, 1) if data_format == "channels_last" else x


class Conv3D(Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid", data_format="channels_last",
                 activation=None, use_bias=True, kernel_initializer="glorot_uniform",
                 kernel_regularizer=None, **kw):
        super().__init__(**kw)
        self.filters, self.k, self.s = filters, kernel_size, strides
        assert padding == "same"
        self.data_format, self.act, self.use_bias = data_format, _activation(activation), use_bias
        self.kinit, self.kreg = kernel_initializer, kernel_regularizer

    def build(self, input_shape):
        cin = input_shape[-1] if self.data_format == "channels_last" else input_shape[1]
        k = self.k
        self.kernel = self.add_weight((k, k, k, cin, self.filters), "kernel", self.kinit, self.kreg,
                                      fan_in=k ** 3 * cin, fan_out=k ** 3 * self.filters)
        self.bias = self.add_weight((self.filters,), "bias", "zeros") if self.use_bias else None

    def call(self, x, training=None):
        xc = _to_cf(x, self.data_format)
        pads = []
        for n in reversed(xc.shape[2:]):
            pb, pa = _same_pads(n, self.k, self.s)
            pads += [pb, pa]
        xc = F.pad(xc, pads)
        y = F.conv3d(xc, self.kernel.permute(4, 3, 0, 1, 2), self.bias, stride=self.s)
        return self.act(_from_cf(y, self.data_format))


class Conv3DTranspose(Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid", data_format="channels_last",
                 kernel_initializer="glorot_uniform", **kw):
        super().__init__(**kw)
        assert padding == "same" and strides == 2 and kernel_size == 3
        self.filters, self.data_format, self.kinit = filters, data_format, kernel_initializer

    def build(self, input_shape):
        cin = input_shape[-1] if self.data_format == "channels_last" else input_shape[1]
        self.kernel = self.add_weight((3, 3, 3, self.filters, cin), "kernel", self.kinit,
                                      fan_in=27 * cin, fan_out=27 * self.filters)
        self.bias = self.add_weight((self.filters,), "bias", "zeros")

    def call(self, x, training=None):
        xc = _to_cf(x, self.data_format)
        d, h, w = xc.shape[2:]
        y = F.conv_transpose3d(xc, self.kernel.permute(4, 3, 0, 1, 2), None, stride=2)
        y = y[:, :, :2 * d, :2 * h, :2 * w] + self.bias.view(1, -1, 1, 1, 1)
        return _from_cf(y, self.data_format)


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer="glorot_uniform",
                 kernel_regularizer=None, **kw):
        super().__init__(**kw)
        self.units, self.act, self.use_bias = units, _activation(activation), use_bias
        self.kinit, self.kreg = kernel_initializer, kernel_regularizer

    def build(self, input_shape):
        self.kernel = self.add_weight((input_shape[-1], self.units), "kernel", self.kinit, self.kreg,
                                      fan_in=input_shape[-1], fan_out=self.units)
        self.bias = self.add_weight((self.units,), "bias", "zeros") if self.use_bias else None

    def call(self, x, training=None):
        y = x @ self.kernel
        if self.bias is not None:
            y = y + self.bias
        return self.act(y)


class GlobalAveragePooling3D(Layer):
    def __init__(self, data_format="channels_last", **kw):
        super().__init__(**kw)
        self.data_format = data_format

    def call(self, x):
        return x.mean(dim=(1, 2, 3) if self.data_format == "channels_last" else (2, 3, 4))


class Reshape(Layer):
    def __init__(self, target_shape, **kw):
        super().__init__(**kw)
        self.target_shape = tuple(target_shape)

    def call(self, x):
        return x.reshape((x.shape[0],) + self.target_shape)


class Flatten(Layer):
    def __init__(self, data_format=None, **kw):
        super().__init__(**kw)
        self.data_format = data_format

    def call(self, x):
        if self.data_format == "channels_first":
            x = x.permute(0, 2, 3, 4, 1)
        return x.reshape(x.shape[0], -1)


class Multiply(Layer):
    def call(self, xs):
        return xs[0] * xs[1]


class Add(Layer):
    def call(self, xs):
        return xs[0] + xs[1]


class Concatenate(Layer):
    def __init__(self, axis=-1, **kw):
        super().__init__(**kw)
        self.axis = axis

    def call(self, xs):
        return torch.cat(list(xs), dim=self.axis)


class Activation(Layer):
    def __init__(self, activation, **kw):
        super().__init__(**kw)
        self.act = _activation(activation)

    def call(self, x):
        return self.act(x)


class Lambda(Layer):
    def __init__(self, fn, **kw):
        super().__init__(**kw)
        self.fn = fn

    def call(self, x):
        return self.fn(x)


class Dropout(Layer):
    # test hook: mask to use when training (deterministic parity); None => draw
    injected_mask = None

    def __init__(self, rate, **kw):
        super().__init__(**kw)
        self.rate = rate

    def call(self, x, training=None):
        if not training:
            return x
        m = Dropout.injected_mask
        if m is None:
            m = (torch.rand(x.shape, dtype=x.dtype) >= self.rate).to(x.dtype)
        return x * m / (1.0 - self.rate)


class MaxPooling3D(Layer):
    def __init__(self, pool_size=2, strides=2, padding="same", data_format="channels_last", **kw):
        super().__init__(**kw)
        self.data_format = data_format

    def call(self, x):
        xc = _to_cf(x, self.data_format)
        pads = []
        for n in reversed(xc.shape[2:]):
            pads += [0, n % 2]
        xc = F.pad(xc, pads, value=float("-inf"))
        return _from_cf(F.max_pool3d(xc, 2, 2), self.data_format)


class UpSampling3D(Layer):
    def __init__(self, size=2, data_format="channels_last", **kw):
        super().__init__(**kw)
        self.data_format, self.size = data_format, size

    def call(self, x):
        xc = _to_cf(x, self.data_format)
        for d in (2, 3, 4):
            xc = xc.repeat_interleave(self.size, dim=d)
        return _from_cf(xc, self.data_format)


for _c in (Conv3D, Conv3DTranspose, Dense, GlobalAveragePooling3D, Reshape, Flatten, Multiply, Add,
           Concatenate, Activation, Lambda, Dropout, MaxPooling3D, UpSampling3D):
    setattr(layers, _c.__name__, _c)


# ------------------------------------------------------------------ optimizer (TF Adam)
class Adam:
    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7, amsgrad=False,
                 name="Adam", **kw):
        self._hyper = {"learning_rate": learning_rate}
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon
        self.iterations = 0
        self._slots = {}

    def _set_hyper(self, k, v):
        self._hyper[k] = float(v)

    @property
    def learning_rate(self):
        return torch.as_tensor(self._hyper["learning_rate"])

    def apply_gradients(self, grads_and_vars):
        self.iterations += 1
        t = self.iterations
        lr = float(self._hyper["learning_rate"])
        alpha = lr * math.sqrt(1 - self.beta_2 ** t) / (1 - self.beta_1 ** t)
        with torch.no_grad():
            for g, v in grads_and_vars:
                m, s = self._slots.setdefault(id(v), (torch.zeros_like(v), torch.zeros_like(v)))
                m.mul_(self.beta_1).add_(g, alpha=1 - self.beta_1)
                s.mul_(self.beta_2).addcmul_(g, g, value=1 - self.beta_2)
                v.sub_(alpha * m / (torch.sqrt(s) + self.epsilon))


optimizers.Adam = Adam
