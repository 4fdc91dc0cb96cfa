"""Deterministic synthetic weights / inputs of SURVEY §8(d) — a NEUTRAL module (numpy + torch tensors as containers):
seeded generators shared by bench.py, __graft_entry__.smoke(), tools/ and the oracle, so that the product arm never
has to import anything under oracle/.  No arithmetic of the hot path lives here.

Names / shapes follow the reference's layer tree in Keras layouts (model.py:9-56, layers/*.py constructors)."""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch

Params = Dict[str, torch.Tensor]


def param_shapes(in_ch=2, out_ch=3, base_filters=16, depth=4, reduction=2, crop=(128, 128, 128),
                 with_vae=True, downsampling="conv", upsampling="conv") -> Dict[str, Tuple[int, ...]]:
    """Shapes of all trainable tensors in Keras layouts, keyed by this repo's names."""
    s: Dict[str, Tuple[int, ...]] = {}

    def block(pre, cin, f):
        s[pre + "ptwise.kernel"] = (1, 1, 1, cin, f)
        s[pre + "ptwise.bias"] = (f,)
        s[pre + "dense_relu.kernel"] = (f, f // reduction)
        s[pre + "dense_sigmoid.kernel"] = (f // reduction, f)
        s[pre + "spatial.kernel"] = (1, 1, 1, f, 1)
        s[pre + "conv1.kernel"] = (3, 3, 3, cin, f)
        s[pre + "conv1.bias"] = (f,)
        s[pre + "gn1.gamma"] = (f,)
        s[pre + "gn1.beta"] = (f,)
        s[pre + "conv2.kernel"] = (3, 3, 3, f, f)
        s[pre + "conv2.bias"] = (f,)
        s[pre + "gn2.gamma"] = (f,)
        s[pre + "gn2.beta"] = (f,)

    def down(pre, cin, f, kind=None):
        if (kind or downsampling) == "max":          # MaxDownsample: no weights, channels kept
            return cin
        s[pre + "conv.kernel"] = (3, 3, 3, cin, f)
        s[pre + "conv.bias"] = (f,)
        s[pre + "norm.gamma"] = (f,)
        s[pre + "norm.beta"] = (f,)
        return f

    def up(pre, cin, f):
        if upsampling == "linear":         # LinearUpsample: 1x1x1 conv (+bias), no norm
            s[pre + "ptwise.kernel"] = (1, 1, 1, cin, f)
            s[pre + "ptwise.bias"] = (f,)
            return
        s[pre + "conv.kernel"] = (3, 3, 3, f, cin)
        s[pre + "conv.bias"] = (f,)
        s[pre + "norm.gamma"] = (f,)
        s[pre + "norm.beta"] = (f,)

    cin = in_ch
    for i in range(depth):
        f = base_filters * 2 ** i
        for j in range(i + 1):
            block(f"enc.L{i}.B{j}.", cin if j == 0 else (j + 1) * f, f)
        cin = f if i == 0 else (i + 1) * f
        if i < depth - 1:
            cin = down(f"enc.L{i}.down.", cin, f)
    bott = cin
    res_ch = [base_filters if i == 0 else (i + 1) * base_filters * 2 ** i for i in range(depth)]
    c = bott
    for i in range(depth - 2, -1, -1):
        f = base_filters * 2 ** i
        up(f"dec.L{i}.up.", c, f)
        block(f"dec.L{i}.block.", res_ch[i] + f, f)
        c = f
    s["dec.out.kernel"] = (1, 1, 1, c, out_ch)
    s["dec.out.bias"] = (out_ch,)
    if with_vae:
        d, h, w = [n // 2 ** (depth - 1) for n in crop]
        # model.py:49-57 does not forward `downsampling` to the VAE: its extra downsample is always the conv variant
        f = down("vae.down.", bott, base_filters // 2, kind="conv")
        flat = (d // 2) * (h // 2) * (w // 2) * f
        s["vae.proj.kernel"] = (flat, base_filters * 2 ** (depth - 1))
        s["vae.proj.bias"] = (base_filters * 2 ** (depth - 1),)
        latent = base_filters * 2 ** (depth - 2)
        s["vae.unproj.kernel"] = (latent, d * h * w // 8)
        s["vae.unproj.bias"] = (d * h * w // 8,)
        up("vae.up.", 1, base_filters * 2 ** (depth - 1))
        c = base_filters * 2 ** (depth - 1)
        for i in range(depth - 2, -1, -1):
            f = base_filters * 2 ** i
            up(f"vae.L{i}.up.", c, f)
            block(f"vae.L{i}.block.", f, f)
            c = f
        s["vae.out.kernel"] = (3, 3, 3, c, in_ch)
        s["vae.out.bias"] = (in_ch,)
    return s


def init_params(shapes: Dict[str, Tuple[int, ...]], seed=2, dtype=torch.float64) -> Params:
    """Synthetic weights of the reference's scale: kernels ~ N(0, 2/fan_in)
    (he-like), GN gamma 1+0.1N (incl. gn2, whose reference init 0 would kill the conv branch, F4),
    beta 0.1N, biases 0.01N."""
    rng = np.random.default_rng(seed)
    p: Params = {}
    for k, shp in shapes.items():
        if k.endswith("kernel"):
            if len(shp) == 5:
                fan_in = shp[0] * shp[1] * shp[2] * shp[3]
                if ".up.conv." in k:  # transpose conv: (k,k,k,Cout,Cin)
                    fan_in = shp[0] * shp[1] * shp[2] * shp[4] / 8.0  # ~27/8 taps hit per output
            else:
                fan_in = shp[0]
            a = rng.standard_normal(shp) * math.sqrt(2.0 / fan_in)
        elif k.endswith("gamma"):
            a = 1.0 + 0.1 * rng.standard_normal(shp)
        elif k.endswith("beta"):
            a = 0.1 * rng.standard_normal(shp)
        else:
            a = 0.01 * rng.standard_normal(shp)
        p[k] = torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
    return p


def synth_batch(shape=(1, 128, 128, 128), in_ch=2, out_ch=3, latent=64, seed=0, dtype=torch.float64):
    """x ~ N(0,1) seed; y: iid labels p=(0.85,0.05,..) one-hot minus background; eps ~ N(0,1);
    dropout mask Bernoulli(0.8)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(shape + (in_ch,))
    pr = [0.85] + [0.15 / out_ch] * out_ch
    lab = np.random.default_rng(seed + 1).choice(out_ch + 1, size=shape, p=pr)
    y = np.stack([(lab == c + 1) for c in range(out_ch)], axis=-1).astype(np.float64)
    eps = np.random.default_rng(seed + 3).standard_normal((shape[0], latent))
    mask = (np.random.default_rng(seed + 4).random(shape + (in_ch,)) < 0.8).astype(np.float64)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
    return t(x), t(y), t(eps), t(mask)
