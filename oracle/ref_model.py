"""ORACLE — test infrastructure only (never imported by the product path).

CPU restatement (torch-CPU used purely as a math library, fp64 by default) of the
training / inference hot path of vliu15/3d-brain-tumor-segmentation.  Every function
cites the reference file:line it follows (paths relative to /root/reference).

Parity status: the reference's real runtime (tensorflow==2.0.0-alpha0, requirements.txt:2)
cannot be installed here, and the reference ships no tests or golden vectors.  The
restatement is therefore pinned in two ways instead:
  1. `oracle/run_reference.py` executes the reference's OWN python files (model.py,
     layers/*.py, util.py) on top of `oracle/tf_shim` (a tiny torch-backed stand-in for
     the handful of TF/Keras symbols they use) and `tests/test_oracle_vs_reference.py`
     checks this file against it -> first-party math (GroupNormalization.call, block /
     encoder / decoder / VAE wiring, losses) is pinned on the reference's own code;
  2. third-party TF op semantics (SAME padding, Conv3DTranspose, Adam) are restated
     from their published definitions => "parity unpinned" for those.

All tensors are channels_last  [B, D, H, W, C];  weights are in Keras layouts:
  Conv3D kernel (kd,kh,kw,Cin,Cout); Conv3DTranspose kernel (kd,kh,kw,Cout,Cin);
  Dense kernel (in,out).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------------
# TF/Keras built-in semantics (third-party; restated from SURVEY.md App. B)
# ----------------------------------------------------------------------------------
def _same_pads(n: int, k: int, s: int) -> Tuple[int, int]:
    """TF 'SAME' padding rule: out=ceil(n/s); total=max((out-1)*s+k-n,0); before=total//2."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    before = total // 2
    return before, total - before


# Operand-rounding emulation ("error model" of the mixed-precision tensor-core path, used only to DERIVE the tolerances
# of the GPU parity tests — tests/test_gpu_baseline_shapes.py): inside `operand_rounding(fwd, bwd, wgrad)` every conv that
# the CUDA path runs on the tensor cores rounds its two operands to the stated type before the (fp32/fp64) contraction:
# forward x,w -> `fwd`; data gradient dy,w -> `bwd`; weight gradient x,dy -> `wgrad`.  Accumulation, bias and everything
# that is not a conv stay in the run's dtype.  With all three None the functions below are the plain restatement.
_ROUND = {"fwd": None, "bwd": None, "wgrad": None}


class operand_rounding:
    def __init__(self, fwd="fp16", bwd="bf16", wgrad="bf16"):
        self.new = {"fwd": fwd, "bwd": bwd, "wgrad": wgrad}

    def __enter__(self):
        self.prev = dict(_ROUND)
        _ROUND.update(self.new)

    def __exit__(self, *a):
        _ROUND.update(self.prev)
        return False


def round_operand(t: torch.Tensor, kind: Optional[str]) -> torch.Tensor:
    """Round-to-nearest-even to the significand / range of `kind` ('fp16' saturating, 'bf16', 'tf32'), same dtype out."""
    if kind is None:
        return t
    if kind == "fp16":
        return t.clamp(-65504.0, 65504.0).to(torch.float16).to(t.dtype)
    if kind == "bf16":
        return t.to(torch.bfloat16).to(t.dtype)
    if kind == "tf32":                     # 10 explicit significand bits, fp32 exponent (cvt.rna: ties away; rn here)
        i = t.to(torch.float32).contiguous().view(torch.int32)
        i = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
        return i.view(torch.float32).to(t.dtype)
    raise ValueError(kind)


class _RoundedConv(torch.autograd.Function):
    """y = fn(rnd_fwd(x), rnd_fwd(w)) + b;  dx from (rnd_bwd(dy), rnd_bwd(w));  dw from (rnd_wg(x), rnd_wg(dy))."""

    @staticmethod
    def forward(ctx, x, w, b, fn):
        ctx.fn, ctx.kinds, ctx.has_b = fn, dict(_ROUND), b is not None
        ctx.save_for_backward(x, w)
        y = fn(round_operand(x, _ROUND["fwd"]), round_operand(w, _ROUND["fwd"]))
        return y if b is None else y + b

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        k = ctx.kinds
        dx = dw = db = None
        with torch.enable_grad():
            if ctx.needs_input_grad[0]:
                x1 = torch.zeros_like(x).requires_grad_(True)          # the conv is linear in x
                dx, = torch.autograd.grad(ctx.fn(x1, round_operand(w, k["bwd"])), x1, round_operand(dy, k["bwd"]))
            if ctx.needs_input_grad[1]:
                w1 = torch.zeros_like(w).requires_grad_(True)
                dw, = torch.autograd.grad(ctx.fn(round_operand(x, k["wgrad"]), w1), w1, round_operand(dy, k["wgrad"]))
        if ctx.has_b and ctx.needs_input_grad[2]:
            db = dy.sum(dim=(0, 1, 2, 3))
        return dx, dw, db, None


def _on_tensor_cores(cin: int, cout: int, stride2: bool) -> bool:
    """Which convs the CUDA path runs with 16-bit operands (csrc/conv_tc.cu tc_conv_supported): all of them except the
    two tiny stride-2-family layers of the VAE bottleneck (16^3x512 -> 8^3x8, 8^3x1 -> 16^3x128: fp32 CUDA cores)."""
    return not (stride2 and min(cin, cout) < 16)


def _conv3d_same_raw(x, w, stride):
    k = w.shape[0]
    xc = x.permute(0, 4, 1, 2, 3)
    pads = []
    for n in reversed(xc.shape[2:]):  # F.pad wants last dim first
        pb, pa = _same_pads(n, k, stride)
        pads += [pb, pa]
    xc = F.pad(xc, pads)
    wc = w.permute(4, 3, 0, 1, 2)
    return F.conv3d(xc, wc, None, stride=stride).permute(0, 2, 3, 4, 1)


def conv3d_same(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], stride: int = 1) -> torch.Tensor:
    """tf.keras.layers.Conv3D(padding='same') on NDHWC input (call sites resnet.py:80-87,
    downsample.py:28-35, decoder.py:55-63, vae.py:92-99).  Cross-correlation, no flip."""
    if any(v is not None for v in _ROUND.values()) and _on_tensor_cores(w.shape[3], w.shape[4], stride == 2):
        return _RoundedConv.apply(x, w, b, lambda a, k: _conv3d_same_raw(a, k, stride))
    y = _conv3d_same_raw(x, w, stride)
    return y if b is None else y + b


def _conv3d_transpose_raw(x, w):
    xc = x.permute(0, 4, 1, 2, 3)
    wc = w.permute(4, 3, 0, 1, 2)  # (Cin, Cout, kd,kh,kw) as conv_transpose3d expects
    y = F.conv_transpose3d(xc, wc, None, stride=2, padding=0)
    d, h, ww = x.shape[1:4]
    return y[:, :, : 2 * d, : 2 * h, : 2 * ww].permute(0, 2, 3, 4, 1)


def conv3d_transpose_same(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    """tf.keras.layers.Conv3DTranspose(kernel 3, strides 2, padding 'same') (upsample.py:28-33):
    the exact adjoint (conv3d_backprop_input) of the SAME/stride-2 conv with filter
    (kd,kh,kw,Cout,Cin); output spatial = 2*in."""
    if any(v is not None for v in _ROUND.values()) and _on_tensor_cores(w.shape[4], w.shape[3], True):
        return _RoundedConv.apply(x, w, b, _conv3d_transpose_raw)
    y = _conv3d_transpose_raw(x, w)
    return y if b is None else y + b


def dense(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    y = x @ w
    return y if b is None else y + b


# ----------------------------------------------------------------------------------
# layers/group_norm.py:83-124
# ----------------------------------------------------------------------------------
# data_format='channels_first' semantics on channels_last-stored tensors: the reference then builds its layers with
# GroupNormalization(axis=1) (resnet.py:89, downsample.py:36, upsample.py:34), i.e. TRUE channel groups, while every
# other op is layout-agnostic.  The oracle keeps its NDHWC storage; inside `channels_first_semantics()` group_norm
# evaluates the literal restatement on the NCDHW view with axis=1.
_CF = {"on": False}


class channels_first_semantics:
    def __enter__(self):
        self.prev = _CF["on"]
        _CF["on"] = True

    def __exit__(self, *a):
        _CF["on"] = self.prev
        return False


def group_norm(x: torch.Tensor, gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor],
               groups: int = 8, eps: float = 1e-5, axis: int = -1) -> torch.Tensor:
    """Literal restatement of GroupNormalization.call: raw reshape of [B,...,C] to
    [B,G,...,C/G] (no transpose, group_norm.py:84-100), moments over axes 2.. (:105),
    division by sqrt(var+eps) (:107), gamma/beta reshaped to broadcast_shape (:114-122)."""
    if _CF["on"] and x.dim() == 5 and axis in (-1, 4):
        out = _group_norm_literal(x.permute(0, 4, 1, 2, 3).contiguous(), gamma, beta, groups, eps, 1)
        return out.permute(0, 2, 3, 4, 1).contiguous()
    return _group_norm_literal(x, gamma, beta, groups, eps, axis)


def _group_norm_literal(x, gamma, beta, groups, eps, axis):
    shape = list(x.shape)
    nd = len(shape)
    ax = axis % nd
    broadcast_shape = [1] * nd
    broadcast_shape[ax] = shape[ax] // groups
    broadcast_shape.insert(1, groups)
    group_axes = list(shape)
    group_axes[ax] = shape[ax] // groups
    group_axes.insert(1, groups)
    g = x.reshape([group_axes[0], groups] + group_axes[2:])
    red = tuple(range(2, len(group_axes)))
    mean = g.mean(dim=red, keepdim=True)
    var = ((g - mean) ** 2).mean(dim=red, keepdim=True)
    g = (g - mean) / torch.sqrt(var + eps)
    if gamma is not None:
        g = g * gamma.reshape(broadcast_shape)
    if beta is not None:
        g = g + beta.reshape(broadcast_shape)
    return g.reshape(shape)


def group_norm_chunk(x, gamma, beta, groups=8, eps=1e-5):
    """Independent second derivation of the channels_last behaviour: group g is the
    g-th contiguous 1/G chunk of each sample's flat buffer; affine index j = g*(C/G) + c % (C/G).
    Used only to cross-check group_norm() above."""
    B = x.shape[0]
    C = x.shape[-1]
    cg = C // groups
    flat = x.reshape(B, groups, -1)
    mean = flat.mean(dim=2, keepdim=True)
    var = ((flat - mean) ** 2).mean(dim=2, keepdim=True)
    xh = ((flat - mean) / torch.sqrt(var + eps))
    n = flat.shape[2]
    c = torch.arange(n) % C  # chunk length is a multiple of C/G but maybe not of C
    # element e of chunk g sits at flat offset g*n+e -> channel (g*n+e) % C
    offs = (torch.arange(groups).view(-1, 1) * n + torch.arange(n).view(1, -1)) % C
    j = torch.arange(groups).view(-1, 1) * cg + offs % cg
    out = xh * gamma[j].unsqueeze(0) + beta[j].unsqueeze(0)
    return out.reshape(x.shape)


# ----------------------------------------------------------------------------------
# layers/resnet.py:116-138
# ----------------------------------------------------------------------------------
def resnet_block(p: Params, pre: str, x: torch.Tensor, groups: int = 8) -> torch.Tensor:
    res = conv3d_same(x, p[pre + "ptwise.kernel"], p[pre + "ptwise.bias"])            # :118
    chse = res.mean(dim=(1, 2, 3))                                                     # :121
    chse = torch.relu(dense(chse, p[pre + "dense_relu.kernel"], None))                 # :122
    chse = torch.sigmoid(dense(chse, p[pre + "dense_sigmoid.kernel"], None))           # :123
    chse = chse.reshape(chse.shape[0], 1, 1, 1, -1)                                    # :124
    spse = torch.sigmoid(conv3d_same(res, p[pre + "spatial.kernel"], None))            # :127
    res = res * (spse + chse)                                                          # :130
    h = x
    for i in (1, 2):                                                                   # :133-136
        h = conv3d_same(h, p[pre + f"conv{i}.kernel"], p[pre + f"conv{i}.bias"])
        h = group_norm(h, p[pre + f"gn{i}.gamma"], p[pre + f"gn{i}.beta"], groups)
        h = torch.relu(h)
    return res + h                                                                     # :137


def max_pool2_same(x):
    """layers/downsample.py:58-62, :64-66: MaxPooling3D(pool_size=2, strides=2, padding='same'), channels_last.
    TF 'SAME' pads max(ceil(n/2)*2 - n, 0) after (with -inf); for even sizes there is no padding."""
    xc = x.permute(0, 4, 1, 2, 3)
    pd, ph, pw = [(-n) % 2 for n in x.shape[1:4]]
    if pd or ph or pw:
        xc = torch.nn.functional.pad(xc, (0, pw, 0, ph, 0, pd), value=float("-inf"))
    return torch.nn.functional.max_pool3d(xc, 2, 2).permute(0, 2, 3, 4, 1)


def upsample2_nearest(x):
    """layers/upsample.py:71-73: UpSampling3D(size=2) repeats every voxel 2x2x2 (nearest neighbour)."""
    return x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def conv_downsample(p: Params, pre: str, x, groups=8):
    """layers/downsample.py:41-45 (ConvDownsample); :64-66 (MaxDownsample: no weights => no keys under `pre`)"""
    if pre + "conv.kernel" not in p:
        return max_pool2_same(x)
    x = conv3d_same(x, p[pre + "conv.kernel"], p[pre + "conv.bias"], stride=2)
    x = group_norm(x, p[pre + "norm.gamma"], p[pre + "norm.beta"], groups)
    return torch.relu(x)


def conv_upsample(p: Params, pre: str, x, groups=8):
    """layers/upsample.py:39-43 (ConvUpsample); :75-78 (LinearUpsample: 1x1x1 conv then UpSampling3D)"""
    if pre + "ptwise.kernel" in p:
        return upsample2_nearest(conv3d_same(x, p[pre + "ptwise.kernel"], p[pre + "ptwise.bias"]))
    x = conv3d_transpose_same(x, p[pre + "conv.kernel"], p[pre + "conv.bias"])
    x = group_norm(x, p[pre + "norm.gamma"], p[pre + "norm.beta"], groups)
    return torch.relu(x)


# ----------------------------------------------------------------------------------
# layers/encoder.py:69-101, layers/decoder.py:65-83, layers/vae.py:114-143, model.py:58-71
# ----------------------------------------------------------------------------------
def encoder(p: Params, x, depth=4, groups=8, dropout_mask=None, dropout=0.2):
    if dropout_mask is not None:                       # encoder.py:71 (training only)
        x = x * dropout_mask / (1.0 - dropout)
    residuals = []
    for i in range(depth):
        cache: List[torch.Tensor] = []
        for j in range(i + 1):
            if j > 0:
                x = torch.cat([x] + cache, dim=-1)     # :85  (x is cache[-1]: duplicated, F3)
            x = resnet_block(p, f"enc.L{i}.B{j}.", x, groups)
            cache.append(x)
        if i > 0:
            x = torch.cat(cache, dim=-1)               # :91
        residuals.append(x)
        if i < depth - 1:
            x = conv_downsample(p, f"enc.L{i}.down.", x, groups)   # :98
    return residuals


def decoder(p: Params, x, residuals, depth=4, groups=8):
    for i, residual in zip(range(depth - 2, -1, -1), residuals[::-1]):
        x = conv_upsample(p, f"dec.L{i}.up.", x, groups)           # :72
        x = torch.cat([residual, x], dim=-1)                       # :75
        x = resnet_block(p, f"dec.L{i}.block.", x, groups)         # :78
    x = conv3d_same(x, p["dec.out.kernel"], p["dec.out.bias"])     # :81
    return torch.sigmoid(x)


def vae(p: Params, x, eps, depth=4, groups=8):
    """eps ~ N(0,1) of shape [B, latent] is injected (vae.py:9-13 draws it internally)."""
    x = conv_downsample(p, "vae.down.", x, groups)                 # :116
    B = x.shape[0]
    x = x.reshape(B, -1)                                           # :119 Flatten (channels_last)
    x = dense(x, p["vae.proj.kernel"], p["vae.proj.bias"])         # :120
    latent = x.shape[1] // 2
    z_mean, z_logvar = x[:, :latent], x[:, latent:]                # :123-124
    z = z_mean + torch.exp(0.5 * z_logvar) * eps                   # :13
    x = torch.relu(dense(z, p["vae.unproj.kernel"], p["vae.unproj.bias"]))   # :128
    return x, z_mean, z_logvar


def vae_full(p: Params, bott, eps, depth=4, groups=8):
    d, h, w = bott.shape[1:4]
    x, z_mean, z_logvar = vae(p, bott, eps, depth, groups)
    x = x.reshape(x.shape[0], d // 2, h // 2, w // 2, 1)           # :129, :109-111
    x = conv_upsample(p, "vae.up.", x, groups)                     # :132
    for i in range(depth - 2, -1, -1):                             # :135-138
        x = conv_upsample(p, f"vae.L{i}.up.", x, groups)
        x = resnet_block(p, f"vae.L{i}.block.", x, groups)
    x = conv3d_same(x, p["vae.out.kernel"], p["vae.out.bias"])     # :141
    return x, z_mean, z_logvar


def model_forward(p: Params, x, eps=None, depth=4, groups=8, inference=False,
                  dropout_mask=None, dropout=0.2):
    """model.py:58-71"""
    res = encoder(p, x, depth, groups, dropout_mask, dropout)
    y_pred = decoder(p, res[-1], res[:-1], depth, groups)
    if inference:
        return y_pred, None, None, None
    y_vae, z_mean, z_logvar = vae_full(p, res[-1], eps, depth, groups)
    return y_pred, y_vae, z_mean, z_logvar


# ----------------------------------------------------------------------------------
# util.py:5-57, train.py:145-146
# ----------------------------------------------------------------------------------
def dice_vae_loss(x, y, y_pred, y_vae, z_mean, z_logvar):
    """util.py:13-24 (channels_last: dice axes (0,1,2,3))."""
    l2 = ((x - y_vae) ** 2).mean()
    kld = (z_mean ** 2 + torch.exp(z_logvar) - z_logvar - 1.0).mean()
    ax = (0, 1, 2, 3)
    inter = (y_pred * y).sum(dim=ax)
    pred = (y_pred ** 2).sum(dim=ax)
    true = (y ** 2).sum(dim=ax)
    dice = (1.0 - (2.0 * inter + 1.0) / (pred + true + 1.0)).mean()
    return dice + 0.1 * l2 + 0.1 * kld


def dice_coefficient(y_true, y_pred, data_format="channels_last"):
    """util.py:35-57 on channels_last-stored tensors.  NB for data_format='channels_last' the macro average reduces
    axes (0,1,2) only, so the ratio is [W, C]-shaped before the mean — kept, it is the reference's
    number; for 'channels_first' the reference reduces (0,2,3,4), i.e. all of batch and space (util.py:36)."""
    mask = (y_pred.max(dim=-1, keepdim=True).values > 0.5).to(y_pred.dtype)
    C = y_pred.shape[-1]
    hard = F.one_hot(y_pred.argmax(dim=-1), C).to(y_pred.dtype) * mask
    axes = (0, 1, 2) if data_format == "channels_last" else (0, 1, 2, 3)
    inter = (hard * y_true).sum(dim=axes)
    pred = hard.sum(dim=axes)
    true = y_true.sum(dim=axes)
    macro = ((2.0 * inter + 1.0) / (pred + true + 1.0)).mean()
    micro = (hard * y_true).sum() / (hard.sum() + y_true.sum())
    return macro, micro


def l2_regularized_names(p: Params) -> List[str]:
    """Names of tensors carrying an L2 regulariser in the reference:
    all Conv3D/Dense kernels except ConvUpsample's Conv3DTranspose (upsample.py:28-33 has none);
    GN gamma/beta only inside ResnetBlocks (resnet.py:93-94,109-110); no biases."""
    names = []
    for k in p:
        if k.endswith(".kernel") or k in ("dec.out.kernel", "vae.out.kernel"):
            if ".up.conv." in k:
                continue
            names.append(k)
        elif (".gn1." in k or ".gn2." in k):
            names.append(k)
    return names


def l2_reg(p: Params, l2_scale=1e-5):
    """train.py:146: tf.reduce_sum(model.losses) with tf.keras.regularizers.l2(l) = l*sum(w^2)."""
    tot = 0.0
    for k in l2_regularized_names(p):
        tot = tot + l2_scale * (p[k] ** 2).sum()
    return tot


def adam_step_tf(theta, m, v, g, t: int, lr: float, b1=0.9, b2=0.999, eps=1e-7):
    """tf.keras.optimizers.Adam dense update (util.py:60-78), t = iterations+1."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    alpha = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    theta.sub_(alpha * m / (torch.sqrt(v) + eps))


def poly_lr(epoch: int, init_lr=1e-4, n_epochs=300.0):
    """util.py:82-84"""
    return init_lr * ((1.0 - epoch / n_epochs) ** 0.9)


def pad_to_spatial_res(res: int, x):
    """test.py:164-178 — trailing zero pad; adds a full `res` when already aligned (App. C)."""
    shape = x.shape[:-1]
    pad = [res - (s % res) for s in shape]
    out = torch.zeros([s + q for s, q in zip(shape, pad)] + [x.shape[-1]], dtype=x.dtype)
    out[: shape[0], : shape[1], : shape[2]] = x
    return out, list(shape)


def tta_inference(p: Params, x, bmask, mean, std, depth=4, groups=8):
    """test.py:105-161 with spatial_tta=True, channel_tta=0: normalise, 8 flip subsets (:96-101), inference
    forward (:133), un-flip (:134), mean (:147-148), brain mask (:151).  x: [D,H,W,C], bmask: [D,H,W,1]."""
    x = (x - mean) / std
    x = x.unsqueeze(0)
    axes = [1, 2, 3]
    augment = [axes, []]
    for a in axes:
        pairs = [b for b in axes if b != a]
        augment += [[a], pairs]
    ys = []
    for flip in augment:
        aug = torch.flip(x, dims=flip) if flip else x
        y = model_forward(p, aug, inference=True, depth=depth, groups=groups)[0]
        ys.append(torch.flip(y, dims=flip) if flip else y)
    y = torch.cat(ys, dim=0).mean(dim=0, keepdim=True)
    return (y * bmask.unsqueeze(0))[0]


# ----------------------------------------------------------------------------------
# Deterministic synthetic weights / inputs live in the neutral top-level module `synthdata`
# (bench.py / smoke() use them without importing oracle/); re-exported here for the tests.
# ----------------------------------------------------------------------------------
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from synthdata import param_shapes, init_params, synth_batch  # noqa: E402,F401


# ----------------------------------------------------------------------------------
# train.py:12-47 (parse_example of the tf.data pipeline), with the random draws as arguments
# ----------------------------------------------------------------------------------
def augment_example(x, y, crop_size, out_ch, shift, scale, offset, flips):
    """x [H,W,D,C], y [H,W,D,1] -> (crop of the intensity-augmented x, one-hot labels without background)."""
    var = x.var(dim=(0, 1, 2), unbiased=False, keepdim=True)            # :20 tf.nn.moments
    x = (x + shift.reshape(1, 1, 1, -1) * torch.sqrt(var)) * scale.reshape(1, 1, 1, -1)   # :23-24
    xy = torch.cat([x, y.to(x.dtype)], dim=-1)                          # :27
    o = offset
    xy = xy[o[0]:o[0] + crop_size[0], o[1]:o[1] + crop_size[1], o[2]:o[2] + crop_size[2]]   # :28 random_crop
    for axis in (0, 1, 2):                                              # :31-35
        if flips[axis]:
            xy = torch.flip(xy, dims=[axis])
    xc, yc = xy[..., :-1], xy[..., -1]                                  # :37
    lab = yc.to(torch.int64)                                            # :41
    onehot = torch.nn.functional.one_hot(lab, out_ch + 1).to(x.dtype)   # :42
    return xc.contiguous(), onehot[..., 1:].contiguous()                # :43
