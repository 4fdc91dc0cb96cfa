"""GPU suite — every C-ABI kernel against the oracle (fp64 CPU restatement) on the same seeded inputs.
Tolerances: fp32 CUDA-core kernels rel-L2 <= 1e-5; tcgen05 TF32 kernels rel-L2 <= 2e-3 (north_star)."""
import numpy as np
import pytest
import torch

from oracle import ref_model as R

pytestmark = pytest.mark.gpu
TOL32 = 2e-5
TOL_TF32 = 2e-3     # tcgen05 kind::tf32 conv forward / dgrad (north_star: <= 2e-3 with TF32)
TOL_BF16 = 1e-2     # tcgen05 kind::f16 bf16-operand weight gradient (north_star: <= 1e-2 with bf16)


def rel(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def t64(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float64) * scale


def dev32(t, dev, grad=False):
    return t.to(torch.float32).to(dev).requires_grad_(grad)


CONV_CASES = [
    # (spatial, Cin, Cout, k, stride, transposed)
    ((8, 8, 8), 2, 16, 3, 1, False),
    ((8, 16, 8), 16, 16, 3, 1, False),
    ((4, 16, 16), 32, 16, 3, 1, False),
    ((6, 10, 12), 5, 7, 3, 1, False),
    ((8, 8, 8), 16, 3, 1, 1, False),
    ((8, 8, 8), 24, 16, 1, 1, False),
    ((8, 8, 8), 16, 16, 3, 2, False),
    ((4, 6, 8), 12, 8, 3, 2, False),
    ((4, 4, 4), 16, 8, 3, 2, True),
    ((2, 2, 2), 1, 128, 3, 2, True),
    # stride-2 family on the tcgen05 path (2x2x2 conv over the coarse grid, space-to-depth addressing)
    ((16, 32, 16), 16, 16, 3, 2, False),
    ((12, 20, 36), 64, 32, 3, 2, False),
    ((8, 8, 8), 192, 64, 3, 2, False),
    ((8, 16, 8), 32, 16, 3, 2, True),
    ((6, 10, 18), 64, 32, 3, 2, True),
    ((4, 4, 4), 512, 64, 3, 2, True),
    ((4, 4, 4), 512, 8, 3, 2, False),          # VAE bottleneck: few outputs, long reduction (split-reduction kernel)
    ((4, 16, 8), 64, 32, 3, 1, False),
]


# CUDA-core fp32 | tcgen05 tf32 operands | tcgen05 bf16 operands | tcgen05 fp16 operands (TF32's significand)
MODES = ["fp32", "tf32", "bf16", "fp16"]
MODE_TOL = {"fp32": TOL32, "tf32": TOL_TF32, "bf16": TOL_BF16, "fp16": TOL_TF32}


def set_mode(b3d, mode):
    b3d.ops.USE_TC["on"] = mode != "fp32"
    if mode == "mixed":                    # the default: fp16 forward, bf16 backward
        b3d.ops.set_conv_precision("fp16", "bf16")
    elif mode != "fp32":
        b3d.ops.set_conv_precision(mode, mode)


def reset_mode(b3d):
    b3d.ops.USE_TC["on"] = True
    b3d.ops.set_conv_precision("fp16", "bf16")


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("mode", MODES)
def test_conv_fwd_bwd(b3d, dev, case, mode):
    sp, cin, cout, k, stride, tr = case
    use_tc = mode != "fp32"
    wshape = (k, k, k) + ((cout, cin) if tr else (cin, cout))
    if use_tc and not b3d.ops.tc_supported(torch.empty(wshape, device="meta"), stride, tr, False):
        pytest.skip("shape not on the tcgen05 path")
    set_mode(b3d, mode)
    try:
        B = 2
        x = t64(B, *sp, cin, seed=1)
        w = t64(k, k, k, *((cout, cin) if tr else (cin, cout)), seed=2, scale=0.2)
        bias = t64(cout, seed=3)
        xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, bias))
        yr = R.conv3d_transpose_same(xr, wr, br) if tr else R.conv3d_same(xr, wr, br, stride)
        gy = t64(*yr.shape, seed=4)
        (yr * gy).sum().backward()
        xd, wd, bd = dev32(x, dev, True), dev32(w, dev, True), dev32(bias, dev, True)
        S = yr.shape[1] * yr.shape[2] * yr.shape[3]
        groups = 8 if (S % 8 == 0 and cout % 8 == 0) else 0
        y, stats, gap = b3d.ops.conv3d(xd, wd, bd, stride, tr, 0, groups, stride == 1)
        (y * dev32(gy, dev)).sum().backward()
        tol = MODE_TOL[mode]
        assert rel(y, yr) < tol
        if stride == 1:
            assert rel(gap, yr.sum(dim=(1, 2, 3))) < max(tol, 1e-4)
        if groups:
            ch = yr.detach().reshape(B, groups, -1)
            assert rel(stats, torch.stack([ch.sum(-1), (ch ** 2).sum(-1)], dim=-1)) < max(tol, 1e-4)
        assert rel(xd.grad, xr.grad) < tol
        assert rel(wd.grad, wr.grad) < (TOL_BF16 if use_tc else tol)
        assert rel(bd.grad, br.grad) < TOL32 * 10
    finally:
        reset_mode(b3d)


@pytest.mark.parametrize("shape", [(2, 8, 8, 8, 16), (1, 20, 6, 4, 16), (2, 5, 3, 3, 8), (1, 4, 4, 4, 32),
                                   (2, 2, 2, 2, 64), (1, 16, 16, 16, 128), (3, 1, 1, 1, 8)])
@pytest.mark.parametrize("relu", [False, True])
def test_group_norm(b3d, dev, shape, relu):
    x, ga, be = t64(*shape, seed=5), 1 + 0.3 * t64(shape[-1], seed=6), 0.3 * t64(shape[-1], seed=7)
    xr, gr, br = (t.clone().requires_grad_(True) for t in (x, ga, be))
    yr = R.group_norm(xr, gr, br)
    yr = torch.relu(yr) if relu else yr
    gy = t64(*shape, seed=8)
    (yr * gy).sum().backward()
    xd, gd, bd = dev32(x, dev, True), dev32(ga, dev, True), dev32(be, dev, True)
    y = b3d.ops.group_norm(xd, gd, bd, None, 8, 1e-5, relu)
    (y * dev32(gy, dev)).sum().backward()
    assert rel(y, yr) < TOL32
    if shape[1] * shape[2] * shape[3] * shape[4] // 8 > 1:      # single-element chunks: dx is 0/0-ish noise
        assert rel(xd.grad, xr.grad) < 2e-4
    assert rel(gd.grad, gr.grad) < 1e-4
    assert rel(bd.grad, br.grad) < 1e-4


def test_group_norm_matches_reference_fixture(b3d, dev):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "gn_cases.npz"))
    for i in range(4):
        x, ga, be = (torch.from_numpy(g[f"{n}{i}"]) for n in ("x", "gamma", "beta"))
        y = b3d.ops.group_norm(dev32(x, dev), dev32(ga, dev), dev32(be, dev), None, 8, 1e-5, False)
        assert rel(y, torch.from_numpy(g[f"y{i}"])) < TOL32


def test_group_norm_value_errors(b3d, dev):
    gn = b3d.GroupNormalization(groups=8)
    with pytest.raises(ValueError, match="cannot be more than the number of channels"):
        gn(torch.zeros(1, 2, 2, 2, 4, device=dev))
    gn = b3d.GroupNormalization(groups=8)
    with pytest.raises(ValueError, match="must be a multiple of the number of channels"):
        gn(torch.zeros(1, 2, 2, 2, 12, device=dev))


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("cfg", [((8, 8, 8), 2, 16, 2), ((4, 8, 16), 32, 32, 2), ((4, 4, 4), 64, 128, 8),
                                 ((3, 3, 3), 16, 16, 2), ((4, 4, 8), 48, 24, 2)])
def test_resnet_block(b3d, dev, cfg, mode):
    """ResnetBlock forward + every gradient vs the fp64 oracle.  fp32 mode (CUDA-core convs) pins the fused
    epilogue / GroupNorm / scSE backward formulas tightly; tensor-core mode (TF32 fwd/dgrad, bf16 wgrad)
    is held to north_star's per-layer tolerance on the output and to a looser bound on gradients, which
    are ill-conditioned here (GroupNorm makes sum(dh) ~ 0, so bias/gamma gradients are small residuals)."""
    sp, cin, f, red = cfg
    shapes = {}
    pre = "b."
    full = R.param_shapes()          # borrow the per-block shape recipe
    for k, v in full.items():
        if k.startswith("enc.L0.B0."):
            name = k[len("enc.L0.B0."):]
            shp = list(v)
            shapes[pre + name] = shp
    # rewrite channel dims
    def fix(name):
        return {"ptwise.kernel": (1, 1, 1, cin, f), "ptwise.bias": (f,), "dense_relu.kernel": (f, f // red),
                "dense_sigmoid.kernel": (f // red, f), "spatial.kernel": (1, 1, 1, f, 1),
                "conv1.kernel": (3, 3, 3, cin, f), "conv1.bias": (f,), "gn1.gamma": (f,), "gn1.beta": (f,),
                "conv2.kernel": (3, 3, 3, f, f), "conv2.bias": (f,), "gn2.gamma": (f,), "gn2.beta": (f,)}[name]
    shapes = {pre + n[len(pre):]: fix(n[len(pre):]) for n in shapes}
    p = R.init_params(shapes, seed=9)
    x = t64(2, *sp, cin, seed=10)
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    xr = x.clone().requires_grad_(True)
    yr = R.resnet_block(pr, pre, xr)
    gy = t64(*yr.shape, seed=11)
    (yr * gy).sum().backward()

    use_tc = mode != "fp32"
    set_mode(b3d, mode)
    blk = b3d.ResnetBlock(f, reduction=red)
    xd = dev32(x, dev, True)
    blk(xd.detach())
    names = {"ptwise.kernel": blk.conv3d_ptwise.kernel, "ptwise.bias": blk.conv3d_ptwise.bias,
             "dense_relu.kernel": blk.dense_relu.kernel, "dense_sigmoid.kernel": blk.dense_sigmoid.kernel,
             "spatial.kernel": blk.spatial.kernel,
             "conv1.kernel": blk.convs[0][0].kernel, "conv1.bias": blk.convs[0][0].bias,
             "gn1.gamma": blk.convs[0][1].gamma, "gn1.beta": blk.convs[0][1].beta,
             "conv2.kernel": blk.convs[1][0].kernel, "conv2.bias": blk.convs[1][0].bias,
             "gn2.gamma": blk.convs[1][1].gamma, "gn2.beta": blk.convs[1][1].beta}
    with torch.no_grad():
        for n, t in names.items():
            t.copy_(p[pre + n].to(torch.float32))
    try:
        y = blk(xd)
        (y * dev32(gy, dev)).sum().backward()
        torch.cuda.synchronize()
    finally:
        reset_mode(b3d)
    errs = {n: rel(t.grad, pr[pre + n].grad) for n, t in names.items()}
    print("resnet_block", cfg, mode, "y", rel(y, yr), "dx", rel(xd.grad, xr.grad),
          {k: f"{v:.1e}" for k, v in errs.items()})
    ytol, gtol = {"fp32": (TOL32, 2e-4), "tf32": (TOL_TF32, 5e-2), "bf16": (TOL_BF16, 2e-1),
                  "fp16": (TOL_TF32, 5e-2)}[mode]
    assert rel(y, yr) < ytol
    assert rel(xd.grad, xr.grad) < gtol
    for n, e in errs.items():
        assert e < gtol, (n, e)


def test_loss_and_dice(b3d, dev):
    for C, shape in ((3, (2, 6, 5, 7)), (1, (1, 4, 4, 4)), (3, (1, 16, 16, 16))):
        x, yv = t64(*shape, 2, seed=1), t64(*shape, 2, seed=2)
        yp = torch.sigmoid(t64(*shape, C, seed=3) * 2)
        y = (t64(*shape, C, seed=4) > 0.8).double()
        mu, lv = t64(shape[0], 64, seed=5), t64(shape[0], 64, seed=6) * 0.5
        ypr, yvr, mur, lvr = (t.clone().requires_grad_(True) for t in (yp, yv, mu, lv))
        lr_ = R.dice_vae_loss(x, y, ypr, yvr, mur, lvr)
        (lr_ * 1.7).backward()
        d = lambda t, g=False: dev32(t, dev, g)
        ypd, yvd, mud, lvd = d(yp, True), d(yv, True), d(mu, True), d(lv, True)
        l = b3d.DiceVAELoss()(d(x), d(y), ypd, yvd, mud, lvd)
        (l * 1.7).backward()
        assert abs(float(l) - float(lr_)) / abs(float(lr_)) < 1e-5
        for a, b in ((ypd, ypr), (yvd, yvr), (mud, mur), (lvd, lvr)):
            assert rel(a.grad, b.grad) < 1e-4
        macro, micro = b3d.DiceCoefficient()(d(y), d(yp))
        mr, ur = R.dice_coefficient(y, yp)
        assert abs(float(macro) - float(mr)) < 1e-5 and abs(float(micro) - float(ur)) < 1e-5


def test_fused_loss_and_dice_pass(b3d, dev):
    """DiceVAELoss followed by DiceCoefficient on the SAME tensors (train.py:143,147): the metric comes out of the loss
    kernel's pass (b3d_loss_dice_fwd); it must equal the oracle and the stand-alone metric kernel, and a changed
    y_pred must not pick up the stale result."""
    ops = b3d.ops
    for C, shape in ((3, (2, 6, 5, 8)), (1, (1, 4, 4, 12)), (3, (1, 16, 16, 16)), (2, (1, 3, 5, 260))):
        x, yv = t64(*shape, 2, seed=1), t64(*shape, 2, seed=2)
        yp = torch.sigmoid(t64(*shape, C, seed=3) * 2)
        y = (t64(*shape, C, seed=4) > 0.8).double()
        mu, lv = t64(shape[0], 64, seed=5), t64(shape[0], 64, seed=6) * 0.5
        d = lambda t, g=False: dev32(t, dev, g)
        yd, ypd = d(y), d(yp, True)
        n0 = ops.LAUNCHES["n"]
        l = b3d.DiceVAELoss()(d(x), yd, ypd, d(yv), d(mu), d(lv))
        n1 = ops.LAUNCHES["n"]
        macro, micro = b3d.DiceCoefficient()(yd, ypd)
        assert ops.LAUNCHES["n"] == n1, "the metric of the same tensors must not launch anything"
        assert n1 - n0 == 1
        lr_ = R.dice_vae_loss(x, y, yp, yv, mu, lv)
        mr, ur = R.dice_coefficient(y, yp)
        assert abs(float(l) - float(lr_)) / abs(float(lr_)) < 1e-5
        assert abs(float(macro) - float(mr)) < 1e-5 and abs(float(micro) - float(ur)) < 1e-5
        m2, u2 = b3d.DiceCoefficient()(yd.clone(), ypd.detach().clone())        # other objects: stand-alone kernel
        assert ops.LAUNCHES["n"] == n1 + 1
        assert abs(float(m2) - float(macro)) < 1e-6 and abs(float(u2) - float(micro)) < 1e-6
        with torch.no_grad():
            ypd.mul_(0.5)                                                        # version bump: cached metric is stale
        m3, u3 = b3d.DiceCoefficient()(yd, ypd)
        mr3, ur3 = R.dice_coefficient(y, yp * 0.5)
        assert abs(float(m3) - float(mr3)) < 1e-5 and abs(float(u3) - float(ur3)) < 1e-5


def test_dense_and_sample(b3d, dev):
    for B, K, N, act in ((1, 4096, 128, 0), (2, 64, 512, 1), (3, 37, 5, 1)):
        x, w, b = t64(B, K, seed=1), t64(K, N, seed=2, scale=0.1), t64(N, seed=3)
        xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
        yr = R.dense(xr, wr, br)
        yr = torch.relu(yr) if act else yr
        gy = t64(B, N, seed=4)
        (yr * gy).sum().backward()
        xd, wd, bd = dev32(x, dev, True), dev32(w, dev, True), dev32(b, dev, True)
        y = b3d.ops.dense(xd, wd, bd, act)
        (y * dev32(gy, dev)).sum().backward()
        assert rel(y, yr) < TOL32 and rel(xd.grad, xr.grad) < TOL32
        assert rel(wd.grad, wr.grad) < TOL32 and rel(bd.grad, br.grad) < TOL32
    proj, eps = t64(2, 128, seed=5), t64(2, 64, seed=6)
    pr = proj.clone().requires_grad_(True)
    zr = pr[:, :64] + torch.exp(0.5 * pr[:, 64:]) * eps
    (zr.sum() * 2 + (pr[:, :64] ** 2).sum() + pr[:, 64:].exp().sum()).backward()
    pd = dev32(proj, dev, True)
    z, zm, zl = b3d.ops.vae_sample(pd, dev32(eps, dev))
    (z.sum() * 2 + (zm ** 2).sum() + zl.exp().sum()).backward()
    assert rel(z, zr) < TOL32 and rel(pd.grad, pr.grad) < TOL32


def test_adam_matches_tf_form(b3d, dev):
    n = 1003
    th, g1, g2 = t64(n, seed=1), t64(n, seed=2), t64(n, seed=3)
    m, v = torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
    ref = th.clone()
    lr = R.poly_lr(7)
    R.adam_step_tf(ref, m, v, g1, 1, lr)
    R.adam_step_tf(ref, m, v, g2, 2, lr)
    opt = b3d.ScheduledOptim(learning_rate=1e-4)
    opt(epoch=7)
    var = dev32(th, dev)
    opt.apply_gradients([(dev32(g1, dev), var)])
    opt.apply_gradients([(dev32(g2, dev), var)])
    assert opt.iterations == 2
    assert float((var.cpu().double() - ref).abs().max()) < 1e-6
    assert rel(var - dev32(th, dev), ref - th) < 5e-3      # update ~2e-4 on values ~1 in fp32


def test_concat_and_dropout(b3d, dev):
    a, b, c = (dev32(t64(2, 3, 4, 5, ch, seed=ch), dev, True) for ch in (4, 6, 16))
    out = b3d.ops.concat([a, b, c])
    ref = torch.cat([a.detach(), b.detach(), c.detach()], dim=-1)
    assert torch.equal(out, ref)
    g = dev32(t64(*out.shape, seed=9), dev)
    (out * g).sum().backward()
    assert torch.equal(a.grad, g[..., :4]) and torch.equal(b.grad, g[..., 4:10]) and torch.equal(c.grad, g[..., 10:])
    x = torch.ones(1, 32, 32, 32, 2, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    y1 = b3d.ops.dropout(x, 0.2, True, None, 123, cnt)
    y2 = b3d.ops.dropout(x, 0.2, True, None, 123, cnt)
    keep = float((y1 > 0).float().mean())
    assert abs(keep - 0.8) < 0.01 and abs(float(y1.max()) - 1.25) < 1e-6
    assert not torch.equal(y1, y2) and int(cnt) == 2
    assert b3d.ops.dropout(x, 0.2, False) is x


TC_CASES = [
    # (B, spatial, Cin, Cout)
    (1, (8, 16, 16), 16, 16),
    (1, (4, 16, 8), 24, 16),       # Cin % 16 != 0: tf32 operands even in bf16 mode
    (2, (4, 16, 8), 8, 32),
    (1, (6, 32, 24), 32, 64),
    (1, (5, 24, 20), 64, 128),      # partial tiles in every dim (inference-like 20x24x20)
    (1, (4, 16, 16), 128, 256),     # N split
    (1, (16, 16, 16), 512, 128),
    (1, (3, 7, 9), 16, 48),
    (1, (32, 32, 32), 16, 16),
]


@pytest.mark.parametrize("case", TC_CASES)
def test_tc_conv_vs_oracle_and_generic(b3d, dev, case):
    """tcgen05 implicit GEMM (TF32) vs the fp64 oracle (<= 2e-3) and vs the fp32 CUDA-core kernel,
    including the fused GroupNorm statistics and the dgrad (flipped/transposed packing)."""
    B, sp, cin, cout = case
    assert b3d.ops.tc_supported(torch.empty(3, 3, 3, cin, cout, device="meta"), 1, False, False)
    x = t64(B, *sp, cin, seed=21)
    w = t64(3, 3, 3, cin, cout, seed=22, scale=(2.0 / (27 * cin)) ** 0.5)
    bias = t64(cout, seed=23)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    yr = R.conv3d_same(xr, wr, bias, 1)
    gy = t64(*yr.shape, seed=24)
    (yr * gy).sum().backward()
    S = sp[0] * sp[1] * sp[2]
    groups = 8 if S % 8 == 0 else 0
    res = {}
    for tc in MODES:
        set_mode(b3d, tc)
        try:
            xd, wd, bd = dev32(x, dev, True), dev32(w, dev, True), dev32(bias, dev, True)
            y, stats, _ = b3d.ops.conv3d(xd, wd, bd, 1, False, 0, groups, False)
            (y * dev32(gy, dev)).sum().backward()
            torch.cuda.synchronize()
            res[tc] = (y.detach(), stats, xd.grad, wd.grad)
        finally:
            reset_mode(b3d)
    for tc, tol in MODE_TOL.items():
        y, stats, dx, dw = res[tc]
        assert rel(y, yr) < tol, ("y", tc, rel(y, yr))
        assert rel(dx, xr.grad) < tol, ("dx", tc, rel(dx, xr.grad))
        assert rel(dw, wr.grad) < (TOL_BF16 if tc != "fp32" else tol), ("dw", tc, rel(dw, wr.grad))
        if groups:
            ch = yr.detach().reshape(B, groups, -1)
            ref = torch.stack([ch.sum(-1), (ch ** 2).sum(-1)], dim=-1)
            assert rel(stats, ref) < max(tol, 1e-4), ("stats", tc)
    assert rel(res["tf32"][0], res["fp32"][0]) < TOL_TF32 and rel(res["bf16"][0], res["fp32"][0]) < TOL_BF16
    assert rel(res["fp16"][0], res["fp32"][0]) < TOL_TF32


KD_FOLD_CASES = [((1, 16, 32, 16), 16, 16, 0), ((1, 5, 24, 20), 32, 16, 0), ((2, 8, 16, 16), 16, 32, 0),
                 ((1, 12, 20, 9), 2, 16, 0), ((1, 9, 17, 11), 16, 3, 1), ((1, 4, 16, 16), 32, 96, 0)]


@pytest.mark.parametrize("case", KD_FOLD_CASES)
def test_kd_folded_conv_equals_unfolded(b3d, dev, case):
    """The kd-folded 3x3x3 tcgen05 kernel (depth taps folded into the MMA N dimension, csrc/conv_tc.cu) against the
    unfolded one on the same operands: only the fp32 summation order differs.  Forward (+bias, sigmoid, GroupNorm
    statistics) and data gradient; full and partial tiles, narrow inputs / outputs, N splits."""
    ops = b3d.ops
    (B, D, H, W), cin, cout, act = case
    x = dev32(t64(B, D, H, W, cin, seed=81), dev)
    w = dev32(t64(3, 3, 3, cin, cout, seed=82, scale=(2.0 / (27 * cin)) ** 0.5), dev)
    bias, dy = dev32(t64(cout, seed=83), dev), dev32(t64(B, D, H, W, cout, seed=84), dev)
    groups = 8 if (D * H * W) % 8 == 0 and cout % 8 == 0 else 0

    def run():
        y, stats, _ = ops.conv3d(x, w, bias, 1, False, act, groups, False)
        dx = torch.empty_like(x)
        ops._call("b3d_conv3d_dgrad", dy, w, dx, 1, 0, 0, ops.pack_weights(w, True, 1, False))
        torch.cuda.synchronize()
        return y, stats, dx

    prev = ops.set_kd_fold(False)
    try:
        ref = run()
        ops.set_kd_fold(True)
        got = run()
    finally:
        ops.set_kd_fold(prev)
    assert rel(got[0], ref[0]) < 1e-5 and rel(got[2], ref[2]) < 1e-5, (rel(got[0], ref[0]), rel(got[2], ref[2]))
    if groups:
        assert rel(got[1], ref[1]) < 1e-6


@pytest.mark.parametrize("k,cin,cout", [(3, 16, 16), (1, 32, 16), (3, 64, 32)])
def test_dgrad_accumulate_epilogue(b3d, dev, k, cin, cout):
    """b3d_conv3d_dgrad(accumulate=1) adds the data gradient to what `dx` already holds (include/b3d.h), on the
    tcgen05 path and on the CUDA-core path."""
    ops = b3d.ops
    dy = dev32(t64(1, 16, 24, 32, cout, seed=71), dev)
    w = dev32(t64(k, k, k, cin, cout, seed=72, scale=0.1), dev)
    base = dev32(t64(1, 16, 24, 32, cin, seed=73), dev)
    for tc in (True, False):
        wp = ops.pack_weights(w, True, 1, False) if tc else None
        if tc:
            assert ops.tc_supported(w, 1, False, True)
        plain = torch.empty_like(base)
        ops._call("b3d_conv3d_dgrad", dy, w, plain, 1, 0, 0, wp)
        acc = base.clone()
        ops._call("b3d_conv3d_dgrad", dy, w, acc, 1, 0, 1, wp)
        torch.cuda.synchronize()
        assert float(plain.abs().max()) > 0
        assert float((acc - (base + plain)).abs().max()) <= 1e-6 * float(plain.abs().max()) + 1e-6, (tc,)


@pytest.mark.parametrize("shape", [(2, 4, 6, 8, 16), (1, 8, 8, 8, 4), (1, 2, 2, 2, 64)])
def test_max_downsample_and_linear_upsample_layers(b3d, dev, shape):
    """MaxDownsample (downsample.py:51-70) and LinearUpsample (upsample.py:49-79) against the oracle, forward and
    gradients; ties in a pooling window send the gradient to the first maximum, as torch / TF do."""
    x = t64(*shape, seed=31)
    x[0, :2, :2, :2, 0] = 1.5                                  # a window of ties
    xr = x.clone().requires_grad_(True)
    yr = R.max_pool2_same(xr)
    gy = t64(*yr.shape, seed=32)
    (yr * gy).sum().backward()
    xd = dev32(x, dev, True)
    y = b3d.MaxDownsample()(xd)
    (y * dev32(gy, dev)).sum().backward()
    assert rel(y, yr) < 1e-7 and rel(xd.grad, xr.grad) < 1e-7
    # LinearUpsample: 1x1x1 conv (fp32 mode for a tight bound) + nearest-neighbour x2
    b3d.ops.USE_TC["on"] = False
    try:
        f = 8
        w, bias = t64(1, 1, 1, shape[-1], f, seed=33, scale=0.3), t64(f, seed=34)
        xr2, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
        yr2 = R.upsample2_nearest(R.conv3d_same(xr2, wr, br))
        gy2 = t64(*yr2.shape, seed=35)
        (yr2 * gy2).sum().backward()
        up = b3d.LinearUpsample(filters=f)
        xd2 = dev32(x, dev, True)
        up(xd2.detach())
        with torch.no_grad():
            up.ptwise.kernel.copy_(w.float())
            up.ptwise.bias.copy_(bias.float())
        y2 = up(xd2)
        (y2 * dev32(gy2, dev)).sum().backward()
        assert rel(y2, yr2) < TOL32 and rel(xd2.grad, xr2.grad) < TOL32
        assert rel(up.ptwise.kernel.grad, wr.grad) < TOL32 and rel(up.ptwise.bias.grad, br.grad) < TOL32
    finally:
        b3d.ops.USE_TC["on"] = True


@pytest.mark.parametrize("case", [((20, 24, 18), 2, 3, (8, 16, 8), (3, 5, 7), (True, False, True)),
                                  ((16, 16, 16), 1, 1, (16, 16, 16), (0, 0, 0), (False, True, False)),
                                  ((12, 10, 14), 4, 3, (8, 8, 8), (4, 2, 6), (False, False, False))])
def test_gpu_example_pipeline_matches_oracle(b3d, dev, case):
    """train.py:12-47 (tf.data parse_example): intensity shift/scale from the channel variance, crop, flips, one-hot
    labels without background — the fused device pipeline against the oracle restatement with the same draws."""
    vol, C, K, crop, off, flips = case
    g = torch.Generator().manual_seed(41)
    x = torch.randn(*vol, C, generator=g, dtype=torch.float64) * 2 + 0.5
    y = torch.randint(0, K + 1, vol + (1,), generator=g).double()
    shift = torch.rand(C, generator=g, dtype=torch.float64) * 0.2 - 0.1
    scale = torch.rand(C, generator=g, dtype=torch.float64) * 0.2 + 0.9
    xr, yr = R.augment_example(x, y, crop, K, shift, scale, off, flips)
    xo, yo = b3d.parse_example(dev32(x, dev), dev32(y, dev), crop, K, shift=shift.float(), scale=scale.float(),
                               offset=off, flips=flips)
    assert xo.shape == xr.shape and yo.shape == yr.shape
    assert rel(xo, xr) < 1e-6
    assert torch.equal(yo.cpu().double(), yr)
    # random draws: shapes, ranges, determinism of the generator
    ds = b3d.VolumeDataset([(dev32(x, dev), dev32(y, dev))] * 3, batch_size=2, crop_size=crop, out_ch=K, seed=5)
    batches = list(ds)
    assert [b[0].shape[0] for b in batches] == [2, 1] and batches[0][0].shape[1:] == tuple(crop) + (C,)
    assert batches[0][1].shape[1:] == tuple(crop) + (K,) and float(batches[0][1].sum(-1).max()) <= 1.0
    ds2 = b3d.VolumeDataset([(dev32(x, dev), dev32(y, dev))] * 3, batch_size=2, crop_size=crop, out_ch=K, seed=5)
    assert torch.equal(next(iter(ds2))[0], batches[0][0])


@pytest.mark.parametrize("shape", [(2, 8, 8, 8, 16), (1, 6, 5, 7, 32), (1, 4, 4, 4, 192), (2, 3, 3, 3, 8)])
@pytest.mark.parametrize("relu", [False, True])
def test_group_norm_channels_first_semantics(b3d, dev, shape, relu):
    """GroupNormalization(axis=1) — true channel groups, the reference's data_format='channels_first' behaviour
    (group_norm.py:83-124) — on NCDHW public tensors, forward and all gradients vs the oracle."""
    x, ga, be = t64(*shape, seed=51), 1 + 0.3 * t64(shape[-1], seed=52), 0.3 * t64(shape[-1], seed=53)
    xr, gr, br = (t.clone().requires_grad_(True) for t in (x, ga, be))
    with R.channels_first_semantics():
        yr = R.group_norm(xr, gr, br)
    yr = torch.relu(yr) if relu else yr
    gy = t64(*shape, seed=54)
    (yr * gy).sum().backward()
    cf = lambda t: t.permute(0, 4, 1, 2, 3).contiguous()
    gn = b3d.GroupNormalization(groups=8, axis=1)
    xd = dev32(cf(x), dev, True)                       # NCDHW in, NCDHW out
    gn(xd.detach())
    with torch.no_grad():
        gn.gamma.copy_(ga.float())
        gn.beta.copy_(be.float())
    y = gn(xd, relu=relu)
    assert tuple(y.shape) == tuple(cf(x).shape)
    (y * dev32(cf(gy), dev)).sum().backward()
    assert rel(y, cf(yr)) < TOL32
    assert rel(xd.grad, cf(xr.grad)) < 2e-4
    assert rel(gn.gamma.grad, gr.grad) < 1e-4 and rel(gn.beta.grad, br.grad) < 1e-4
