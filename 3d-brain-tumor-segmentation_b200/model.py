"""Model — surface of /root/reference/model.py (:9-56 ctor, :58-71 call).

    Encoder -> Decoder(last, rest) -> optional VAE(last);  returns (y_pred, y_vae, z_mean, z_logvar),
    with Nones after y_pred when `inference=True`.

B200-side additions (not in the reference): all trainable tensors live in ONE flat fp32 buffer
(regularised tensors first) with a matching flat gradient buffer, so that the L2 penalties
(train.py:146), the TF-form Adam update (util.py:60-84) and the data-parallel gradient all-reduce
are each a single pass over HBM / a single collective.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from .keras_compat import Layer
from . import ops
from .layers.encoder import Encoder
from .layers.decoder import Decoder
from .layers.vae import VariationalAutoencoder


class _EpochVariable:
    """tf.Variable(0, name='epoch', trainable=False) (model.py:29)."""

    def __init__(self, v=0):
        self._v = int(v)
        self.name = 'epoch'
        self.trainable = False

    def value(self):
        return self

    def numpy(self):
        return self._v

    def assign(self, v):
        self._v = int(v)

    def __int__(self):
        return self._v


class FlatParams:
    """All trainable tensors of a model as views of one buffer; `grad` is the matching flat gradient."""

    def __init__(self, variables, device):
        # regularised tensors first, grouped by their L2 coefficient (stable: attribute order inside a group), so that
        # every group is one contiguous range of the flat buffer.  Model(l2_scale=X) gives two groups: the VAE's
        # ConvDownsample swallows its kernel_regularizer kwarg and keeps the default 1e-5 (vae.py:53-57)
        reg = sorted((v for v in variables if v.regularizer is not None), key=lambda v: v.regularizer.l)
        unreg = [v for v in variables if v.regularizer is None]
        self.order = reg + unreg
        pad = lambda n: (n + 3) // 4 * 4          # keep every tensor 16-byte aligned
        sizes = [pad(v.tensor.numel()) for v in self.order]
        total = sum(sizes)
        self.theta = torch.zeros(total, device=device, dtype=torch.float32)
        self.grad = torch.zeros(total, device=device, dtype=torch.float32)
        self.spans = {}
        off = 0
        seg = [0]
        for v, sz in zip(self.order, sizes):
            n = v.tensor.numel()
            view = self.theta[off:off + n].view(v.tensor.shape)
            view.copy_(v.tensor.detach())
            v.tensor.data = view                      # same leaf object, storage now inside the flat buffer
            v.tensor.grad = self.grad[off:off + n].view(v.tensor.shape)
            self.spans[id(v.tensor)] = (off, n)
            v.tensor._b3d_flat = self
            off += sz
            if v.regularizer is not None:
                seg.append(off)
        self.n_reg = len(reg)
        self.reg_end = seg[-1]
        self.reg_tensors = [v.tensor for v in reg]
        # (first tensor, one past the last tensor, coefficient) of every run of equal coefficients
        self.l2_groups = []
        for i, v in enumerate(reg):
            if self.l2_groups and self.l2_groups[-1][2] == v.regularizer.l:
                self.l2_groups[-1][1] = i + 1
            else:
                self.l2_groups.append([i, i + 1, v.regularizer.l])
        self.l2 = self.l2_groups[0][2] if len(self.l2_groups) == 1 else None     # single coefficient: fused fast paths
        # segment table of the regularised tensors (padding belongs to the preceding tensor; it stays 0)
        self.offsets = torch.tensor(seg, dtype=torch.int64, device=device)
        self._seg = seg
        self.total = total
        # persistent packed conv operands (ops.pack_weights / ops.repack_all) and their staleness epoch
        self.packs = {}
        self.pack_table = None
        self.epoch = 0
        self.folds = []                # (kernel, folded kernel, F) of the dense-connection convs (ops.FoldDupFn)

    def zero_grad(self):
        ops._call("b3d_zero", self.grad)           # a memset node under graph capture, not a fill kernel

    def notify(self, tensor):
        """A backward kernel has written `tensor`'s gradient straight into the flat buffer (ops.direct_param_grads):
        tell the data-parallel bucket scheduler, which autograd's post-accumulate hooks would otherwise do."""
        cb = getattr(self, "grad_ready_cb", None)
        if cb is not None:
            cb(tensor)

    def add_l2_grad(self, scale=1.0):
        """grad += scale*2*l*w over the regularised tensors (what autograd does through model.losses, train.py:146)."""
        seg = self._seg
        for t0, t1, l in self.l2_groups:
            # a group is a contiguous range of the flat buffer and its padding is zero: one axpy over it
            lo, hi = seg[t0], seg[t1]
            ops._call("b3d_axpy", self.theta[lo:hi], self.grad[lo:hi], int(hi - lo), 2.0 * float(l) * scale, None)

    def attach_grads(self):
        for v in self.order:
            off, n = self.spans[id(v.tensor)]
            v.tensor.grad = self.grad[off:off + n].view(v.tensor.shape)


class _LossList(list):
    """model.losses: the per-tensor penalties as a list (Keras), remembering the vector they are views of so that
    train.reduce_sum can add them up with one kernel."""

    def __init__(self, vec):
        super().__init__(vec.unbind(0))
        self.vector = vec


class _L2LossesFn(torch.autograd.Function):
    """model.losses: one l*sum(w^2) per regularised tensor, computed by ONE kernel over the flat buffer.
    Backward accumulates 2*l*w*gout straight into the flat gradient buffer (the tensors' .grad views)."""

    @staticmethod
    def forward(ctx, flat: FlatParams, *params):
        out = torch.empty(flat.n_reg, device=flat.theta.device, dtype=torch.float32)
        for t0, t1, l in flat.l2_groups:
            ops._call("b3d_l2_losses", flat.theta, flat.offsets[t0:t1 + 1], out[t0:t1], float(l))
        ctx.flat = flat
        return out

    @staticmethod
    def backward(ctx, gout):
        flat = ctx.flat
        gout = gout.contiguous()
        for t0, t1, l in flat.l2_groups:
            ops._call("b3d_l2_grad", flat.theta, flat.grad, flat.offsets[t0:t1 + 1], gout[t0:t1], 2.0 * float(l))
        return (None,) * (1 + flat.n_reg)


class Model(Layer):
    def __init__(self,
                 data_format='channels_last',
                 groups=8,
                 reduction=2,
                 l2_scale=1e-5,
                 dropout=0.2,
                 downsampling='conv',
                 upsampling='conv',
                 base_filters=16,
                 depth=4,
                 in_ch=2,
                 out_ch=3):
        super().__init__()
        from .keras_compat import check_data_format
        self.data_format = check_data_format(data_format)
        self.epoch = _EpochVariable(0)
        self.encoder = Encoder(data_format=data_format, groups=groups, reduction=reduction, l2_scale=l2_scale,
                               dropout=dropout, downsampling=downsampling, base_filters=base_filters, depth=depth)
        self.decoder = Decoder(data_format=data_format, groups=groups, reduction=reduction, l2_scale=l2_scale,
                               upsampling=upsampling, base_filters=base_filters, depth=depth, out_ch=out_ch)
        self.vae = VariationalAutoencoder(data_format=data_format, groups=groups, reduction=reduction,
                                          l2_scale=l2_scale, upsampling=upsampling, base_filters=base_filters,
                                          depth=depth, out_ch=in_ch)
        self._flat: Optional[FlatParams] = None

    # ---- Keras protocol
    def __call__(self, inputs, training=None, inference=None, **kw):
        from .keras_compat import internal_layout, _internal_layout, map5d
        if self.data_format == 'channels_first' and not _internal_layout():
            # public tensors are NCDHW (model.py:58 under args.py's GPU default); storage inside is NDHWC
            inputs = ops.to_channels_last(inputs)
            kw = {k: map5d(v, ops.to_channels_last) for k, v in kw.items()}
            with internal_layout():
                out = self.__call__(inputs, training=training, inference=inference, **kw)
            return map5d(out, ops.to_channels_first)
        if not self.built:
            # the reference builds all weights with a first call on zeros (train.py:96); here the first
            # call does a weight-creating dry run, then re-homes every tensor into the flat buffer
            with torch.no_grad():
                self.call(inputs, training=False, inference=False)
            self.built = True
            self.flatten_parameters()
        return self.call(inputs, training=training, inference=inference, **kw)

    def call(self, inputs, training=None, inference=None, dropout_mask=None, eps=None):
        # Inference mode does not evaluate the VAE branch (model.py:60)
        assert (not inference or not training), \
            'Cannot run training and inference modes simultaneously.'
        # inside the model every layer output is consumed by convs only: blocks and resampling layers hand over 16-bit
        # operand twins instead of fp32 tensors and channel concatenation is a list of sources (ops.fused_scope)
        with ops.fused_scope():
            residuals = self.encoder(inputs, training=training, dropout_mask=dropout_mask)
            y_pred = self.decoder((residuals[-1], residuals[:-1]), training=training)
            if inference:
                return (y_pred, None, None, None)
            y_vae, z_mean, z_logvar = self.vae(residuals[-1], training=training, eps=eps)
        return (y_pred, y_vae, z_mean, z_logvar)

    # ---- flat parameter storage
    def flatten_parameters(self) -> FlatParams:
        if self._flat is None:
            vs = self.variables()
            self._flat = FlatParams(vs, vs[0].tensor.device)
        return self._flat

    @property
    def flat(self) -> Optional[FlatParams]:
        return self._flat

    @property
    def losses(self) -> List[torch.Tensor]:
        """L2 penalties of all regularised tensors (168 for the default model; train.py:146 sums them)."""
        flat = self.flatten_parameters()
        out = _L2LossesFn.apply(flat, *flat.reg_tensors)
        return _LossList(out)

    # ---- weight exchange by this repo's structural names (oracle.ref_model.param_shapes)
    def named_variables(self):
        out = {}

        def blk(b, pre):
            out[pre + "ptwise.kernel"] = b.conv3d_ptwise.kernel
            out[pre + "ptwise.bias"] = b.conv3d_ptwise.bias
            out[pre + "dense_relu.kernel"] = b.dense_relu.kernel
            out[pre + "dense_sigmoid.kernel"] = b.dense_sigmoid.kernel
            out[pre + "spatial.kernel"] = b.spatial.kernel
            for i in (0, 1):
                conv, norm, _ = b.convs[i]
                out[pre + f"conv{i+1}.kernel"] = conv.kernel
                out[pre + f"conv{i+1}.bias"] = conv.bias
                out[pre + f"gn{i+1}.gamma"] = norm.gamma
                out[pre + f"gn{i+1}.beta"] = norm.beta

        def rs(l, pre):
            if not hasattr(l, "conv"):                       # MaxDownsample (no weights) / LinearUpsample
                if hasattr(l, "ptwise"):
                    out[pre + "ptwise.kernel"], out[pre + "ptwise.bias"] = l.ptwise.kernel, l.ptwise.bias
                return
            out[pre + "conv.kernel"] = l.conv.kernel
            out[pre + "conv.bias"] = l.conv.bias
            out[pre + "norm.gamma"] = l.norm.gamma
            out[pre + "norm.beta"] = l.norm.beta

        depth = len(self.encoder.levels)
        for i, (convs, _, down) in enumerate(self.encoder.levels):
            for j, (b, _) in enumerate(convs):
                blk(b, f"enc.L{i}.B{j}.")
            if down is not None:
                rs(down, f"enc.L{i}.down.")
        for i, (up, _, b) in zip(range(depth - 2, -1, -1), self.decoder.levels):
            rs(up, f"dec.L{i}.up.")
            blk(b, f"dec.L{i}.block.")
        out["dec.out.kernel"], out["dec.out.bias"] = self.decoder.out.kernel, self.decoder.out.bias
        v = self.vae
        if v.built:
            rs(v.downsample, "vae.down.")
            out["vae.proj.kernel"], out["vae.proj.bias"] = v.proj.kernel, v.proj.bias
            out["vae.unproj.kernel"], out["vae.unproj.bias"] = v.unproj.kernel, v.unproj.bias
            rs(v.upsample, "vae.up.")
            for i, (up, b) in zip(range(depth - 2, -1, -1), v.levels):
                rs(up, f"vae.L{i}.up.")
                blk(b, f"vae.L{i}.block.")
            out["vae.out.kernel"], out["vae.out.bias"] = v.out.kernel, v.out.bias
        return out

    # ---- checkpoints (train.py:99-100 load_weights, :199-201 save_weights)
    # The reference calls Keras' `save_weights('chkpt.hdf5')` on a subclassed model: an HDF5 file with one group per
    # top-level layer (`model.layers` = encoder, decoder, variational_autoencoder; attribute `layer_names`), each with
    # its `weight_names` attribute and one dataset per weight, written / read back TOPOLOGICALLY (layers in order,
    # weights in `layer.weights` order; hdf5_format.save_weights_to_hdf5_group / load_weights_from_hdf5_group).  The
    # model's own `epoch` variable (model.py:29) is not part of `model.layers` and so not in Keras' file.
    # `keras_weight_groups()` reproduces that structure — layer names by Keras' per-class counters, weight names
    # `<scope path>/<var>:0`, Keras layouts, fp32.  With h5py it is written as that HDF5 file; h5py is NOT in this image, so
    # here the same structure goes into a NumPy .npz (keys '<layer>/<weight name>', '__layer_names__',
    # '__weight_names__/<layer>', plus '__epoch__'), which tools/npz_to_keras_h5.py turns into the .hdf5 on a machine
    # that has h5py.  TF itself cannot run here, so the names are restated from Keras' rules, not verified against it;
    # loading is topological, like Keras', so it does not depend on them.
    def keras_weight_groups(self):
        """[(top-level layer name, [(Keras weight name, tensor), ...]), ...] in `model.layers` / `layer.weights` order."""
        groups = []
        for top in self._sublayers():
            items = []

            def walk(layer, scope):
                for v in layer._vars:
                    items.append((f"{scope}/{v.name}:0", v.tensor))
                for sub in layer._sublayers():
                    walk(sub, f"{scope}/{sub.name}")

            walk(top, f"{self.name}/{top.name}")
            if items:
                groups.append((top.name, items))
        return groups

    def save_weights(self, filepath):
        import numpy as np
        groups = self.keras_weight_groups()
        if str(filepath).endswith((".h5", ".hdf5", ".keras")):
            try:
                import h5py
            except ImportError:
                h5py = None
            if h5py is not None:
                with h5py.File(filepath, "w") as f:
                    f.attrs["layer_names"] = [n.encode("utf8") for n, _ in groups]
                    f.attrs["backend"] = b"tensorflow"
                    f.attrs["keras_version"] = b"2.2.4-tf"
                    for lname, items in groups:
                        g = f.create_group(lname)
                        g.attrs["weight_names"] = [wn.encode("utf8") for wn, _ in items]
                        for wn, t in items:
                            g.create_dataset(wn, data=t.detach().cpu().numpy())
                return
            raise ImportError("b3d: writing Keras' HDF5 container needs h5py (not in this image); save to a '.npz' path "
                              "and convert with tools/npz_to_keras_h5.py")
        arrs = {"__layer_names__": np.array([n for n, _ in groups]),
                "__epoch__": np.asarray(int(self.epoch), dtype=np.int64)}
        for lname, items in groups:
            arrs[f"__weight_names__/{lname}"] = np.array([wn for wn, _ in items])
            for wn, t in items:
                arrs[f"{lname}/{wn}"] = t.detach().cpu().numpy()
        with open(filepath, "wb") as f:
            np.savez(f, **arrs)

    def load_weights(self, filepath):
        """Topological load (Keras: by_name=False): the file's layers in order onto this model's weighted top-level
        layers, each layer's weights in order; shapes are checked.  Also reads the round-1 structural-name .npz."""
        import numpy as np
        groups = self.keras_weight_groups()
        saved = None
        try:
            z = np.load(filepath, allow_pickle=False)
        except Exception:        # not an .npz: Keras' HDF5
            import h5py          # raises ImportError with a clear message when absent
            with h5py.File(filepath, "r") as f:
                lnames = [n.decode("utf8") if isinstance(n, bytes) else n for n in f.attrs["layer_names"]]
                saved = []
                for ln in lnames:
                    wns = [n.decode("utf8") if isinstance(n, bytes) else n for n in f[ln].attrs["weight_names"]]
                    if wns:
                        saved.append((ln, [torch.from_numpy(np.asarray(f[ln][wn])) for wn in wns]))
        else:
            with z:
                if "__layer_names__" not in z.files:          # round-1 file: this repo's structural names
                    self.load_named_weights({k: torch.from_numpy(z[k]) for k in z.files if k != "__epoch__"})
                    if "__epoch__" in z.files:
                        self.epoch.assign(int(z["__epoch__"]))
                    return
                saved = [(str(ln), [torch.from_numpy(z[f"{ln}/{wn}"]) for wn in z[f"__weight_names__/{ln}"]])
                         for ln in z["__layer_names__"]]
                if "__epoch__" in z.files:
                    self.epoch.assign(int(z["__epoch__"]))
        if len(saved) != len(groups):
            raise ValueError(f"You are trying to load a weight file containing {len(saved)} layers into a model with "
                             f"{len(groups)} layers.")
        with torch.no_grad():
            for (ln, vals), (mn, items) in zip(saved, groups):
                if len(vals) != len(items):
                    raise ValueError(f"Layer {mn} expects {len(items)} weights, but the saved weights have {len(vals)} "
                                     f"elements.")
                for src, (wn, t) in zip(vals, items):
                    if tuple(src.shape) != tuple(t.shape):
                        raise ValueError(f"{wn}: shape {tuple(src.shape)} != {tuple(t.shape)}")
                    t.copy_(src.to(device=t.device, dtype=t.dtype))
        if self._flat is not None:
            ops.repack_all(self._flat)

    def load_named_weights(self, params):
        nv = self.named_variables()
        missing = set(nv) - set(params)
        if missing:
            raise KeyError(f"missing weights: {sorted(missing)[:5]} ...")
        with torch.no_grad():
            for k, t in nv.items():
                src = params[k]
                if tuple(src.shape) != tuple(t.shape):
                    raise ValueError(f"{k}: shape {tuple(src.shape)} != {tuple(t.shape)}")
                t.copy_(src.to(device=t.device, dtype=t.dtype))
        if self._flat is not None:
            ops.repack_all(self._flat)       # captured inference graphs read the persistent packed operands
