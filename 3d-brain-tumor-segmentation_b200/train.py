"""Training-step glue with the shape of the reference's hot loop (/root/reference/train.py:140-152):

    with tf.GradientTape() as tape:
        y_pred, y_vae, z_mean, z_logvar = model(x, training=True, inference=False)
        loss = loss_fn(x, y, y_pred, y_vae, z_mean, z_logvar)
        loss += tf.reduce_sum(model.losses)
    macro_dice, micro_dice = dice_fn(y, y_pred)
    grads = tape.gradient(loss, model.trainable_variables)
    optimizer.apply_gradients(zip(grads, model.trainable_variables))

plus the two B200-side execution modes the reference has no counterpart for: whole-step CUDA-graph
replay (the step is ~600 small launches; eager dispatch would dominate) and one-process-per-GPU data
parallelism with a bucketed NCCL all-reduce of the flat gradient buffer overlapped with backward.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .model import Model
from .util import DiceVAELoss, DiceCoefficient, ScheduledOptim


class GradientTape:
    """Minimal tf.GradientTape look-alike over torch autograd."""

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def gradient(self, loss, variables, model: Optional[Model] = None, direct: bool = False):
        """`direct`: the backward kernels write parameter gradients straight into the model's flat gradient
        buffer (ops.direct_param_grads) instead of handing them to autograd's accumulation."""
        flat = getattr(variables[0], "_b3d_flat", None)
        if flat is not None:
            flat.zero_grad()
            flat.attach_grads()
        else:
            for v in variables:
                v.grad = None
        if direct and flat is not None:
            with ops.direct_param_grads(flat):
                loss.backward()
        else:
            loss.backward()
        return [v.grad for v in variables]


class _SumAddFn(torch.autograd.Function):
    """sum(vec) (+ addend) by one kernel; differentiable (d/dvec = gout broadcast, d/daddend = gout)."""

    @staticmethod
    def forward(ctx, vec, addend):
        out = torch.empty((), device=vec.device, dtype=torch.float32)
        ops._call("b3d_sum_add", vec.contiguous(), None if addend is None else addend.reshape(1), out.reshape(1))
        ctx.n = vec.shape
        return out

    @staticmethod
    def backward(ctx, g):
        return g.expand(ctx.n), (g if ctx.needs_input_grad[1] else None)


def reduce_sum(xs, addend=None):
    """tf.reduce_sum over a list of scalars (train.py:146), optionally + `addend` (the data loss).  `model.losses`
    remembers the vector its entries are views of, which makes this one kernel."""
    vec = getattr(xs, "vector", None)
    if vec is None:
        s = torch.stack(list(xs)).sum() if len(xs) else 0.0
        return s if addend is None else addend + s
    return _SumAddFn.apply(vec, addend)


def train_step(model: Model, optimizer: ScheduledOptim, loss_fn: DiceVAELoss, dice_fn: DiceCoefficient, x, y,
               dropout_mask=None, eps=None, dp: "DataParallel | None" = None):
    """One iteration of train.py:140-152.  Returns (loss, macro_dice, micro_dice) as 0-d device tensors.

    With `dp` (one process per GPU) the data gradients are all-reduced bucket by bucket while backward is
    still running, and the L2-regulariser gradient 2*l*w — identical on every rank — is added inside the
    fused Adam kernel instead of by autograd, so it is applied once and not averaged."""
    with GradientTape() as tape:
        y_pred, y_vae, z_mean, z_logvar = model(x, training=True, inference=False,
                                                dropout_mask=dropout_mask, eps=eps)
        data_loss = loss_fn(x, y, y_pred, y_vae, z_mean, z_logvar)
        with torch.no_grad():      # reported loss only: the regulariser's gradient is applied as one axpy below
            loss = reduce_sum(model.losses, addend=data_loss)
    macro_dice, micro_dice = dice_fn(y, y_pred)
    variables = model.trainable_variables
    if dp is not None:
        dp.begin_backward()
    # the regulariser's gradient 2*l*w is a single pass over the flat buffer: after backward (one GPU), or inside
    # the Adam kernel (data parallel: it must not be averaged over ranks)
    grads = tape.gradient(data_loss, variables, direct=True)
    if dp is not None:
        dp.finish_backward()
        optimizer.apply_flat(model.flat, l2_in_step=True)
    else:
        model.flat.add_l2_grad()
        optimizer.apply_gradients(zip(grads, variables), flat=model.flat)
    return loss.detach(), macro_dice, micro_dice


class DataParallel:
    """One-process-per-GPU data parallelism for train_step: every rank holds a replica (42.5 MB of fp32
    weights + Adam state for the default model) and its own crop; the flat gradient buffer is cut into
    `n_buckets` contiguous ranges that are all-reduced (sum; the 1/N is folded into the Adam kernel) over
    NCCL / NVLink on a side stream as soon as autograd has produced every gradient of the range, i.e.
    overlapped with the rest of backward.  GroupNorm and scSE are per-sample, so no other collective exists
    on the path.  Works under CUDA-graph capture (the collectives are captured on the side stream)."""

    def __init__(self, model: Model, optimizer: ScheduledOptim, world_size: int, n_buckets: int = 4,
                 process_group=None, overlap: bool = True, objective: str = "replica_mean", loss_fn=None,
                 sync_weights: bool = True):
        """objective (SURVEY F6 — the reference's Dice sums over the batch axis, util.py:11,18-20):
          'replica_mean'  every rank optimises the loss of its own crop(s); gradients are averaged — the mean of
                          per-crop DiceVAE losses (what `--batch_size 1` training sees, averaged over N crops);
          'global_batch'  the reference's `--batch_size N` objective: `loss_fn` all-reduces its 3C+2 partial sums in the
                          forward (one tiny collective), every rank holds the loss of the whole batch, gradients are
                          SUMMED.  Pass the DiceVAELoss instance as `loss_fn`.
        sync_weights: broadcast rank 0's parameters (and check nothing else diverged) so that replicas cannot start
        from different weights silently."""
        import torch.distributed as dist
        if objective not in ("replica_mean", "global_batch"):
            raise ValueError(objective)
        self.dist, self.group, self.world = dist, process_group, world_size
        self.model, self.flat = model, model.flatten_parameters()
        self.objective = objective
        if objective == "global_batch":
            if loss_fn is None:
                raise ValueError("objective='global_batch' needs the DiceVAELoss instance (loss_fn=)")
            loss_fn.dp = (process_group, world_size)
            optimizer.grad_scale = 1.0
        else:
            optimizer.grad_scale = 1.0 / world_size
        if sync_weights and world_size > 1 and dist.is_initialized():
            dist.broadcast(self.flat.theta, 0, group=process_group)
            for mv in getattr(optimizer, "_slots", {}).get(id(self.flat), ()):
                dist.broadcast(mv, 0, group=process_group)
            if getattr(optimizer, "_state", None) is not None:
                dist.broadcast(optimizer._state, 0, group=process_group)
            if self.flat.theta.is_cuda:
                ops.repack_all(self.flat)
            # different crops need different dropout masks: fold the rank into the dropout seed
            drop = getattr(getattr(model, "encoder", None), "dropout", None)
            if drop is not None:
                drop.seed = (int(drop.seed) + 7919 * dist.get_rank(process_group)) & 0x7FFFFFFF
        self.overlap = overlap and self.flat.grad.is_cuda
        self.side = torch.cuda.Stream() if self.flat.grad.is_cuda else None
        self.buckets = self.plan_buckets(self.flat, n_buckets)
        self._pending = [0] * len(self.buckets)
        self._seen = set()
        self._owner = {}
        for bi, (lo, hi, members) in enumerate(self.buckets):
            for t in members:
                self._owner[id(t)] = bi
        if self.overlap:
            for v in self.flat.order:
                v.tensor.register_post_accumulate_grad_hook(self._on_grad)
            self.flat.grad_ready_cb = self._on_grad      # gradients written directly by the backward kernels
        self._active = False

    @staticmethod
    def plan_buckets(flat, n_buckets):
        """Contiguous [lo, hi) ranges of the flat buffer with ~equal sizes, cut at tensor boundaries."""
        spans = sorted(((flat.spans[id(v.tensor)][0], v.tensor) for v in flat.order), key=lambda t: t[0])
        target = flat.total / n_buckets
        buckets, lo, members = [], 0, []
        for i, (off, t) in enumerate(spans):
            members.append(t)
            end = spans[i + 1][0] if i + 1 < len(spans) else flat.total
            if end - lo >= target or i + 1 == len(spans):
                buckets.append((lo, end, members))
                lo, members = end, []
        return buckets

    def begin_backward(self):
        self._pending = [len(m) for _, _, m in self.buckets]
        self._seen = set()
        self._active = True

    def _reduce(self, bi):
        lo, hi, _ = self.buckets[bi]
        view = self.flat.grad[lo:hi]
        if self.side is not None:
            self.side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.side):
                self.dist.all_reduce(view, op=self.dist.ReduceOp.SUM, group=self.group)
        else:
            self.dist.all_reduce(view, op=self.dist.ReduceOp.SUM, group=self.group)

    def _on_grad(self, param):
        if not self._active:
            return
        # a parameter can be announced twice in one backward — by the kernel that wrote its gradient straight into the
        # flat buffer (FlatParams.notify) AND by autograd's post-accumulate hook (measured on torch 2.11: every
        # parameter, although those backward functions return None for it); counting both made every bucket's
        # all-reduce start when only HALF its gradients existed and run a second time in finish_backward
        if id(param) in self._seen:
            return
        self._seen.add(id(param))
        bi = self._owner[id(param)]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._reduce(bi)

    def finish_backward(self):
        self._active = False
        for bi, left in enumerate(self._pending):
            if left != 0 or not self.overlap:        # no-overlap mode, or a tensor that got no gradient
                self._reduce(bi)
        self._pending = [0] * len(self.buckets)
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)


class GraphedTrainStep:
    """Captures train_step into a CUDA graph: static input buffers, one graph launch per step."""

    def __init__(self, model, optimizer, loss_fn, dice_fn, x, y, warmup=2, dp=None):
        self.model, self.optimizer = model, optimizer
        self.x, self.y = x.clone(), y.clone()
        self._args = (model, optimizer, loss_fn, dice_fn)
        # everything that allocates persistent state must exist before capture
        flat = model.flatten_parameters() if model.built else None
        if flat is not None:
            optimizer._ensure_state(flat.theta.device)
            optimizer._mv(id(flat), flat.theta)
        # the warm-up steps (allocator / lazy-initialisation warm-up before capture) must not train: weights, Adam
        # moments, step count and the dropout counter are snapshotted and restored, so that building the graphed step
        # leaves the training state exactly where it was
        snap = self._snapshot(flat, optimizer, model) if (flat is not None and warmup > 0) else None
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                train_step(*self._args, self.x, self.y, dp=dp)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if snap is not None:
            self._restore(snap, flat, optimizer, model)
        if model.flat is not None and model.flat.packs:
            ops.ensure_pack_table(model.flat)
        n0 = ops.LAUNCHES["n"]
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = train_step(*self._args, self.x, self.y, dp=dp)
        self.launches_per_step = ops.LAUNCHES["n"] - n0

    @staticmethod
    def _snapshot(flat, optimizer, model):
        m, v = optimizer._mv(id(flat), flat.theta)
        drop = getattr(getattr(model, "encoder", None), "dropout", None)
        cnt = getattr(drop, "_counter", None)
        return (flat.theta.clone(), m.clone(), v.clone(), optimizer._state.clone(), None if cnt is None else cnt.clone())

    @staticmethod
    def _restore(snap, flat, optimizer, model):
        th, m0, v0, st, cnt = snap
        m, v = optimizer._mv(id(flat), flat.theta)
        with torch.no_grad():
            flat.theta.copy_(th); m.copy_(m0); v.copy_(v0); optimizer._state.copy_(st)
            drop = getattr(getattr(model, "encoder", None), "dropout", None)
            if cnt is not None and getattr(drop, "_counter", None) is not None:
                drop._counter.copy_(cnt)
        ops.repack_all(flat)
        torch.cuda.synchronize()

    def __call__(self, x=None, y=None):
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if y is not None:
            self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        return self.out

    # ---- input pipelining: the host -> device copy of the NEXT batch overlaps the current step
    def prefetch(self, x_host, y_host):
        """Start copying the next batch (pinned host tensors) into device staging buffers on a side stream; it
        overlaps whatever the compute stream is running.  `step_prefetched()` then consumes it."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream()
            self._stage = (torch.empty_like(self.x), torch.empty_like(self.y))
            self._staged = torch.cuda.Event()
            self._consumed = torch.cuda.Event()
            self._consumed.record()
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._consumed)        # the staging buffers are free again
            self._stage[0].copy_(x_host, non_blocking=True)
            self._stage[1].copy_(y_host, non_blocking=True)
            self._staged.record()

    def step_prefetched(self):
        """One step on the batch handed to the last prefetch() (a 42 MB device-to-device copy, then the graph)."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged)
        self.x.copy_(self._stage[0], non_blocking=True)
        self.y.copy_(self._stage[1], non_blocking=True)
        self._consumed.record()
        self.graph.replay()
        return self.out


class TrainLog:
    """The CSV log of the reference's training loop (train.py:117-127 header, :184-195 one row per epoch) and its
    best-validation-Dice checkpoint / patience rule (train.py:197-208)."""
    HEADER = ['epoch', 'lr', 'train_loss', 'train_macro_dice', 'train_micro_dice', 'val_loss', 'val_macro_dice',
              'val_micro_dice']

    def __init__(self, save_folder=None, patience=10):
        import os
        self.save_folder, self.patience_limit = save_folder, patience
        self.best_val_dice, self.patience = 0.0, 0
        if save_folder:
            os.makedirs(save_folder, exist_ok=True)
            with open(os.path.join(save_folder, 'train.log'), 'w') as f:
                f.write(','.join(self.HEADER) + '\n')

    @staticmethod
    def checkpoint_name() -> str:
        """'chkpt.hdf5' (train.py:201) when h5py can write Keras' container, else the same structure as 'chkpt.npz'
        (Model.save_weights)."""
        try:
            import h5py  # noqa: F401
            return 'chkpt.hdf5'
        except ImportError:
            return 'chkpt.npz'

    def end_epoch(self, epoch, lr, train, val, model=None):
        """train / val: (loss, macro_dice, micro_dice).  Returns False when training should stop (patience)."""
        import os
        if self.save_folder:
            with open(os.path.join(self.save_folder, 'train.log'), 'a') as f:
                f.write(','.join(str(float(v)) if i else str(int(v))
                                 for i, v in enumerate([epoch, lr, *train, *val])) + '\n')
        if float(val[1]) > self.best_val_dice:
            self.best_val_dice, self.patience = float(val[1]), 0
            if self.save_folder and model is not None:
                model.epoch.assign(epoch)
                model.save_weights(os.path.join(self.save_folder, self.checkpoint_name()))
            return True
        if self.patience == self.patience_limit:
            return False
        self.patience += 1
        return True
