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

    def gradient(self, loss, variables, model: Optional[Model] = None):
        flat = getattr(variables[0], "_b3d_flat", None)
        if flat is not None:
            flat.zero_grad()
            flat.attach_grads()
        else:
            for v in variables:
                v.grad = None
        loss.backward()
        return [v.grad for v in variables]


def reduce_sum(xs):
    """tf.reduce_sum over a list of scalars (train.py:146)."""
    return torch.stack(list(xs)).sum() if len(xs) else 0.0


def train_step(model: Model, optimizer: ScheduledOptim, loss_fn: DiceVAELoss, dice_fn: DiceCoefficient, x, y,
               dropout_mask=None, eps=None, grad_hook=None):
    """One iteration of train.py:140-152.  Returns (loss, macro_dice, micro_dice) as 0-d device tensors."""
    with GradientTape() as tape:
        y_pred, y_vae, z_mean, z_logvar = model(x, training=True, inference=False,
                                                dropout_mask=dropout_mask, eps=eps)
        loss = loss_fn(x, y, y_pred, y_vae, z_mean, z_logvar)
        loss = loss + reduce_sum(model.losses)
    macro_dice, micro_dice = dice_fn(y, y_pred)
    variables = model.trainable_variables
    grads = tape.gradient(loss, variables)
    if grad_hook is not None:
        grad_hook()
    optimizer.apply_gradients(zip(grads, variables), flat=model.flat)
    return loss.detach(), macro_dice, micro_dice


class GraphedTrainStep:
    """Captures train_step into a CUDA graph: static input buffers, one graph launch per step."""

    def __init__(self, model, optimizer, loss_fn, dice_fn, x, y, warmup=2, grad_hook=None):
        self.model, self.optimizer = model, optimizer
        self.x, self.y = x.clone(), y.clone()
        self._args = (model, optimizer, loss_fn, dice_fn)
        self._hook = grad_hook
        # everything that allocates persistent state must exist before capture
        flat = model.flatten_parameters() if model.built else None
        if flat is not None:
            optimizer._ensure_state(flat.theta.device)
            optimizer._mv(id(flat), flat.theta)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                train_step(*self._args, self.x, self.y, grad_hook=grad_hook)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        n0 = ops.LAUNCHES["n"]
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = train_step(*self._args, self.x, self.y, grad_hook=grad_hook)
        self.launches_per_step = ops.LAUNCHES["n"] - n0

    def __call__(self, x=None, y=None):
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if y is not None:
            self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        return self.out
