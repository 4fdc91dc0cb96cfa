"""Loss, metric and optimiser with the reference's names and call signatures
(/root/reference/util.py:5-24 DiceVAELoss, :27-57 DiceCoefficient, :60-84 ScheduledOptim)."""
from __future__ import annotations

import torch

from . import ops


class DiceVAELoss(object):
    """soft-Dice (summed over batch and space, squared denominators, +1 smoothing) + 0.1*MSE(x, y_vae)
    + 0.1*mean(mu^2 + exp(logvar) - logvar - 1)  — one reduction pass (csrc/misc.cu)."""

    def __init__(self, name='custom_loss', data_format='channels_last', **kwargs):
        from .keras_compat import check_data_format
        self.data_format = check_data_format(data_format)
        self.axis = (0, 1, 2, 3) if data_format == 'channels_last' else (0, 2, 3, 4)
        # data parallelism with the reference's batch-global Dice: (process group, world size), set by
        # train.DataParallel(objective='global_batch'); None = this process's batch is the whole batch
        self.dp = None

    def __call__(self, x, y, y_pred, y_vae, z_mean, z_logvar, sample_weight=None):
        if self.data_format == 'channels_first':       # NCDHW arguments: the sums are layout-independent
            from .keras_compat import map5d
            x, y, y_pred, y_vae = (map5d(t, ops.to_channels_last) for t in (x, y, y_pred, y_vae))
        return ops.dice_vae_loss(x, y, y_pred, y_vae, z_mean, z_logvar, self.dp)


class DiceCoefficient(object):
    """Hard dice of one_hot(argmax)*[max>0.5]: (macro, micro); macro keeps the reference's
    un-reduced W axis (util.py:36,50-54)."""

    def __init__(self, name='dice_coefficient', data_format='channels_last'):
        from .keras_compat import check_data_format
        self.name = name
        self.data_format = check_data_format(data_format)

    def __call__(self, y_true, y_pred):
        if self.data_format == 'channels_first':       # util.py:36: axes (0,2,3,4) -> one ratio per class
            y_true, y_pred = ops.to_channels_last(y_true), ops.to_channels_last(y_pred.detach())
            return ops.dice_coefficient(y_true, y_pred, reduce_w=True)
        return ops.dice_coefficient(y_true, y_pred)


class ScheduledOptim(object):
    """tf.keras.optimizers.Adam semantics:
        alpha_t = lr*sqrt(1-b2^t)/(1-b1^t);  theta -= alpha_t * m / (sqrt(v) + eps),  eps = 1e-7,
    with the per-epoch polynomial schedule of util.py:82-84.  When the variables are a model's flat
    parameter buffer the whole update is ONE kernel over (theta, m, v, g)."""

    def __init__(self, learning_rate=1e-4, beta_1=0.9, beta_2=0.999, epsilon=1e-7, amsgrad=False,
                 name='Adam', n_epochs=300, **kwargs):
        if amsgrad:
            raise NotImplementedError("amsgrad")
        self.init_lr = float(learning_rate)
        self.beta_1, self.beta_2, self.epsilon = float(beta_1), float(beta_2), float(epsilon)
        self.n_epochs = float(n_epochs)
        self._lr = float(learning_rate)
        self._state = None        # device fp64 [2] = (iterations, lr)
        self._slots = {}          # id(tensor-or-flat) -> (m, v)
        self.grad_scale = 1.0     # data-parallel averaging factor folded into the update

    def __call__(self, epoch):
        new_lr = self.init_lr * ((1.0 - epoch / self.n_epochs) ** 0.9)
        self._set_hyper('learning_rate', new_lr)

    def _set_hyper(self, name, value):
        assert name == 'learning_rate'
        self._lr = float(value)
        if self._state is not None:
            self._state[1] = self._lr

    @property
    def learning_rate(self):
        return torch.tensor(self._lr, dtype=torch.float64)

    @property
    def iterations(self):
        return 0 if self._state is None else int(self._state[0].item())

    def _ensure_state(self, device):
        if self._state is None:
            self._state = torch.tensor([0.0, self._lr], dtype=torch.float64, device=device)

    def _mv(self, key, like):
        if key not in self._slots:
            self._slots[key] = (torch.zeros_like(like), torch.zeros_like(like))
        return self._slots[key]

    def apply_flat(self, flat, l2_in_step=False):
        """Fused update of a model's flat parameter buffer from its flat gradient buffer.  With
        `l2_in_step` the regulariser gradient 2*l*w of the first `reg_end` elements is added in the kernel
        (data-parallel mode: it must not be averaged over ranks)."""
        self._ensure_state(flat.theta.device)
        m, v = self._mv(id(flat), flat.theta)
        fused = l2_in_step and flat.l2 is not None
        if l2_in_step and not fused:
            # several L2 coefficients (Model(l2_scale != 1e-5)): the kernel fuses one range only, so the regulariser
            # gradient is added per group before the step, pre-multiplied by 1/grad_scale (the kernel scales g)
            flat.add_l2_grad(1.0 / float(self.grad_scale))
        ops._call("b3d_adam_step", flat.theta, m, v, flat.grad, self._state, self.beta_1, self.beta_2,
                  self.epsilon, float(self.grad_scale), 2.0 * float(flat.l2) if fused else 0.0,
                  int(flat.reg_end) if fused else 0, 1)
        ops.repack_all(flat)          # packed conv operands follow the weights (one launch)

    def apply_gradients(self, grads_and_vars, flat=None):
        gv = list(grads_and_vars)
        if flat is not None or _is_flat_group(gv):
            flat = flat or gv[0][1]._b3d_flat
            return self.apply_flat(flat)
        self._ensure_state(gv[0][1].device)
        for i, (g, var) in enumerate(gv):
            m, v = self._mv(id(var), var)
            th = var.detach().view(-1)
            ops._call("b3d_adam_step", th, m.view(-1), v.view(-1), g.contiguous().view(-1), self._state,
                      self.beta_1, self.beta_2, self.epsilon, float(self.grad_scale), 0.0, 0,
                      int(i == len(gv) - 1))
            vf = getattr(var, "_b3d_flat", None)
            if vf is not None:
                vf.epoch += 1         # packed conv operands of this model are stale (re-made at next use)


def _is_flat_group(gv):
    """True when (grads, vars) are exactly a model's flat parameter group with its own .grad views."""
    flat = getattr(gv[0][1], "_b3d_flat", None)
    if flat is None or len(gv) != len(flat.order):
        return False
    for g, v in gv:
        if getattr(v, "_b3d_flat", None) is not flat or g is None:
            return False
        off, n = flat.spans[id(v)]
        if g.data_ptr() != flat.grad.data_ptr() + 4 * off:
            return False
    return True
