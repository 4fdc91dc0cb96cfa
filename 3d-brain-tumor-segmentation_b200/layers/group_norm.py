"""GroupNormalization with the reference's constructor and call surface
(/root/reference/layers/group_norm.py:10-40 ctor, :42-81 build, :83-124 call), backed by the fused
chunk-norm kernels of csrc/norm.cu.

NB: with axis=-1 the reference reshapes [B,D,H,W,C] -> [B,G,D,H,W,C/G] WITHOUT a transpose,
so "group g" is the g-th contiguous 1/G chunk of each sample's flat buffer.  That behaviour is the
parity target and is what the kernels implement.
"""
from ..keras_compat import Layer
from .. import ops


class GroupNormalization(Layer):
    def __init__(self,
                 groups=8,
                 axis=-1,
                 epsilon=1e-5,
                 center=True,
                 scale=True,
                 beta_initializer='zeros',
                 gamma_initializer='ones',
                 beta_regularizer=None,
                 gamma_regularizer=None,
                 beta_constraint=None,
                 gamma_constraint=None,
                 **kwargs):
        super().__init__(**kwargs)
        self.supports_masking = True
        self.groups, self.axis, self.epsilon = groups, axis, epsilon
        # axis=1 = the reference's channels_first construction (true channel groups, NCDHW public tensors)
        self.channel_mode = axis == 1
        self.data_format = 'channels_first' if axis == 1 else 'channels_last'
        self.center, self.scale = center, scale
        self.beta_initializer, self.gamma_initializer = beta_initializer, gamma_initializer
        self.beta_regularizer, self.gamma_regularizer = beta_regularizer, gamma_regularizer
        self.beta_constraint, self.gamma_constraint = beta_constraint, gamma_constraint
        self.gamma = self.beta = None

    def build(self, input_shape, device):
        if self.axis not in (-1, 1, len(input_shape) - 1):
            raise NotImplementedError("b3d GroupNormalization: axis must be -1 (channels_last) or 1 (channels_first)")
        dim = input_shape[-1]                      # storage is NDHWC in both cases (Layer.__call__ converts)
        if dim is None:
            raise ValueError('Axis ' + str(self.axis) + ' of input tensor should have a defined dimension '
                             'but the layer received an input with shape ' + str(input_shape) + '.')
        if dim < self.groups:
            raise ValueError('Number of groups (' + str(self.groups) + ') cannot be '
                             'more than the number of channels (' + str(dim) + ').')
        if dim % self.groups != 0:
            raise ValueError('Number of groups (' + str(self.groups) + ') must be a '
                             'multiple of the number of channels (' + str(dim) + ').')
        if self.scale:
            self.gamma = self.add_weight('gamma', (dim,), self.gamma_initializer, device, self.gamma_regularizer)
        else:
            import torch
            self._gamma_const = torch.ones(dim, device=device)
        if self.center:
            self.beta = self.add_weight('beta', (dim,), self.beta_initializer, device, self.beta_regularizer)
        else:
            import torch
            self._beta_const = torch.zeros(dim, device=device)
        self.built = True

    def call(self, inputs, training=None, stats=None, relu=False, operand_only=False, **kwargs):
        """operand_only: every consumer of the result is a conv, which reads its 16-bit P16 twin — the fp32 form is not
        materialised (ops.virtual).  Only the fused callers (ResnetBlock, resampling layers inside a Model) ask for it."""
        gamma = self.gamma if self.scale else self._gamma_const
        beta = self.beta if self.center else self._beta_const
        return ops.group_norm(inputs, gamma, beta, None if self.channel_mode else stats, self.groups, self.epsilon,
                              relu, channel_mode=self.channel_mode, operand_only=operand_only)

    def get_config(self):
        config = {
            'groups': self.groups, 'axis': self.axis, 'epsilon': self.epsilon,
            'center': self.center, 'scale': self.scale,
            'beta_initializer': self.beta_initializer, 'gamma_initializer': self.gamma_initializer,
            'beta_regularizer': None if self.beta_regularizer is None else {'l2': self.beta_regularizer.l},
            'gamma_regularizer': None if self.gamma_regularizer is None else {'l2': self.gamma_regularizer.l},
            'beta_constraint': self.beta_constraint, 'gamma_constraint': self.gamma_constraint,
        }
        base = super().get_config()
        return dict(list(base.items()) + list(config.items()))

    def compute_output_shape(self, input_shape):
        return input_shape
