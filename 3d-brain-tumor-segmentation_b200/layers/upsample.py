"""Upsampling layers — surface of /root/reference/layers/upsample.py (:7-11 factory, :14-46 ConvUpsample)."""
from ..keras_compat import Layer, Conv3D, Conv3DTranspose, L2
from .group_norm import GroupNormalization


def get_upsampling(upsampling):
    if upsampling == 'linear':
        return LinearUpsample
    elif upsampling == 'conv':
        return ConvUpsample


class ConvUpsample(Layer):
    """Conv3DTranspose(k3, s2, 'same'; bias, glorot_uniform, no L2) -> GroupNorm -> ReLU.
    The transposed conv is the exact adjoint of the SAME/stride-2 conv."""

    def __init__(self,
                 filters,
                 groups=8,
                 data_format='channels_last',
                 l2_scale=1e-5,
                 **kwargs):
        super().__init__()
        self.config = super().get_config()
        self.config.update({'filters': filters,
                            'data_format': data_format,
                            'groups': groups,
                            'l2_scale': l2_scale})
        self.groups = groups
        self.data_format = data_format
        self.conv = Conv3DTranspose(filters=filters, kernel_size=3, strides=2, padding='same',
                                    data_format=data_format)
        self.norm = GroupNormalization(groups=groups, axis=-1 if data_format == 'channels_last' else 1)

    def build(self, input_shape, device):
        self.conv.build(input_shape, device)
        self.norm.build([input_shape[0]] + [2 * s for s in input_shape[1:-1]] + [self.conv.filters], device)
        self.built = True

    def call(self, inputs, training=None):
        h, st = self.conv.call(inputs, gn_groups=0 if self.norm.channel_mode else self.groups, aux=True)
        from .. import ops
        return self.norm.call(h, stats=st, relu=True, operand_only=ops.FUSED["on"])

    def get_config(self):
        return self.config


class LinearUpsample(Layer):
    """Conv3D(filters, 1x1x1; he_normal, L2, bias) -> UpSampling3D(size=2) (upsample.py:49-79).  Keras'
    UpSampling3D repeats voxels (nearest neighbour) despite the layer's name; no GroupNorm, no activation."""

    def __init__(self,
                 filters,
                 data_format='channels_last',
                 l2_scale=1e-5,
                 **kwargs):
        super().__init__()
        self.config = super().get_config()
        self.config.update({'filters': filters,
                            'data_format': data_format,
                            'l2_scale': l2_scale})
        self.data_format = data_format
        self.ptwise = Conv3D(filters=filters, kernel_size=1, strides=1, padding='same', data_format=data_format,
                             kernel_regularizer=L2(l2_scale), kernel_initializer='he_normal')

    def build(self, input_shape, device):
        self.ptwise.build(input_shape, device)
        self.built = True

    def call(self, inputs, training=None):
        from .. import ops
        return ops.upsample2(self.ptwise.call(inputs))

    def get_config(self):
        return self.config
