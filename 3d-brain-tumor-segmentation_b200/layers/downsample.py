"""Downsampling layers — surface of /root/reference/layers/downsample.py (:7-11 factory, :14-48 ConvDownsample)."""
from ..keras_compat import Layer, Conv3D, L2
from .group_norm import GroupNormalization


def get_downsampling(downsampling):
    if downsampling == 'max':
        return MaxDownsample
    elif downsampling == 'conv':
        return ConvDownsample


class ConvDownsample(Layer):
    """Conv3D(k3, s2, TF 'same' => pad_before 0 / pad_after 1) -> GroupNorm -> ReLU."""

    def __init__(self,
                 filters,
                 data_format='channels_last',
                 groups=8,
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
        self.conv = Conv3D(filters=filters, kernel_size=3, strides=2, padding='same', data_format=data_format,
                           kernel_regularizer=L2(l2_scale), kernel_initializer='he_normal')
        self.norm = GroupNormalization(groups=groups, axis=-1 if data_format == 'channels_last' else 1)

    def build(self, input_shape, device):
        self.conv.build(input_shape, device)
        self.norm.build([input_shape[0]] + [s // 2 for s in input_shape[1:-1]] + [self.conv.filters], device)
        self.built = True

    def call(self, inputs, training=None):
        h, st, _ = self.conv.call(inputs, gn_groups=0 if self.norm.channel_mode else self.groups, aux=True)
        from .. import ops
        return self.norm.call(h, stats=st, relu=True, operand_only=ops.FUSED["on"])

    def get_config(self):
        return self.config


class MaxDownsample(Layer):
    """MaxPooling3D(pool_size=2, strides=2, padding='same') (downsample.py:51-70): no weights, the channel count is
    kept (the `filters` the Encoder passes is swallowed by **kwargs, as in the reference)."""

    def __init__(self,
                 data_format='channels_last',
                 **kwargs):
        super().__init__()
        from ..keras_compat import check_data_format
        self.data_format = check_data_format(data_format)
        self.config = super().get_config()
        self.config.update({'data_format': data_format})

    def call(self, inputs, training=None):
        from .. import ops
        return ops.max_pool2(inputs)

    def get_config(self):
        return self.config
