"""Layer classes with the reference's names and constructor/call surfaces (/root/reference/layers/)."""
from . import group_norm, resnet, downsample, upsample, encoder, decoder, vae  # noqa: F401
from .group_norm import GroupNormalization  # noqa: F401
from .resnet import ResnetBlock  # noqa: F401
from .downsample import ConvDownsample, MaxDownsample, get_downsampling  # noqa: F401
from .upsample import ConvUpsample, LinearUpsample, get_upsampling  # noqa: F401
from .encoder import Encoder  # noqa: F401
from .decoder import Decoder  # noqa: F401
from .vae import VariationalAutoencoder  # noqa: F401
