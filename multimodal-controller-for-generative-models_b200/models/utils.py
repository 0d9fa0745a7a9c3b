"""Parameter init and spectral normalisation of the hot-path models
(src/models/utils.py:7-21)."""
import torch
import torch.nn as nn

from ..config import cfg

_GAN_NAMES = ('cgan', 'mcgan')
_NORMS = (nn.BatchNorm1d, nn.BatchNorm2d)
_DENSE = (nn.Linear, nn.Conv2d, nn.ConvTranspose2d)


def init_param(m):
    """BN: weight ~ N(1, 0.02), bias 0.  GAN models: xavier-uniform (gain 1) on every
    Linear / Conv weight — through ``.weight.data``, so spectral-normed layers get it on the
    storage they share with ``weight_orig`` (src/models/utils.py:7-14)."""
    if isinstance(m, _NORMS):
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.zero_()
    if cfg['model_name'] in _GAN_NAMES and isinstance(m, _DENSE):
        nn.init.xavier_uniform_(m.weight.data, 1.)
    return m


def make_SpectralNormalization(m):
    """Legacy ``torch.nn.utils.spectral_norm`` on Linear / Conv layers (src/models/utils.py:17-21):
    creates ``weight_orig``, ``weight_u``, ``weight_v`` — the state_dict keys of the reference."""
    return torch.nn.utils.spectral_norm(m) if isinstance(m, _DENSE) else m
