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


# --------------------------------------------------------------------------- sampling utilities (codebook side)
# The callers of the MultimodalController boundary that REPLACE its ``codebook`` buffer:
# generate.py / create.py / transit.py of the reference go through src/models/utils.py:34-152.  For the
# mc* families the only state they touch is the codebook; the packed device copy the kernels read is re-derived
# by ``MultimodalController.packed_bits()`` whenever the buffer object changes, so these work unchanged on the B200
# path.  (The embedding branches of the reference serve the c* baselines, which are out of scope.)
def _is_controller(module):
    return module.__class__.__name__ == 'MultimodalController'       # by NAME, as the reference selects them


def create_codebook(codebook):
    """A fresh random binary codebook ``[cfg['classes_size'], K]`` for "creating" unseen modes
    (src/models/utils.py:34-44): Bernoulli(0.5) blocks of classes_size rows from the global RNG, rows
    de-duplicated through a Python set, the first classes_size entries of ``list(set)`` — bit-identical to
    the reference for the same seed (tests/test_sampling_host.py)."""
    n, width = cfg['classes_size'], codebook.size(1)
    draw = torch.distributions.bernoulli.Bernoulli(probs=0.5)
    rows = set()
    while len(rows) < n:
        rows.update(tuple(r) for r in draw.sample((n, width)).tolist())
    return torch.tensor(list(rows)[:n], dtype=torch.float).to(cfg['device'])


def transit_codebook(codebook, root, alpha):
    """Move every mode towards mode ``root`` (src/models/utils.py:100-110): the first
    round((1 - alpha) * K) channels of each row take the root row's bits; the root row is unchanged."""
    cross = int(round((1 - alpha) * codebook.size(1)))
    out = codebook.detach().clone().cpu()
    out[:, :cross] = out[root, :cross].clone()
    return out.to(cfg['device'])


def create(model):
    """Give every MultimodalController of ``model`` a fresh random codebook (src/models/utils.py:47-97)."""
    for module in model.modules():
        if _is_controller(module):
            module.register_buffer('codebook', create_codebook(module.codebook))


def transit(model, root, alpha):
    """Transit every controller's codebook towards mode ``root`` (src/models/utils.py:113-152); the trained
    codebook is kept in ``codebook_orig`` so repeated calls always start from it."""
    for module in model.modules():
        if _is_controller(module):
            if not hasattr(module, 'codebook_orig'):
                module.register_buffer('codebook_orig', module.codebook.data)
            module.register_buffer('codebook', transit_codebook(module.codebook_orig, root, alpha))
