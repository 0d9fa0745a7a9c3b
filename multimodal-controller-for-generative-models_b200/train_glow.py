"""The MCGlow training step and its data-dependent initialisation (src/train_glow.py) on the B200 kernels.

``init_actnorm`` restates src/train_glow.py:60-67: before the optimiser exists, the first
``cfg['num_init_batches']`` (8) training batches are concatenated and pushed through the model once under
``no_grad`` in training mode — every ActNorm (144 of them at K=16, L=3) sets ``loc = -mean`` and
``scale = 1 / (std + 1e-6)`` from the activations it sees (src/models/mcglow.py:32-41).  The step itself
(:116-121: zero_grad → forward → backward → clip_grad_norm_(1) → Adam) is ``train_vae.train_step``.
"""
from itertools import islice

import torch

from .train_vae import GraphedTrainStep, make_optimizer, train_step  # noqa: F401  (same step as the VAE's)

NUM_INIT_BATCHES = 8          # src/config.yml: num_init_batches


def init_actnorm(model, batches, num_init_batches=NUM_INIT_BATCHES, noise=None):
    """batches: an iterable of ``{'img': [n,C,H,W], 'label': [n]}`` dicts (a data loader) or one such dict.
    Runs the data-dependent ActNorm initialisation on the first ``num_init_batches`` of them, concatenated
    (src/train_glow.py:60-67).  ``noise`` injects the dequantisation noise of that forward (parity tests).
    Returns the model output of the initialisation pass (its loss is the 'init loss' of the reference run)."""
    if isinstance(batches, dict):
        batches = [batches]
    taken = list(islice(iter(batches), None, num_init_batches))
    if not taken:
        raise ValueError('Not valid init batches')
    dev = next(model.parameters()).device
    init = {k: torch.cat([b[k].to(dev) for b in taken], 0) for k in ('img', 'label')}
    was_training = model.training
    model.train(True)
    try:
        with torch.no_grad():
            out = model(init, noise) if noise is not None else model(init)
    finally:
        model.train(was_training)
    pending = [m for m in model.modules() if hasattr(m, 'needs_init') and m.needs_init()]
    if pending:
        raise RuntimeError(f'{len(pending)} ActNorm layers were not reached by the initialisation pass')
    return out
