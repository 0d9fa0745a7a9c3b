"""The MCVAE training step (src/train_vae.py:98-126, core :106-111) on the B200 kernels:
zero_grad → forward (BCE + KLD) → backward → clip_grad_norm_(1) → Adam(3e-4)."""
import torch

VAE_HYPER = {"lr": 3e-4, "weight_decay": 0, "clip": 1.0}      # src/train_vae.py:30-36, 110


def make_optimizer(model, fused=None):
    """Adam lr 3e-4 (src/train_vae.py:33, 160-173).  On CUDA: ``optim.FusedAdam`` with the
    ``clip_grad_norm_(1)`` of :110 folded into the update (no host round trip)."""
    params = list(model.parameters())
    if fused is None:
        fused = bool(params) and params[0].is_cuda
    if fused:
        from .optim import FusedAdam
        return FusedAdam(params, lr=VAE_HYPER["lr"], weight_decay=VAE_HYPER["weight_decay"], clip_norm=VAE_HYPER["clip"])
    return torch.optim.Adam(params, lr=VAE_HYPER["lr"], weight_decay=VAE_HYPER["weight_decay"])


def train_step(model, optimizer, input, eps=None):
    """One reference step on ``{'img' in [-1,1] [N,C,H,W], 'label' [N]}``; ``eps`` injects the
    reparameterisation noise.  Returns the model's output dict (device tensors, no host sync)."""
    optimizer.zero_grad()
    output = model(input, eps) if eps is not None else model(input)
    output["loss"].backward()
    if getattr(optimizer, "clip_norm", None) is None:        # FusedAdam clips inside its step
        torch.nn.utils.clip_grad_norm_(model.parameters(), VAE_HYPER["clip"])
    optimizer.step()
    return output
