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


class GraphedTrainStep:
    """``train_step`` of MCVAE / MCGlow captured ONCE into a CUDA graph and replayed per batch.

    A Glow step is ~48 flow steps x ~40 launches forward and as many backward, a VAE step a few hundred
    launches of 5-50 us kernels: issued eagerly both are bound by the host, not the B200.  The step has
    static shapes and — with ``optim.FusedAdam`` (device-side step counter and clip coefficient) — never
    synchronises with the host, so it is captured after a few eager warm-up steps (which build every workspace,
    packed codebook and TMA descriptor; the model / optimiser state they touch is rolled back afterwards, so
    Glow's data-dependent ActNorm initialisation must already have happened: ``train_glow.init_actnorm``) and
    replayed with new data copied into static input buffers.  ``noise``: optional static buffer for the
    injected reparameterisation (VAE) / dequantisation (Glow) noise of parity runs; otherwise the noise is
    drawn inside the graph (torch's graph-safe Philox stream).  No autograd graph of an earlier eager step on
    another stream may still be alive (drop the step's output dict first): its AccumulateGrad nodes would tie
    the capture to that stream.
    """

    def __init__(self, model, optimizer, example_input, warmup=3, inject_noise=None):
        if not hasattr(optimizer, "flat"):
            raise ValueError("Not valid optimizer: graph capture needs optim.FusedAdam (no host sync in the step)")
        self.img = example_input["img"].detach().clone()
        self.label = example_input["label"].detach().clone()
        self.noise = inject_noise.detach().clone() if inject_noise is not None else None
        self.model, self.optimizer = model, optimizer
        batch = {"img": self.img, "label": self.label}
        from .optim import StateGuard
        for m in model.modules():
            if hasattr(m, "needs_init") and m.needs_init():
                raise ValueError("Not valid model state: run the data-dependent ActNorm initialisation first "
                                 "(train_glow.init_actnorm, src/train_glow.py:60-67) — the warm-up steps of the "
                                 "capture are rolled back and must not be the ones that initialise it")
        guard = StateGuard(model, [optimizer])
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                train_step(model, optimizer, batch, self.noise)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import _lib
        c0 = _lib.launch_count
        from . import ops as _ops
        _ops.clear_pack_caches()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = train_step(model, optimizer, batch, self.noise)
        self.native_launches = _lib.launch_count - c0
        guard.restore()

    def __call__(self, input, noise=None):
        """input: {'img','label'} on the host (pinned) or the device; returns the output dict of device
        tensors, overwritten by the next call."""
        self.img.copy_(input["img"], non_blocking=True)
        self.label.copy_(input["label"], non_blocking=True)
        if self.noise is not None and noise is not None:
            self.noise.copy_(noise, non_blocking=True)
        self.optimizer.sync_hyper()          # lr schedulers: the captured Adam reads lr from a staged device buffer
        self.graph.replay()
        torch.autograd.graph.increment_version(self.optimizer.flat.params)     # packed-weight caches: re-pack on eager use
        return self.out
