"""ORACLE — TEST INFRASTRUCTURE ONLY.

Derives the bf16 tolerance the GPU tests state: runs the fp32 oracle with a bf16
round-trip (forward value AND incoming gradient) at every tensor the B200 pipeline
stores in bf16 — conv outputs, gated activations, conv weights — and reports how far that
storage format alone moves losses and parameter gradients away from the fp32 oracle.
Nothing here is a kernel; it only answers "what can bf16 storage cost on this network?".
"""
import torch
import torch.nn.functional as F

from . import mcgan_oracle as O


class _Round(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def _wide(w):
    return w.shape[0] > 4 and w.shape[1] > 4


class _Patched:
    """Context manager that makes the oracle round like the bf16 pipeline."""

    def __enter__(self):
        self.conv2d, self.mask = F.conv2d, O.mc_mask

        def conv2d(x, w, b=None, padding=0):
            w = _Round.apply(w)
            return _Round.apply(self.conv2d(x, w, b, padding=padding)) if _wide(w) or w.shape[0] > 4 \
                else self.conv2d(x, w, b, padding=padding)

        ns = type("FPatched", (), {k: getattr(F, k) for k in dir(F) if not k.startswith("__")})
        ns.conv2d = staticmethod(conv2d)
        O.F = ns
        O.mc_mask = lambda x, i, c: _Round.apply(self.mask(x, i, c))
        return self

    def __exit__(self, *a):
        O.F = F
        O.mc_mask = self.mask


def gradients(sd, z, label, num_mode, strides, simulate):
    """-D(G(z)).mean() and its parameter gradients, with or without simulated bf16 storage."""
    sd = O.make_trainable(sd)
    ind = O.one_hot(label, num_mode)
    if simulate:
        with _Patched():
            loss = -O.discriminator(O.generator(z, ind, sd), ind, sd, strides).mean()
            loss.backward()
    else:
        loss = -O.discriminator(O.generator(z, ind, sd), ind, sd, strides).mean()
        loss.backward()
    return float(loss.detach()), {k: v.grad for k, v in sd.items() if v.grad is not None}
