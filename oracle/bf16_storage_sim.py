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


class _Patched:
    """Context manager that makes the oracles (MCGAN, MCVAE, PixelCNN) round like the bf16 pipeline:
    conv weights, conv inputs coming from an fp32 image (<= 4 channels: the im2col buffer is bf16),
    conv / conv-transpose outputs with more than 4 channels (RGB outputs stay fp32), and every
    4-D masked activation."""

    def __enter__(self):
        from . import mcpixelcnn_oracle as P
        from . import mcvae_oracle as V
        self.V, self.P = V, P
        self.conv2d, self.convt, self.mask_o, self.mask_v = F.conv2d, F.conv_transpose2d, O.mc_mask, V.mc_mask

        def conv2d(x, w, b=None, stride=1, padding=0):
            if w.shape[1] <= 4:
                x = _Round.apply(x)
            y = self.conv2d(x, _Round.apply(w), b, stride=stride, padding=padding)
            return _Round.apply(y) if w.shape[0] > 4 else y

        def conv_transpose2d(x, w, b=None, stride=1, padding=0):
            y = self.convt(x, _Round.apply(w), b, stride=stride, padding=padding)
            return _Round.apply(y) if w.shape[1] > 4 else y

        ns = type("FPatched", (), {k: getattr(F, k) for k in dir(F) if not k.startswith("__")})
        ns.conv2d = staticmethod(conv2d)
        ns.conv_transpose2d = staticmethod(conv_transpose2d)
        mask = lambda x, i, c: _Round.apply(self.mask_o(x, i, c)) if x.dim() > 2 else self.mask_o(x, i, c)  # noqa: E731
        O.F = V.F = P.F = ns
        O.mc_mask = V.mc_mask = P.mc_mask = mask
        return self

    def __exit__(self, *a):
        O.F = self.V.F = self.P.F = F
        O.mc_mask, self.V.mc_mask, self.P.mc_mask = self.mask_o, self.mask_v, self.mask_o


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


def vae_gradients(sd, img, label, num_mode, eps, simulate):
    """MCVAE loss and parameter gradients, with or without simulated bf16 storage."""
    from . import mcvae_oracle as V
    sd = V.make_trainable(sd)
    if simulate:
        with _Patched():
            out = V.forward(img, label, sd, num_mode, eps)
            out["loss"].backward()
    else:
        out = V.forward(img, label, sd, num_mode, eps)
        out["loss"].backward()
    return float(out["loss"].detach()), {k: v.grad for k, v in sd.items() if v.grad is not None}
