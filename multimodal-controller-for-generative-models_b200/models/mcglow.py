"""MCGlow on the B200 kernels — same module tree, ``state_dict`` keys and
``forward / reverse / generate`` API as the reference (src/models/mcglow.py).

Where the work is.  Per flow step the coupling network
    Conv3x3(C/2→512) → ActNorm → ReLU → MC → Conv1x1(512→512) → ActNorm → ReLU → MC → ZeroConv3x3(512→C)
carries > 99.9 % of the flops and bytes; it runs on the tcgen05 conv kernels with
ActNorm → ReLU → MC fused into ONE gate kernel pair (``ops.ActNormGateFn``).  The few-channel
(C ≤ 48) arithmetic around it runs on fp32 NHWC tensors through two native kernel pairs (csrc/flow.cu):
``ops.FlowFrontFn`` = ActNorm(C) → invertible 1x1 conv, which also emits the coupling network's bf16
input, and ``ops.FlowTailFn`` = ZeroConv2d's exp(3·scale) → sigmoid → affine coupling → per-sample
logdet.  Only parameter-sized work (the C×C LU product, Σ log|scale|), the squeeze / split reshapes
and the three Gaussian priors per forward stay in torch.  The sampling path (``reverse``) is the
reference's NCHW formulation on the same coupling-network kernels.

Channel padding: the tensor-core kernels take input channels in multiples of 64 and output channels
in {<=16, 32, 64k}; C/2 ∈ {6, 12, 24} inputs are zero-padded to 64 channels and C ∈ {12, 24, 48}
outputs are computed as 64 channels (zero weight rows) and sliced.

ActNorm's data-dependent initialisation (mcglow.py:32-41) is kept; its per-forward
``initialized.item()`` host synchronisation (144 per step in the reference) is replaced by a host-side
flag that is refreshed from the buffer only when the module has not seen an initialised forward yet.
"""
from math import log, pi

import numpy as np
import scipy.linalg as la
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..config import cfg
from ..modules import MultimodalController, Wrapper
from .mcgan import _Ctx, _labels_from_indicator
from .utils import init_param


def logabs(x):
    return torch.log(torch.abs(x))


def gaussian_log_p(x, mean, log_sd):
    return -0.5 * log(2 * pi) - log_sd - 0.5 * (x - mean) ** 2 / torch.exp(2 * log_sd)


def gaussian_sample(eps, mean, log_sd):
    return mean + torch.exp(log_sd) * eps


class ActNorm(nn.Module):
    """Per-channel affine with data-dependent init (src/models/mcglow.py:24-55)."""

    def __init__(self, input_size, logdet=True):
        super().__init__()
        self.loc = nn.Parameter(torch.zeros(1, input_size, 1, 1))
        self.scale = nn.Parameter(torch.ones(1, input_size, 1, 1))
        self.register_buffer('initialized', torch.tensor(0, dtype=torch.uint8))
        self.logdet = logdet
        self._init_seen = False

    def needs_init(self):
        if self._init_seen:
            return False
        if int(self.initialized.item()) != 0:           # one host sync, then never again
            self._init_seen = True
            return False
        return True

    def initialize_from(self, mean, std):
        """mean / std: per-channel statistics [C] of the tensor that enters this ActNorm."""
        if not self.training:
            return
        with torch.no_grad():
            self.loc.data.copy_(-mean.view_as(self.loc))
            self.scale.data.copy_(1 / (std.view_as(self.scale) + 1e-6))
            self.initialized.fill_(1)
        self._init_seen = True

    def forward(self, input):
        """NCHW fp32 path (the few-channel flow variables)."""
        height, width = input.size(2), input.size(3)
        if self.needs_init():
            self.initialize_from(torch.mean(input, dim=[0, 2, 3]), torch.std(input, dim=[0, 2, 3]))
        out = self.scale * (input + self.loc)
        if self.logdet:
            return out, height * width * torch.sum(logabs(self.scale))
        return out

    def reverse(self, output):
        return output / self.scale - self.loc


class InvConv2dLU(nn.Module):
    """Invertible 1x1 conv, W = P (L∘l_mask + I) (U∘u_mask + diag(sign·exp(w_s))) (mcglow.py:76-116)."""

    def __init__(self, input_size):
        super().__init__()
        q, _ = la.qr(np.random.randn(input_size, input_size))
        w_p, w_l, w_u = la.lu(q.astype(np.float32))
        w_s = np.diag(w_u)
        w_u = np.triu(w_u, 1)
        u_mask = np.triu(np.ones_like(w_u), 1)
        l_mask = u_mask.T
        w_s = torch.from_numpy(w_s)
        self.register_buffer('w_p', torch.from_numpy(w_p))
        self.register_buffer('u_mask', torch.from_numpy(u_mask))
        self.register_buffer('l_mask', torch.from_numpy(l_mask))
        self.register_buffer('s_sign', torch.sign(w_s))
        self.register_buffer('l_eye', torch.eye(l_mask.shape[0]))
        self.w_l = nn.Parameter(torch.from_numpy(w_l))
        self.w_s = nn.Parameter(logabs(w_s))
        self.w_u = nn.Parameter(torch.from_numpy(w_u))

    def calc_weight(self):
        lower = self.w_l * self.l_mask + self.l_eye
        upper = self.w_u * self.u_mask + torch.diag(self.s_sign * torch.exp(self.w_s))
        return (self.w_p @ lower @ upper).unsqueeze(2).unsqueeze(3)

    def forward(self, input):
        height, width = input.size(2), input.size(3)
        out = F.conv2d(input, self.calc_weight())
        return out, height * width * torch.sum(self.w_s)

    def reverse(self, output):
        return F.conv2d(output, self.calc_weight().squeeze().inverse().unsqueeze(2).unsqueeze(3))


class ZeroConv2d(nn.Module):
    """Zero-initialised Conv3x3 times exp(3·scale) (src/models/mcglow.py:119-130)."""

    def __init__(self, input_size, output_size, kernel_size=3, stride=1, padding=1):
        super().__init__()
        self.conv = nn.Conv2d(input_size, output_size, kernel_size, stride, padding)
        self.conv.weight.data.zero_()
        self.conv.bias.data.zero_()
        self.scale = nn.Parameter(torch.zeros(1, output_size, 1, 1))

    def forward(self, input):
        """NCHW reference formulation (kept for outside callers; the model itself uses ``forward_nhwc``)."""
        return self.conv(input) * torch.exp(self.scale * 3)

    def forward_nhwc(self, x):
        """fp32 NHWC [N,H,W,Cin] (Cin <= 64) -> fp32 NHWC [N,H,W,Cout] on the tcgen05 conv: input channels
        zero-padded to 64 (bf16 operand), output channels computed as a multiple of 64 (zero weight rows) and
        sliced; autograd reaches conv.weight / conv.bias / scale through the pads."""
        cout = self.conv.out_channels
        cp = (cout + 63) // 64 * 64
        xa = _pad_to(x, 3, 64).to(torch.bfloat16).contiguous()
        w = _pad_to(_pad_to(self.conv.weight, 1, 64), 0, cp)
        h = _conv(xa, w, _pad_to(self.conv.bias, 0, cp), True)[..., :cout]
        return h * torch.exp(self.scale.view(1, 1, 1, -1) * 3)

    def of_zeros(self, n, h, w):
        """``forward(zeros)`` in NHWC without a conv: a conv of an all-zero input is its bias."""
        v = self.conv.bias.view(1, 1, 1, -1) * torch.exp(self.scale.view(1, 1, 1, -1) * 3)
        return v.expand(n, h, w, -1)


def _pad_to(t, dim, size):
    if t.shape[dim] == size:
        return t
    shape = list(t.shape)
    shape[dim] = size - t.shape[dim]
    return torch.cat([t, t.new_zeros(shape)], dim)


def _conv(x, w, b, out_f32=False):
    return ops.ConvFn.apply(x, w, b, None, False, None, False, False, None, None, out_f32)


class AffineCoupling(nn.Module):
    def __init__(self, input_size, hidden_size=512, affine=True, num_mode=None, controller_rate=None):
        super().__init__()
        self.affine = affine
        self.net = nn.Sequential(
            Wrapper(nn.Conv2d(input_size // 2, hidden_size, 3, padding=1)),
            Wrapper(ActNorm(hidden_size, logdet=False)),
            Wrapper(nn.ReLU(inplace=True)),
            MultimodalController(hidden_size, num_mode, controller_rate),
            Wrapper(nn.Conv2d(hidden_size, hidden_size, 1)),
            Wrapper(ActNorm(hidden_size, logdet=False)),
            Wrapper(nn.ReLU(inplace=True)),
            MultimodalController(hidden_size, num_mode, controller_rate),
            Wrapper(ZeroConv2d(hidden_size, input_size if self.affine else input_size // 2)),
        )
        self.net[0].module.weight.data.normal_(0, 0.05)
        self.net[0].module.bias.data.zero_()
        self.net[4].module.weight.data.normal_(0, 0.05)
        self.net[4].module.bias.data.zero_()

    def _gate(self, h, actnorm, mc, ctx):
        """ActNorm(logdet=False) → ReLU → MC as one fused kernel; data-dependent init on first use."""
        if actnorm.needs_init():
            hf = h.detach().float()
            actnorm.initialize_from(hf.mean(dim=(0, 1, 2)), hf.std(dim=(0, 1, 2)))
        return ops.ActNormGateFn.apply(h, actnorm.scale, actnorm.loc, ctx.labels, mc.packed_bits())

    def net_nhwc(self, xa, ctx):
        """xa: bf16 NHWC [N,H,W,64] (C/2 live channels) -> RAW zero-conv output fp32 NHWC [N,H,W,64]
        (C or C/2 live channels; ZeroConv2d's exp(3*scale) is applied by the consumer)."""
        conv1, conv2, zero = self.net[0].module, self.net[4].module, self.net[8].module
        h = _conv(xa, _pad_to(conv1.weight, 1, 64), conv1.bias)
        h = self._gate(h, self.net[1].module, self.net[3], ctx)
        h = _conv(h, conv2.weight, conv2.bias)
        h = self._gate(h, self.net[5].module, self.net[7], ctx)
        return _conv(h, _pad_to(zero.conv.weight, 0, 64), _pad_to(zero.conv.bias, 0, 64), True)

    def run_net(self, in_a, ctx):
        """in_a: NCHW fp32 [N, C/2, H, W] -> net output NCHW fp32 [N, C_out, H, W] (sampling path)."""
        zero = self.net[8].module
        c_out = zero.conv.out_channels
        x = _pad_to(in_a.permute(0, 2, 3, 1), 3, 64).to(torch.bfloat16).contiguous()      # NHWC, Cin -> 64
        out = self.net_nhwc(x, ctx)[..., :c_out].permute(0, 3, 1, 2)
        return out * torch.exp(zero.scale * 3)

    def forward(self, y, xa, ctx):
        """y: fp32 NHWC [N,H,W,C] (the flow variable after the invertible conv), xa: its first half as the
        network's bf16 input (both from ``ops.FlowFrontFn``) -> (out NHWC, logdet [N] or None)."""
        out, logdet = ops.FlowTailFn.apply(y, self.net_nhwc(xa, ctx), self.net[8].module.scale, self.affine)
        return out, (logdet if self.affine else None)

    def reverse(self, output, ctx):
        """NCHW reference formulation (src/models/mcglow.py:166-175) on the coupling-network kernels."""
        out_a, out_b = output.chunk(2, 1)
        if self.affine:
            log_s, t = self.run_net(out_a, ctx).chunk(2, 1)
            in_b = out_b / torch.sigmoid(log_s + 2) - t
        else:
            in_b = out_b - self.run_net(out_a, ctx)
        return torch.cat([out_a, in_b], 1)

    def reverse_nhwc(self, output, xa, ctx):
        """output fp32 NHWC [N,H,W,C], xa = its first half as the network's bf16 input -> y fp32 NHWC."""
        return ops.flow_tail_rev(output, self.net_nhwc(xa, ctx), self.net[8].module.scale.detach().reshape(-1).contiguous(),
                                 self.affine)


class Flow(nn.Module):
    def __init__(self, input_size, hidden_size, affine=True, conv_lu=True, num_mode=None, controller_rate=None):
        super().__init__()
        if not conv_lu:
            raise ValueError('Not valid flow: only the LU-parameterised invertible conv (cfg conv_lu=True) is built')
        self.actnorm = ActNorm(input_size)
        self.invconv = InvConv2dLU(input_size)
        self.coupling = AffineCoupling(input_size, hidden_size, affine, num_mode, controller_rate)

    def forward(self, input, ctx):
        """input: fp32 NHWC [N,H,W,C] -> (out NHWC, logdet [N])   (src/models/mcglow.py:186-193)."""
        an, hw = self.actnorm, input.size(1) * input.size(2)
        if an.needs_init():
            xf = input.detach()
            an.initialize_from(xf.mean(dim=(0, 1, 2)), xf.std(dim=(0, 1, 2)))
        y, xa = ops.FlowFrontFn.apply(input, an.loc, an.scale, self.invconv.calc_weight()[:, :, 0, 0])
        out, det2 = self.coupling(y, xa, ctx)
        logdet = hw * (torch.sum(logabs(an.scale)) + torch.sum(self.invconv.w_s))
        if det2 is not None:
            logdet = logdet + det2
        return out, logdet

    def reverse(self, output, ctx):
        return self.actnorm.reverse(self.invconv.reverse(self.coupling.reverse(output, ctx)))

    def reverse_nhwc(self, output, xa, ctx):
        """Flow.reverse (src/models/mcglow.py:195-201) on fp32 NHWC through the native kernels (no autograd: this
        is the sampling path).  -> (x, bf16 network input of the flow that is inverted next)."""
        y = self.coupling.reverse_nhwc(output, xa, ctx)
        an = self.actnorm
        W = self.invconv.calc_weight()[:, :, 0, 0].detach()
        scale, loc = an.scale.detach().reshape(-1), an.loc.detach().reshape(-1)
        # x = (W^-1 y) / scale - loc  ==  W' (1 * (y + loc'))   with  W' = diag(1/scale) W^-1,  loc' = -W (scale * loc)
        w_rev = (torch.linalg.inv(W) / scale[:, None]).contiguous()
        loc_rev = (-(W @ (scale * loc))).contiguous()
        return ops.flow_front(y.contiguous(), loc_rev, torch.ones_like(scale), w_rev)


class Block(nn.Module):
    def __init__(self, input_size, hidden_size, K, split=True, affine=True, conv_lu=True, num_mode=None,
                 controller_rate=None):
        super().__init__()
        squeeze_dim = input_size * 4
        self.flows = nn.ModuleList([Flow(squeeze_dim, hidden_size, affine, conv_lu, num_mode, controller_rate)
                                    for _ in range(K)])
        self.split = split
        self.prior = ZeroConv2d(input_size * 2, input_size * 4) if split else ZeroConv2d(input_size * 4, input_size * 8)

    def forward(self, input, ctx):
        """input: fp32 NHWC [N,H,W,c] -> (out NHWC, logdet, log_p, z_new NCHW)   (mcglow.py:213-239).
        Squeeze keeps the reference's channel order c*4 + i*2 + j."""
        b, h, w, c = input.shape
        out = input.reshape(b, h // 2, 2, w // 2, 2, c).permute(0, 1, 3, 5, 2, 4).reshape(b, h // 2, w // 2, c * 4)
        logdet = 0
        for flow in self.flows:
            out, det = flow(out, ctx)
            logdet = logdet + det
        if self.split:
            out, z_nhwc = out.chunk(2, 3)
            mean, log_sd = self.prior.forward_nhwc(out).chunk(2, 3)        # ZeroConv2d on the tcgen05 conv
        else:
            z_nhwc = out
            mean, log_sd = self.prior.of_zeros(b, h // 2, w // 2).chunk(2, 3)   # prior(zeros) = its bias
        log_p = gaussian_log_p(z_nhwc, mean, log_sd).reshape(b, -1).sum(1)
        return out, logdet, log_p, z_nhwc.permute(0, 3, 1, 2)             # z_new: NCHW view of the (small) NHWC half

    def reverse(self, output, ctx, eps=None, reconstruct=False):
        """Block.reverse (src/models/mcglow.py:241-265) on fp32 NHWC tensors: output / eps [N,h,w,c] ->
        [N,2h,2w,c'] (un-squeezed).  Sampling path: runs without autograd."""
        x = output
        if reconstruct:
            x = torch.cat([output, eps], 3) if self.split else eps
        elif self.split:
            mean, log_sd = self.prior.forward_nhwc(x).chunk(2, 3)
            x = torch.cat([output, gaussian_sample(eps, mean, log_sd)], 3)
        else:
            mean, log_sd = self.prior.of_zeros(*x.shape[:3]).chunk(2, 3)
            x = gaussian_sample(eps, mean, log_sd)
        x = x.contiguous()
        xa = _pad_to(x[..., :x.shape[3] // 2], 3, ops.FLOW_XA_CH).to(torch.bfloat16).contiguous()
        for flow in self.flows[::-1]:
            x, xa = flow.reverse_nhwc(x, xa, ctx)
        b, h, w, c = x.shape
        return x.reshape(b, h, w, c // 4, 2, 2).permute(0, 1, 4, 2, 5, 3).reshape(b, h * 2, w * 2, c // 4)


class MCGlow(nn.Module):
    def __init__(self, data_shape, hidden_size, K, L, affine=True, conv_lu=True, num_mode=None, controller_rate=0.5):
        super().__init__()
        self.data_shape = data_shape
        self.K = K
        self.L = L
        self.num_mode = num_mode
        self.controller_rate = controller_rate
        self.blocks = nn.ModuleList()
        input_size = self.data_shape[0]
        for _ in range(L - 1):
            self.blocks.append(Block(input_size, hidden_size, K, True, affine, conv_lu, num_mode, controller_rate))
            input_size *= 2
        self.blocks.append(Block(input_size, hidden_size, K, False, affine, conv_lu, num_mode, controller_rate))

    def loss_fn(self, log_p, logdet):
        """bits/dim NLL with NaN per-sample losses zeroed in training (src/models/mcglow.py:283-293)."""
        n_pixel = np.prod(self.data_shape)
        loss = -float(log(256.) * n_pixel) + logdet + log_p
        loss = -loss / float(log(2.) * n_pixel)
        if self.training:
            loss = torch.where(torch.isnan(loss), torch.zeros_like(loss), loss)
        else:
            loss = loss[~torch.isnan(loss)]
        return loss.mean()

    def forward(self, input, noise=None):
        """input {'img','label'} -> {'loss','z'} (src/models/mcglow.py:295-312).  ``noise``: optional
        injected dequantisation noise U[0,1) of the image shape (parity tests)."""
        ctx = _Ctx(_labels_from_indicator(input['label']))
        x = input['img'] * 0.5
        x = x + (torch.rand_like(x) if noise is None else noise) / 256
        x = x.permute(0, 2, 3, 1).contiguous()                     # the flow variables travel as fp32 NHWC
        z, log_p_sum, logdet = [], 0, 0
        for block in self.blocks:
            x, det, log_p, z_new = block(x, ctx)
            z.append(z_new)
            logdet = logdet + det
            if log_p is not None:
                log_p_sum = log_p_sum + log_p
        return {'loss': self.loss_fn(log_p_sum, logdet), 'z': z}

    def reverse(self, input):
        """{'z': [NCHW...], 'label', 'reconstruct'} -> {'img'} (src/models/mcglow.py:314-325).  The chain runs in
        NHWC on the native kernels, without autograd (it is the sampling path)."""
        ctx = _Ctx(_labels_from_indicator(input['label']))
        with torch.no_grad():
            z = [t.detach().float().permute(0, 2, 3, 1).contiguous() for t in input['z']]
            x = None
            for i, block in enumerate(self.blocks[::-1]):
                x = block.reverse(z[-1] if i == 0 else x, ctx, eps=z[-(i + 1)], reconstruct=input['reconstruct'])
            return {'img': (torch.clamp(x, -.5, .5) * 2).permute(0, 3, 1, 2)}

    def make_z_shapes(self):
        C, H, W = self.data_shape
        z_shapes = []
        for _ in range(self.L - 1):
            H, W, C = H // 2, W // 2, C * 2
            z_shapes.append((C, H, W))
        z_shapes.append((C * 4, H // 2, W // 2))
        return z_shapes

    def generate(self, C, x=None, temperature=1):
        if x is None:
            x = [torch.randn([C.size(0), *s], device=cfg['device']) * temperature for s in self.make_z_shapes()]
        return self.reverse({'z': x, 'reconstruct': False, 'label': C})['img']


def mcglow():
    """Factory reading the global ``cfg`` (src/models/mcglow.py:353-364)."""
    model = MCGlow(cfg['data_shape'], cfg['glow']['hidden_size'], cfg['glow']['K'], cfg['glow']['L'],
                   cfg['glow']['affine'], cfg['glow']['conv_lu'], cfg['classes_size'], cfg['controller_rate'])
    model.apply(init_param)
    return model
