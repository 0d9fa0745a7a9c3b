"""MCVAE on the B200 kernels — same module tree, ``state_dict`` keys and
``forward / encode / decode / generate`` API as the reference (src/models/mcvae.py).

Kernels per layer (activations NHWC bf16, fp32 accumulate):
  Conv2d(3, 64, 4, 2, 1)            stride-2 im2col of the 3-channel image + 1x1 tcgen05 GEMM
  Conv2d(C, K, 4, 2, 1)             tcgen05 conv, strided pixels gathered by TMA (traversal stride 2)
  BN -> ReLU -> MC / BN -> MC       fused gate kernel (mcgm_bn_stats + mcgm_gate_fwd/bwd)
  ResBlock convs 3x3                tcgen05 conv
  ConvTranspose2d(C, K, 4, 2, 1)    4-phase 2x2 tcgen05 conv (no zero-insertion, no scatter)
Left to torch (tiny 2-D tensors, < 1 % of the work): the three Linear layers (cuBLAS), BatchNorm1d on
[N, 4096], the ResBlock's ``relu(y + x)`` on [N,4,4,256], sigmoid / BCE / KLD on the image.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .._lib import GATE_BN, GATE_MASK, GATE_RELU
from ..config import cfg
from ..modules import MultimodalController, Wrapper
from .mcgan import _Ctx, _labels_from_indicator, _packs
from .utils import init_param


def loss(input, output):
    """BCE(sum) + KLD over numel (src/models/mcvae.py:10-14)."""
    bce = F.binary_cross_entropy(output['img'], input['img'], reduction='sum')
    kld = 0.5 * torch.sum(output['mu'].pow(2) + output['logvar'].exp() - 1 - output['logvar'])
    return (bce + kld) / input['img'].numel()


def _w(module):
    return Wrapper(module)


def _gate(x, bn, mc, ctx, training, relu=True):
    flags = GATE_BN | GATE_MASK | (GATE_RELU if relu else 0)
    if training:
        y, _ = ops.GateFn.apply(x, bn.weight, bn.bias, ctx.labels, mc.packed_bits(), flags, bn.running_mean,
                                bn.running_var, bn.momentum, bn.eps, False)
        bn.num_batches_tracked += 1
        return y
    rstd = torch.rsqrt(bn.running_var + bn.eps)
    return ops.gate_fwd(x, flags, ctx.labels, mc.packed_bits(), bn.running_mean, rstd, bn.weight.detach(),
                        bn.bias.detach(), False)[0]


def _conv3(x, conv):
    return ops.ConvFn.apply(x, conv.weight, conv.bias, None, False, None, False, False, _packs(conv.weight), None)


class ResBlock(nn.Module):
    def __init__(self, hidden_size, num_mode, controller_rate):
        super().__init__()
        self.conv = nn.Sequential(
            _w(nn.Conv2d(hidden_size, hidden_size, 3, 1, 1)), _w(nn.BatchNorm2d(hidden_size)),
            _w(nn.ReLU(inplace=True)), MultimodalController(hidden_size, num_mode, controller_rate),
            _w(nn.Conv2d(hidden_size, hidden_size, 3, 1, 1)), _w(nn.BatchNorm2d(hidden_size)),
            MultimodalController(hidden_size, num_mode, controller_rate))
        self.activation = _w(nn.ReLU(inplace=True))

    def fused(self, x, ctx):
        """Conv → BN → ReLU → MC → Conv → BN → MC; + x; ReLU (src/models/mcvae.py:17-35)."""
        t = self.training
        h = _gate(_conv3(x, self.conv[0].module), self.conv[1].module, self.conv[3], ctx, t)
        h = _gate(_conv3(h, self.conv[4].module), self.conv[5].module, self.conv[6], ctx, t, relu=False)
        return F.relu(h + x)


class Encoder(nn.Module):
    def __init__(self, data_shape, hidden_size, latent_size, num_res_block, num_mode, controller_rate):
        super().__init__()
        widths = [data_shape[0]] + list(hidden_size)
        blocks = []
        for i in range(len(hidden_size)):
            blocks += [_w(nn.Conv2d(widths[i], widths[i + 1], 4, 2, 1)), _w(nn.BatchNorm2d(widths[i + 1])),
                       _w(nn.ReLU(inplace=True)), MultimodalController(widths[i + 1], num_mode, controller_rate)]
        blocks += [ResBlock(hidden_size[-1], num_mode, controller_rate) for _ in range(num_res_block)]
        self.blocks = nn.Sequential(*blocks)
        self.n_down = len(hidden_size)
        down = 2 ** len(hidden_size)
        self.encoded_shape = (hidden_size[-1], data_shape[1] // down, data_shape[2] // down)
        self.mu = nn.Linear(np.prod(self.encoded_shape).item(), latent_size)
        self.logvar = nn.Linear(np.prod(self.encoded_shape).item(), latent_size)

    def reparameterize(self, mu, logvar, eps=None):
        std = torch.exp(0.5 * logvar)
        eps = torch.randn_like(std) if eps is None else eps
        return mu + eps * std

    def forward(self, input, eps=None):
        x, ctx = input[0], _Ctx(_labels_from_indicator(input[1]))
        t = self.training
        img = x.permute(0, 2, 3, 1).contiguous().float()                       # NHWC fp32 image
        c0 = self.blocks[0].module
        h = ops.WideFromThinFn.apply(img, c0.weight, c0.bias, None, False, None, 2, 1)
        h = _gate(h, self.blocks[1].module, self.blocks[3], ctx, t)
        for i in range(1, self.n_down):
            conv = self.blocks[4 * i].module
            h = ops.ConvS2Fn.apply(h, conv.weight, conv.bias)
            h = _gate(h, self.blocks[4 * i + 1].module, self.blocks[4 * i + 3], ctx, t)
        for blk in list(self.blocks)[4 * self.n_down:]:
            h = blk.fused(h, ctx)
        flat = h.permute(0, 3, 1, 2).reshape(h.size(0), -1).float()            # (C,H,W) order of the reference view
        mu, logvar = self.mu(flat), self.logvar(flat)
        z = self.reparameterize(mu, logvar, eps) if self.training else mu
        return z, mu, logvar


class Decoder(nn.Module):
    def __init__(self, data_shape, hidden_size, latent_size, num_res_block, num_mode, controller_rate):
        super().__init__()
        down = 2 ** len(hidden_size)
        self.encoded_shape = (hidden_size[-1], data_shape[1] // down, data_shape[2] // down)
        self.linear = nn.Sequential(
            MultimodalController(latent_size, num_mode, controller_rate),
            _w(nn.Linear(latent_size, np.prod(self.encoded_shape).item())),
            _w(nn.BatchNorm1d(np.prod(self.encoded_shape).item())), _w(nn.ReLU(inplace=True)))
        blocks = [MultimodalController(hidden_size[-1], num_mode, controller_rate)]
        blocks += [ResBlock(hidden_size[-1], num_mode, controller_rate) for _ in range(num_res_block)]
        for i in range(len(hidden_size) - 1, 0, -1):
            blocks += [_w(nn.ConvTranspose2d(hidden_size[i], hidden_size[i - 1], 4, 2, 1)),
                       _w(nn.BatchNorm2d(hidden_size[i - 1])), _w(nn.ReLU(inplace=True)),
                       MultimodalController(hidden_size[i - 1], num_mode, controller_rate)]
        blocks += [_w(nn.ConvTranspose2d(hidden_size[0], data_shape[0], 4, 2, 1)), _w(nn.Sigmoid())]
        self.blocks = nn.Sequential(*blocks)
        self.num_res_block = num_res_block

    def forward(self, input):
        z, indicator = input[0], input[1]
        ctx = _Ctx(_labels_from_indicator(indicator))
        t = self.training
        h = self.linear[0]([z, ctx.labels])[0]                                  # MC on the latent (2-D)
        lin, bn1 = self.linear[1].module, self.linear[2].module
        h = F.relu(bn1(F.linear(h, lin.weight, lin.bias)))                      # [N, 4096]: cuBLAS + ATen BN1d
        h = h.view(h.size(0), *self.encoded_shape).permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()
        h = ops.MCMaskFn.apply(h, ctx.labels, self.blocks[0].packed_bits(), None, False)
        i = 1
        for _ in range(self.num_res_block):
            h = self.blocks[i].fused(h, ctx)
            i += 1
        while isinstance(self.blocks[i + 1].module, nn.BatchNorm2d):
            ct = self.blocks[i].module
            h = ops.ConvT4Fn.apply(h, ct.weight, ct.bias)
            h = _gate(h, self.blocks[i + 1].module, self.blocks[i + 3], ctx, t)
            i += 4
        ct = self.blocks[i].module
        img = ops.ConvT4Fn.apply(h, ct.weight, ct.bias)                         # fp32 [N,H,W,3]
        return torch.sigmoid(img).permute(0, 3, 1, 2)


class MCVAE(nn.Module):
    def __init__(self, data_shape=(3, 32, 32), hidden_size=(64, 128, 256), latent_size=128, num_res_block=2,
                 num_mode=None, controller_rate=0.5):
        super().__init__()
        self.data_shape = data_shape
        self.hidden_size = hidden_size
        self.latent_size = latent_size
        self.num_res_block = num_res_block
        self.num_mode = num_mode
        self.controller_rate = controller_rate
        self.encoder = Encoder(data_shape, hidden_size, latent_size, num_res_block, num_mode, controller_rate)
        self.decoder = Decoder(data_shape, hidden_size, latent_size, num_res_block, num_mode, controller_rate)

    def encode(self, input, eps=None):
        return self.encoder(input, eps)

    def decode(self, input):
        return self.decoder(input)

    def generate(self, C, z=None):
        if z is None:
            z = torch.randn([C.size(0), self.latent_size], device=cfg['device'])
        return self.decode([z, C]) * 2 - 1

    def forward(self, input, eps=None):
        """input {'img' in [-1,1], 'label'} -> {'loss','img','mu','logvar'} (src/models/mcvae.py:133-144).
        ``eps``: optional injected reparameterisation noise [N, latent] (parity tests)."""
        output = {}
        input['img'] = (input['img'] + 1) / 2
        z, output['mu'], output['logvar'] = self.encode([input['img'], input['label']], eps)
        output['img'] = self.decode([z, input['label']])
        output['loss'] = loss(input, output)
        input['img'] = input['img'] * 2 - 1
        output['img'] = output['img'] * 2 - 1
        return output


def mcvae():
    """Factory reading the global ``cfg`` (src/models/mcvae.py:147-157)."""
    model = MCVAE(data_shape=cfg['data_shape'], hidden_size=cfg['vae']['hidden_size'],
                  latent_size=cfg['vae']['latent_size'], num_res_block=cfg['vae']['num_res_block'],
                  num_mode=cfg['classes_size'], controller_rate=cfg['controller_rate'])
    model.apply(init_param)
    return model
