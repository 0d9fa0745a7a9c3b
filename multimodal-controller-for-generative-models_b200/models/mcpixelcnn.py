"""The MC-gated PixelCNN prior (src/models/mcpixelcnn.py) on the B200 kernels — SURVEY.md §8 f3.

Same module tree, constructor order (hence RNG consumption) and ``state_dict`` keys as the reference;
``forward`` runs in NHWC bf16:

  MCGatedMaskedConv2d (mcpixelcnn.py:23-61), x_v / x_h : [N,H,W,C]
    explicit-tap tcgen05 conv  vert_stack   (k/2+1) x k kernel, taps dh in [-k/2,0], dw in [-k/2,k/2]  -> h_vert [.,2C]
    gated-activation kernel    out_v = mc(relu(bn(h_vert[:C])) * sigmoid(h_vert[C:]))        (ops.SGateFn)
    explicit-tap tcgen05 conv  horiz_stack  1 x (k/2+1) kernel, taps dw in [-k/2,0]                     -> h_horiz
    tcgen05 1x1 conv           vert_to_horiz(h_vert) + h_horiz (residual fused in the epilogue)
    gated-activation kernel    out_h
    tcgen05 1x1 conv -> BN -> MC gate kernel (+ x_h for the residual layers)
  The crops ``[:, :, :H, :]`` / ``[:, :, :, :W]`` of the reference are the choice of tap offsets: no padded
  rows are ever computed.  Mask 'A' (first layer) zeroes the last kernel row / column in place before every
  forward, exactly as ``make_causal`` does — the stored weights stay bit-compatible with the reference's.
Left to torch: the embedding gather and the cross entropy over the 512 logits.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .._lib import GATE_BN, GATE_MASK, GATE_RELU
from ..config import cfg
from ..modules import MultimodalController, Wrapper
from .mcgan import _Ctx, _labels_from_indicator, _packs
from .utils import init_param


class MCGatedActivation(nn.Module):
    def __init__(self, hidden_size, num_mode, controller_rate):
        super().__init__()
        self.bn = nn.BatchNorm2d(hidden_size)
        self.activation = nn.ReLU(inplace=True)
        self.mc = MultimodalController(hidden_size, num_mode, controller_rate)

    def fused(self, h, ctx):
        """h: bf16 NHWC [N,H,W,2C] -> [N,H,W,C]."""
        bn = self.bn
        out = ops.SGateFn.apply(h, bn.weight, bn.bias, ctx.labels, self.mc.packed_bits(), bn.running_mean, bn.running_var,
                                bn.momentum, bn.eps, self.training)
        if self.training:
            bn.num_batches_tracked += 1
        return out


def _bn_mc(x, bn, mc, ctx, training):
    """BatchNorm2d -> MC (no activation) through the gate kernels."""
    flags = GATE_BN | GATE_MASK
    if training:
        y, _ = ops.GateFn.apply(x, bn.weight, bn.bias, ctx.labels, mc.packed_bits(), flags, bn.running_mean, bn.running_var,
                                bn.momentum, bn.eps, False)
        bn.num_batches_tracked += 1
        return y
    rstd = torch.rsqrt(bn.running_var + bn.eps)
    return ops.gate_fwd(x, flags, ctx.labels, mc.packed_bits(), bn.running_mean, rstd, bn.weight.detach(), bn.bias.detach(),
                        False)[0]


def _conv1(x, conv, res=None, out_f32=False):
    return ops.ConvFn.apply(x, conv.weight, conv.bias, res, False, None, False, False, _packs(conv.weight), None, out_f32)


class MCGatedMaskedConv2d(nn.Module):
    def __init__(self, mask_type, hidden_size, kernel, residual, num_mode, controller_rate):
        super().__init__()
        if kernel % 2 != 1:
            raise ValueError('Not valid kernel size (must be odd)')
        self.mask_type = mask_type
        self.residual = residual
        self.kernel = kernel
        self.vert_stack = nn.Conv2d(hidden_size, 2 * hidden_size, (kernel // 2 + 1, kernel), 1, (kernel // 2, kernel // 2))
        self.vert_to_horiz = nn.Conv2d(2 * hidden_size, 2 * hidden_size, 1)
        self.horiz_stack = nn.Conv2d(hidden_size, 2 * hidden_size, (1, kernel // 2 + 1), 1, (0, kernel // 2))
        self.gate_v = MCGatedActivation(hidden_size, num_mode, controller_rate)
        self.gate_h = MCGatedActivation(hidden_size, num_mode, controller_rate)
        self.horiz_resid = nn.Sequential(
            Wrapper(nn.Conv2d(hidden_size, hidden_size, 1)),
            Wrapper(nn.BatchNorm2d(hidden_size)),
            MultimodalController(hidden_size, num_mode, controller_rate))

    def make_causal(self):
        self.vert_stack.weight.data[:, :, -1].zero_()        # mask the final row      (mcpixelcnn.py:43-45)
        self.horiz_stack.weight.data[:, :, :, -1].zero_()    # mask the final column

    def fused(self, x_v, x_h, ctx):
        if self.mask_type == 'A':
            self.make_causal()
        k = self.kernel
        h_vert = ops.ConvTapsFn.apply(x_v, self.vert_stack.weight, self.vert_stack.bias, (k // 2, k // 2), None)
        out_v = self.gate_v.fused(h_vert, ctx)
        h_horiz = ops.ConvTapsFn.apply(x_h, self.horiz_stack.weight, self.horiz_stack.bias, (0, k // 2), None)
        v2h = _conv1(h_vert, self.vert_to_horiz, res=h_horiz)                     # v2h + h_horiz
        out_h = self.gate_h.fused(v2h, ctx)
        r = _conv1(out_h, self.horiz_resid[0].module)
        r = _bn_mc(r, self.horiz_resid[1].module, self.horiz_resid[2], ctx, self.training)
        return out_v, (r + x_h if self.residual else r)


class MCGatedPixelCNN(nn.Module):
    def __init__(self, input_size=256, hidden_size=64, num_layer=15, num_mode=10, controller_rate=0.5):
        super().__init__()
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.embedding = nn.Embedding(input_size, hidden_size)
        self.layers = nn.ModuleList()
        for i in range(num_layer):                 # first block: mask A, 7x7, no residual; the rest: mask B, 3x3
            self.layers.append(MCGatedMaskedConv2d('A' if i == 0 else 'B', hidden_size, 7 if i == 0 else 3, i != 0,
                                                   num_mode, controller_rate))
        self.output_conv = nn.Sequential(
            Wrapper(nn.Conv2d(hidden_size, 512, 1)),
            Wrapper(nn.BatchNorm2d(512)),
            Wrapper(nn.ReLU(inplace=True)),
            MultimodalController(512, num_mode, controller_rate),
            Wrapper(nn.Conv2d(512, input_size, 1)))

    def forward(self, input):
        """{'img': int64 code map [N,H,W], 'label': [N]} -> {'loss', 'logits' [N, input_size, H, W]}
        (src/models/mcpixelcnn.py:86-101)."""
        ctx = _Ctx(_labels_from_indicator(input['label']))
        code = input['img']
        x = F.embedding(code.reshape(-1), self.embedding.weight).view(*code.shape, -1)          # NHWC already
        x = x.to(torch.bfloat16).contiguous()
        x_v = x_h = x
        for layer in self.layers:
            x_v, x_h = layer.fused(x_v, x_h, ctx)
        oc = self.output_conv
        h = _conv1(x_h, oc[0].module)
        bn, mc = oc[1].module, oc[3]
        flags = GATE_BN | GATE_RELU | GATE_MASK
        if self.training:
            h, _ = ops.GateFn.apply(h, bn.weight, bn.bias, ctx.labels, mc.packed_bits(), flags, bn.running_mean,
                                    bn.running_var, bn.momentum, bn.eps, False)
            bn.num_batches_tracked += 1
        else:
            h = ops.gate_fwd(h, flags, ctx.labels, mc.packed_bits(), bn.running_mean, torch.rsqrt(bn.running_var + bn.eps),
                             bn.weight.detach(), bn.bias.detach(), False)[0]
        logits = _conv1(h, oc[4].module, out_f32=True).permute(0, 3, 1, 2)                     # NCHW view
        return {'loss': F.cross_entropy(logits, code), 'logits': logits}

    def generate(self, C, x=None):
        """Ancestral sampling of an 8x8 code map, one full forward per position (mcpixelcnn.py:103-112)."""
        if x is None:
            x = torch.zeros((C.size(0), 8, 8), dtype=torch.long, device=cfg['device'])
        input = {'img': x, 'label': C}
        for i in range(x.size(1)):
            for j in range(x.size(2)):
                output = self.forward(input)
                probs = F.softmax(output['logits'][:, :, i, j], -1)
                input['img'].data[:, i, j].copy_(probs.multinomial(1).squeeze().data)
        return input['img']


def mcpixelcnn():
    """Factory reading the global ``cfg`` (src/models/mcpixelcnn.py:115-124)."""
    p = cfg['pixelcnn']
    model = MCGatedPixelCNN(input_size=p['num_embedding'], hidden_size=p['hidden_size'], num_layer=p['num_layer'],
                            num_mode=cfg['classes_size'], controller_rate=cfg['controller_rate'])
    model.apply(init_param)
    return model
