"""MCGAN on the B200 kernels — same module tree, constructor signatures, ``state_dict``
keys and ``generate`` / ``discriminate`` API as the reference (src/models/mcgan.py),
with every block's forward replaced by fused sm_100a kernels.

The module tree is built from small specs so that parameter / buffer names
(``blocks.0.conv.4.module.weight``, ``blocks.0.shortcut.1.codebook``,
``...module.weight_orig`` …) and the order in which the global RNG is consumed match the
reference exactly: ``torch.manual_seed(s); models.mcgan()`` yields a bit-identical
``state_dict`` (tests/test_host_parity.py).  The ``nn.Conv2d`` / ``nn.BatchNorm2d``
children are parameter containers only; their own ``forward`` is never called.

Data flow inside a block (all activations NHWC bf16, fp32 accumulate):

  GenResBlock (mcgan.py:9-44), x:[N,h,w,C]
    gate kernel : BN1 stats+apply → ReLU → mask(mc_1) → a [N,h,w,C]   (mask and nearest-upsample commute)
                  and mask(mc_1)·x                     → s [N,h,w,C]
    tcgen05 conv: 1x1(s) at LOW resolution (a 1x1 conv commutes with nearest upsampling)
    tcgen05 conv: Up(2)→Conv3x3 as 4 output phases x 2x2 taps on a (no upsampled tensor, 2.25x fewer MACs) → c1
    gate kernel : BN2 → ReLU → mask(mc_2) → h2
    tcgen05 conv: 3x3(h2)+b + upsample2(shortcut) fused in the epilogue

  DisResBlock (mcgan.py:96-138), no BN:
    gate kernel : mask(mc_1)·ReLU(x) (and mask(mc_1)·x for a conv shortcut)
    tcgen05 conv: 3x3 + b → ReLU → mask(mc_2) fused in the epilogue
    tcgen05 conv: 3x3 + b + shortcut fused; AvgPool2d(2) kernel when stride 2
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .._lib import GATE_BN, GATE_MASK, GATE_RELU, GATE_UP2
from ..config import cfg
from ..modules import MultimodalController, Wrapper
from .utils import init_param, make_SpectralNormalization


# --------------------------------------------------------------------------- tree building
def _layer(spec):
    """spec -> module: an nn.Module is passed through, a tuple/str builds a wrapped torch layer."""
    if isinstance(spec, nn.Module):
        return spec
    kind, args = (spec, ()) if isinstance(spec, str) else (spec[0], spec[1:])
    if kind == 'bn':
        inner = nn.BatchNorm2d(*args)
    elif kind == 'relu':
        inner = nn.ReLU()
    elif kind == 'up':
        inner = nn.Upsample(scale_factor=args[0], mode='nearest')
    elif kind == 'conv':
        cin, cout, k = args
        inner = nn.Conv2d(cin, cout, k, 1, (k - 1) // 2)
    elif kind == 'pool':
        inner = nn.AvgPool2d(2)
    elif kind == 'tanh':
        inner = nn.Tanh()
    elif kind == 'sumpool':
        inner = GlobalSumPooling()
    elif kind == 'linear':
        inner = nn.Linear(*args)
    else:
        raise ValueError('Not valid layer spec')
    return Wrapper(inner)


def _chain(*specs):
    return nn.Sequential(*[_layer(s) for s in specs])


class GlobalSumPooling(nn.Module):
    """Sum over the spatial dims of an NCHW tensor (src/models/mcgan.py:141-147)."""

    def forward(self, input):
        return input.sum(dim=[-2, -1]).view(input.size(0), -1)


# --------------------------------------------------------------------------- fused helpers
class _Ctx:
    """Per-forward context: int32 mode labels on the device."""

    def __init__(self, labels, sn=None, sn_packs=None):
        self.labels = ops.as_labels(labels).contiguous()
        self.sn = sn            # {module: spectral-norm snapshot} for this forward
        self.sn_packs = sn_packs or {}      # {module: bf16 operand copies of W/sigma} (ops.sn_pack_all)


def _labels_from_indicator(indicator):
    """The public API passes a one-hot float indicator [N, M] (mcgan.py:196); the fused path
    gathers codebook rows by integer label, which is identical for a one-hot indicator."""
    if indicator.dim() == 1 and not indicator.is_floating_point():
        return indicator
    if indicator.dim() != 2:
        raise ValueError('Not valid indicator')
    # The reference computes code = indicator @ codebook (src/modules/modules.py:73): only for a ONE-HOT
    # indicator is that the row gather the fused blocks perform.  A soft / multi-hot indicator would silently
    # give a different network, so it is rejected (one host sync; the training steps pass integer labels and
    # never come here; skipped while a CUDA graph is being captured, where a sync is illegal).
    if not (indicator.is_cuda and torch.cuda.is_current_stream_capturing()):
        ones = (indicator == 1).sum(dim=1)
        nonzero = (indicator != 0).sum(dim=1)
        if not bool(((ones == 1) & (nonzero == 1)).all()):
            raise ValueError('Not valid indicator: the fused blocks take a one-hot [N, num_mode] indicator or integer '
                             'labels [N] (use MultimodalController.forward for a dense code)')
    return indicator.argmax(dim=1)


def _sync_act_dtype():
    """cfg['compute_dtype']: 'bf16' (default, the measured path) or 'fp32' (parity pipeline: fp32 storage,
    3xTF32 forward / dgrad and bf16x3 weight gradients on the same tensor-core kernels)."""
    want = torch.float32 if cfg.get('compute_dtype', 'bf16') == 'fp32' else torch.bfloat16
    if ops.ACT_DTYPE != want:
        ops.ACT_DTYPE = want


_pack_cache = {}
# ops.PACKS_FORCE is set while re-playable graph SEGMENTS are captured: the cache check below happens at capture time
# only, so a segment must re-pack inside the graph (the parameters change between replays of different segments)


def _packs(weight):
    """bf16 (KRSC, flipped CRSK) copies of a parameter, cached until the parameter changes."""
    key = id(weight)
    hit = _pack_cache.get(key)
    if hit is not None and hit[0] == weight._version and hit[1] is weight and not ops.PACKS_FORCE:
        return hit[2]
    packs = ops.pack_weight(weight.detach())
    _pack_cache[key] = (weight._version, weight, packs)
    return packs


_up2_cache = {}
ops.PACK_CACHES += [_pack_cache, _up2_cache]


def _up2_packs(weight):
    """Phase / adjoint weights of an upsample-fused 3x3 conv, cached like ``_packs``."""
    key = id(weight)
    hit = _up2_cache.get(key)
    if hit is not None and hit[0] == weight._version and hit[1] is weight and not ops.PACKS_FORCE:
        return hit[2]
    packs = ops.up2_conv_packs(weight.detach())
    _up2_cache[key] = (weight._version, weight, packs)
    return packs


def _conv(x, conv, ctx, res=None, res_up=False, gate=None, premask_dx=False, grad_premasked=False, want_stats=False):
    """tcgen05 conv through ``conv`` (an nn.Conv2d container, possibly spectral-normed: then
    ``ctx.sn[conv]`` holds this forward's [1/sigma | u | v] and the weight is packed as W/sigma)."""
    if hasattr(conv, 'weight_orig'):
        return ops.ConvFn.apply(x, conv.weight_orig, conv.bias, res, res_up, gate, premask_dx, grad_premasked,
                                ctx.sn_packs.get(conv), ctx.sn[conv])
    w = conv.weight
    return ops.ConvFn.apply(x, w, conv.bias, res, res_up, gate, premask_dx, grad_premasked, _packs(w), None, False,
                            want_stats)


def _bn_gate(x, bn, mc, ctx, training, up=False, want_raw=False, pre_stats=None):
    """pre_stats: ops.ConvStats of x from the epilogue of the conv that produced it (training only)."""
    flags = GATE_BN | GATE_RELU | GATE_MASK | (GATE_UP2 if up else 0)
    if training:
        y, raw = ops.GateFn.apply(x, bn.weight, bn.bias, ctx.labels, mc.packed_bits(), flags, bn.running_mean,
                                  bn.running_var, bn.momentum, bn.eps, want_raw, pre_stats)
        bn.num_batches_tracked += 1
        return y, raw
    # eval: running statistics (inference path; no autograd through the gate)
    rstd = torch.rsqrt(bn.running_var + bn.eps)
    return ops.gate_fwd(x, flags, ctx.labels, mc.packed_bits(), bn.running_mean, rstd, bn.weight.detach(),
                        bn.bias.detach(), want_raw)


def _relu_gate(x, mc, ctx, want_raw=False):
    return ops.GateFn.apply(x, None, None, ctx.labels, mc.packed_bits(), GATE_RELU | GATE_MASK, None, None, 0.0, 0.0,
                            want_raw)


# --------------------------------------------------------------------------- generator
class GenResBlock(nn.Module):
    def __init__(self, input_size, output_size, num_mode, controller_rate, stride):
        super().__init__()
        self.stride = stride
        self.mc_1 = MultimodalController(input_size, num_mode, controller_rate)
        self.mc_2 = MultimodalController(output_size, num_mode, controller_rate)
        self.conv = _chain(('bn', input_size), 'relu', ('up', stride), self.mc_1,
                           ('conv', input_size, output_size, 3),
                           ('bn', output_size), 'relu', self.mc_2,
                           ('conv', output_size, output_size, 3))
        if stride > 1:
            self.shortcut = _chain(('up', stride), self.mc_1, ('conv', input_size, output_size, 1))
        elif input_size != output_size:
            self.shortcut = nn.Sequential(self.mc_1, nn.Conv2d(input_size, output_size, 1, 1, 0))
        else:
            self.shortcut = nn.Sequential()

    def fused(self, x, ctx, x_stats=None):
        """x_stats: batch statistics of x from the conv that produced it.  Returns (out, stats of out): every
        BatchNorm of the generator but the first reads a tensor a conv has just written, and that conv's
        epilogue accumulates the statistics (no separate pass over the tensor)."""
        if self.stride != 2:
            raise ValueError('Not valid stride (the fused generator block upsamples by 2)')
        t = self.training
        # epilogue statistics (default ON): the staged epilogue sums the tile it has just written to shared memory —
        # two columns per lane, no shuffles — so the statistics pass over the tensor disappears (-1.07 ms / iteration
        # measured; round 1's shuffle-based column sums cost more than they saved).  cfg['bn_stats_in_conv'] = False
        # switches back to the separate statistics kernels (parity-tested either way).
        ws = t and bool(cfg.get('bn_stats_in_conv', True))
        a, s = _bn_gate(x, self.conv[0].module, self.mc_1, ctx, t, want_raw=True, pre_stats=x_stats)   # LOW resolution
        sc = _conv(s, self.shortcut[2].module, ctx)                      # 1x1 commutes with the upsample
        c1m = self.conv[4].module                                        # Up(2) -> Conv3x3 as a 4-phase 2x2 conv
        c1 = ops.ConvUp2Fn.apply(a, c1m.weight, c1m.bias, _up2_packs(c1m.weight), ws)
        h2, _ = _bn_gate(c1, self.conv[5].module, self.mc_2, ctx, t, pre_stats=ops.take_stats())
        out = _conv(h2, self.conv[8].module, ctx, res=sc, res_up=True, want_stats=ws)
        return out, ops.take_stats()


class Generator(nn.Module):
    def __init__(self, data_shape, latent_size, hidden_size, num_mode, controller_rate):
        super().__init__()
        self.latent_size = latent_size
        self.data_shape = data_shape
        self.linear = Wrapper(nn.Linear(latent_size, hidden_size[0] * 4 * 4))
        blocks = [GenResBlock(hidden_size[i], hidden_size[i + 1], num_mode, controller_rate, stride=2)
                  for i in range(len(hidden_size) - 1)]
        blocks += [_layer(('bn', hidden_size[-1])), _layer('relu'),
                   MultimodalController(hidden_size[-1], num_mode, controller_rate),
                   _layer(('conv', hidden_size[-1], data_shape[0], 3)), _layer('tanh')]
        self.blocks = nn.Sequential(*blocks)

    def forward(self, input, indicator):
        """input z [N, latent]; indicator one-hot [N, M] (or int labels [N]) -> image [N,C,H,W] fp32
        (src/models/mcgan.py:64-69).  The returned tensor is an NCHW view of NHWC storage."""
        ctx = _Ctx(_labels_from_indicator(indicator))
        _sync_act_dtype()
        lin = self.linear.module
        n_res = len(self.blocks) - 5
        c0 = lin.out_features // 16
        x = F.linear(input, lin.weight, lin.bias)                        # [N, C0*16]  (tiny GEMM, cuBLAS)
        x = x.view(-1, c0, 4, 4).permute(0, 2, 3, 1).to(ops.ACT_DTYPE).contiguous()
        st = None
        for i in range(n_res):
            x, st = self.blocks[i].fused(x, ctx, st)
        bn, mc, conv = self.blocks[n_res].module, self.blocks[n_res + 2], self.blocks[n_res + 3].module
        h, _ = _bn_gate(x, bn, mc, ctx, self.training, pre_stats=st)
        img = ops.ThinFromWideFn.apply(h, conv.weight, conv.bias, True)   # 3x3 -> RGB, tanh fused
        return img.permute(0, 3, 1, 2)


# --------------------------------------------------------------------------- discriminator
class FirstDisResBlock(nn.Module):
    def __init__(self, input_size, output_size, num_mode, controller_rate):
        super().__init__()
        self.mc_1 = MultimodalController(output_size, num_mode, controller_rate)
        self.conv = _chain(('conv', input_size, output_size, 3), 'relu', self.mc_1,
                           ('conv', output_size, output_size, 3), 'pool')
        self.shortcut = _chain(('conv', input_size, output_size, 1), 'pool')

    def fused(self, img, ctx):
        """img: fp32 NHWC [N,H,W,3] (mcgan.py:72-93).  AvgPool is linear, so conv2 + shortcut are
        summed at full resolution in conv2's epilogue and pooled once."""
        c1, c2, sc = self.conv[0].module, self.conv[3].module, self.shortcut[0].module
        h = ops.WideFromThinFn.apply(img, c1.weight_orig, c1.bias, (ctx.labels, self.mc_1.packed_bits()), True, ctx.sn[c1],
                                     1, None, ctx.sn_packs.get(c1))
        s = ops.WideFromThinFn.apply(img, sc.weight_orig, sc.bias, None, False, ctx.sn[sc], 1, None, ctx.sn_packs.get(sc))
        y = _conv(h, c2, ctx, res=s, premask_dx=True)
        return ops.Pool2Fn.apply(y)


class DisResBlock(nn.Module):
    def __init__(self, input_size, output_size, num_mode, controller_rate, stride):
        super().__init__()
        self.stride = stride
        self.mc_1 = MultimodalController(input_size, num_mode, controller_rate)
        self.mc_2 = MultimodalController(output_size, num_mode, controller_rate)
        body = ['relu', self.mc_1, ('conv', input_size, output_size, 3),
                'relu', self.mc_2, ('conv', output_size, output_size, 3)]
        if stride > 1:
            self.conv = _chain(*body, 'pool')
            self.shortcut = _chain(self.mc_1, ('conv', input_size, output_size, 1), 'pool')
        else:
            self.conv = _chain(*body)
            if input_size != output_size:
                self.shortcut = _chain(self.mc_1, ('conv', input_size, output_size, 1))
            else:
                self.shortcut = nn.Sequential()

    def fused(self, x, ctx):
        has_sc = len(self.shortcut) > 0
        a, s = _relu_gate(x, self.mc_1, ctx, want_raw=has_sc)
        h = _conv(a, self.conv[2].module, ctx, gate=(ctx.labels, self.mc_2.packed_bits()), grad_premasked=True)
        res = _conv(s, self.shortcut[1].module, ctx) if has_sc else x
        y = _conv(h, self.conv[5].module, ctx, res=res, premask_dx=True)
        return ops.Pool2Fn.apply(y) if self.stride > 1 else y


class Discriminator(nn.Module):
    def __init__(self, data_shape, hidden_size, num_mode, controller_rate):
        super().__init__()
        self.data_shape = data_shape
        h = hidden_size
        blocks = [FirstDisResBlock(data_shape[0], h[0], num_mode, controller_rate)]
        # CIFAR: (len-3) stride-2 blocks then two stride-1 blocks; otherwise (len-2) + one (mcgan.py:154-175)
        n_down = len(h) - 3 if cfg['data_name'] in ['CIFAR10', 'CIFAR100'] else len(h) - 2
        blocks += [DisResBlock(h[i], h[i + 1], num_mode, controller_rate, stride=2) for i in range(n_down)]
        blocks += [DisResBlock(h[i], h[i + 1], num_mode, controller_rate, stride=1)
                   for i in range(n_down, len(h) - 1)]
        blocks += [_layer('relu'), MultimodalController(h[-1], num_mode, controller_rate),
                   _layer('sumpool'), _layer(('linear', h[-1], 1))]
        self.blocks = nn.Sequential(*blocks)
        self._sn_batch = None       # built lazily: spectral norm is applied after construction (mcgan.py:191)

    def forward(self, input, indicator):
        """input image [N,C,H,W] fp32; indicator one-hot [N,M] (or int labels) -> logits [N,1]
        (src/models/mcgan.py:178-181)."""
        if self._sn_batch is None:
            mods = [m for m in self.modules() if hasattr(m, 'weight_orig')]
            self._sn_batch = [ops.SNBatch(mods[i:i + 16]) for i in range(0, len(mods), 16)]
        sn = {}
        for batch in self._sn_batch:      # one launch per <= 16 layers: power iteration + [1/sigma|u|v] snapshot
            sn.update(batch.run(self.training))
        _sync_act_dtype()
        packs = ops.sn_pack_all(sn) if ops.ACT_DTYPE == torch.bfloat16 else None     # one launch for all conv layers
        ctx = _Ctx(_labels_from_indicator(indicator), sn, packs)
        img = input.permute(0, 2, 3, 1).contiguous().float()
        n_res = len(self.blocks) - 4
        x = self.blocks[0].fused(img, ctx)
        for i in range(1, n_res):
            x = self.blocks[i].fused(x, ctx)
        g, _ = _relu_gate(x, self.blocks[n_res + 1], ctx)
        pooled = g.float().sum(dim=(1, 2))                                # global SUM pooling [N, C]
        lin = self.blocks[n_res + 3].module
        w = ops.SNScaleFn.apply(lin.weight_orig, sn[lin]) if lin in sn else lin.weight
        return F.linear(pooled, w, lin.bias)


# --------------------------------------------------------------------------- MCGAN
class MCGAN(nn.Module):
    def __init__(self, data_shape, latent_size, generator_hidden_size, discriminator_hidden_size, num_mode,
                 controller_rate):
        super().__init__()
        self.latent_size = latent_size
        self.generator = Generator(data_shape, latent_size, generator_hidden_size, num_mode, controller_rate)
        self.discriminator = Discriminator(data_shape, discriminator_hidden_size, num_mode, controller_rate)
        self.discriminator.apply(make_SpectralNormalization)

    def generate(self, C, x=None):
        if x is None:
            x = torch.randn([C.size(0), self.latent_size], device=cfg['device'])
        return self.generator(x, C)

    def discriminate(self, x, C):
        return self.discriminator(x, C)

    def forward(self, input):
        x = torch.randn(input['img'].size(0), self.latent_size, device=cfg['device'])
        return self.discriminate(self.generate(input['label'], x), input['label'])


def mcgan():
    """Factory reading the global ``cfg`` (src/models/mcgan.py:212-221)."""
    model = MCGAN(cfg['data_shape'], cfg['gan']['latent_size'], cfg['gan']['generator_hidden_size'],
                  cfg['gan']['discriminator_hidden_size'], cfg['classes_size'], cfg['controller_rate'])
    model.apply(init_param)
    return model
