"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU restatement (fp32, NCHW, torch.nn.functional) of the reference's MCGlow path as pure functions
over a ``state_dict``-shaped mapping:

    actnorm        src/models/mcglow.py:24-55    data-dependent init on the first training forward
    invconv_weight src/models/mcglow.py:104-110  W = P (L∘l_mask + I) (U∘u_mask + diag(sign·exp(w_s)))
    coupling       src/models/mcglow.py:133-164  net = Conv3x3→ActNorm→ReLU→MC→Conv1x1→ActNorm→ReLU→MC→ZeroConv
    flow / block   src/models/mcglow.py:178-239  squeeze, K flows, split + Gaussian prior
    forward        src/models/mcglow.py:283-312  dequantisation noise, bits/dim NLL, NaN→0
    train_step     src/train_glow.py:116-121     backward → clip_grad_norm_(1) → Adam(3e-4)

Pinned against fixtures generated from the unmodified reference (oracle/make_golden.py →
tests/golden/glow_small.npz, mcgan_golden.json["cases"]["glow_small"]).
"""
from math import log, pi

import torch
import torch.nn.functional as F

from .mcgan_oracle import Adam, make_trainable as _mt, mc_mask, one_hot  # noqa: F401
from .mcvae_oracle import clip_grad_norm_, trainable  # noqa: F401


def make_trainable(sd):
    """Parameters of Glow are everything floating point that is not a registered buffer."""
    buffers = ('codebook', 'w_p', 'u_mask', 'l_mask', 's_sign', 'l_eye', 'initialized')
    out = {}
    for k, t in sd.items():
        t = t.detach().clone()
        if t.is_floating_point() and k.rsplit('.', 1)[1] not in buffers:
            t.requires_grad_(True)
        out[k] = t
    return out


def actnorm(x, sd, pre, training, with_logdet):
    if int(sd[pre + 'initialized']) == 0 and training:
        with torch.no_grad():
            mean = torch.mean(x, dim=[0, 2, 3], keepdim=True)
            std = torch.std(x, dim=[0, 2, 3], keepdim=True)
            sd[pre + 'loc'].copy_(-mean)
            sd[pre + 'scale'].copy_(1 / (std + 1e-6))
            sd[pre + 'initialized'].fill_(1)
    out = sd[pre + 'scale'] * (x + sd[pre + 'loc'])
    if with_logdet:
        return out, x.size(2) * x.size(3) * torch.sum(torch.log(torch.abs(sd[pre + 'scale'])))
    return out


def invconv_weight(sd, pre):
    lower = sd[pre + 'w_l'] * sd[pre + 'l_mask'] + sd[pre + 'l_eye']
    upper = sd[pre + 'w_u'] * sd[pre + 'u_mask'] + torch.diag(sd[pre + 's_sign'] * torch.exp(sd[pre + 'w_s']))
    return (sd[pre + 'w_p'] @ lower @ upper).unsqueeze(2).unsqueeze(3)


def zero_conv(x, sd, pre):
    return F.conv2d(x, sd[pre + 'conv.weight'], sd[pre + 'conv.bias'], padding=1) * torch.exp(sd[pre + 'scale'] * 3)


def coupling(x, ind, sd, pre, training):
    in_a, in_b = x.chunk(2, 1)
    h = F.conv2d(in_a, sd[pre + 'net.0.module.weight'], sd[pre + 'net.0.module.bias'], padding=1)
    h = F.relu(actnorm(h, sd, pre + 'net.1.module.', training, False))
    h = mc_mask(h, ind, sd[pre + 'net.3.codebook'])
    h = F.conv2d(h, sd[pre + 'net.4.module.weight'], sd[pre + 'net.4.module.bias'])
    h = F.relu(actnorm(h, sd, pre + 'net.5.module.', training, False))
    h = mc_mask(h, ind, sd[pre + 'net.7.codebook'])
    log_s, t = zero_conv(h, sd, pre + 'net.8.module.').chunk(2, 1)
    s = torch.sigmoid(log_s + 2)
    out_b = (in_b + t) * s
    return torch.cat([in_a, out_b], 1), torch.sum(torch.log(s).view(x.size(0), -1), 1)


def flow(x, ind, sd, pre, training):
    out, logdet = actnorm(x, sd, pre + 'actnorm.', training, True)
    out = F.conv2d(out, invconv_weight(sd, pre + 'invconv.'))
    det1 = x.size(2) * x.size(3) * torch.sum(sd[pre + 'invconv.w_s'])
    out, det2 = coupling(out, ind, sd, pre + 'coupling.', training)
    return out, logdet + det1 + det2


def block(x, ind, sd, pre, training, split):
    b, c, h, w = x.shape
    out = x.view(b, c, h // 2, 2, w // 2, 2).permute(0, 1, 3, 5, 2, 4).reshape(b, c * 4, h // 2, w // 2)
    logdet = 0
    k = 0
    while (pre + f'flows.{k}.actnorm.loc') in sd:
        out, det = flow(out, ind, sd, pre + f'flows.{k}.', training)
        logdet = logdet + det
        k += 1
    if split:
        out, z_new = out.chunk(2, 1)
        mean, log_sd = zero_conv(out, sd, pre + 'prior.').chunk(2, 1)
    else:
        mean, log_sd = zero_conv(torch.zeros_like(out), sd, pre + 'prior.').chunk(2, 1)
        z_new = out
    log_p = (-0.5 * log(2 * pi) - log_sd - 0.5 * (z_new - mean) ** 2 / torch.exp(2 * log_sd)).view(b, -1).sum(1)
    return out, logdet, log_p, z_new


def forward(img, label, sd, num_mode, noise, training=True):
    """MCGlow.forward (src/models/mcglow.py:295-312); ``noise`` ~ U[0,1) replaces torch.rand_like."""
    ind = one_hot(label, num_mode)
    x = img * 0.5 + noise / 256
    n_blocks = 0
    while f'blocks.{n_blocks}.prior.scale' in sd:
        n_blocks += 1
    z, log_p_sum, logdet = [], 0, 0
    for i in range(n_blocks):
        x, det, log_p, z_new = block(x, ind, sd, f'blocks.{i}.', training, split=i < n_blocks - 1)
        z.append(z_new)
        logdet = logdet + det
        log_p_sum = log_p_sum + log_p
    n_pixel = img[0].numel()
    loss = -float(log(256.) * n_pixel) + logdet + log_p_sum
    loss = -loss / float(log(2.) * n_pixel)
    if training:
        loss = torch.where(torch.isnan(loss), torch.zeros_like(loss), loss)
    return {'loss': loss.mean(), 'z': z}


def train_step(sd, opt, img, label, num_mode, noise):
    opt.zero_grad()
    out = forward(img, label, sd, num_mode, noise)
    out['loss'].backward()
    clip_grad_norm_(opt.params, 1.0)
    opt.step()
    return out


def make_adam(sd, lr=3e-4):
    return Adam(trainable(sd), lr, (0.9, 0.999))


# ----------------------------------------------------------------------------- sampling path (reverse)
def coupling_reverse(out, ind, sd, pre):
    """AffineCoupling.reverse (src/models/mcglow.py:166-175): in_b = out_b / sigmoid(log_s + 2) - t."""
    out_a, out_b = out.chunk(2, 1)
    h = F.conv2d(out_a, sd[pre + 'net.0.module.weight'], sd[pre + 'net.0.module.bias'], padding=1)
    h = F.relu(actnorm(h, sd, pre + 'net.1.module.', False, False))
    h = mc_mask(h, ind, sd[pre + 'net.3.codebook'])
    h = F.conv2d(h, sd[pre + 'net.4.module.weight'], sd[pre + 'net.4.module.bias'])
    h = F.relu(actnorm(h, sd, pre + 'net.5.module.', False, False))
    h = mc_mask(h, ind, sd[pre + 'net.7.codebook'])
    log_s, t = zero_conv(h, sd, pre + 'net.8.module.').chunk(2, 1)
    in_b = out_b / torch.sigmoid(log_s + 2) - t
    return torch.cat([out_a, in_b], 1)


def flow_reverse(out, ind, sd, pre):
    """Flow.reverse (src/models/mcglow.py:195-201): coupling^-1, invconv^-1 (conv with W^-1, :114-116),
    actnorm^-1 (output / scale - loc, :54-55)."""
    x = coupling_reverse(out, ind, sd, pre + 'coupling.')
    w = invconv_weight(sd, pre + 'invconv.')
    x = F.conv2d(x, w.squeeze().inverse().unsqueeze(2).unsqueeze(3))
    return x / sd[pre + 'actnorm.scale'] - sd[pre + 'actnorm.loc']


def block_reverse(output, ind, sd, pre, split, eps, reconstruct):
    """Block.reverse (src/models/mcglow.py:241-265)."""
    x = output
    if reconstruct:
        x = torch.cat([output, eps], 1) if split else eps
    elif split:
        mean, log_sd = zero_conv(x, sd, pre + 'prior.').chunk(2, 1)
        x = torch.cat([output, mean + torch.exp(log_sd) * eps], 1)
    else:
        mean, log_sd = zero_conv(torch.zeros_like(x), sd, pre + 'prior.').chunk(2, 1)
        x = mean + torch.exp(log_sd) * eps
    k = 0
    while (pre + f'flows.{k}.actnorm.loc') in sd:
        k += 1
    for i in range(k - 1, -1, -1):
        x = flow_reverse(x, ind, sd, pre + f'flows.{i}.')
    b, c, h, w = x.shape
    return x.view(b, c // 4, 2, 2, h, w).permute(0, 1, 4, 2, 5, 3).reshape(b, c // 4, h * 2, w * 2)


def reverse(z, label, sd, num_mode, reconstruct):
    """MCGlow.reverse (src/models/mcglow.py:314-325); ``generate`` is reverse(z ~ N(0, T), reconstruct=False)."""
    ind = one_hot(label, num_mode)
    n_blocks = 0
    while f'blocks.{n_blocks}.prior.scale' in sd:
        n_blocks += 1
    x = None
    for i in range(n_blocks):
        b = n_blocks - 1 - i
        x = block_reverse(z[-1] if i == 0 else x, ind, sd, f'blocks.{b}.', split=b < n_blocks - 1,
                          eps=z[-(i + 1)], reconstruct=reconstruct)
    return torch.clamp(x, -.5, .5) * 2
