"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU restatement (fp32, NCHW, torch.nn.functional) of the reference's MCVAE path as pure functions
over a ``state_dict``-shaped mapping:

    res_block     src/models/mcvae.py:17-35     Conv3x3→BN→ReLU→MC→Conv3x3→BN→MC; +x; ReLU
    encoder       src/models/mcvae.py:38-68     [Conv4x4s2→BN→ReLU→MC]xL, ResBlocks, mu/logvar, reparameterise
    decoder       src/models/mcvae.py:71-101    MC→Linear→BN1d→ReLU, MC, ResBlocks, [ConvT4x4s2→BN→ReLU→MC], ConvT→Sigmoid
    forward       src/models/mcvae.py:133-144   rescale to [0,1], BCE(sum)+KLD over numel (:10-14)
    train_step    src/train_vae.py:106-111      backward → clip_grad_norm_(1) → Adam(3e-4)

Pinned against fixtures generated from the unmodified reference (oracle/make_golden.py →
tests/golden/tiny_vae.npz, mcgan_golden.json["cases"]["tiny_vae"]).
"""
import torch
import torch.nn.functional as F

from .mcgan_oracle import Adam, _bn, make_trainable, mc_mask, one_hot  # noqa: F401


def _count(sd, fmt):
    i = 0
    while fmt.format(i) in sd:
        i += 1
    return i


def res_block(x, ind, sd, pre, training):
    h = F.conv2d(x, sd[pre + 'conv.0.module.weight'], sd[pre + 'conv.0.module.bias'], padding=1)
    h = F.relu(_bn(h, sd, pre + 'conv.1.module.', training))
    h = mc_mask(h, ind, sd[pre + 'conv.3.codebook'])
    h = F.conv2d(h, sd[pre + 'conv.4.module.weight'], sd[pre + 'conv.4.module.bias'], padding=1)
    h = _bn(h, sd, pre + 'conv.5.module.', training)
    h = mc_mask(h, ind, sd[pre + 'conv.6.codebook'])
    return F.relu(h + x)


def _layout(sd, pre):
    """(number of stride-2 stages, index of the first ResBlock, number of ResBlocks) of a blocks list."""
    idx = sorted({int(k[len(pre):].split('.')[0]) for k in sd if k.startswith(pre)})
    res = [i for i in idx if (pre + f'{i}.conv.0.module.weight') in sd]
    return res


def encoder(x, ind, sd, training, eps, pre='encoder.'):
    res = _layout(sd, pre + 'blocks.')
    n_down = res[0] // 4
    h = x
    for i in range(n_down):
        b = pre + f'blocks.{4 * i}.module.'
        h = F.conv2d(h, sd[b + 'weight'], sd[b + 'bias'], stride=2, padding=1)
        h = F.relu(_bn(h, sd, pre + f'blocks.{4 * i + 1}.module.', training))
        h = mc_mask(h, ind, sd[pre + f'blocks.{4 * i + 3}.codebook'])
    for i in res:
        h = res_block(h, ind, sd, pre + f'blocks.{i}.', training)
    h = h.view(h.size(0), -1)
    mu = F.linear(h, sd[pre + 'mu.weight'], sd[pre + 'mu.bias'])
    logvar = F.linear(h, sd[pre + 'logvar.weight'], sd[pre + 'logvar.bias'])
    z = mu + eps * torch.exp(0.5 * logvar) if training else mu
    return z, mu, logvar


def decoder(z, ind, sd, training, pre='decoder.'):
    h = mc_mask(z, ind, sd[pre + 'linear.0.codebook'])
    h = F.linear(h, sd[pre + 'linear.1.module.weight'], sd[pre + 'linear.1.module.bias'])
    b = pre + 'linear.2.module.'
    if training:
        sd[b + 'num_batches_tracked'] += 1
    h = F.relu(F.batch_norm(h, sd[b + 'running_mean'], sd[b + 'running_var'], sd[b + 'weight'], sd[b + 'bias'],
                            training, 0.1, 1e-5))
    res = _layout(sd, pre + 'blocks.')
    c = sd[pre + f'blocks.{res[0]}.conv.0.module.weight'].shape[0]
    side = int((h.size(1) // c) ** 0.5)
    h = h.view(h.size(0), c, side, side)
    h = mc_mask(h, ind, sd[pre + 'blocks.0.codebook'])
    for i in res:
        h = res_block(h, ind, sd, pre + f'blocks.{i}.', training)
    i = res[-1] + 1
    while (pre + f'blocks.{i + 1}.module.running_mean') in sd:
        h = F.conv_transpose2d(h, sd[pre + f'blocks.{i}.module.weight'], sd[pre + f'blocks.{i}.module.bias'],
                               stride=2, padding=1)
        h = F.relu(_bn(h, sd, pre + f'blocks.{i + 1}.module.', training))
        h = mc_mask(h, ind, sd[pre + f'blocks.{i + 3}.codebook'])
        i += 4
    h = F.conv_transpose2d(h, sd[pre + f'blocks.{i}.module.weight'], sd[pre + f'blocks.{i}.module.bias'],
                           stride=2, padding=1)
    return torch.sigmoid(h)


def forward(img, label, sd, num_mode, eps, training=True):
    """MCVAE.forward (src/models/mcvae.py:133-144); img in [-1, 1]."""
    x = (img + 1) / 2
    ind = one_hot(label, num_mode)
    z, mu, logvar = encoder(x, ind, sd, training, eps)
    dec = decoder(z, ind, sd, training)
    bce = F.binary_cross_entropy(dec, x, reduction='sum')
    kld = 0.5 * torch.sum(mu.pow(2) + logvar.exp() - 1 - logvar)
    return {'loss': (bce + kld) / x.numel(), 'img': dec * 2 - 1, 'mu': mu, 'logvar': logvar}


def clip_grad_norm_(params, max_norm):
    """torch.nn.utils.clip_grad_norm_ (L2): scale all grads by min(1, max_norm / (total + 1e-6))."""
    grads = [p.grad for p in params if p.grad is not None]
    total = torch.sqrt(sum((g.detach() ** 2).sum() for g in grads))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return total


def trainable(sd):
    return [v for v in sd.values() if v.requires_grad]


def train_step(sd, opt, img, label, num_mode, eps):
    """src/train_vae.py:106-111."""
    opt.zero_grad()
    out = forward(img, label, sd, num_mode, eps)
    out['loss'].backward()
    clip_grad_norm_(opt.params, 1.0)
    opt.step()
    return out


def make_adam(sd, lr=3e-4):
    return Adam(trainable(sd), lr, (0.9, 0.999))
