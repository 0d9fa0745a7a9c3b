"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU restatement (fp32, NCHW, plain ``torch.nn.functional`` calls) of the reference's MCGAN
hot path, written as pure functions over a ``state_dict``-shaped mapping so that the same
weights can be pushed through the reference, this oracle and the B200 kernels:

    mc_mask            src/modules/modules.py:71-76   (code = indicator @ codebook; x * code)
    make_codebook      src/modules/modules.py:58-69
    sn_weight          src/models/utils.py:17-21 -> torch/nn/utils/spectral_norm.py compute_weight
    gen_block          src/models/mcgan.py:9-44       (GenResBlock)
    generator          src/models/mcgan.py:47-69
    first_dis_block    src/models/mcgan.py:72-93
    dis_block          src/models/mcgan.py:96-138
    discriminator      src/models/mcgan.py:150-181
    train_iteration    src/train_gan.py:139-176       (5 D updates + 1 G update, hinge, Adam)

The arithmetic of the reference lives in PyTorch itself (unpinned; README.md:8 cites a
requirements.txt that is not shipped), so the oracle calls the same torch CPU ops.  Parity
pinning: the reference has no tests or golden vectors; this oracle is
pinned against the reference imported from /root/reference in this container
(tests/test_oracle_golden.py) through the fixtures in
tests/golden/ that oracle/make_golden.py generated FROM THE REFERENCE.
"""
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- operator
def make_codebook(num_mode, input_size, controller_rate):
    """src/modules/modules.py:58-69 — consumes torch's global CPU RNG exactly like the reference."""
    if controller_rate == 1:
        return torch.ones(num_mode, input_size)
    dist = torch.distributions.bernoulli.Bernoulli(probs=controller_rate)
    seen = set()
    while len(seen) < num_mode:
        block = dist.sample((num_mode, input_size)).tolist()
        seen.update(tuple(row) for row in block)
    return torch.tensor(list(seen)[:num_mode], dtype=torch.float)


def mc_mask(x, indicator, codebook):
    """src/modules/modules.py:71-76."""
    code = indicator.matmul(codebook)
    code = code.view(*code.shape, *([1] * (x.dim() - 2)))
    return x * code.detach()


def one_hot(labels, num_mode):
    return F.one_hot(labels, num_mode).float()


# ----------------------------------------------------------------------------- layers
def _bn(x, sd, key, training):
    """nn.BatchNorm2d forward (train: biased batch var, running stats momentum 0.1 with the
    unbiased var, num_batches_tracked += 1)."""
    if training:
        sd[key + 'num_batches_tracked'] += 1
    return F.batch_norm(x, sd[key + 'running_mean'], sd[key + 'running_var'], sd[key + 'weight'],
                        sd[key + 'bias'], training, 0.1, 1e-5)


def sn_weight(sd, key, training):
    """torch legacy spectral_norm ``compute_weight``: one power iteration (in-place on the
    u / v buffers, eps 1e-12) when training, sigma = u.(W v), W / sigma."""
    w, u, v = sd[key + 'weight_orig'], sd[key + 'weight_u'], sd[key + 'weight_v']
    mat = w.reshape(w.shape[0], -1)
    if training:
        with torch.no_grad():
            v.copy_(F.normalize(torch.mv(mat.t(), u), dim=0, eps=1e-12))
            u.copy_(F.normalize(torch.mv(mat, v), dim=0, eps=1e-12))
        u, v = u.clone(), v.clone()
    sigma = torch.dot(u, torch.mv(mat, v))
    return w / sigma


def _up(x):
    return F.interpolate(x, scale_factor=2, mode='nearest')


# ----------------------------------------------------------------------------- generator
def gen_block(x, ind, sd, pre, training):
    """GenResBlock (src/models/mcgan.py:9-44): shortcut = Conv1x1(mc_1(Up(x))) on the raw input;
    main = BN→ReLU→Up→mc_1→Conv3x3→BN→ReLU→mc_2→Conv3x3; sum."""
    cb1, cb2 = sd[pre + 'mc_1.codebook'], sd[pre + 'mc_2.codebook']
    sc = F.conv2d(mc_mask(_up(x), ind, cb1), sd[pre + 'shortcut.2.module.weight'], sd[pre + 'shortcut.2.module.bias'])
    h = _up(F.relu(_bn(x, sd, pre + 'conv.0.module.', training)))
    h = F.conv2d(mc_mask(h, ind, cb1), sd[pre + 'conv.4.module.weight'], sd[pre + 'conv.4.module.bias'], padding=1)
    h = F.relu(_bn(h, sd, pre + 'conv.5.module.', training))
    h = F.conv2d(mc_mask(h, ind, cb2), sd[pre + 'conv.8.module.weight'], sd[pre + 'conv.8.module.bias'], padding=1)
    return h + sc


def generator(z, ind, sd, training=True, pre='generator.'):
    """Generator.forward (src/models/mcgan.py:64-69)."""
    x = F.linear(z, sd[pre + 'linear.module.weight'], sd[pre + 'linear.module.bias'])
    x = x.view(x.size(0), -1, 4, 4)
    i = 0
    while (pre + f'blocks.{i}.mc_1.codebook') in sd:
        x = gen_block(x, ind, sd, pre + f'blocks.{i}.', training)
        i += 1
    x = F.relu(_bn(x, sd, pre + f'blocks.{i}.module.', training))
    x = mc_mask(x, ind, sd[pre + f'blocks.{i + 2}.codebook'])
    x = F.conv2d(x, sd[pre + f'blocks.{i + 3}.module.weight'], sd[pre + f'blocks.{i + 3}.module.bias'], padding=1)
    return torch.tanh(x)


# ----------------------------------------------------------------------------- discriminator
def _sn_conv(x, sd, key, training, padding):
    return F.conv2d(x, sn_weight(sd, key, training), sd[key + 'bias'], padding=padding)


def first_dis_block(x, ind, sd, pre, training):
    """FirstDisResBlock (src/models/mcgan.py:72-93).  The reference evaluates the shortcut
    before the main branch (``forward`` :88-93) — the order matters for nothing here because
    each spectral-normed layer owns its u/v."""
    sc = F.avg_pool2d(_sn_conv(x, sd, pre + 'shortcut.0.module.', training, 0), 2)
    h = F.relu(_sn_conv(x, sd, pre + 'conv.0.module.', training, 1))
    h = mc_mask(h, ind, sd[pre + 'mc_1.codebook'])
    h = F.avg_pool2d(_sn_conv(h, sd, pre + 'conv.3.module.', training, 1), 2)
    return h + sc


def dis_block(x, ind, sd, pre, training, stride):
    """DisResBlock (src/models/mcgan.py:96-138): the conv shortcut gates the RAW input with mc_1;
    stride-2 blocks end both branches with AvgPool2d(2)."""
    cb1, cb2 = sd[pre + 'mc_1.codebook'], sd[pre + 'mc_2.codebook']
    if (pre + 'shortcut.1.module.weight_orig') in sd:
        sc = _sn_conv(mc_mask(x, ind, cb1), sd, pre + 'shortcut.1.module.', training, 0)
    else:
        sc = x
    h = mc_mask(F.relu(x), ind, cb1)
    h = F.relu(_sn_conv(h, sd, pre + 'conv.2.module.', training, 1))
    h = _sn_conv(mc_mask(h, ind, cb2), sd, pre + 'conv.5.module.', training, 1)
    if stride == 2:
        h = F.avg_pool2d(h, 2)
        sc = F.avg_pool2d(sc, 2)
    return h + sc


def discriminator(x, ind, sd, strides, training=True, pre='discriminator.'):
    """Discriminator.forward (src/models/mcgan.py:178-181).  ``strides``: stride of each
    DisResBlock after the first block (AvgPool has no parameters, so the state_dict alone
    cannot tell) — CIFAR10: [2, 1, 1]; other 32x32 data: [2, 2, 1] (mcgan.py:154-175)."""
    x = first_dis_block(x, ind, sd, pre + 'blocks.0.', training)
    for i, s in enumerate(strides):
        x = dis_block(x, ind, sd, pre + f'blocks.{i + 1}.', training, s)
    i = len(strides) + 1
    x = mc_mask(F.relu(x), ind, sd[pre + f'blocks.{i + 1}.codebook'])
    x = x.sum(dim=[-2, -1]).view(x.size(0), -1)
    key = pre + f'blocks.{i + 3}.module.'
    return F.linear(x, sn_weight(sd, key, training), sd[key + 'bias'])


# ----------------------------------------------------------------------------- training step
def split_params(sd, prefix):
    """Names of the trainable tensors of one net, in ``nn.Module.parameters()`` order.  For the
    reference's module tree that is state_dict order restricted to weight/bias/weight_orig."""
    names = []
    for k in sd:
        if not k.startswith(prefix):
            continue
        leaf = k.rsplit('.', 1)[1]
        if leaf in ('weight', 'bias', 'weight_orig'):
            names.append(k)
    return names


class Adam:
    """torch.optim.Adam (defaults eps=1e-8, no amsgrad, no weight decay) restated per tensor
    (src/train_gan.py:223-236 -> torch/optim/adam.py _single_tensor_adam)."""

    def __init__(self, params, lr, betas):
        self.params, self.lr, self.b1, self.b2 = params, lr, betas[0], betas[1]
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.t = 0

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    def step(self):
        self.t += 1
        bc1 = 1 - self.b1 ** self.t
        bc2 = 1 - self.b2 ** self.t
        with torch.no_grad():
            for p, m, v in zip(self.params, self.m, self.v):
                if p.grad is None:
                    continue
                g = p.grad
                m.lerp_(g, 1 - self.b1)
                v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                denom = (v.sqrt() / (bc2 ** 0.5)).add_(1e-8)
                p.addcdiv_(m, denom, value=-self.lr / bc1)


def make_trainable(sd):
    """Clone a state_dict into leaf tensors (params require grad, buffers do not)."""
    out = {}
    for k, t in sd.items():
        t = t.detach().clone()
        leaf = k.rsplit('.', 1)[1]
        if leaf in ('weight', 'bias', 'weight_orig') and t.is_floating_point():
            t.requires_grad_(True)
        out[k] = t
    return out


def train_iteration(sd, opt, img, label, zs, num_mode, strides, n_d=5, n_g=1):
    """One reference iteration (src/train_gan.py:139-176) with injected noise ``zs``
    (n_d + n_g tensors).  ``opt`` = {'generator': Adam, 'discriminator': Adam}.
    Returns (loss_D, loss_G) of the last D / G update."""
    ind = one_hot(label, num_mode)
    zi = iter(zs)
    d_loss = g_loss = None
    for _ in range(n_d):
        opt['discriminator'].zero_grad()
        opt['generator'].zero_grad()
        d_x = discriminator(img, ind, sd, strides)
        fake = generator(next(zi), ind, sd)
        d_g = discriminator(fake.detach(), ind, sd, strides)
        d_loss = F.relu(1.0 - d_x).mean() + F.relu(1.0 + d_g).mean()
        d_loss.backward()
        opt['discriminator'].step()
    for _ in range(n_g):
        opt['discriminator'].zero_grad()
        opt['generator'].zero_grad()
        fake = generator(next(zi), ind, sd)
        g_loss = -discriminator(fake, ind, sd, strides).mean()
        g_loss.backward()
        opt['generator'].step()
    return d_loss.detach(), g_loss.detach()


def make_adams(sd, lr=2e-4, betas=(0.5, 0.999)):
    return {
        'generator': Adam([sd[k] for k in split_params(sd, 'generator.')], lr, betas),
        'discriminator': Adam([sd[k] for k in split_params(sd, 'discriminator.')], lr, betas),
    }
