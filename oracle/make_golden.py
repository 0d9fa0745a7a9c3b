"""Generate tests/golden/*.{json,npz} FROM THE UNMODIFIED REFERENCE (test infrastructure).

Run in the build container, where the reference is mounted read-only:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

It imports /root/reference/src exactly as SURVEY.md Appendix C describes (cwd = src so that
config.py:5 finds config.yml), builds MCGAN models from fixed seeds, runs the restated
train() body (src/train_gan.py:139-176, with injected z so every consumer sees the same
noise) through the reference modules and torch.optim.Adam, and records small fixtures:
state_dict hashes / key lists, codebooks, forward outputs, losses and per-parameter update
norms.  Nothing here is imported by the product; the GPU box never needs /root/reference.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, 'tests', 'golden')
REF_SRC = '/root/reference/src'

CASES = {
    # name: (data_name, classes, G widths, D widths, batch)
    'tiny_cifar': ('CIFAR10', 10, [64, 64, 64, 64], [64, 64, 64, 64], 8),
    'tiny_omni': ('Omniglot', 50, [128, 64, 64, 64], [64, 64, 128, 128], 8),
}


def sd_hash(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def build(cfg, models, data_name, classes, g_w, d_w, seed=0):
    from utils import process_control
    cfg.update(device='cpu', data_name=data_name, model_name='mcgan', control={'controller_rate': '0.5'})
    process_control()
    cfg['classes_size'] = classes
    if g_w is not None:
        cfg['gan']['generator_hidden_size'] = list(g_w)
        cfg['gan']['discriminator_hidden_size'] = list(d_w)
    torch.manual_seed(seed)
    return models.mcgan()


def make_inputs(cfg, classes, batch, n_z=6, seed=1):
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(batch, *cfg['data_shape'], generator=g) * 2 - 1
    label = torch.randint(0, classes, (batch,), generator=g)
    zs = [torch.randn(batch, cfg['gan']['latent_size'], generator=g) for _ in range(n_z)]
    return img, label, zs


def reference_iteration(model, img, label, zs):
    """src/train_gan.py:139-176 restated over the reference model (hinge, Adam 2e-4, (0.5, 0.999))."""
    opt = {'generator': torch.optim.Adam(model.generator.parameters(), lr=2e-4, weight_decay=0, betas=(0.5, 0.999)),
           'discriminator': torch.optim.Adam(model.discriminator.parameters(), lr=2e-4, weight_decay=0,
                                             betas=(0.5, 0.999))}
    zi = iter(zs)
    model.train(True)
    for _ in range(5):
        opt['discriminator'].zero_grad()
        opt['generator'].zero_grad()
        d_x = model.discriminate(img, label)
        generated = model.generate(label, next(zi))
        d_g = model.discriminate(generated.detach(), label)
        d_loss = torch.nn.functional.relu(1.0 - d_x).mean() + torch.nn.functional.relu(1.0 + d_g).mean()
        d_loss.backward()
        opt['discriminator'].step()
    for _ in range(1):
        opt['discriminator'].zero_grad()
        opt['generator'].zero_grad()
        generated = model.generate(label, next(zi))
        g_loss = -model.discriminate(generated, label).mean()
        g_loss.backward()
        opt['generator'].step()
    return float(d_loss), float(g_loss)


def main():
    os.makedirs(OUT, exist_ok=True)
    os.chdir(REF_SRC)
    sys.path.insert(0, '.')
    sys.dont_write_bytecode = True
    from config import cfg
    cfg.update(device='cpu', data_name='CIFAR10', model_name='mcgan', control={'controller_rate': '0.5'})
    import models
    torch.set_num_threads(8)
    meta = {'torch': torch.__version__, 'cases': {}}

    # --- full-width CIFAR10 / Omniglot models: hashes only (weights are re-derived from the seed)
    for name, data_name, classes in (('cifar10_full', 'CIFAR10', 10), ('omniglot_full', 'Omniglot', 1623)):
        m = build(cfg, models, data_name, classes, None, None, seed=0)
        sd = m.state_dict()
        h = hashlib.sha256()
        for n, b in m.named_buffers():
            if n.endswith('codebook'):
                h.update(b.numpy().tobytes())
        meta['cases'][name] = {
            'data_name': data_name, 'classes': classes, 'seed': 0,
            'state_dict_sha256': sd_hash(sd), 'codebook_sha256': h.hexdigest(),
            'keys': [[k, list(v.shape)] for k, v in sd.items()],
            'n_params_g': sum(p.numel() for p in m.generator.parameters()),
            'n_params_d': sum(p.numel() for p in m.discriminator.parameters()),
        }
        if name == 'cifar10_full':
            np.savez_compressed(os.path.join(OUT, 'cifar10_full_codebooks.npz'),
                                **{n: np.packbits(b.numpy().astype(np.uint8), axis=1)
                                   for n, b in m.named_buffers() if n.endswith('codebook')})

    # --- tiny models: forward outputs + one full reference iteration
    for name, (data_name, classes, g_w, d_w, batch) in CASES.items():
        m = build(cfg, models, data_name, classes, g_w, d_w, seed=0)
        img, label, zs = make_inputs(cfg, classes, batch)
        sd0 = {k: v.clone() for k, v in m.state_dict().items()}
        case = {'data_name': data_name, 'classes': classes, 'g_widths': g_w, 'd_widths': d_w, 'batch': batch,
                'seed': 0, 'input_seed': 1, 'state_dict_sha256': sd_hash(sd0),
                'keys': [[k, list(v.shape)] for k, v in sd0.items()]}
        arrays = {'img': img.numpy(), 'label': label.numpy(), 'zs': torch.stack(zs).numpy()}
        # forward-only outputs in train mode on a scratch copy (train-mode forwards mutate BN / SN buffers)
        import copy
        scratch = copy.deepcopy(m)
        scratch.train(True)
        with torch.no_grad():
            gen = scratch.generate(label, zs[0])
            d_real = scratch.discriminate(img, label)
        arrays['gen_out'] = gen.numpy()
        arrays['d_real'] = d_real.numpy()
        # gradients of one G+D forward/backward (no optimiser): -D(G(z)).mean()
        scratch = copy.deepcopy(m)
        scratch.train(True)
        loss = -scratch.discriminate(scratch.generate(label, zs[0]), label).mean()
        loss.backward()
        case['g_loss_fwd'] = float(loss)
        case['grad_norms'] = {k: float(p.grad.norm()) for k, p in scratch.named_parameters() if p.grad is not None}
        # one full iteration
        d_loss, g_loss = reference_iteration(m, img, label, zs)
        sd1 = m.state_dict()
        case['loss_D'], case['loss_G'] = d_loss, g_loss
        case['delta_norms'] = {k: float((sd1[k].float() - sd0[k].float()).norm()) for k in sd0
                               if sd0[k].is_floating_point()}
        case['after_norms'] = {k: float(sd1[k].float().norm()) for k in sd0 if sd0[k].is_floating_point()}
        for k in sd1:
            if k.endswith('running_mean') or k.endswith('running_var') or k.endswith('weight_u'):
                arrays['after/' + k] = sd1[k].numpy()
        np.savez_compressed(os.path.join(OUT, name + '.npz'), **arrays)
        meta['cases'][name] = case
        print(name, 'loss_D', d_loss, 'loss_G', g_loss)

    # --- MCVAE (COIL100-shaped, 100 modes): forward, gradients and one train step through the reference
    from utils import process_control
    cfg.update(device='cpu', data_name='COIL100', model_name='mcvae', control={'controller_rate': '0.5'})
    process_control()
    cfg['classes_size'] = 100
    torch.manual_seed(0)
    m = models.mcvae()
    meta['cases']['vae_full'] = {'data_name': 'COIL100', 'classes': 100, 'seed': 0, 'model': 'mcvae',
                                 'state_dict_sha256': sd_hash(m.state_dict()),
                                 'keys': [[k, list(v.shape)] for k, v in m.state_dict().items()]}
    batch = 16
    g = torch.Generator().manual_seed(1)
    img = torch.rand(batch, 3, 32, 32, generator=g) * 2 - 1
    label = torch.randint(0, 100, (batch,), generator=g)
    eps = torch.randn(batch, cfg['vae']['latent_size'], generator=g)
    _randn_like = torch.randn_like
    torch.randn_like = lambda t: eps.clone()                  # inject the reparameterisation noise (mcvae.py:60)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    import copy
    scratch = copy.deepcopy(m)
    scratch.train(True)
    out = scratch({'img': img.clone(), 'label': label})
    out['loss'].backward()
    case = {'data_name': 'COIL100', 'classes': 100, 'seed': 0, 'input_seed': 1, 'batch': batch, 'model': 'mcvae',
            'state_dict_sha256': sd_hash(sd0), 'loss_fwd': float(out['loss'].detach()),
            'grad_norms': {k: float(p.grad.norm()) for k, p in scratch.named_parameters() if p.grad is not None}}
    arrays = {'img': img.numpy(), 'label': label.numpy(), 'eps': eps.numpy(),
              'recon': out['img'].detach().numpy(), 'mu': out['mu'].detach().numpy(),
              'logvar': out['logvar'].detach().numpy()}
    opt = torch.optim.Adam(m.parameters(), lr=3e-4, weight_decay=0)
    m.train(True)
    opt.zero_grad()
    out = m({'img': img.clone(), 'label': label})
    out['loss'].backward()
    torch.nn.utils.clip_grad_norm_(m.parameters(), 1)
    opt.step()
    out2 = m({'img': img.clone(), 'label': label})
    sd1 = m.state_dict()
    case['loss_step1'] = float(out['loss'].detach())
    case['loss_step2'] = float(out2['loss'].detach())
    case['delta_norms'] = {k: float((sd1[k].float() - sd0[k].float()).norm()) for k in sd0 if sd0[k].is_floating_point()}
    np.savez_compressed(os.path.join(OUT, 'vae_full.npz'), **arrays)
    meta['cases']['vae_step'] = case
    torch.randn_like = _randn_like
    print('vae loss', case['loss_fwd'], case['loss_step2'])

    # --- MCGlow (CIFAR10-shaped; K=2 flows per level and hidden 128 keep the fixture small — the full
    #     K=16 / hidden 512 model is the same code path): data-dependent ActNorm init, loss, gradients, one step
    cfg.update(device='cpu', data_name='CIFAR10', model_name='mcglow', control={'controller_rate': '0.5'})
    process_control()
    cfg['classes_size'] = 10
    cfg['glow']['K'] = 2
    cfg['glow']['hidden_size'] = 128
    torch.manual_seed(0)
    np.random.seed(0)
    m = models.mcglow()
    batch = 16
    g = torch.Generator().manual_seed(1)
    img = torch.rand(batch, 3, 32, 32, generator=g) * 2 - 1
    label = torch.randint(0, 10, (batch,), generator=g)
    noise = torch.rand(batch, 3, 32, 32, generator=g)
    _rand_like = torch.rand_like
    torch.rand_like = lambda t: noise.clone()                 # inject the dequantisation noise (mcglow.py:299)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    case = {'data_name': 'CIFAR10', 'classes': 10, 'seed': 0, 'np_seed': 0, 'input_seed': 1, 'batch': batch,
            'model': 'mcglow', 'K': 2, 'hidden_size': 128, 'state_dict_sha256': sd_hash(sd0),
            'keys': [[k, list(v.shape)] for k, v in sd0.items()]}
    m.train(True)
    with torch.no_grad():                                     # train_glow.py:60-67: the init pass
        out_init = m({'img': img, 'label': label})
    case['loss_init'] = float(out_init['loss'])
    sd_init = {k: v.clone() for k, v in m.state_dict().items()}
    out = m({'img': img, 'label': label})
    out['loss'].backward()
    case['loss_fwd'] = float(out['loss'].detach())
    case['grad_norms'] = {k: float(p.grad.norm()) for k, p in m.named_parameters() if p.grad is not None}
    arrays = {'img': img.numpy(), 'label': label.numpy(), 'noise': noise.numpy()}
    for i, zt in enumerate(out['z']):
        arrays[f'z{i}'] = zt.detach().numpy()
    for k in ('blocks.0.flows.0.actnorm.loc', 'blocks.0.flows.0.coupling.net.1.module.scale',
              'blocks.2.flows.1.coupling.net.5.module.loc'):
        arrays['init/' + k] = sd_init[k].numpy()
    opt = torch.optim.Adam(m.parameters(), lr=3e-4, weight_decay=0)
    opt.zero_grad()
    out = m({'img': img, 'label': label})
    out['loss'].backward()
    torch.nn.utils.clip_grad_norm_(m.parameters(), 1)
    opt.step()
    out2 = m({'img': img, 'label': label})
    case['loss_step2'] = float(out2['loss'].detach())
    torch.rand_like = _rand_like
    np.savez_compressed(os.path.join(OUT, 'glow_small.npz'), **arrays)
    meta['cases']['glow_small'] = case
    print('glow loss', case['loss_init'], case['loss_fwd'], case['loss_step2'])

    with open(os.path.join(OUT, 'mcgan_golden.json'), 'w') as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print('wrote', OUT)


if __name__ == '__main__':
    main()
