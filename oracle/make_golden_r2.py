"""Round-2 fixtures FROM THE UNMODIFIED REFERENCE (test infrastructure; run in the build container only):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_r2.py

    eval_generate.npz     the sampling path: eval-mode ``generate`` of MCGAN / MCVAE / MCGlow
                          (src/train_gan.py:197-220, src/models/mcvae.py:126-131, src/models/mcglow.py:314-350)
                          after train-mode forwards have moved the BatchNorm running statistics, plus Glow's
                          ``reverse(reconstruct=True)``
    pixelcnn_small.npz    the MC-gated PixelCNN (§8 f3, src/models/mcpixelcnn.py): logits, loss, gradient norms
    r2_golden.json        scalars, key lists and state_dict hashes of the above

The reference is imported exactly as SURVEY.md Appendix C describes.  Nothing here is imported by the product.
"""
import copy
import hashlib
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, 'tests', 'golden')
REF_SRC = '/root/reference/src'


def sd_hash(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def main():
    os.chdir(REF_SRC)
    sys.path.insert(0, '.')
    sys.dont_write_bytecode = True
    from config import cfg
    cfg.update(device='cpu', data_name='CIFAR10', model_name='mcgan', control={'controller_rate': '0.5'})
    import models
    from utils import process_control
    torch.set_num_threads(8)
    meta, arrays = {'torch': torch.__version__}, {}

    # ---- MCGAN (tiny_cifar widths): two train-mode generator forwards, then eval-mode generate
    process_control()
    cfg['classes_size'] = 10
    cfg['gan']['generator_hidden_size'] = [64, 64, 64, 64]
    cfg['gan']['discriminator_hidden_size'] = [64, 64, 64, 64]
    torch.manual_seed(0)
    m = models.mcgan()
    g = torch.Generator().manual_seed(11)
    label = torch.randint(0, 10, (8,), generator=g)
    zs = [torch.randn(8, 128, generator=g) for _ in range(3)]
    m.train(True)
    with torch.no_grad():
        m.generate(label, zs[0])
        m.generate(label, zs[1])
    m.train(False)
    with torch.no_grad():
        out = m.generate(label, zs[2])
        d = m.discriminate(out, label)
    arrays.update({'gan/label': label.numpy(), 'gan/zs': torch.stack(zs).numpy(), 'gan/generated': out.numpy(),
                   'gan/d_eval': d.numpy()})

    # ---- MCVAE (COIL100-shaped): one train-mode forward, then eval-mode generate and eval-mode forward
    cfg.update(device='cpu', data_name='COIL100', model_name='mcvae', control={'controller_rate': '0.5'})
    process_control()
    cfg['classes_size'] = 100
    torch.manual_seed(0)
    m = models.mcvae()
    g = torch.Generator().manual_seed(12)
    img = torch.rand(16, 3, 32, 32, generator=g) * 2 - 1
    label = torch.randint(0, 100, (16,), generator=g)
    eps = torch.randn(16, 128, generator=g)
    z = torch.randn(16, 128, generator=g)
    _randn_like = torch.randn_like
    torch.randn_like = lambda t: eps.clone()
    m.train(True)
    with torch.no_grad():
        m({'img': img.clone(), 'label': label})
    m.train(False)
    with torch.no_grad():
        gen = m.generate(label, z)
        ev = m({'img': img.clone(), 'label': label})
    torch.randn_like = _randn_like
    arrays.update({'vae/img': img.numpy(), 'vae/label': label.numpy(), 'vae/eps': eps.numpy(), 'vae/z': z.numpy(),
                   'vae/generated': gen.numpy(), 'vae/eval_recon': ev['img'].numpy()})
    meta['vae_eval_loss'] = float(ev['loss'])

    # ---- MCGlow (K=2, hidden 128 fixture model): init pass, eval-mode generate and reconstruct
    cfg.update(device='cpu', data_name='CIFAR10', model_name='mcglow', control={'controller_rate': '0.5'})
    process_control()
    cfg['classes_size'] = 10
    cfg['glow']['K'] = 2
    cfg['glow']['hidden_size'] = 128
    torch.manual_seed(0)
    np.random.seed(0)
    m = models.mcglow()
    g = torch.Generator().manual_seed(13)
    img = torch.rand(16, 3, 32, 32, generator=g) * 2 - 1
    label = torch.randint(0, 10, (16,), generator=g)
    noise = torch.rand(16, 3, 32, 32, generator=g)
    _rand_like = torch.rand_like
    torch.rand_like = lambda t: noise.clone()
    m.train(True)
    with torch.no_grad():
        m({'img': img, 'label': label})                       # data-dependent init
    # make the zero-initialised layers non-trivial (a trained model has them non-zero): deterministic fill
    gw = torch.Generator().manual_seed(14)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith('net.8.module.conv.weight') or n.endswith('prior.conv.weight'):
                p.copy_(torch.randn(p.shape, generator=gw) * 0.01)
            elif n.endswith('net.8.module.scale') or n.endswith('prior.scale') or n.endswith('conv.bias') and 'prior' in n:
                p.copy_(torch.randn(p.shape, generator=gw) * 0.05)
    meta['glow_fill'] = "zero-conv weights ~ 0.01 N(0,1), their scales / prior biases ~ 0.05 N(0,1), Generator seed 14, named_parameters order"
    m.train(False)
    zsh = m.make_z_shapes()
    zlist = [torch.randn(16, *s, generator=g) * 0.7 for s in zsh]
    with torch.no_grad():
        gen = m.generate(label, [t.clone() for t in zlist])
        fz = m({'img': img, 'label': label})['z']
        rec = m.reverse({'z': fz, 'label': label, 'reconstruct': True})['img']
    torch.rand_like = _rand_like
    arrays.update({'glow/img': img.numpy(), 'glow/label': label.numpy(), 'glow/noise': noise.numpy(),
                   'glow/generated': gen.numpy(), 'glow/reconstructed': rec.numpy()})
    for i, t in enumerate(zlist):
        arrays[f'glow/z{i}'] = t.numpy()
    np.savez_compressed(os.path.join(OUT, 'eval_generate.npz'), **arrays)

    # ---- MC-gated PixelCNN (3 layers, hidden 64, 512 embeddings; the full model has 15 layers x hidden 128)
    cfg.update(device='cpu', data_name='CIFAR10', model_name='mcpixelcnn', control={'controller_rate': '0.5'})
    process_control()
    cfg['classes_size'] = 10
    cfg['pixelcnn']['num_layer'] = 3
    cfg['pixelcnn']['hidden_size'] = 64
    torch.manual_seed(0)
    m = models.mcpixelcnn()
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(15)
    code = torch.randint(0, 512, (16, 8, 8), generator=g)
    label = torch.randint(0, 10, (16,), generator=g)
    m.train(True)
    out = m({'img': code, 'label': label})
    out['loss'].backward()
    pa = {'img': code.numpy(), 'label': label.numpy(), 'logits': out['logits'].detach().numpy()[:4]}       # first 4 samples keep the fixture small
    meta['pixelcnn_small'] = {
        'num_layer': 3, 'hidden_size': 64, 'num_embedding': 512, 'classes': 10, 'seed': 0, 'input_seed': 15, 'batch': 16,
        'state_dict_sha256': sd_hash(sd0), 'keys': [[k, list(v.shape)] for k, v in sd0.items()],
        'loss_fwd': float(out['loss'].detach()),
        'grad_norms': {k: float(p.grad.norm()) for k, p in m.named_parameters() if p.grad is not None}}
    m.train(False)
    with torch.no_grad():
        ev = m({'img': code, 'label': label})
    pa['logits_eval'] = ev['logits'].numpy()[:4]
    meta['pixelcnn_small']['loss_eval'] = float(ev['loss'])
    np.savez_compressed(os.path.join(OUT, 'pixelcnn_small.npz'), **pa)

    with open(os.path.join(OUT, 'r2_golden.json'), 'w') as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print('wrote', OUT, {k: v for k, v in meta.items() if not isinstance(v, dict)})


if __name__ == '__main__':
    main()
