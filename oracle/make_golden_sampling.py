"""Generate tests/golden/sampling.npz FROM THE UNMODIFIED REFERENCE (test infrastructure).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_sampling.py

Fixtures for the codebook side of the reference's sampling utilities (src/models/utils.py:34-44
create_codebook, :100-110 transit_codebook, and create()/transit() :47-97, :113-152 applied to an MCGAN):
the callers of the MultimodalController boundary that swap its ``codebook`` buffer.
Nothing here is imported by the product; the GPU box never needs /root/reference."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, 'tests', 'golden', 'sampling.npz')
REF_SRC = '/root/reference/src'


def main():
    os.chdir(REF_SRC)
    sys.path.insert(0, REF_SRC)
    from config import cfg
    from utils import process_control
    import models
    from models import utils as U
    out = {}
    cfg.update(device='cpu', data_name='CIFAR10', model_name='mcgan', control={'controller_rate': '0.5'})
    process_control()
    # create_codebook: (seed, classes_size, K)
    for i, (seed, n, K) in enumerate([(0, 10, 64), (3, 50, 32), (7, 100, 128), (11, 4, 2)]):
        cfg['classes_size'] = n
        torch.manual_seed(seed)
        out[f'create/{i}/args'] = np.array([seed, n, K])
        out[f'create/{i}/codebook'] = U.create_codebook(torch.zeros(n, K)).numpy()
    # transit_codebook: fixed input codebook, several (root, alpha)
    g = torch.Generator().manual_seed(5)
    cb = (torch.rand(10, 48, generator=g) > 0.5).float()
    out['transit/input'] = cb.numpy()
    for i, (root, alpha) in enumerate([(0, 0.0), (3, 0.25), (9, 0.5), (4, 0.99), (2, 1.0), (5, 0.3333)]):
        out[f'transit/{i}/args'] = np.array([root, alpha])
        out[f'transit/{i}/codebook'] = U.transit_codebook(cb, root, alpha).numpy()
    # create() / transit() on a small MCGAN: every controller's codebook afterwards
    cfg['classes_size'] = 10
    cfg['gan']['generator_hidden_size'] = [64, 64, 64, 64]
    cfg['gan']['discriminator_hidden_size'] = [64, 64, 64, 64]
    torch.manual_seed(0)
    model = models.mcgan()
    torch.manual_seed(21)
    cfg['classes_size'] = 6                       # create(): a different number of modes than trained
    U.create(model)
    names = [k for k, _ in model.named_buffers() if k.endswith('codebook')]
    out['model/create/names'] = np.array(names)
    for k, v in model.named_buffers():
        if k.endswith('codebook'):
            out['model/create/' + k] = v.numpy()
    cfg['classes_size'] = 10
    torch.manual_seed(0)
    model = models.mcgan()
    U.transit(model, 2, 0.4)
    U.transit(model, 7, 0.8)                      # second call restarts from codebook_orig
    for k, v in model.named_buffers():
        if k.endswith('codebook') or k.endswith('codebook_orig'):
            out['model/transit/' + k] = v.numpy()
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, len(out), 'arrays')


if __name__ == '__main__':
    main()
