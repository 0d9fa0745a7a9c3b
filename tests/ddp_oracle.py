"""Test helper: the data-parallel training iteration restated over the ORACLE functions.

SURVEY.md §8(e): "parity for multi-GPU is checked against the oracle run with the same per-rank shards and
grads averaged".  Every rank holds a full copy of the state (as nn.DataParallel replicas / our processes do):
its own BatchNorm running statistics and its own spectral-norm u / v (identical across ranks because they only
depend on the shared weights), runs forward/backward on ITS shard, then the parameter gradients are averaged
and every rank applies the same Adam update (replaces src/train_gan.py:96-98)."""
import torch
import torch.nn.functional as F


def ddp_train_iteration(O, sds, opts, imgs, labels, zss, num_mode, strides, n_d=5, n_g=1):
    """sds / opts / imgs / labels / zss: one entry per rank (zss[r] = the n_d + n_g latent draws of rank r).
    Returns per-rank (loss_D, loss_G) of the last updates."""
    world = len(sds)
    inds = [O.one_hot(labels[r], num_mode) for r in range(world)]
    d_names = O.split_params(sds[0], 'discriminator.')
    g_names = O.split_params(sds[0], 'generator.')
    d_loss, g_loss = [None] * world, [None] * world

    def average(names):
        for k in names:
            grads = [sds[r][k].grad for r in range(world)]
            if any(g is None for g in grads):
                continue
            mean = sum(grads) / world
            for r in range(world):
                sds[r][k].grad = mean.clone()

    for i in range(n_d):
        for r in range(world):
            opts[r]['discriminator'].zero_grad()
            opts[r]['generator'].zero_grad()
            d_x = O.discriminator(imgs[r], inds[r], sds[r], strides)
            fake = O.generator(zss[r][i], inds[r], sds[r])
            d_g = O.discriminator(fake.detach(), inds[r], sds[r], strides)
            d_loss[r] = F.relu(1.0 - d_x).mean() + F.relu(1.0 + d_g).mean()
            d_loss[r].backward()
        average(d_names)
        for r in range(world):
            opts[r]['discriminator'].step()
    for i in range(n_g):
        for r in range(world):
            opts[r]['discriminator'].zero_grad()
            opts[r]['generator'].zero_grad()
            fake = O.generator(zss[r][n_d + i], inds[r], sds[r])
            g_loss[r] = -O.discriminator(fake, inds[r], sds[r], strides).mean()
            g_loss[r].backward()
        average(g_names)
        for r in range(world):
            opts[r]['generator'].step()
    return [(float(d_loss[r].detach()), float(g_loss[r].detach())) for r in range(world)]


def shard_inputs(case_arrays, world, rank):
    """Per-rank shard of a golden case's inputs: rank r takes the batch rolled by r samples with a rank-specific
    sign pattern on z — distinct data per rank from the committed fixture, no new files."""
    img = torch.roll(case_arrays["img"], shifts=rank, dims=0) * (1.0 if rank % 2 == 0 else -1.0)
    label = torch.roll(case_arrays["label"], shifts=rank, dims=0)
    zs = [torch.roll(z, shifts=rank, dims=0) * (1.0 if rank % 2 == 0 else -1.0) for z in case_arrays["zs"]]
    return img, label, zs
