"""The MCGAN training step (src/train_gan.py:128-194, core :139-176) on the B200 kernels,
plus one-process-per-GPU data parallelism (replaces nn.DataParallel, train_gan.py:96-98).

``train_step`` restates the reference loop body: ``iter['discriminator']`` D updates
(zero_grad both nets; D(real); G(z1); D(G.detach()); hinge; backward; Adam) followed by
``iter['generator']`` G updates (G(z2); D(G); -mean; backward; Adam).  Noise can be injected
(``z=[z1.., z2]``) so that a parity run and the oracle see identical randomness.  Losses stay
on the device; nothing in the step synchronises with the host.
"""
import torch
import torch.distributed as dist
import torch.nn.functional as F

from .config import GAN_HYPER, cfg


def make_optimizer(model, lr, betas, capturable=False, fused=None):
    """Adam with torch defaults eps=1e-8, no amsgrad (src/train_gan.py:223-236).
    On CUDA this is ``optim.FusedAdam`` (one native pass over a flat buffer, device-side step counter:
    always graph-capturable); ``fused=False`` or CPU parameters give ``torch.optim.Adam``, where
    ``capturable`` keeps the step counter on the device."""
    params = list(model.parameters())
    if fused is None:
        fused = bool(params) and params[0].is_cuda
    if fused:
        from .optim import FusedAdam
        return FusedAdam(params, lr=lr, weight_decay=GAN_HYPER['weight_decay'], betas=betas)
    return torch.optim.Adam(params, lr=lr, weight_decay=GAN_HYPER['weight_decay'], betas=betas,
                            capturable=capturable)


def make_optimizers(model, capturable=False, fused=None):
    return {
        'generator': make_optimizer(model.generator, GAN_HYPER['lr']['generator'], GAN_HYPER['betas']['generator'],
                                    capturable, fused),
        'discriminator': make_optimizer(model.discriminator, GAN_HYPER['lr']['discriminator'],
                                        GAN_HYPER['betas']['discriminator'], capturable, fused),
    }


class GradAllReduce:
    """Data-parallel gradient averaging over NCCL (or gloo in CPU tests).

    Each net's gradients live in ONE flat fp32 buffer (``p.grad`` are views), cut into
    buckets in reverse parameter order.  A post-accumulate hook counts the parameters of a
    bucket as autograd finishes them; when a bucket is complete its all-reduce is issued
    asynchronously on the communication stream, overlapping the rest of backward.
    ``finish()`` waits for the outstanding work and divides by the world size.
    BatchNorm statistics stay per-rank, as with the reference's DataParallel replicas.
    """

    def __init__(self, module, bucket_bytes=8 << 20, group=None, overlap=True, average=True):
        """overlap=True: hook-driven bucketed all-reduce during backward (eager mode).
        overlap=False: one all-reduce of the whole flat buffer in ``finish()`` — the form that is
        captured inside a CUDA graph (the payload is 4-17 MB, i.e. tens of microseconds on NVLink).
        average=False: ``finish()`` leaves the SUM in the buffer; the 1/world factor then rides on the
        optimiser update (``FusedAdam.grad_scale``) instead of costing a pass over the gradients."""
        self.group = group
        self.overlap = overlap
        self.average = average
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        from .optim import flat_buffers
        self.params = [p for p in module.parameters() if p.requires_grad]
        fb = flat_buffers(self.params)          # shared with FusedAdam when the optimiser was built first
        self.flat, self.views = fb.grad, fb.grad_views
        total = fb.numel
        # buckets over the flat buffer, filled from the END (backward reaches late layers first)
        self.buckets = []           # (start, stop, n_params)
        self.bucket_of = {}
        stop = total
        count = 0
        for p, start in zip(reversed(self.params), reversed(fb.offsets)):
            count += 1
            self.bucket_of[id(p)] = len(self.buckets)
            if (stop - start) * 4 >= bucket_bytes:
                self.buckets.append([start, stop, count])
                stop, count = start, 0
        if count:
            self.buckets.append([0, stop, count])
        self.pending = [0] * len(self.buckets)
        self.launched = [False] * len(self.buckets)
        self.handles = []
        self.enabled = True
        if self.world > 1 and overlap:
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._hook)

    def zero_grad(self):
        """Replacement for optimizer.zero_grad(): one memset, grads stay views of the flat buffer."""
        self.flat.zero_()
        for p, g in zip(self.params, self.views):
            p.grad = g
        self.pending = [0] * len(self.buckets)
        self.launched = [False] * len(self.buckets)
        self.handles = []

    def _hook(self, p):
        if not self.enabled:
            return
        from . import ops
        b = self.bucket_of[id(p)]
        self.pending[b] += 1
        # A bucket goes out during backward only when every gradient in it has really been written.  With deferred
        # end-of-backward work (ops.SN_DEFER: the D update's spectral-norm backwards and grouped weight gradients are
        # added to .grad by ONE callback after the pass) the hook count says nothing — autograd fires the hook for those
        # parameters too (with an undefined gradient) — so such a pass is reduced in finish().  Found by the 2-rank
        # parity test: the eager overlapped form all-reduced D's bucket before the deferred gradients existed.
        if self.pending[b] == self.buckets[b][2] and not ops.SN_DEFER:
            s, e, _ = self.buckets[b]
            self.launched[b] = True
            self.handles.append(dist.all_reduce(self.flat[s:e], group=self.group, async_op=True))

    def finish(self):
        if self.world == 1:
            return
        if not self.overlap:
            dist.all_reduce(self.flat, group=self.group)
            if self.average:
                self.flat.div_(self.world)
            return
        # buckets that did not go out during backward (unused parameters, deferred gradients)
        for b, (s, e, n) in enumerate(self.buckets):
            if not self.launched[b]:
                self.launched[b] = True
                self.handles.append(dist.all_reduce(self.flat[s:e], group=self.group, async_op=True))
        for h in self.handles:
            h.wait()
        self.handles = []
        if self.average:
            self.flat.div_(self.world)


def hinge_d(d_real, d_fake):
    """relu(1 - D(x)).mean() + relu(1 + D(G(z))).mean()   (src/train_gan.py:154)."""
    return F.relu(1.0 - d_real).mean() + F.relu(1.0 + d_fake).mean()


def d_forward_backward(model, img, label, z, generated=None):
    """Forward + backward of one D update (src/train_gan.py:143-157); returns the hinge loss.
    G(z) is detached by the reference (:148), so the generator runs without an autograd tape.
    generated: G(z) computed by the caller (``train_step_pipelined``); z is then unused."""
    from . import ops
    d_x = model.discriminate(img, label)
    if generated is None:
        with torch.no_grad():
            generated = model.generate(label, z)
    d_g = model.discriminate(generated, label)
    d_loss = hinge_d(d_x, d_g)
    ops.reset_deferred()
    ops.SN_DEFER = True            # spectral-norm backwards and small-layer weight gradients: batched at the end of the pass
    try:
        d_loss.backward()
    finally:
        ops.SN_DEFER = False
    return d_loss


def g_forward_backward(model, label, z):
    """Forward + backward of one G update (src/train_gan.py:165-175); returns -D(G(z)).mean().

    The reference lets autograd fill D's parameter gradients here and throws them away at the next
    ``zero_grad`` (:141 / :163) before anyone reads them.  D's parameters are therefore frozen for this
    backward: the discriminator only propagates the gradient to its input (dgrad), its eleven weight
    gradients and spectral-norm backward passes are not computed.  Every value the reference ever uses
    is unchanged."""
    d_params = [p for p in model.discriminator.parameters() if p.requires_grad]
    for p in d_params:
        p.requires_grad_(False)
    try:
        generated = model.generate(label, z)
        g_loss = -model.discriminate(generated, label).mean()
        g_loss.backward()
    finally:
        for p in d_params:
            p.requires_grad_(True)
    return g_loss


def train_step(model, optimizer, input, z=None, reducers=None):
    """One reference iteration on a batch ``{'img': [N,C,H,W] fp32, 'label': [N] int64}``.

    z: optional list of ``iter_D + iter_G`` latent tensors [N, latent] (injected noise).
    reducers: optional ``{'generator': GradAllReduce, 'discriminator': GradAllReduce}``.
    Returns ``{'loss_D', 'loss_G', 'loss'}`` as device tensors (no host sync).
    """
    img, label = input['img'], input[cfg.get('subset', 'label')]
    n = img.size(0)
    n_d, n_g = GAN_HYPER['iter']['discriminator'], GAN_HYPER['iter']['generator']
    zi = iter(z) if z is not None else None

    def draw():
        if zi is not None:
            return next(zi)
        return torch.randn(n, model.latent_size, device=img.device)

    def zero():
        if reducers is not None:
            reducers['discriminator'].zero_grad()
            reducers['generator'].zero_grad()
        else:
            optimizer['discriminator'].zero_grad()
            optimizer['generator'].zero_grad()

    d_loss = g_loss = None
    for _ in range(n_d):                                       # (1) update D   train_gan.py:139-158
        zero()
        if reducers is not None:
            reducers['generator'].enabled = False              # G receives no gradient in the D step
        d_loss = d_forward_backward(model, img, label, draw())
        if reducers is not None:
            reducers['discriminator'].finish()
            reducers['generator'].enabled = True
        optimizer['discriminator'].step()
    for _ in range(n_g):                                       # (2) update G   train_gan.py:162-176
        zero()
        if reducers is not None:
            reducers['discriminator'].enabled = False          # D's grads are discarded by the next zero_grad
        g_loss = g_forward_backward(model, label, draw())
        if reducers is not None:
            reducers['generator'].finish()
            reducers['discriminator'].enabled = True
        optimizer['generator'].step()
    total = (d_loss - g_loss).abs() if d_loss is not None and g_loss is not None else None
    return {'loss': total, 'loss_D': d_loss, 'loss_G': g_loss}


def train_step_pipelined(model, optimizer, input, z=None, reducers=None, side=None):
    """``train_step`` with the SAME operations in a schedule that hides each D update's gradient all-reduce and Adam
    update: G(z) of the NEXT update (and the generator forward of the G update) does not depend on the discriminator, so
    it is issued on the current stream while ``side`` runs all-reduce + Adam of the update just finished; the streams
    join before the discriminator is used again.  Every tensor gets the value ``train_step`` gives it (the generator
    forward moves in front of D(real) inside an update — the two are independent); inside a CUDA graph the fork / join
    become graph edges.  What stays exposed is the G update's own all-reduce at the end of the iteration."""
    img, label = input['img'], input[cfg.get('subset', 'label')]
    n = img.size(0)
    n_d, n_g = GAN_HYPER['iter']['discriminator'], GAN_HYPER['iter']['generator']
    if side is None or n_d < 1 or n_g != 1:
        return train_step(model, optimizer, input, z=z, reducers=reducers)
    zi = iter(z) if z is not None else None
    main = torch.cuda.current_stream()

    def draw():
        if zi is not None:
            return next(zi)
        return torch.randn(n, model.latent_size, device=img.device)

    def zero():
        if reducers is not None:
            reducers['discriminator'].zero_grad()
            reducers['generator'].zero_grad()
        else:
            optimizer['discriminator'].zero_grad()
            optimizer['generator'].zero_grad()

    with torch.no_grad():
        generated = model.generate(label, draw())
    d_loss = g_loss = None
    for k in range(n_d):
        zero()                                                 # after the join: D's Adam has consumed the old gradients
        if reducers is not None:
            reducers['generator'].enabled = False
        d_loss = d_forward_backward(model, img, label, None, generated=generated)
        if reducers is not None:
            reducers['generator'].enabled = True
        side.wait_stream(main)
        with torch.cuda.stream(side):                          # all-reduce + Adam of this update ...
            if reducers is not None:
                reducers['discriminator'].finish()
            optimizer['discriminator'].step()
        if k + 1 < n_d:                                        # ... under the next update's G(z)
            with torch.no_grad():
                generated = model.generate(label, draw())
        else:                                                  # ... or under the G update's generator forward (taped)
            generated = model.generate(label, draw())
        main.wait_stream(side)
    zero()
    if reducers is not None:
        reducers['discriminator'].enabled = False
    d_params = [p for p in model.discriminator.parameters() if p.requires_grad]
    for p in d_params:
        p.requires_grad_(False)
    try:
        g_loss = -model.discriminate(generated, label).mean()
        g_loss.backward()
    finally:
        for p in d_params:
            p.requires_grad_(True)
    if reducers is not None:
        reducers['generator'].finish()
        reducers['discriminator'].enabled = True
    optimizer['generator'].step()
    total = (d_loss - g_loss).abs()
    return {'loss': total, 'loss_D': d_loss, 'loss_G': g_loss}


def _bump_versions(optimizers):
    """A replayed graph updates the parameters without Python noticing: bump their version counters so that the
    packed-weight caches (keyed on ``Parameter._version``) re-pack on the next EAGER use (evaluation, sampling)."""
    for o in optimizers:
        if hasattr(o, 'flat'):
            torch.autograd.graph.increment_version(o.flat.params)


class GraphedTrainStep:
    """The whole training iteration (5 D + 1 G updates: forwards, backwards, Adam, noise draws)
    captured ONCE into a CUDA graph and replayed per batch.

    One iteration is ~6000 kernel launches of 10-400 us each; issued eagerly the host can barely
    stay ahead of the B200.  The step has static shapes and never synchronises with the host, so
    it is captured after a few eager warm-up iterations (which also build every lazily created
    workspace, packed codebook and TMA descriptor) and replayed with new data copied into static
    input buffers.  Optimisers must be built with ``capturable=True``.  The warm-up iterations run on an
    all-zero batch; parameters, buffers and optimiser state are snapshotted before and restored after them
    (``optim.StateGuard``), so the caller's model starts real training exactly where it was.
    """

    def __init__(self, model, optimizer, batch_size, warmup=3, reducers=None, inject_noise=False):
        dev = next(model.parameters()).device
        shape = model.generator.data_shape
        self.img = torch.zeros(batch_size, *shape, device=dev)
        self.label = torch.zeros(batch_size, dtype=torch.int64, device=dev)
        self.model, self.optimizer = model, optimizer
        batch = {'img': self.img, 'label': self.label}
        # inject_noise: latent draws come from static buffers the caller fills (parity tests)
        n_z = GAN_HYPER['iter']['discriminator'] + GAN_HYPER['iter']['generator']
        self.z = ([torch.zeros(batch_size, model.latent_size, device=dev) for _ in range(n_z)]
                  if inject_noise else None)
        from .optim import StateGuard
        guard = StateGuard(model, optimizer.values())
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                train_step(model, optimizer, batch, z=self.z, reducers=reducers)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import _lib
        c0 = _lib.launch_count
        from . import ops as _ops
        _ops.clear_pack_caches()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = train_step(model, optimizer, batch, z=self.z, reducers=reducers)
        self.native_launches = _lib.launch_count - c0      # libmcgm kernels inside one replay
        guard.restore()

    def __call__(self, input):
        """input: {'img','label'} on the host (pinned) or the device; returns device loss tensors
        that are overwritten by the next call."""
        self.img.copy_(input['img'], non_blocking=True)
        self.label.copy_(input['label'], non_blocking=True)
        if self.z is not None:
            for dst, src in zip(self.z, input['z']):
                dst.copy_(src, non_blocking=True)
        for o in self.optimizer.values():
            if hasattr(o, 'sync_hyper'):
                o.sync_hyper()              # the captured Adam reads lr & co from a staged device buffer
        self.graph.replay()
        _bump_versions(self.optimizer.values())
        return self.out


class GraphedDDPTrainStep:
    """Data-parallel variant with the collectives BETWEEN graph segments: the forward+backward of a D update
    and of a G update are captured as two CUDA graphs; after each replay the flat gradient buffer is
    all-reduced over NCCL/NVLink (sum; the 1/world factor rides on the Adam kernel) and Adam steps eagerly.
    Per iteration the host issues 6 graph launches + 6 all-reduces + 6 optimiser steps instead of ~6000 kernel
    launches.  Weight packing of G is forced inside the graphs (its cache check happens at capture time).
    ``GraphedDDPIngraphTrainStep`` (collectives captured too) is the default; this form is the fallback."""

    launch_form = "CUDA graph per D/G forward+backward, NCCL all-reduce + Adam between graphs (1/world folded into Adam)"

    def __init__(self, model, optimizer, reducers, batch_size, warmup=2, inject_noise=False):
        from . import ops
        from .optim import StateGuard
        dev = next(model.parameters()).device
        shape = model.generator.data_shape
        self.img = torch.zeros(batch_size, *shape, device=dev)
        self.label = torch.zeros(batch_size, dtype=torch.int64, device=dev)
        self.model, self.optimizer, self.reducers = model, optimizer, reducers
        self.n_d, self.n_g = GAN_HYPER['iter']['discriminator'], GAN_HYPER['iter']['generator']
        latent = model.latent_size
        # one latent buffer per segment graph; with inject_noise the caller's z_k is copied in before replay k
        self.z_d = torch.zeros(batch_size, latent, device=dev)
        self.z_g = torch.zeros(batch_size, latent, device=dev)
        self.inject = inject_noise

        def d_part():
            reducers['discriminator'].zero_grad()
            reducers['generator'].zero_grad()
            z = self.z_d if inject_noise else torch.randn(batch_size, latent, device=dev)
            return d_forward_backward(model, self.img, self.label, z)

        def g_part():
            reducers['discriminator'].zero_grad()
            reducers['generator'].zero_grad()
            z = self.z_g if inject_noise else torch.randn(batch_size, latent, device=dev)
            return g_forward_backward(model, self.label, z)

        for r in reducers.values():
            if r.overlap:
                raise ValueError('Not valid reducer: graph mode needs GradAllReduce(overlap=False)')
        _fold_world_into_adam(optimizer, reducers)
        guard = StateGuard(model, optimizer.values())
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                d_part()
                g_part()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import _lib
        # (round 1 set this flag on ``models.mcgan`` — which is the FACTORY FUNCTION re-exported by models/__init__, not
        # the module: the D segment then replayed with the generator weights packed at capture time.  Found by
        # tests/test_gpu_ddp.py; the flag now lives in ops.)
        ops.PACKS_FORCE = True
        try:
            c0 = _lib.launch_count
            self.graph_d = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_d):
                self.d_loss = d_part()
            c1 = _lib.launch_count
            self.graph_g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_g, pool=self.graph_d.pool()):
                self.g_loss = g_part()
            c2 = _lib.launch_count
        finally:
            ops.PACKS_FORCE = False
        self.native_launches = (c1 - c0 + 2) * self.n_d + (c2 - c1 + 2) * self.n_g      # + the eager Adam launches
        self.out = {'loss_D': self.d_loss, 'loss_G': self.g_loss, 'loss': None}
        guard.restore()

    def close(self):
        """Release the captured graphs (call before ``dist.destroy_process_group()``)."""
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        self.graph_d = self.graph_g = None

    def __call__(self, input):
        self.img.copy_(input['img'], non_blocking=True)
        self.label.copy_(input['label'], non_blocking=True)
        zs = iter(input['z']) if self.inject else None
        for _ in range(self.n_d):
            if zs is not None:
                self.z_d.copy_(next(zs), non_blocking=True)
            self.graph_d.replay()
            self.reducers['discriminator'].finish()
            self.optimizer['discriminator'].step()
        for _ in range(self.n_g):
            if zs is not None:
                self.z_g.copy_(next(zs), non_blocking=True)
            self.graph_g.replay()
            self.reducers['generator'].finish()
            self.optimizer['generator'].step()
        self.out['loss'] = (self.d_loss - self.g_loss).abs()
        return self.out


def _fold_world_into_adam(optimizer, reducers):
    """Sum-reduce + (1/world inside the Adam kernel) instead of sum-reduce + a division pass."""
    for k, r in reducers.items():
        o = optimizer[k]
        if not hasattr(o, 'grad_scale'):
            raise ValueError('Not valid optimizer: the data-parallel graph steps need optim.FusedAdam')
        if r.average:
            r.average = False
            o.grad_scale = o.grad_scale / r.world


class GraphedDDPIngraphTrainStep:
    """Data-parallel iteration as ONE CUDA graph per rank: forwards, backwards, the NCCL all-reduce of each
    update's flat gradient buffer and the Adam updates are all captured (NCCL collectives are capturable; the
    ranks replay the same graph in lock step).  The host issues one ``cudaGraphLaunch`` per iteration at any
    world size — no launch gaps around the collectives, no eager optimiser launches, 1/world folded into Adam.
    The D updates' all-reduce (4.2 MB each) + Adam run on a forked stream under the next update's G(z) forward
    (``train_step_pipelined``; MCGM_DDP_PIPELINE=0 restores the serial schedule).  What stays exposed: the G update's
    all-reduce (17.1 MB) + Adam at the end of the iteration."""

    launch_form = ("one CUDA graph per iteration, NCCL all-reduce + Adam captured inside (1/world folded into Adam), "
                   "D's all-reduce + Adam overlapped with the next G(z)")

    def __init__(self, model, optimizer, reducers, batch_size, warmup=2, inject_noise=False):
        from .optim import StateGuard
        for r in reducers.values():
            if r.overlap:
                raise ValueError('Not valid reducer: graph mode needs GradAllReduce(overlap=False)')
        _fold_world_into_adam(optimizer, reducers)
        dev = next(model.parameters()).device
        shape = model.generator.data_shape
        self.img = torch.zeros(batch_size, *shape, device=dev)
        self.label = torch.zeros(batch_size, dtype=torch.int64, device=dev)
        self.model, self.optimizer, self.reducers = model, optimizer, reducers
        n_z = GAN_HYPER['iter']['discriminator'] + GAN_HYPER['iter']['generator']
        self.z = ([torch.zeros(batch_size, model.latent_size, device=dev) for _ in range(n_z)]
                  if inject_noise else None)
        batch = {'img': self.img, 'label': self.label}
        guard = StateGuard(model, optimizer.values())
        import os
        self.pipelined = os.environ.get('MCGM_DDP_PIPELINE', '1') != '0'
        self.comm_stream = torch.cuda.Stream() if self.pipelined else None      # all-reduce + Adam of the D updates
        if not self.pipelined:
            self.launch_form = "one CUDA graph per iteration, NCCL all-reduce + Adam captured inside (1/world folded into Adam)"
        step_fn = (lambda: train_step_pipelined(model, optimizer, batch, z=self.z, reducers=reducers, side=self.comm_stream)) \
            if self.pipelined else (lambda: train_step(model, optimizer, batch, z=self.z, reducers=reducers))
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):                     # also brings up the NCCL communicator before the capture
                step_fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if dist.is_initialized():
            dist.barrier()
        from . import _lib
        c0 = _lib.launch_count
        from . import ops as _ops
        _ops.clear_pack_caches()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode='thread_local'):
            self.out = step_fn()
        self.native_launches = _lib.launch_count - c0
        guard.restore()

    def close(self):
        """Release the captured graph.  It holds NCCL kernels: destroy it BEFORE ``dist.destroy_process_group()`` —
        tearing the communicator down under a live graph that references it blocks (measured: both ranks hang in
        ``destroy_process_group`` until killed)."""
        torch.cuda.synchronize()
        self.graph = None

    def __call__(self, input):
        self.img.copy_(input['img'], non_blocking=True)
        self.label.copy_(input['label'], non_blocking=True)
        if self.z is not None:
            for dst, src in zip(self.z, input['z']):
                dst.copy_(src, non_blocking=True)
        for o in self.optimizer.values():
            o.sync_hyper()
        self.graph.replay()
        _bump_versions(self.optimizer.values())
        return self.out


def make_ddp_step(model, optimizer, reducers, batch_size, mode='auto', inject_noise=False):
    """The data-parallel graph step: 'ingraph' (collectives captured), 'segmented' (collectives between graph
    segments), or 'auto' = ingraph, falling back to segmented if this torch / NCCL build refuses the capture
    (every rank takes the same decision: the outcome is agreed through an all-reduce)."""
    import sys
    if mode == 'segmented':
        return GraphedDDPTrainStep(model, optimizer, reducers, batch_size, inject_noise=inject_noise)
    err = None
    step = None
    try:
        step = GraphedDDPIngraphTrainStep(model, optimizer, reducers, batch_size, inject_noise=inject_noise)
    except Exception as e:                           # noqa: BLE001 - capture failures surface as several types
        if mode == 'ingraph':
            raise
        err = e
    if dist.is_initialized() and mode == 'auto':
        dev = next(model.parameters()).device
        try:
            torch.cuda.synchronize()
        except Exception:                            # noqa: BLE001
            pass
        flag = torch.tensor([0 if step is not None else 1], device=dev)
        dist.all_reduce(flag)
        if int(flag) > 0:
            step = None
    if step is not None:
        return step
    print(f'[train_gan] in-graph NCCL capture unavailable ({type(err).__name__ if err else "peer rank"}: {err}); '
          f'using segmented graphs', file=sys.stderr, flush=True)
    return GraphedDDPTrainStep(model, optimizer, reducers, batch_size, inject_noise=inject_noise)


def shutdown_process_group(*steps, grace_s=20.0):
    """Orderly end of a data-parallel run: release the captured graphs (they reference the communicator), synchronise,
    then destroy the process group.  A watchdog ends the process if the teardown itself blocks (a NCCL teardown that
    waits forever would otherwise keep the launcher — and its GPUs — busy after all results have been written)."""
    import gc
    import os
    import sys
    import threading
    for s in steps:
        if s is not None and hasattr(s, 'close'):
            s.close()
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if not dist.is_initialized():
        return
    sys.stdout.flush()
    sys.stderr.flush()
    t = threading.Timer(grace_s, lambda: os._exit(0))
    t.daemon = True
    t.start()
    try:
        dist.barrier()
        dist.destroy_process_group()
    finally:
        t.cancel()
