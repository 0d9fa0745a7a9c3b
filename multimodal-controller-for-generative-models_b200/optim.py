"""Optimiser side of the training steps on the B200 kernels.

``FusedAdam`` is ``torch.optim.Adam`` (src/train_gan.py:223-236: eps 1e-8, no amsgrad) and, with
``clip_norm``, ``clip_grad_norm_`` + Adam (src/train_vae.py:110-111, src/train_glow.py:120-121) as ONE
pass of ``mcgm_adam_step`` over a flat fp32 buffer per network: parameters, gradients and both moments
are views of four flat buffers, the step counter and the clip coefficient live on the device, so a
step is two or four kernel launches, never synchronises with the host and can be captured in a CUDA
graph.  torch's multi-tensor path needs ~10 launches per step (plus a host round trip for the clip).

The parameters stay the reference's ``nn.Parameter`` objects with the reference's names and shapes
(``state_dict`` is unchanged, ``load_state_dict`` copies into the views); only their storage moves into
the flat buffer.  Build the optimiser AFTER the model is on its device: ``model.to(...)`` afterwards
would re-allocate the parameters and detach them from the buffers.
"""
import torch

from . import _lib

_ALIGN = 64                      # floats: every parameter segment starts on a 256-byte boundary
_FLAT_OF = {}                    # id(param) -> FlatBuffers


class FlatBuffers:
    """Parameters and gradients of one network as views of two flat fp32 buffers."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError('Not valid parameters: nothing to optimise')
        dev = self.params[0].device
        self.offsets = []
        off = 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError('Not valid parameters: flat buffers need fp32 parameters on one device')
            self.offsets.append(off)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = off
        self.param = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad_views = []
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                v = self.param[o:o + p.numel()].view_as(p)
                v.copy_(p.data)
                p.data = v
                g = self.grad[o:o + p.numel()].view_as(p)
                if p.grad is not None:
                    g.copy_(p.grad)
                p.grad = g
                self.grad_views.append(g)
        for p in self.params:
            _FLAT_OF[id(p)] = self

    def zero_grad(self):
        """One memset; ``p.grad`` stay views of the flat buffer (autograd accumulates in place)."""
        self.grad.zero_()
        for p, g in zip(self.params, self.grad_views):
            if p.grad is not g:
                p.grad = g


def flat_buffers(params):
    """The FlatBuffers that already holds exactly these parameters, else a new one."""
    params = [p for p in params if p.requires_grad]
    fb = _FLAT_OF.get(id(params[0])) if params else None
    if fb is not None and len(fb.params) == len(params) and all(a is b for a, b in zip(fb.params, params)) \
            and all(p.data.data_ptr() == fb.param.data_ptr() + 4 * o for p, o in zip(fb.params, fb.offsets)):
        return fb
    return FlatBuffers(params)


class StateGuard:
    """Snapshot of everything a training step mutates — parameters, buffers (BatchNorm running statistics,
    spectral-norm u / v, ActNorm flags) and the optimisers' moments / step counters — restored IN PLACE, so
    pointers captured by a CUDA graph stay valid.  The graph-capturing step classes run their warm-up
    iterations between ``StateGuard(...)`` and ``restore()``: the caller's model is not trained by them."""

    def __init__(self, model, optimizers):
        self.model = model
        self.opts = [o for o in optimizers if hasattr(o, 'snapshot')]
        with torch.no_grad():
            self.tensors = [(t, t.detach().clone()) for t in model.state_dict().values()]
            self.opt_snaps = [o.snapshot() for o in self.opts]

    def restore(self):
        with torch.no_grad():
            for t, saved in self.tensors:
                t.copy_(saved)
            for o, snap in zip(self.opts, self.opt_snaps):
                o.restore(snap)
            params = [p for p in self.model.parameters()]
            if params:
                torch.autograd.graph.increment_version(params)       # packed-weight caches key on the version


class FusedAdam(torch.optim.Optimizer):
    """Adam on a flat buffer through ``mcgm_adam_step`` (CUDA only; there is no CPU path).

    clip_norm: fold ``clip_grad_norm_(params, clip_norm)`` into the step (device-side coefficient).
    grad_scale: constant factor on the gradient (e.g. 1/world_size for summed data-parallel gradients).
    Parameters that never receive a gradient see g = 0, which leaves them (and their zero moments)
    unchanged — the same outcome as torch skipping ``grad is None`` parameters.
    """

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, clip_norm=None,
                 grad_scale=1.0):
        params = list(params)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise ValueError('Not valid param groups: FusedAdam takes one group')
        ps = self.param_groups[0]['params']
        if not ps or not ps[0].is_cuda:
            raise ValueError('Not valid device: FusedAdam needs CUDA parameters (no CPU fallback)')
        self.flat = flat_buffers(ps)
        dev = self.flat.param.device
        self.exp_avg = torch.zeros_like(self.flat.param)
        self.exp_avg_sq = torch.zeros_like(self.flat.param)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.clip_norm, self.grad_scale = clip_norm, float(grad_scale)
        self.clip_out = torch.zeros(2, dtype=torch.float32, device=dev)       # [coef, total_norm]
        self.clip_ws = torch.zeros(_lib.query("mcgm_grad_clip_workspace_bytes") // 4, dtype=torch.float32, device=dev)
        # [lr, beta1, beta2, eps, weight_decay, grad_scale]: read by the update kernel from DEVICE memory, staged
        # through a pinned host buffer by a copy that is part of every step — also of a captured one, so a CUDA
        # graph replay follows ``param_groups[0]['lr']`` (ReduceLROnPlateau, src/train_vae.py:35,78-81)
        self.hyper_host = torch.zeros(6, dtype=torch.float32).pin_memory()
        self.hyper_dev = torch.zeros(6, dtype=torch.float32, device=dev)
        self._hyper_vals = None
        self._bind_state(0.0)

    # ---- torch.optim.Adam's state layout over the flat buffers ---------------------------------------------
    def _views(self):
        f = self.flat
        for p, o in zip(f.params, f.offsets):
            yield p, self.exp_avg[o:o + p.numel()].view_as(p), self.exp_avg_sq[o:o + p.numel()].view_as(p)

    def _bind_state(self, step):
        """``self.state[p] = {'step', 'exp_avg', 'exp_avg_sq'}`` exactly as torch.optim.Adam keeps it; the moments
        are VIEWS of the flat buffers, 'step' is a host copy of the device counter (refreshed by state_dict)."""
        for p, m, v in self._views():
            self.state[p] = {'step': torch.tensor(float(step)), 'exp_avg': m, 'exp_avg_sq': v}

    def state_dict(self):
        """Same structure as ``torch.optim.Adam.state_dict()`` (the reference checkpoints it as
        'optimizer_dict' and resumes from it, src/train_vae.py:86 / resume()): loads into torch's Adam and back."""
        self._bind_state(int(self.step_dev.item()))
        sd = super().state_dict()
        # a snapshot, not live views: torch's load_state_dict keeps the tensors it is handed when dtype / device already
        # match, so handing out views would alias this optimiser's moments into whoever loads the dict
        sd['state'] = {k: {kk: (vv.detach().clone() if torch.is_tensor(vv) else vv) for kk, vv in st.items()}
                       for k, st in sd['state'].items()}
        return sd

    def load_state_dict(self, state_dict):
        """Accepts a ``torch.optim.Adam`` (or FusedAdam) state_dict: moments are copied into the flat buffers,
        the per-parameter step counts (all equal for one-group Adam) into the device counter."""
        super().load_state_dict(state_dict)
        steps = set()
        with torch.no_grad():
            for p, m, v in self._views():
                st = self.state.get(p, None)
                if st and 'exp_avg' in st:
                    m.copy_(st['exp_avg'])
                    v.copy_(st['exp_avg_sq'])
                    steps.add(int(float(st['step'])))
                else:                                  # torch keeps no state for parameters that never saw a gradient
                    m.zero_()
                    v.zero_()
        if len(steps) > 1:
            raise ValueError('Not valid optimizer state: parameters carry different step counts ' + str(sorted(steps)))
        step = steps.pop() if steps else 0
        self.step_dev.fill_(step)
        self._hyper_vals = None
        self._bind_state(step)

    def zero_grad(self, set_to_none=False):
        self.flat.zero_grad()

    def reset_state(self):
        """Back to step 0 with zero moments (what a freshly built optimiser holds)."""
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self.step_dev.zero_()

    def snapshot(self):
        """Copy of the optimiser's device state (moments + step counter) — see ``restore``."""
        return self.exp_avg.clone(), self.exp_avg_sq.clone(), self.step_dev.clone()

    def restore(self, snap):
        self.exp_avg.copy_(snap[0])
        self.exp_avg_sq.copy_(snap[1])
        self.step_dev.copy_(snap[2])

    @property
    def total_norm(self):
        """Gradient norm seen by the last clipped step (device tensor)."""
        return self.clip_out[1]

    def sync_hyper(self):
        """Stage ``param_groups[0]``'s hyper-parameters for the next step / graph replay (host write only)."""
        g = self.param_groups[0]
        vals = (float(g['lr']), float(g['betas'][0]), float(g['betas'][1]), float(g['eps']), float(g['weight_decay']),
                float(self.grad_scale))
        if vals != self._hyper_vals:
            if self._hyper_vals is not None and not torch.cuda.is_current_stream_capturing():
                torch.cuda.current_stream().synchronize()     # an earlier step's staging copy may still be in flight
            self.hyper_host.copy_(torch.tensor(vals, dtype=torch.float32))
            self._hyper_vals = vals

    def _sync_grads(self):
        """``p.grad`` must be the views of the flat gradient buffer.  ``model.zero_grad()`` (set_to_none) makes
        autograd allocate fresh ``.grad`` tensors: copy those in (a missing gradient counts as zero) and rebind."""
        f = self.flat
        for p, g in zip(f.params, f.grad_views):
            if p.grad is g:
                continue
            if p.grad is None:
                g.zero_()
            else:
                g.copy_(p.grad)
            p.grad = g

    @torch.no_grad()
    def step(self, closure=None):
        st = torch.cuda.current_stream().cuda_stream
        f = self.flat
        self._sync_grads()
        self.sync_hyper()
        self.hyper_dev.copy_(self.hyper_host, non_blocking=True)
        coef = None
        if self.clip_norm is not None:
            _lib.call("mcgm_grad_clip_coef", f.grad.data_ptr(), f.numel, float(self.clip_norm),
                      self.clip_out.data_ptr(), self.clip_ws.data_ptr(), st)
            _lib.launch_count += 1
            coef = self.clip_out.data_ptr()
        _lib.call("mcgm_adam_step_dev", f.param.data_ptr(), f.grad.data_ptr(), self.exp_avg.data_ptr(),
                  self.exp_avg_sq.data_ptr(), f.numel, self.step_dev.data_ptr(), self.hyper_dev.data_ptr(), coef, st)
        _lib.launch_count += 1
        # the update wrote through raw pointers: tell autograd (and the packed-weight caches keyed on
        # ``Parameter._version``) that the parameters changed
        torch.autograd.graph.increment_version(f.params)
        return None
