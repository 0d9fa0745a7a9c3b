"""Tensor-level wrappers and ``torch.autograd.Function``s over the C ABI.

Internal activation layout is NHWC: tensors of shape ``[N, H, W, C]``, contiguous,
bf16 (or fp32 for the gate kernels).  PyTorch is used for device memory, streams and
autograd bookkeeping only; every kernel launched here comes from libmcgm.so.
Shape / dtype / contiguity problems raise ``ValueError`` (the reference's error
style, e.g. src/utils.py:126); a failing native call raises ``RuntimeError``.
"""
import os

import torch

from . import _lib
from ._lib import (BF16, CONV_PHASE2, CONV_S1, CONV_S2K4, EPI_BIAS, EPI_MASK, EPI_POSMASK, EPI_RELU, EPI_RES, EPI_RES_UP2, EPI_TANH, F32,
                   GATE_AFFINE, GATE_BN, GATE_MASK, GATE_RELU, GATE_UP2)


# --------------------------------------------------------------------------- helpers
def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _dtype_code(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise ValueError(f"Not valid dtype {t.dtype} (float32 or bfloat16)")


def _check_cuda(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("mcgm_b200 kernels run on CUDA tensors only; there is no CPU fallback")
        if not t.is_contiguous():
            raise ValueError("Not valid layout: tensor must be contiguous")


# Optional per-kernel timing (bench.py): PROFILE = {} enables CUDA-event pairs around the
# tagged native calls; each record is (tag, start_event, end_event, work) with work in
# algorithmic bytes (HBM-bound kernels) or flops (tensor-core kernels).
PROFILE = None


def _prof_begin():
    if PROFILE is None:
        return None
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    return ev


def _prof_end(ev, tag, work, desc="", ref_work=None):
    """work: EXECUTED algorithmic work of the launch (flops the tensor cores run / bytes the kernel must move);
    ref_work: the same launch booked at the reference op's count when that differs (Up2->Conv3x3 executes 4 of
    the reference's 9 taps per output) — reported separately as ``frac_reference_op``, never as the roofline."""
    if ev is None:
        return
    end = torch.cuda.Event(enable_timing=True)
    end.record()
    PROFILE.setdefault(tag, []).append((ev, end, work, desc, work if ref_work is None else ref_work))


_workspaces = {}
_retired_workspaces = []      # outgrown buffers stay alive: a captured CUDA graph may still hold their address


def _workspace(tag, nbytes, device):
    """Persistent per-(tag, device) scratch buffer, grown on demand (never shrinks, never freed: a CUDA graph
    captured while a smaller buffer was current keeps replaying into that buffer's address)."""
    key = (tag, device)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _retired_workspaces.append(buf)
        buf = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


class PackedCodebook:
    """Bit-packed device copy of a MultimodalController codebook (mcgm_pack_codebook)."""

    def __init__(self, codebook):
        _check_cuda(codebook)
        if codebook.dim() != 2 or codebook.dtype != torch.float32:
            raise ValueError("Not valid codebook (float32 [num_mode, input_size])")
        self.M, self.C = codebook.shape
        self.words = (self.C + 31) // 32
        self.bits = torch.empty(self.M, self.words, dtype=torch.int32, device=codebook.device)
        _lib.call("mcgm_pack_codebook", codebook.data_ptr(), self.M, self.C, self.bits.data_ptr(), _stream())


def as_labels(labels):
    """int32 device labels from int64/int32 labels."""
    if labels.dtype == torch.int32:
        return labels
    return labels.to(torch.int32)


# --------------------------------------------------------------------------- raw calls
def mc_mask(x, labels=None, packed=None, code=None, nchw=False):
    """y = x * code(n, c) — MultimodalController.forward (src/modules/modules.py:71-76)."""
    _check_cuda(x, labels, code)
    N = x.shape[0]
    if nchw:
        C = x.shape[1]
        P = x.numel() // max(N * C, 1)
    else:
        C = x.shape[-1]
        P = x.numel() // max(N * C, 1)
    y = torch.empty_like(x)
    if N == 0:
        return y
    if labels is not None:
        _lib.call("mcgm_mc_forward", x.data_ptr(), y.data_ptr(), _dtype_code(x), int(nchw), N, P, C,
                  labels.data_ptr(), packed.bits.data_ptr(), packed.M, 0, _stream())
    else:
        if code.dtype != torch.float32 or code.shape != (N, C):
            raise ValueError("Not valid code (float32 [N, C])")
        _lib.call("mcgm_mc_forward", x.data_ptr(), y.data_ptr(), _dtype_code(x), int(nchw), N, P, C,
                  0, 0, 0, code.data_ptr(), _stream())
    return y


def bn_stats(x, running_mean=None, running_var=None, momentum=0.1, eps=1e-5):
    """Train-mode batch statistics of an NHWC tensor -> (mean, rstd), updates running stats in place."""
    _check_cuda(x, running_mean, running_var)
    C = x.shape[-1]
    P = x.numel() // C
    mean = torch.empty(C, dtype=torch.float32, device=x.device)
    rstd = torch.empty(C, dtype=torch.float32, device=x.device)
    ws = _workspace("bn", _lib.query("mcgm_bn_stats_workspace_bytes", C), x.device)
    ev = _prof_begin()
    _lib.call("mcgm_bn_stats", x.data_ptr(), _dtype_code(x), P, C, mean.data_ptr(), rstd.data_ptr(),
              _ptr(running_mean), _ptr(running_var), float(momentum), float(eps), ws.data_ptr(), _stream())
    _lib.launch_count += 1                       # partial + finalize kernels
    _prof_end(ev, "gate", x.numel() * x.element_size(), f"bn_stats P{P} C{C}")
    return mean, rstd


class ConvStats:
    """Per-channel sum / sum-of-squares partials of a tensor, written by the epilogue of the conv that produced
    it (``conv_fwd_bnstats``): the BatchNorm gate that consumes the tensor finalises them instead of reading
    the tensor once more for its batch statistics."""
    __slots__ = ("partial", "slots", "count")

    def __init__(self, partial, slots, count):
        self.partial, self.slots, self.count = partial, slots, count


# statistics of the most recent conv output (set by ConvFn / ConvUp2Fn when asked; consumed by take_stats)
_last_stats = None


def take_stats():
    global _last_stats
    st, _last_stats = _last_stats, None
    return st


def stats_supported(x, Cout, res=None):
    """Epilogue statistics ride on the staged (TMA-store) epilogue: bf16 tensors whose Cout is one channel tile."""
    return x.dtype == torch.bfloat16 and Cout in (64, 128, 256)


def bn_finalize(stats, C, running_mean=None, running_var=None, momentum=0.1, eps=1e-5):
    """(mean, rstd) + running-stat update from conv-epilogue partials (no pass over the tensor)."""
    dev = stats.partial.device
    mean = torch.empty(C, dtype=torch.float32, device=dev)
    rstd = torch.empty(C, dtype=torch.float32, device=dev)
    ev = _prof_begin()
    _lib.call("mcgm_bn_finalize", stats.partial.data_ptr(), stats.slots, C, stats.count, mean.data_ptr(), rstd.data_ptr(),
              _ptr(running_mean), _ptr(running_var), float(momentum), float(eps), _stream())
    _prof_end(ev, "gate", 0, f"bn_finalize slots{stats.slots} C{C}")
    return mean, rstd


def conv_fwd_bnstats(kind, x, w, bias=None, res=None, flags=0):
    """conv_fwd / conv_fwd_ex (bf16, linear epilogue) that also returns the ConvStats of its output."""
    import ctypes
    _check_cuda(x, w, bias, res)
    N, H, W, Cin = x.shape
    if kind == CONV_S1:
        Cout, R, S, Ho, Wo, taps = w.shape[0], w.shape[1], w.shape[2], H, W, w.shape[1] * w.shape[2]
    elif kind == CONV_PHASE2:
        Cout, R, S, Ho, Wo, taps = w.shape[1], 0, 0, 2 * H, 2 * W, 9
    else:
        raise ValueError("Not valid conv kind for epilogue statistics")
    out = torch.empty(N, Ho, Wo, Cout, dtype=torch.bfloat16, device=x.device)
    partial = torch.empty(_lib.query("mcgm_conv_fwd_bnstats_bytes", Cout) // 4, dtype=torch.float32, device=x.device)
    slots = ctypes.c_int(0)
    if bias is not None:
        flags |= EPI_BIAS
    ev = _prof_begin()
    _lib.call("mcgm_conv_fwd_bnstats", kind, x.data_ptr(), w.data_ptr(), out.data_ptr(), N, H, W, Cin, Cout, R, S,
              _ptr(bias), _ptr(res), flags, partial.data_ptr(), ctypes.byref(slots), _stream())
    ex_taps = 4 if kind == CONV_PHASE2 else taps
    _prof_end(ev, "conv", 2.0 * N * Ho * Wo * Cout * Cin * ex_taps,
              f"conv+bnstats kind{kind} {N}x{H}x{W} {Cin}->{Cout} f{flags:#x}", 2.0 * N * Ho * Wo * Cout * Cin * taps)
    return out, ConvStats(partial, slots.value, N * Ho * Wo)


def gate_fwd(x, flags, labels=None, packed=None, mean=None, rstd=None, gamma=None, beta=None, want_raw=False):
    _check_cuda(x, labels, mean, rstd, gamma, beta)
    N, H, W, C = x.shape
    up = 2 if flags & GATE_UP2 else 1
    y = torch.empty(N, H * up, W * up, C, dtype=x.dtype, device=x.device)
    y_raw = torch.empty_like(x) if want_raw else None
    ev = _prof_begin()
    _lib.call("mcgm_gate_fwd", x.data_ptr(), y.data_ptr(), _ptr(y_raw), _dtype_code(x), N, H, W, C,
              _ptr(labels), _ptr(packed.bits if packed else None), packed.M if packed else 0,
              _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(beta), flags, _stream())
    _prof_end(ev, "gate", (x.numel() + y.numel() + (x.numel() if want_raw else 0)) * x.element_size(),
              f"gate_fwd {N}x{H}x{W}x{C} f{flags:#x} raw{int(want_raw)}")
    return y, y_raw


def gate_bwd(dy, d_raw, x, flags, labels=None, packed=None, mean=None, rstd=None, gamma=None, beta=None):
    _check_cuda(dy, d_raw, x, labels)
    N, H, W, C = x.shape
    dx = torch.empty_like(x)
    dgamma = dbeta = None
    ws = None
    if flags & GATE_BN:
        dgamma = torch.empty(C, dtype=torch.float32, device=x.device)
        dbeta = torch.empty(C, dtype=torch.float32, device=x.device)
        ws = _workspace("gate_bwd", _lib.query("mcgm_gate_bwd_workspace_bytes", C), x.device)
    ev = _prof_begin()
    _lib.call("mcgm_gate_bwd", _ptr(dy), _ptr(d_raw), x.data_ptr(), dx.data_ptr(), _dtype_code(x), N, H, W, C,
              _ptr(labels), _ptr(packed.bits if packed else None), packed.M if packed else 0,
              _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(beta), _ptr(dgamma), _ptr(dbeta), flags,
              _ptr(ws), _stream())
    if flags & GATE_BN:
        _lib.launch_count += 2
    passes = 2 if flags & GATE_BN else 1          # BN: reduce pass + apply pass both read dy and x
    nbytes = passes * ((dy.numel() if dy is not None else 0) + x.numel()) + x.numel()
    nbytes += d_raw.numel() if d_raw is not None else 0
    _prof_end(ev, "gate", nbytes * x.element_size(), f"gate_bwd {N}x{H}x{W}x{C} f{flags:#x} raw{int(d_raw is not None)}")
    return dx, dgamma, dbeta


def pack_weight(w, inv_sigma=None, want_fwd=True, want_flip=True):
    """fp32 OIHW -> (bf16 KRSC, bf16 CRSK with flipped taps)."""
    _check_cuda(w, inv_sigma)
    if w.dtype != torch.float32 or w.dim() != 4:
        raise ValueError("Not valid weight (float32 [K, C, R, S])")
    K, C, R, S = w.shape
    fwd = torch.empty(K, R, S, C, dtype=torch.bfloat16, device=w.device) if want_fwd else None
    flip = torch.empty(C, R, S, K, dtype=torch.bfloat16, device=w.device) if want_flip else None
    _lib.call("mcgm_pack_weight", w.data_ptr(), K, C, R, S, _ptr(inv_sigma), _ptr(fwd), _ptr(flip), _stream())
    return fwd, flip


def pack_weight_padded(w, inv_sigma, row_stride):
    """bf16 [K, 1, 1, row_stride] <- W * inv_sigma, columns in (r,s,c) order, zero padded."""
    _check_cuda(w, inv_sigma)
    K, C, R, S = w.shape
    out = torch.empty(K, 1, 1, row_stride, dtype=torch.bfloat16, device=w.device)
    _lib.call("mcgm_pack_weight_padded", w.data_ptr(), K, C, R, S, _ptr(inv_sigma), out.data_ptr(), row_stride, _stream())
    return out


class SNBatch:
    """All spectral-normed layers of one network: ONE launch runs every layer's power iteration
    (in place on weight_u / weight_v) and snapshots [1/sigma | u | v] per layer."""

    def __init__(self, modules):
        import ctypes
        self.modules = list(modules)
        n = len(self.modules)
        if n == 0 or n > 16:
            raise ValueError("Not valid number of spectral-normed layers (1..16 per batch)")
        self.shapes = []
        offs, off = [], 0
        for m in self.modules:
            w = m.weight_orig
            K, CRS = w.shape[0], w.numel() // w.shape[0]
            self.shapes.append((K, CRS))
            offs.append(off)
            off += (1 + K + CRS + 3) // 4 * 4
        self.total, self.offs = off, offs
        self._c = ctypes
        self._key = None

    def _arrays(self):
        key = tuple((m.weight_orig.data_ptr(), m.weight_u.data_ptr(), m.weight_v.data_ptr()) for m in self.modules)
        if key != self._key:
            c, n = self._c, len(self.modules)
            self._W = (c.c_void_p * n)(*[k[0] for k in key])
            self._u = (c.c_void_p * n)(*[k[1] for k in key])
            self._v = (c.c_void_p * n)(*[k[2] for k in key])
            self._K = (c.c_int * n)(*[s[0] for s in self.shapes])
            self._CRS = (c.c_int * n)(*[s[1] for s in self.shapes])
            self._off = (c.c_long * n)(*self.offs)
            self._key = key
        return self._W, self._K, self._CRS, self._u, self._v, self._off

    def run(self, do_iter):
        """-> {module: state view [1/sigma | u | v]}"""
        dev = self.modules[0].weight_orig.device
        for m in self.modules:
            _check_cuda(m.weight_orig, m.weight_u, m.weight_v)
        state = torch.empty(self.total, dtype=torch.float32, device=dev)
        W, K, CRS, u, v, off = self._arrays()
        _lib.call("mcgm_sn_power_iter", len(self.modules), W, K, CRS, u, v, state.data_ptr(), off, int(do_iter),
                  1e-12, _stream())
        return {m: state[o:o + 1 + k + crs] for m, o, (k, crs) in zip(self.modules, self.offs, self.shapes)}


# Deferred end-of-backward work.  With SN_DEFER set (train_gan's D update does), two kinds of parameter-gradient
# work do not run where autograd reaches them but are queued and flushed by ONE engine callback at the end of the
# backward pass (autograd sees ``None`` for those inputs; the results are added to ``param.grad`` in place):
#   * the weight gradients of the small-spatial (H*W <= 256) spectral-normed convs — independent of the rest of
#     the pass, run as grouped launches of up to four same-shape layers (``mcgm_conv_wgrad_grouped``);
#   * the spectral-norm backward of every layer — one pair of launches for all of them.
# Only valid for ``loss.backward()`` on leaf parameters, which is what the training step does.
SN_DEFER = False
WGRAD_GROUP = os.environ.get("MCGM_NO_WGRAD_GROUP") is None      # developer A/B knob (tools/): grouped small-layer wgrads
_sn_pending = []
_wg_pending = []
_flush_queued = False


def reset_deferred():
    """Drop queued end-of-backward work (left behind only when an earlier backward pass raised)."""
    global _flush_queued
    _sn_pending.clear()
    _wg_pending.clear()
    _flush_queued = False


def _queue_flush():
    global _flush_queued
    if not _flush_queued:
        _flush_queued = True
        torch.autograd.Variable._execution_engine.queue_callback(_deferred_flush)


def _deferred_flush():
    global _flush_queued
    _flush_queued = False
    _wgrad_flush()
    _sn_flush()


def _wgrad_chunks(items, limit=4):
    """Split same-shape deferred items into launches of <= `limit`, never the same parameter (item[4]) twice in
    one launch: its bias gradient is accumulated by ONE multi-tensor add per launch, which must not alias."""
    chunks = []
    for it in items:
        for ch in chunks:
            if len(ch) < limit and all(o[4] is not it[4] for o in ch):
                ch.append(it)
                break
        else:
            chunks.append([it])
    return chunks


def _wgrad_flush():
    """Grouped weight gradients of the queued (x, dout, R, S, weight_orig, bias, sn_state) items."""
    import ctypes as c
    items, _wg_pending[:] = list(_wg_pending), []
    groups = {}
    for it in items:
        x, dout, R, S = it[0], it[1], it[2], it[3]
        groups.setdefault((tuple(x.shape), dout.shape[-1], R, S, it[5] is not None), []).append(it)
    for (xs, Cout, R, S, has_bias), its in groups.items():
        N, H, W, Cin = xs
        for chunk in _wgrad_chunks(its):
            n = len(chunk)
            dev = chunk[0][0].device
            dw = torch.empty(n, Cout, R, S, Cin, dtype=torch.float32, device=dev)
            db = torch.empty(n, Cout, dtype=torch.float32, device=dev) if has_bias else None
            nbytes = _lib.query("mcgm_conv_wgrad_grouped_workspace_bytes", n, N, H, W, Cin, Cout, R, S)
            if nbytes == 0:
                raise ValueError(f"Not valid grouped wgrad shape N={N} H={H} W={W} Cin={Cin} Cout={Cout} R={R}")
            ws = _workspace("wgrad_grouped", nbytes, dev)
            ev = _prof_begin()
            _lib.call("mcgm_conv_wgrad_grouped", n, (c.c_void_p * n)(*[it[0].data_ptr() for it in chunk]),
                      (c.c_void_p * n)(*[it[1].data_ptr() for it in chunk]), (c.c_void_p * n)(*[dw[i].data_ptr() for i in range(n)]),
                      (c.c_void_p * n)(*[db[i].data_ptr() for i in range(n)]) if has_bias else None,
                      N, H, W, Cin, Cout, R, S, ws.data_ptr(), _stream())
            _lib.launch_count += 1
            _prof_end(ev, "wgrad", 2.0 * n * N * H * W * Cout * Cin * R * S, f"wgrad_grouped x{n} {N}x{H}x{W} {Cin}->{Cout} k{R}")
            for i, it in enumerate(chunk):
                _sn_pending.append((dw[i], R * S * Cin, 1, it[4], it[6]))
            if has_bias:
                grads = []
                for it in chunk:
                    b = it[5]
                    if b.grad is None:
                        b.grad = torch.zeros_like(b)
                    grads.append(b.grad)
                torch._foreach_add_(grads, list(db.unbind(0)))


def defer_wgrad(x, dout, R, S, w, bias, sn_state):
    """Queue a small-spatial SN conv's weight gradient for the grouped end-of-backward launch."""
    _wg_pending.append((x, dout, R, S, w, bias, sn_state))
    _queue_flush()


def _sn_flush():
    import ctypes as c
    items, _sn_pending[:] = list(_sn_pending), []
    if not items:
        return
    dev = items[0][3].device
    for start in range(0, len(items), 32):
        chunk = items[start:start + 32]
        n = len(chunk)
        grads = []
        for (_, _, _, w, _) in chunk:
            if w.grad is None:
                w.grad = torch.zeros_like(w)
            if not w.grad.is_contiguous():
                raise ValueError("Not valid gradient layout: spectral-normed weights need contiguous .grad")
            grads.append(w.grad)
        shp = [(w.shape if w.dim() == 4 else (w.shape[0], w.shape[1], 1, 1)) for (_, _, _, w, _) in chunk]
        ws = _workspace("sn_bwd_batched", _lib.query("mcgm_sn_weight_bwd_batched_workspace_bytes"), dev)
        _lib.call("mcgm_sn_weight_bwd_batched", n,
                  (c.c_void_p * n)(*[it[0].data_ptr() for it in chunk]), (c.c_long * n)(*[it[1] for it in chunk]),
                  (c.c_int * n)(*[int(it[2]) for it in chunk]), (c.c_void_p * n)(*[it[3].data_ptr() for it in chunk]),
                  (c.c_void_p * n)(*[it[4].data_ptr() for it in chunk]), (c.c_int * n)(*[s_[0] for s_ in shp]),
                  (c.c_int * n)(*[s_[1] for s_ in shp]), (c.c_int * n)(*[s_[2] * s_[3] for s_ in shp]),
                  (c.c_void_p * n)(*[g.data_ptr() for g in grads]), ws.data_ptr(), _stream())
        _lib.launch_count += 1


def sn_pack_all(states):
    """Operand copies of every spectral-normed CONV layer of one forward in one launch.
    states: {module: [1/sigma | u | v]} from SNBatch.run.  -> {module: (KRSC, flipped CRSK)} for wide layers,
    {module: padded [K,1,1,32|64]} for thin (Cin <= 4) layers.  Linear layers are left to SNScaleFn."""
    import ctypes as c
    mods = [m for m in states if m.weight_orig.dim() == 4]
    if not mods:
        return {}
    dev = mods[0].weight_orig.device
    sizes, metas = [], []
    for m in mods:
        K, C, R, S = m.weight_orig.shape
        if C <= 4:
            rs = _KU                    # 64 columns = 128-byte rows: the UMMA operand layout of the thin forward kernel
            sizes.append(K * rs)
            metas.append((1, rs))
        else:
            sizes.append(2 * K * C * R * S)
            metas.append((0, 0))
    offs, off = [], 0
    for n_ in sizes:
        offs.append(off)
        off += (n_ + 63) // 64 * 64
    flat = torch.empty(off, dtype=torch.bfloat16, device=dev)
    out, o0, o1 = {}, [], []
    for m, o, (mode, rs) in zip(mods, offs, metas):
        K, C, R, S = m.weight_orig.shape
        if mode == 1:
            out[m] = flat[o:o + K * rs].view(K, 1, 1, rs)
            o0.append(out[m].data_ptr()); o1.append(None)
        else:
            n_ = K * C * R * S
            out[m] = (flat[o:o + n_].view(K, R, S, C), flat[o + n_:o + 2 * n_].view(C, R, S, K))
            o0.append(out[m][0].data_ptr()); o1.append(out[m][1].data_ptr())
    n = len(mods)
    shp = [m.weight_orig.shape for m in mods]
    _lib.call("mcgm_pack_weight_batched", n, (c.c_void_p * n)(*[m.weight_orig.data_ptr() for m in mods]),
              (c.c_void_p * n)(*[states[m].data_ptr() for m in mods]), (c.c_int * n)(*[s_[0] for s_ in shp]),
              (c.c_int * n)(*[s_[1] for s_ in shp]), (c.c_int * n)(*[s_[2] for s_ in shp]), (c.c_int * n)(*[s_[3] for s_ in shp]),
              (c.c_int * n)(*[mt[0] for mt in metas]), (c.c_int * n)(*[mt[1] for mt in metas]),
              (c.c_void_p * n)(*o0), (c.c_void_p * n)(*o1), _stream())
    return out


def sn_weight_bwd(dwbar, row_stride, krsc, w, state):
    """dW (OIHW, fp32) = (dW_bar - <dW_bar, W_bar> u v^T) / sigma.  With SN_DEFER: queued for the batched
    end-of-backward launch (returns None: the result goes straight into ``w.grad``)."""
    _check_cuda(w, state)
    if SN_DEFER and w.is_leaf and w.requires_grad:
        _sn_pending.append((dwbar, row_stride, krsc, w, state))
        _queue_flush()
        return None
    K, C, R, S = w.shape if w.dim() == 4 else (w.shape[0], w.shape[1], 1, 1)
    dw = torch.empty_like(w)
    ws = _workspace("sn_bwd", _lib.query("mcgm_sn_weight_bwd_workspace_bytes"), w.device)
    _lib.call("mcgm_sn_weight_bwd", dwbar.data_ptr(), row_stride, int(krsc), w.data_ptr(), state.data_ptr(),
              K, C, R, S, dw.data_ptr(), ws.data_ptr(), _stream())
    _lib.launch_count += 1
    return dw


class SNScaleFn(torch.autograd.Function):
    """W_bar = W / sigma for a small (Linear) weight; backward through sigma with u, v constant."""

    @staticmethod
    def forward(ctx, w, state):
        ctx.save_for_backward(w, state)
        return w * state[0]

    @staticmethod
    def backward(ctx, dwbar):
        w, state = ctx.saved_tensors
        dwbar = dwbar.contiguous()
        return sn_weight_bwd(dwbar, dwbar.numel() // dwbar.shape[0], 0, w, state), None


def conv_fwd(x, w_krsc, out_dtype=torch.bfloat16, bias=None, res=None, aux=None, labels=None, packed=None,
             flags=0, thin=False):
    """Implicit-GEMM conv (tcgen05): x [N,H,W,Cin] bf16, w [Cout,R,S,Cin] bf16 -> [N,H,W,Cout].
    thin: the launch is the GEMM half of an RGB-side (3-channel) conv — HBM-bound; the value is the ALGORITHMIC
    byte count of the whole op (wide tensor once + the 3-channel image), the im2col / col2im launches around it
    are booked with zero work in the same family."""
    _check_cuda(x, w_krsc, bias, res, aux, labels)
    if x.dtype != torch.bfloat16 or w_krsc.dtype != torch.bfloat16:
        raise ValueError("Not valid dtype: conv operands must be bfloat16")
    N, H, W, Cin = x.shape
    Cout, R, S, Cin2 = w_krsc.shape
    if Cin2 != Cin:
        raise ValueError(f"Not valid weight: Cin {Cin2} != input channels {Cin}")
    out = torch.empty(N, H, W, Cout, dtype=out_dtype, device=x.device)
    if bias is not None:
        flags |= EPI_BIAS
    ev = _prof_begin()
    _lib.call("mcgm_conv_fwd", x.data_ptr(), w_krsc.data_ptr(), out.data_ptr(),
              F32 if out_dtype == torch.float32 else BF16, N, H, W, Cin, Cout, R, S,
              _ptr(bias), _ptr(res), _ptr(aux), _ptr(labels), _ptr(packed.bits if packed else None),
              packed.M if packed else 0, flags, _stream())
    desc = f"conv {N}x{H}x{W} {Cin}->{Cout} k{R} f{flags:#x} {'f32' if out_dtype == torch.float32 else 'bf16'}"
    if thin:
        _prof_end(ev, "thin", float(thin), desc)
    else:
        _prof_end(ev, "conv", 2.0 * N * H * W * Cout * Cin * R * S, desc)
    return out


def conv_fwd_ex(kind, x, w, out_dtype=torch.bfloat16, bias=None, res=None, aux=None, labels=None, packed=None,
                flags=0, ref_phase_taps=4):
    """Generalised geometry on the tcgen05 conv kernel (include/mcgm.h):
    CONV_S2K4  : w [Cout,4,4,Cin]        x [N,H,W,Cin] -> [N,H/2,W/2,Cout]   (4x4, stride 2, pad 1)
    CONV_PHASE2: w [4,Cout,2,2,Cin]      x [N,H,W,Cin] -> [N,2H,2W,Cout]     (4 phases x 2x2 taps)"""
    _check_cuda(x, w, bias, res, aux, labels)
    if x.dtype != torch.bfloat16 or w.dtype != torch.bfloat16:
        raise ValueError("Not valid dtype: conv operands must be bfloat16")
    N, H, W, Cin = x.shape
    if kind == CONV_S2K4:
        Cout, Ho, Wo = w.shape[0], H // 2, W // 2
        ok = tuple(w.shape) == (Cout, 4, 4, Cin)
    elif kind == CONV_PHASE2:
        Cout, Ho, Wo = w.shape[1], 2 * H, 2 * W
        ok = tuple(w.shape) == (4, Cout, 2, 2, Cin)
    else:
        raise ValueError("Not valid conv kind")
    if not ok:
        raise ValueError(f"Not valid weight shape {tuple(w.shape)} for conv kind {kind}")
    out = torch.empty(N, Ho, Wo, Cout, dtype=out_dtype, device=x.device)
    if bias is not None:
        flags |= EPI_BIAS
    ev = _prof_begin()
    _lib.call("mcgm_conv_fwd_ex", kind, BF16, x.data_ptr(), w.data_ptr(), out.data_ptr(),
              F32 if out_dtype == torch.float32 else BF16, N, H, W, Cin, Cout, 0, 0,
              _ptr(bias), _ptr(res), _ptr(aux), _ptr(labels), _ptr(packed.bits if packed else None),
              packed.M if packed else 0, flags, _stream())
    # EXECUTED MACs: the phase conv runs 2x2 taps per output pixel (the reference op it replaces, Up2 -> Conv3x3,
    # has 9; ConvTranspose2d(4,2,1) has 4 as well), the strided conv all 16
    ex_taps, ref_taps = (4, ref_phase_taps) if kind == CONV_PHASE2 else (16, 16)
    _prof_end(ev, "conv", 2.0 * N * Ho * Wo * Cout * Cin * ex_taps,
              f"conv_ex kind{kind} {N}x{H}x{W} {Cin}->{Cout} f{flags:#x}", 2.0 * N * Ho * Wo * Cout * Cin * ref_taps)
    return out


def conv_fwd_tf32(kind, x, w, R=0, S=0, bias=None, res=None, aux=None, labels=None, packed=None, flags=0):
    """fp32 parity pipeline: fp32 NHWC x, fp32 weights (KRSC / phase layout), tensor cores in TF32."""
    _check_cuda(x, w, bias, res, aux, labels)
    if x.dtype != torch.float32 or w.dtype != torch.float32:
        raise ValueError("Not valid dtype: the TF32 path takes float32 operands")
    N, H, W, Cin = x.shape
    if kind == CONV_S1:
        Cout, Ho, Wo = w.shape[0], H, W
    elif kind == CONV_S2K4:
        Cout, Ho, Wo = w.shape[0], H // 2, W // 2
    else:
        Cout, Ho, Wo = w.shape[1], 2 * H, 2 * W
    out = torch.empty(N, Ho, Wo, Cout, dtype=torch.float32, device=x.device)
    if bias is not None:
        flags |= EPI_BIAS
    _lib.call("mcgm_conv_fwd_ex", kind, F32, x.data_ptr(), w.data_ptr(), out.data_ptr(), F32, N, H, W, Cin, Cout, R, S,
              _ptr(bias), _ptr(res), _ptr(aux), _ptr(labels), _ptr(packed.bits if packed else None),
              packed.M if packed else 0, flags, _stream())
    return out


def split_tf32(x):
    """x (fp32) -> (big, small), both exactly TF32-representable, big + small = x up to 2^-21."""
    _check_cuda(x)
    big, small = torch.empty_like(x), torch.empty_like(x)
    _lib.call("mcgm_split_tf32", x.data_ptr(), big.data_ptr(), small.data_ptr(), x.numel(), _stream())
    return big, small


def split_bf16x3(x):
    """x (fp32) -> three bf16 tensors whose sum is x up to 2^-24."""
    _check_cuda(x)
    ps = [torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) for _ in range(3)]
    _lib.call("mcgm_split_bf16x3", x.data_ptr(), ps[0].data_ptr(), ps[1].data_ptr(), ps[2].data_ptr(), x.numel(),
              _stream())
    return ps


_NONLIN = EPI_RELU | EPI_MASK | EPI_POSMASK | EPI_TANH


def conv_f32(kind, x, w, R=0, S=0, bias=None, res=None, res_up=False, aux=None, labels=None, packed=None, flags=0):
    """fp32-accurate convolution of the parity pipeline: 3xTF32 on the tensor cores.
    x, w, res, aux fp32 (w in the kernel layout of `kind`); the linear part is accumulated through the
    fp32 residual input, the non-linear epilogue flags run on the last pass only."""
    xb, xs = split_tf32(x)
    wb, ws = split_tf32(w)
    first = (EPI_RES_UP2 if res_up else EPI_RES) if res is not None else 0
    out = conv_fwd_tf32(kind, xb, wb, R, S, bias, res, None, None, None, first)
    out = conv_fwd_tf32(kind, xb, ws, R, S, None, out, None, None, None, EPI_RES)
    return conv_fwd_tf32(kind, xs, wb, R, S, None, out, aux, labels, packed, EPI_RES | (flags & _NONLIN))


def wgrad_f32(kind, x, dy, R=0, S=0, want_bias=True):
    """fp32-accurate weight gradient: bf16x3 operand pieces, the six products of total order <= 4 on the
    bf16 MN-major wgrad kernel.  Returns dw fp32 [Cout, taps..., Cin] and dbias."""
    xs, ds = split_bf16x3(x), split_bf16x3(dy)
    dw = db = None
    for i, j in ((0, 0), (0, 1), (1, 0), (0, 2), (1, 1), (2, 0)):
        wb = want_bias and i == 0
        d, b = (conv_wgrad(xs[i], ds[j], R, S, want_bias=wb) if kind == CONV_S1
                else conv_wgrad_ex(kind, xs[i], ds[j], want_bias=wb))
        dw = d if dw is None else dw + d
        if b is not None:
            db = b if db is None else db + b
    return dw, db


def conv_wgrad_ex(kind, x, dy, want_bias=True, ref_phase_taps=4):
    """dw fp32 [Cout, 4, 4, Cin] (tap index p*4+q, see include/mcgm.h) and dbias; x dims are [N,H,W,Cin]."""
    _check_cuda(x, dy)
    N, H, W, Cin = x.shape
    Cout = dy.shape[-1]
    nbytes = _lib.query("mcgm_conv_wgrad_ex_workspace_bytes", kind, BF16, N, H, W, Cin, Cout, 0, 0)
    if nbytes == 0:
        raise ValueError(f"Not valid wgrad shape N={N} H={H} W={W} Cin={Cin} Cout={Cout} kind={kind}")
    ws = _workspace("wgrad", nbytes, x.device)
    dw = torch.empty(Cout, 4, 4, Cin, dtype=torch.float32, device=x.device)
    db = torch.empty(Cout, dtype=torch.float32, device=x.device) if want_bias else None
    ev = _prof_begin()
    _lib.call("mcgm_conv_wgrad_ex", kind, BF16, x.data_ptr(), dy.data_ptr(), dw.data_ptr(), _ptr(db), N, H, W, Cin,
              Cout, 0, 0, ws.data_ptr(), _stream())
    _lib.launch_count += 1
    # EXECUTED MACs.  PHASE2: 16 taps over the N*H*W input-grid pixels (= 4 per output pixel; the reference op
    # Up2 -> Conv3x3 has 9 per output pixel); S2K4: 16 taps per output pixel
    npix = dy.shape[0] * dy.shape[1] * dy.shape[2]
    ex = 2.0 * npix * Cout * Cin * (4 if kind == CONV_PHASE2 else 16)
    ref = 2.0 * npix * Cout * Cin * (ref_phase_taps if kind == CONV_PHASE2 else 16)
    _prof_end(ev, "wgrad", ex, f"wgrad_ex kind{kind} {N}x{H}x{W} {Cin}->{Cout}", ref)
    return dw, db


def conv_wgrad(x, dy, R, S, want_bias=True, thin=False):
    """dw (fp32 KRSC [Cout,R,S,Cin]) and dbias from NHWC bf16 x and dy."""
    _check_cuda(x, dy)
    N, H, W, Cin = x.shape
    Cout = dy.shape[-1]
    nbytes = _lib.query("mcgm_conv_wgrad_workspace_bytes", N, H, W, Cin, Cout, R, S)
    if nbytes == 0:
        raise ValueError(f"Not valid wgrad shape N={N} H={H} W={W} Cin={Cin} Cout={Cout} R={R}")
    ws = _workspace("wgrad", nbytes, x.device)
    dw = torch.empty(Cout, R, S, Cin, dtype=torch.float32, device=x.device)
    db = torch.empty(Cout, dtype=torch.float32, device=x.device) if want_bias else None
    ev = _prof_begin()
    _lib.call("mcgm_conv_wgrad", x.data_ptr(), dy.data_ptr(), dw.data_ptr(), _ptr(db), N, H, W, Cin, Cout, R, S,
              ws.data_ptr(), _stream())
    _lib.launch_count += 1                       # wgrad + split-K / bias reduce
    if thin:
        _prof_end(ev, "thin", float(thin), f"wgrad {N}x{H}x{W} {Cin}->{Cout} k{R}")
    else:
        _prof_end(ev, "wgrad", 2.0 * N * H * W * Cout * Cin * R * S, f"wgrad {N}x{H}x{W} {Cin}->{Cout} k{R}")
    return dw, db


def pool2(x, scale):
    """y[n,h,w,c] = scale * sum of the 2x2 window (AvgPool2d(2) with scale=.25; upsample backward with 1)."""
    _check_cuda(x)
    N, H, W, C = x.shape
    if H % 2 or W % 2:
        raise ValueError("Not valid shape for 2x2 pooling")
    y = torch.empty(N, H // 2, W // 2, C, dtype=x.dtype, device=x.device)
    _lib.call("mcgm_pool2", x.data_ptr(), y.data_ptr(), _dtype_code(x), N, H // 2, W // 2, C, float(scale), _stream())
    return y


def unpool2(dy, scale):
    """dx[n,2h+i,2w+j,c] = scale * dy[n,h,w,c] (AvgPool2d backward with .25; nearest upsample with 1)."""
    _check_cuda(dy)
    N, H, W, C = dy.shape
    dx = torch.empty(N, 2 * H, 2 * W, C, dtype=dy.dtype, device=dy.device)
    _lib.call("mcgm_unpool2", dy.data_ptr(), dx.data_ptr(), _dtype_code(dy), N, H, W, C, float(scale), _stream())
    return dx


# --------------------------------------------------------------------------- autograd Functions
class MCMaskFn(torch.autograd.Function):
    """x * code.detach() (src/modules/modules.py:75): dx = dy * code, nothing flows to the indicator."""

    @staticmethod
    def forward(ctx, x, labels, packed, code, nchw):
        ctx.labels, ctx.packed, ctx.code, ctx.nchw = labels, packed, code, nchw
        return mc_mask(x, labels, packed, code, nchw)

    @staticmethod
    def backward(ctx, dy):
        return mc_mask(dy.contiguous(), ctx.labels, ctx.packed, ctx.code, ctx.nchw), None, None, None, None


class GateFn(torch.autograd.Function):
    """BN(train) -> ReLU -> [2x nearest upsample] -> MC mask, plus the optional raw-gated output.

    Restates src/models/mcgan.py:15-18 / 20-22 / 56-58 (with BN) and :104-105 (without).
    Returns (y, y_raw or None).
    """

    @staticmethod
    def forward(ctx, x, gamma, beta, labels, packed, flags, running_mean, running_var, momentum, eps, want_raw,
                pre_stats=None):
        mean = rstd = None
        if flags & GATE_BN:
            if pre_stats is not None:            # statistics already accumulated by the producing conv's epilogue
                mean, rstd = bn_finalize(pre_stats, x.shape[-1], running_mean, running_var, momentum, eps)
            else:
                mean, rstd = bn_stats(x, running_mean, running_var, momentum, eps)
        y, y_raw = gate_fwd(x, flags, labels, packed, mean, rstd, gamma, beta, want_raw)
        ctx.save_for_backward(x, gamma, beta, mean, rstd)
        ctx.labels, ctx.packed, ctx.flags, ctx.want_raw = labels, packed, flags, want_raw
        if want_raw:
            return y, y_raw
        ctx.mark_non_differentiable()
        return y, None

    @staticmethod
    def backward(ctx, dy, d_raw):
        x, gamma, beta, mean, rstd = ctx.saved_tensors
        dy = dy.contiguous() if dy is not None else None
        d_raw = d_raw.contiguous() if d_raw is not None else None
        if dy is None and d_raw is None:
            return (None,) * 12
        if dy is None and (ctx.flags & GATE_BN):
            up = 2 if ctx.flags & GATE_UP2 else 1
            N, H, W, C = x.shape
            dy = torch.zeros(N, H * up, W * up, C, dtype=x.dtype, device=x.device)
        dx, dgamma, dbeta = gate_bwd(dy, d_raw, x, ctx.flags, ctx.labels, ctx.packed, mean, rstd, gamma, beta)
        return dx, dgamma, dbeta, None, None, None, None, None, None, None, None, None


class ActNormGateFn(torch.autograd.Function):
    """Glow coupling-net gate: ActNorm(logdet=False) -> ReLU -> MC mask (src/models/mcglow.py:139-146)
    in one kernel pair: y = mask * relu(scale * (x + loc)); dx = mask*relu' * scale * dy,
    dscale = sum g*(x+loc), dloc = scale * sum g."""

    @staticmethod
    def forward(ctx, x, scale, loc, labels, packed):
        C = x.shape[-1]
        sc, lc = scale.detach().reshape(C).contiguous(), loc.detach().reshape(C).contiguous()
        mean, rstd, beta = -lc, torch.ones_like(sc), torch.zeros_like(sc)
        flags = GATE_BN | GATE_AFFINE | GATE_RELU | GATE_MASK
        y, _ = gate_fwd(x, flags, labels, packed, mean, rstd, sc, beta, False)
        ctx.save_for_backward(x, sc, mean, rstd, beta)
        ctx.labels, ctx.packed, ctx.flags, ctx.pshape = labels, packed, flags, scale.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        x, sc, mean, rstd, beta = ctx.saved_tensors
        dx, dgamma, dbeta = gate_bwd(dy.contiguous(), None, x, ctx.flags, ctx.labels, ctx.packed, mean, rstd, sc, beta)
        return dx, dgamma.reshape(ctx.pshape), (sc * dbeta).reshape(ctx.pshape), None, None


class ConvFn(torch.autograd.Function):
    """3x3 / 1x1 stride-1 conv on tensor cores with fused epilogue.

    w is the fp32 OIHW master weight (state_dict layout).  ``gate`` = (labels, packed):
    out = mask * relu(conv + bias) — D's Conv→ReLU→MC (src/models/mcgan.py:103-108).
    ``res`` / ``res_up``: residual add of a same-resolution / half-resolution tensor (the
    ResBlock sum mcgan.py:42, with the shortcut's nearest upsample folded in).
    ``premask_dx``: the input x is itself a ReLU-gated activation with no other consumer,
    so dgrad emits dx * 1[x > 0] (masked-gradient epilogue) and the producer of x is told
    (``grad_premasked``) not to mask again.
    """

    @staticmethod
    def forward(ctx, x, w, bias, res, res_up, gate, premask_dx, grad_premasked, packs, sn, out_f32=False,
                want_stats=False):
        # sn: [1/sigma | u | v] snapshot of a spectral-normed layer (w is then weight_orig)
        f32 = x.dtype == torch.float32           # fp32 parity pipeline (cfg['compute_dtype'] = 'fp32')
        if f32:
            weff = w.detach() * sn[0] if sn is not None else w.detach()
            wf = weff.permute(0, 2, 3, 1).contiguous()
            wflip = weff.flip(2, 3).permute(1, 2, 3, 0).contiguous()
        elif packs is not None:
            wf, wflip = packs
        else:
            wf, wflip = pack_weight(w.detach(), sn[0:1] if sn is not None else None)
        flags = 0
        labels = packed = None
        if gate is not None:
            labels, packed = gate
            flags |= EPI_RELU | EPI_MASK
        if f32:
            out = conv_f32(CONV_S1, x, wf, w.shape[2], w.shape[3], bias, res, res_up, None, labels, packed, flags)
        else:
            if res is not None:
                flags |= EPI_RES_UP2 if res_up else EPI_RES
            global _last_stats
            if want_stats and gate is None and not out_f32 and stats_supported(x, w.shape[0]):
                out, _last_stats = conv_fwd_bnstats(CONV_S1, x, wf, bias, res, flags)
            else:
                out = conv_fwd(x, wf, torch.float32 if out_f32 else torch.bfloat16, bias, res, None, labels, packed, flags)
        ctx.save_for_backward(x, out if gate is not None else None, w if sn is not None else None, sn,
                              bias if sn is not None else None)
        ctx.wflip, ctx.wshape = wflip, tuple(w.shape)
        ctx.has_bias, ctx.has_res, ctx.res_up = bias is not None, res is not None, res_up
        ctx.gated, ctx.premask_dx, ctx.grad_premasked = gate is not None, premask_dx, grad_premasked
        return out

    @staticmethod
    def backward(ctx, dout):
        x, out, w_orig, sn, bias_p = ctx.saved_tensors
        dout = dout.contiguous()
        if dout.dtype != x.dtype:
            dout = dout.to(x.dtype)                 # fp32-output convs: gradients re-enter the bf16 pipeline
        K, C, R, S = ctx.wshape
        if ctx.gated and not ctx.grad_premasked:
            # d(pre) = dout * 1[out > 0]   (mask already folded: out == 0 where masked)
            dout = torch.where(out > 0, dout, torch.zeros_like(dout))
        dx = dw = db = dres = None
        f32 = x.dtype == torch.float32
        if ctx.needs_input_grad[0]:
            if f32:
                dx = conv_f32(CONV_S1, dout, ctx.wflip, R, S, None, None, False, x if ctx.premask_dx else None, None,
                              None, EPI_POSMASK if ctx.premask_dx else 0)
            elif ctx.premask_dx:
                dx = conv_fwd(dout, ctx.wflip, torch.bfloat16, None, None, x, None, None, EPI_POSMASK)
            else:
                dx = conv_fwd(dout, ctx.wflip, torch.bfloat16)
        if (SN_DEFER and WGRAD_GROUP and not f32 and sn is not None and ctx.needs_input_grad[1] and x.shape[1] * x.shape[2] <= 256
                and w_orig.is_leaf and (bias_p is None or (bias_p.is_leaf and ctx.needs_input_grad[2]))):
            # small-spatial layer: its weight gradient joins the grouped launch at the end of the backward pass
            defer_wgrad(x, dout, R, S, w_orig, bias_p if ctx.has_bias else None, sn)
        elif ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            if f32:
                dw_krsc, db = wgrad_f32(CONV_S1, x, dout, R, S, want_bias=ctx.has_bias)
                dw_krsc = dw_krsc.reshape(K, R, S, C)
            else:
                dw_krsc, db = conv_wgrad(x, dout, R, S, want_bias=ctx.has_bias)
            if sn is not None:
                dw = sn_weight_bwd(dw_krsc, R * S * C, 1, w_orig, sn)
            else:
                dw = dw_krsc.permute(0, 3, 1, 2)      # KRSC -> OIHW view
        if ctx.has_res and ctx.needs_input_grad[3]:
            dres = pool2(dout, 1.0) if ctx.res_up else dout
        return dx, dw, db, dres, None, None, None, None, None, None, None, None


def thin_im2col(thin, R, S, flip=False, KU=64, stride=1, pad=None, dtype=torch.bfloat16):
    """U[p, t*Ct + j] = thin[stride*p +/- off(t), j] as bf16 (or fp32) [N,Ho,Wo,KU] (zero padded)."""
    _check_cuda(thin)
    if thin.dtype != torch.float32:
        raise ValueError("Not valid dtype: thin tensors are float32")
    N, H, W, Ct = thin.shape
    pad = (R - 1) // 2 if pad is None else pad
    Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
    U = torch.empty(N, Ho, Wo, KU, dtype=dtype, device=thin.device)
    ev = _prof_begin()
    _lib.call("mcgm_thin_im2col", thin.data_ptr(), U.data_ptr(), F32 if dtype == torch.float32 else BF16, N, H, W,
              Ct, R, S, stride, pad, int(flip), KU, _stream())
    _prof_end(ev, "thin", 0.0, f"thin_im2col {N}x{H}x{W}x{Ct} k{R} s{stride}")
    return U


def thin_col2im(T, Ct, R, S, bias=None, tanh=False):
    """thin[p, j] = bias[j] + sum_t T[p + off(t), t*Ct + j], optional tanh; T fp32 [N,H,W,CT]."""
    _check_cuda(T, bias)
    N, H, W, CT = T.shape
    out = torch.empty(N, H, W, Ct, dtype=torch.float32, device=T.device)
    flags = (EPI_BIAS if bias is not None else 0) | (EPI_TANH if tanh else 0)
    ev = _prof_begin()
    _lib.call("mcgm_thin_col2im", T.data_ptr(), CT, out.data_ptr(), N, H, W, Ct, R, S, _ptr(bias), flags, _stream())
    _prof_end(ev, "thin", 0.0, f"thin_col2im {N}x{H}x{W}x{Ct} k{R}")
    return out


def _pad_rows(m, rows):
    """[r, c] -> [rows, c] zero padded (tiny weight matrices)."""
    if m.shape[0] == rows:
        return m
    return torch.cat([m, m.new_zeros(rows - m.shape[0], m.shape[1])], 0)


def _pad_cols(m, cols):
    if m.shape[1] == cols:
        return m
    return torch.cat([m, m.new_zeros(m.shape[0], cols - m.shape[1])], 1)


PACKS_FORCE = False             # see models/mcgan.py:_packs (graph-segment capture)
PACK_CACHES = []                # packed-weight caches of the model files (cleared before a CUDA-graph capture)


def clear_pack_caches():
    """Forget every cached packed weight: the next use re-packs.  Called right before a CUDA graph is captured so that
    the packing kernels are part of the graph no matter what ran before the capture."""
    for c in PACK_CACHES:
        c.clear()
ACT_DTYPE = torch.bfloat16      # storage type of activations; torch.float32 selects the fp32 parity pipeline
_TK = 32          # patch columns of the direct thin-conv kernels (csrc/thin_mma.cu)
_KU = 64          # im2col width of the thin tensor (R*S*Ct <= 36 taps*channels, padded to one 64-wide K block)


# Nearest-up2 followed by a 3x3 conv collapses, per output phase (ph, pw), to a 2x2 conv over the
# LOW-resolution input with pre-summed taps: rows {r} of the 3x3 kernel that read the same input row
# are added.  _RM[ph, a, r] = 1 if kernel row r reads input row (i + a - 1 + ph) for output row 2i+ph.
_RM = torch.tensor([[[1., 0., 0.], [0., 1., 1.]], [[1., 1., 0.], [0., 0., 1.]]])
# The adjoint is a 4x4 stride-2 conv over the output gradient; _QM[x, r] = 1 if kernel row r
# contributes at offset x - 1 (x = 3 - 2a - ph).
_QM = torch.tensor([[0., 0., 1.], [0., 1., 1.], [1., 1., 0.], [1., 0., 0.]])
_mats = {}


def _up2_mats(device):
    m = _mats.get(device)
    if m is None:
        m = _mats[device] = (_RM.to(device), _QM.to(device))
    return m


def up2_conv_packs(w, dtype=torch.bfloat16):
    """fp32 OIHW 3x3 weight -> (phase weights [4,Cout,2,2,Cin], adjoint 4x4/s2 weights [Cin,4,4,Cout])."""
    rm, qm = _up2_mats(w.device)
    wp = torch.einsum('par,qbs,kcrs->pqkabc', rm, rm, w)
    K, C = w.shape[0], w.shape[1]
    wp = wp.reshape(4, K, 2, 2, C).to(dtype).contiguous()
    kd = torch.einsum('xr,ys,kcrs->cxyk', qm, qm, w).to(dtype).contiguous()
    return wp, kd


class ConvUp2Fn(torch.autograd.Function):
    """nn.Upsample(2, 'nearest') -> Conv2d(C, K, 3, 1, 1) (src/models/mcgan.py:17-19) WITHOUT
    materialising the upsampled tensor: 4 output phases x 2x2 taps on the low-resolution input
    (2.25x fewer MACs, and the gate kernel writes T instead of 4T).  Backward: the data gradient is a
    4x4 stride-2 conv of dout (TMA traversal stride 2), the weight gradient a 16-tap strided wgrad
    folded back onto the 3x3 kernel."""

    @staticmethod
    def forward(ctx, x, w, bias, packs, want_stats=False):
        if x.dtype == torch.float32:             # fp32 parity pipeline
            wp, kd = up2_conv_packs(w.detach(), torch.float32)
            out = conv_f32(CONV_PHASE2, x, wp, bias=bias)
        else:
            wp, kd = packs if packs is not None else up2_conv_packs(w.detach())
            global _last_stats
            if want_stats and stats_supported(x, w.shape[0]):
                out, _last_stats = conv_fwd_bnstats(CONV_PHASE2, x, wp, bias)
            else:
                out = conv_fwd_ex(CONV_PHASE2, x, wp, torch.bfloat16, bias, ref_phase_taps=9)
        ctx.save_for_backward(x)
        ctx.kd, ctx.has_bias, ctx.dev = kd, bias is not None, w.device
        return out

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        dout = dout.contiguous()
        dx = dw = db = None
        f32 = x.dtype == torch.float32
        if ctx.needs_input_grad[0]:
            dx = conv_f32(CONV_S2K4, dout, ctx.kd) if f32 else conv_fwd_ex(CONV_S2K4, dout, ctx.kd, torch.bfloat16)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dk, db = (wgrad_f32(CONV_PHASE2, x, dout, want_bias=ctx.has_bias) if f32
                      else conv_wgrad_ex(CONV_PHASE2, x, dout, want_bias=ctx.has_bias, ref_phase_taps=9))     # [K,4,4,C]
            _, qm = _up2_mats(ctx.dev)
            dw = torch.einsum('xr,ys,kxyc->kcrs', qm, qm, dk)
        return dx, dw, db, None, None


_PIDX = ((3, 1), (2, 0))     # kernel index of (phase ph, tap a): p = 3 - 2a - ph


_pidx_cache = {}


def _pidx_tensor(device):
    """Device copy of _PIDX, made once per device (a host->device copy is not allowed during graph capture)."""
    t = _pidx_cache.get(device)
    if t is None:
        t = _pidx_cache[device] = torch.tensor(_PIDX, device=device)
    return t


def _phase_weights(w_oc_ic_44):
    """[O, I, 4, 4] (O = output channels of the PHASE conv) -> [4, O, 2, 2, I] with
    out[ph*2+pw, o, a, b, i] = w[o, i, 3-2a-ph, 3-2b-pw]."""
    idx = _pidx_tensor(w_oc_ic_44.device)
    parts = [w_oc_ic_44[:, :, idx[ph]][:, :, :, idx[pw]] for ph in range(2) for pw in range(2)]   # [O,I,2,2] each
    return torch.stack(parts).permute(0, 1, 3, 4, 2).contiguous()


class ConvS2Fn(torch.autograd.Function):
    """Conv2d(C, K, 4, 2, 1) (src/models/mcvae.py:41-49) on the tensor cores: the strided input pixels
    are gathered by TMA with traversal stride 2.  Backward: data gradient = 4-phase 2x2 conv of dout,
    weight gradient = 16-tap strided wgrad."""

    @staticmethod
    def forward(ctx, x, w, bias):
        wk = w.detach().permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()                 # KRSC
        out = conv_fwd_ex(CONV_S2K4, x, wk, torch.bfloat16, bias)
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w = ctx.saved_tensors
        dout = dout.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            wph = _phase_weights(w.detach().permute(1, 0, 2, 3)).to(torch.bfloat16)          # [4, Cin, 2, 2, Cout]
            dx = conv_fwd_ex(CONV_PHASE2, dout, wph, torch.bfloat16)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dk, db = conv_wgrad_ex(CONV_S2K4, x, dout, want_bias=ctx.has_bias)             # [K,4,4,C]
            dw = dk.permute(0, 3, 1, 2)
        return dx, dw, db


class ConvT4Fn(torch.autograd.Function):
    """ConvTranspose2d(C, K, 4, 2, 1) (src/models/mcvae.py:87-92; weight layout [C, K, 4, 4]) as a
    4-phase 2x2 conv.  K <= 4 (the RGB output layer) produces an fp32 image; its gradients go
    through the stride-2 im2col of the (thin) output gradient and 1x1 GEMMs."""

    @staticmethod
    def forward(ctx, x, w, bias):
        C, K = w.shape[0], w.shape[1]
        wph = _phase_weights(w.detach().permute(1, 0, 2, 3)).to(torch.bfloat16)              # [4, K, 2, 2, C]
        thin = K <= 4
        out = conv_fwd_ex(CONV_PHASE2, x, wph, torch.float32 if thin else torch.bfloat16, bias)
        ctx.save_for_backward(x, w)
        ctx.has_bias, ctx.thin = bias is not None, thin
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w = ctx.saved_tensors
        C, K = w.shape[0], w.shape[1]
        dout = dout.contiguous()
        dx = dw = db = None
        if ctx.thin:
            U = thin_im2col(dout, 4, 4, False, _KU, 2, 1)                                   # [N,H,W,64], cols (p,q,j)
            if ctx.needs_input_grad[0]:
                wm = _pad_cols(w.detach().permute(0, 2, 3, 1).reshape(C, 16 * K), _KU)
                dx = conv_fwd(U, wm.to(torch.bfloat16).reshape(C, 1, 1, _KU).contiguous(), torch.bfloat16,
                              thin=x.numel() * 2 + dout.numel() * 4)
            if ctx.needs_input_grad[1]:
                dwm, _ = conv_wgrad(x, U, 1, 1, want_bias=False, thin=x.numel() * 2 + dout.numel() * 4)   # [KU,1,1,C]
                dw = dwm.reshape(_KU, C)[:16 * K].reshape(4, 4, K, C).permute(3, 2, 0, 1)
            if ctx.has_bias and ctx.needs_input_grad[2]:
                db = dout.sum(dim=(0, 1, 2))
            return dx, dw, db
        if ctx.needs_input_grad[0]:
            kd = w.detach().permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()             # [C,4,4,K]
            dx = conv_fwd_ex(CONV_S2K4, dout, kd, torch.bfloat16)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dk, db = conv_wgrad_ex(CONV_PHASE2, x, dout, want_bias=ctx.has_bias)           # [K,4,4,C]
            dw = dk.permute(3, 0, 1, 2)
        return dx, dw, db


class WideFromThinFn(torch.autograd.Function):
    """Conv2d(Ct<=4 -> Cw) on an fp32 NHWC image, optional ReLU+mask epilogue
    (src/models/mcgan.py:77-79, 87).  im2col of the 3-channel image (tiny) turns it into a
    1x1 tcgen05 GEMM with K = 64; the weight gradient is the 1x1 wgrad GEMM; the image
    gradient (needed in the G update) is a 1x1 GEMM to [P, 32] followed by col2im."""

    @staticmethod
    def forward(ctx, thin, w, bias, gate, grad_premasked, sn, stride=1, pad=None, wm_packed=None):
        K, C, R, S = w.shape                                   # OIHW: K = Cw, C = Ct
        if stride != 1 and thin.requires_grad:
            raise ValueError("Not valid: the strided thin conv has no image gradient (first encoder layer only)")
        f32 = ACT_DTYPE == torch.float32
        labels, packed = gate if gate is not None else (None, None)
        if (not f32 and stride == 1 and pad is None and R == S and R in (1, 3) and R * S * C <= _TK - 1 and K % 16 == 0
                and 128 % thin.shape[2] == 0 and (thin.shape[1] * thin.shape[2]) % 128 == 0):
            # direct path (csrc/thin_mma.cu): the patch matrix never leaves shared memory
            wm = wm_packed if (wm_packed is not None and wm_packed.shape[-1] == _KU) else \
                pack_weight_padded(w.detach(), sn[0:1] if sn is not None else None, _KU)
            N, H, Wd, _ = thin.shape
            out = torch.empty(N, H, Wd, K, dtype=torch.bfloat16, device=thin.device)
            ev = _prof_begin()
            _lib.call("mcgm_thin_conv_fwd", thin.data_ptr(), wm.data_ptr(), _KU, _ptr(bias), _ptr(labels),
                      _ptr(packed.bits if packed else None), packed.M if packed else 0, int(gate is not None),
                      out.data_ptr(), N, H, Wd, C, R, S, K, _stream())
            _prof_end(ev, "thin", thin.numel() * 4 + out.numel() * 2, f"thin_conv {N}x{H}x{Wd} {C}->{K} k{R} gate{int(gate is not None)}")
            ctx.save_for_backward(thin, w, out if gate is not None else None, sn)
            ctx.gated, ctx.has_bias, ctx.grad_premasked, ctx.direct = gate is not None, bias is not None, grad_premasked, True
            return out
        ctx.direct = False
        U = thin_im2col(thin, R, S, False, _KU, stride, pad, ACT_DTYPE)
        flags = 0
        if gate is not None:
            flags = EPI_RELU | EPI_MASK
        if f32:
            weff = w.detach() * sn[0] if sn is not None else w.detach()
            wm = _pad_cols(weff.permute(0, 2, 3, 1).reshape(K, R * S * C), _KU).reshape(K, 1, 1, _KU).contiguous()
            out = conv_f32(CONV_S1, U, wm, 1, 1, bias, None, False, None, labels, packed, flags)
        else:
            wm = pack_weight_padded(w.detach(), sn[0:1] if sn is not None else None, _KU)
            out = conv_fwd(U, wm, torch.bfloat16, bias, None, None, labels, packed, flags,
                           thin=thin.numel() * 4 + U.shape[0] * U.shape[1] * U.shape[2] * K * 2)
        ctx.save_for_backward(U, w, out if gate is not None else None, sn)
        ctx.gated, ctx.has_bias, ctx.grad_premasked = gate is not None, bias is not None, grad_premasked
        return out

    @staticmethod
    def backward(ctx, dout):
        U, w, out, sn = ctx.saved_tensors
        dout = dout.contiguous()
        K, C, R, S = w.shape
        if ctx.gated and not ctx.grad_premasked:
            dout = torch.where(out > 0, dout, torch.zeros_like(dout))
        dthin = dw = db = None
        if ctx.needs_input_grad[0]:
            # dthin[q,j] = sum_t T[q + off(t')][(t',j)],  T[p][(t',j)] = sum_co dout[p,co] w[co,j,flip(t')]
            CT = 32 if R * S * C <= 32 else 64
            weff = w.detach() * sn[0] if sn is not None else w.detach()
            wd = _pad_rows(weff.flip(2, 3).permute(2, 3, 1, 0).reshape(R * S * C, K), CT)
            if dout.dtype == torch.float32:
                T = conv_f32(CONV_S1, dout, wd.reshape(CT, 1, 1, K).contiguous(), 1, 1)
            else:
                T = conv_fwd(dout, wd.to(torch.bfloat16).reshape(CT, 1, 1, K).contiguous(), torch.float32,
                             thin=dout.numel() * 2 + dout.numel() // K * C * 4)
            dthin = thin_col2im(T, C, R, S)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            if ctx.direct:
                thin = U                                       # the direct path saved the image itself
                N, H, Wd, _ = thin.shape
                dwm = torch.empty(K, _TK, dtype=torch.float32, device=thin.device)
                ws = _workspace("thin_wgrad", _lib.query("mcgm_thin_conv_wgrad_workspace_bytes", K), thin.device)
                ev = _prof_begin()
                _lib.call("mcgm_thin_conv_wgrad", thin.data_ptr(), dout.data_ptr(), dwm.data_ptr(), N, H, Wd, C, R, S, K,
                          ws.data_ptr(), _stream())
                _lib.launch_count += 1
                _prof_end(ev, "thin", thin.numel() * 4 + dout.numel() * 2, f"thin_wgrad {N}x{H}x{Wd} {C}->{K} k{R}")
                db = dwm[:, R * S * C].contiguous() if ctx.has_bias else None
                ku = _TK
            else:
                if dout.dtype == torch.float32:
                    dwm, db = wgrad_f32(CONV_S1, U, dout, 1, 1, want_bias=ctx.has_bias)
                else:
                    dwm, db = conv_wgrad(U, dout, 1, 1, want_bias=ctx.has_bias,           # [K,1,1,KU], columns (r,s,c)
                                         thin=dout.numel() * 2 + dout.numel() // K * C * 4)
                ku = _KU
            if sn is not None:
                dw = sn_weight_bwd(dwm, ku, 1, w, sn)
            else:
                dw = dwm.reshape(K, ku)[:, :R * S * C].reshape(K, R, S, C).permute(0, 3, 1, 2)
        return dthin, dw, db, None, None, None, None, None, None


class ThinFromWideFn(torch.autograd.Function):
    """Conv2d(Cw -> Ct<=4) producing an fp32 NHWC image, optional tanh (mcgan.py:59-60):
    a 1x1 tcgen05 GEMM of the wide tensor against all R*S*Ct tap filters ([P, 32] fp32),
    then a col2im gather-sum.  Backward: im2col of the image gradient, then 1x1 GEMMs."""

    @staticmethod
    def forward(ctx, wide, w, bias, tanh):
        K, C, R, S = w.shape                                   # OIHW: K = Ct, C = Cw
        CT = 32 if R * S * K <= 32 else 64
        wm = _pad_rows(w.detach().permute(2, 3, 0, 1).reshape(R * S * K, C), CT)
        if wide.dtype == torch.float32:
            T = conv_f32(CONV_S1, wide, wm.reshape(CT, 1, 1, C).contiguous(), 1, 1)
        else:
            T = conv_fwd(wide, wm.to(torch.bfloat16).reshape(CT, 1, 1, C).contiguous(), torch.float32,
                         thin=wide.numel() * 2 + wide.numel() // C * K * 4)
        out = thin_col2im(T, K, R, S, bias, tanh)
        ctx.save_for_backward(wide, w, out if tanh else None)
        ctx.tanh, ctx.has_bias = tanh, bias is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        wide, w, out = ctx.saved_tensors
        K, C, R, S = w.shape
        dout = dout.contiguous()
        if ctx.tanh:
            dout = dout * (1.0 - out * out)                     # fp32 [N,H,W,Ct]: tiny
        dwide = dw = db = None
        f32 = wide.dtype == torch.float32
        Ub = thin_im2col(dout, R, S, True, _KU, 1, None, wide.dtype)          # Ub[q,(t,j)] = dout[q - off(t), j]
        if ctx.needs_input_grad[0]:
            wb = _pad_cols(w.detach().permute(1, 2, 3, 0).reshape(C, R * S * K), _KU)
            if f32:
                dwide = conv_f32(CONV_S1, Ub, wb.reshape(C, 1, 1, _KU).contiguous(), 1, 1)
            else:
                dwide = conv_fwd(Ub, wb.to(torch.bfloat16).reshape(C, 1, 1, _KU).contiguous(), torch.bfloat16,
                                 thin=wide.numel() * 2 + dout.numel() * 4)
        if ctx.needs_input_grad[1]:
            dwm, _ = (wgrad_f32(CONV_S1, wide, Ub, 1, 1, want_bias=False) if f32
                      else conv_wgrad(wide, Ub, 1, 1, want_bias=False, thin=wide.numel() * 2 + dout.numel() * 4))   # [KU,1,1,C]
            dw = dwm.reshape(_KU, C)[:R * S * K].reshape(R, S, K, C).permute(2, 3, 0, 1)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dout.sum(dim=(0, 1, 2))
        return dwide, dw, db, None


class Pool2Fn(torch.autograd.Function):
    """AvgPool2d(2) on NHWC (src/models/mcgan.py:81,88,109,114)."""

    @staticmethod
    def forward(ctx, x):
        return pool2(x, 0.25)

    @staticmethod
    def backward(ctx, dy):
        return unpool2(dy.contiguous(), 0.25)


# ----------------------------------------------------------------------------- Glow flow step (few-channel part)
FLOW_XA_CH = 64          # channels of the coupling network's bf16 input (tcgen05 conv operand granularity)


class FlowFrontFn(torch.autograd.Function):
    """ActNorm(C) -> InvConv2dLU of one flow step (src/models/mcglow.py:24-55, 76-116) on fp32 NHWC:
    y = W (scale*(x+loc)); second output xa = bf16 y[..., :C/2] zero-padded to 64 channels, the coupling
    network's input.  W is the CxC matrix from ``InvConv2dLU.calc_weight()`` (autograd carries the
    gradient on through the LU parameters)."""

    @staticmethod
    def forward(ctx, x, loc, scale, W):
        _check_cuda(x, loc, scale, W)
        N, H, Wd, C = x.shape
        lc, sc, Wc = loc.detach().reshape(C).contiguous(), scale.detach().reshape(C).contiguous(), W.detach().contiguous()
        y = torch.empty_like(x)
        xa = torch.empty(N, H, Wd, FLOW_XA_CH, dtype=torch.bfloat16, device=x.device)
        _lib.call("mcgm_flow_front_fwd", x.data_ptr(), lc.data_ptr(), sc.data_ptr(), Wc.data_ptr(), y.data_ptr(),
                  xa.data_ptr(), N * H * Wd, C, FLOW_XA_CH, _stream())
        ctx.save_for_backward(x, lc, sc, Wc)
        ctx.pshape = loc.shape
        return y, xa

    @staticmethod
    def backward(ctx, dy, dxa):
        x, lc, sc, Wc = ctx.saved_tensors
        N, H, Wd, C = x.shape
        if dy is None:
            dy = torch.zeros_like(x)
        dy = dy.contiguous()
        dxa = dxa.contiguous() if dxa is not None else None
        dx = torch.empty_like(x)
        G = torch.empty(C, C + 1, dtype=torch.float32, device=x.device)
        ws = _workspace("flow_front", _lib.query("mcgm_flow_front_bwd_workspace_bytes", C), x.device)
        _lib.call("mcgm_flow_front_bwd", dy.data_ptr(), _ptr(dxa), x.data_ptr(), lc.data_ptr(), sc.data_ptr(),
                  Wc.data_ptr(), dx.data_ptr(), G.data_ptr(), N * H * Wd, C, FLOW_XA_CH, ws.data_ptr(), _stream())
        _lib.launch_count += 2
        Gm, S = G[:, :C], G[:, C]
        dW = Gm * sc[None, :]
        dscale = (Wc * Gm).sum(0)
        dloc = sc * (Wc.t() @ S)
        return dx, dloc.reshape(ctx.pshape), dscale.reshape(ctx.pshape), dW


class FlowTailFn(torch.autograd.Function):
    """The affine (or additive) coupling around its network plus ZeroConv2d's exp(3*scale)
    (src/models/mcglow.py:119-130, 153-164): y fp32 NHWC [N,H,W,C], net fp32 NHWC [N,H,W,>=C] (the raw
    zero-conv output) -> out [N,H,W,C], logdet [N]."""

    @staticmethod
    def forward(ctx, y, net, zscale, affine):
        _check_cuda(y, net, zscale)
        N, H, Wd, C = y.shape
        y, net = y.contiguous(), net.contiguous()
        zs = zscale.detach().reshape(-1).contiguous()
        out = torch.empty_like(y)
        logdet = torch.zeros(N, dtype=torch.float32, device=y.device)
        _lib.call("mcgm_flow_tail_fwd", y.data_ptr(), net.data_ptr(), zs.data_ptr(), out.data_ptr(), logdet.data_ptr(),
                  N, H * Wd, C, net.shape[-1], 1 if affine else 0, _stream())
        ctx.save_for_backward(y, net, zs)
        ctx.affine, ctx.zshape = affine, zscale.shape
        return out, logdet

    @staticmethod
    def backward(ctx, dout, dlogdet):
        y, net, zs = ctx.saved_tensors
        N, H, Wd, C = y.shape
        if dout is None:
            dout = torch.zeros_like(y)
        dout = dout.contiguous()
        dlogdet = dlogdet.contiguous() if dlogdet is not None else None
        dy = torch.empty_like(y)
        dnet = torch.empty_like(net)
        nz = C if ctx.affine else C // 2
        dzp = torch.empty(N, nz, dtype=torch.float32, device=y.device)
        _lib.call("mcgm_flow_tail_bwd", dout.data_ptr(), _ptr(dlogdet), y.data_ptr(), net.data_ptr(), zs.data_ptr(),
                  dy.data_ptr(), dnet.data_ptr(), dzp.data_ptr(), N, H * Wd, C, net.shape[-1],
                  1 if ctx.affine else 0, _stream())
        return dy, dnet, dzp.sum(0).reshape(ctx.zshape), None


def flow_front(x, loc, scale, W, want_xa=True):
    """The front kernel without autograd: y = W (scale * (x + loc)) on fp32 NHWC, plus (optionally) the bf16
    zero-padded first half of y.  Used by the sampling path, where Flow.reverse is the SAME kernel with the
    parameters W' = diag(1/scale) W^-1, scale' = 1, loc' = -W (scale * loc)."""
    _check_cuda(x, loc, scale, W)
    N, H, Wd, C = x.shape
    y = torch.empty_like(x)
    xa = torch.empty(N, H, Wd, FLOW_XA_CH, dtype=torch.bfloat16, device=x.device) if want_xa else None
    _lib.call("mcgm_flow_front_fwd", x.data_ptr(), loc.data_ptr(), scale.data_ptr(), W.data_ptr(), y.data_ptr(),
              _ptr(xa), N * H * Wd, C, FLOW_XA_CH, _stream())
    return y, xa


def flow_tail_rev(out, net, zscale, affine):
    """AffineCoupling.reverse around its network (src/models/mcglow.py:166-175): out fp32 NHWC [N,H,W,C], net the RAW
    zero-conv output fp32 NHWC [N,H,W,>=C] -> y = [out_a, out_b / sigmoid(log_s + 2) - t]."""
    _check_cuda(out, net, zscale)
    N, H, Wd, C = out.shape
    y = torch.empty_like(out)
    _lib.call("mcgm_flow_tail_rev", out.data_ptr(), net.data_ptr(), zscale.data_ptr(), y.data_ptr(), N, H * Wd, C,
              net.shape[-1], 1 if affine else 0, _stream())
    return y


# ----------------------------------------------------------------------------- gated PixelCNN
def _tap_arrays(taps):
    import ctypes as c
    n = len(taps)
    return n, (c.c_byte * n)(*[int(t[0]) for t in taps]), (c.c_byte * n)(*[int(t[1]) for t in taps])


def conv_fwd_taps(x, w_ktc, taps, out_dtype=torch.bfloat16, bias=None, res=None, flags=0):
    """Explicit-tap stride-1 conv on the tcgen05 kernel: x [N,H,W,Cin] bf16, w [Cout, T, Cin] bf16,
    taps = [(dh, dw)] * T -> out[n,i,j] = sum_t x[n, i+dh_t, j+dw_t] . w[:, t, :]."""
    _check_cuda(x, w_ktc, bias, res)
    N, H, W, Cin = x.shape
    Cout, T = w_ktc.shape[0], w_ktc.shape[1]
    if T != len(taps) or w_ktc.shape[2] != Cin:
        raise ValueError("Not valid weight for the tap list")
    out = torch.empty(N, H, W, Cout, dtype=out_dtype, device=x.device)
    n, dh, dw = _tap_arrays(taps)
    if bias is not None:
        flags |= EPI_BIAS
    if res is not None:
        flags |= EPI_RES
    ev = _prof_begin()
    _lib.call("mcgm_conv_fwd_taps", x.data_ptr(), w_ktc.data_ptr(), out.data_ptr(), F32 if out_dtype == torch.float32 else BF16,
              N, H, W, Cin, Cout, n, dh, dw, _ptr(bias), _ptr(res), None, None, None, 0, flags, _stream())
    _prof_end(ev, "conv", 2.0 * N * H * W * Cout * Cin * T, f"conv_taps {N}x{H}x{W} {Cin}->{Cout} t{T} f{flags:#x}")
    return out


def conv_wgrad_taps(x, dy, taps, want_bias=True):
    """dw fp32 [Cout, T, Cin] (and dbias) of the explicit-tap conv."""
    _check_cuda(x, dy)
    N, H, W, Cin = x.shape
    Cout = dy.shape[-1]
    n, dh, dw_ = _tap_arrays(taps)
    nbytes = _lib.query("mcgm_conv_wgrad_taps_workspace_bytes", N, H, W, Cin, Cout, n)
    if nbytes == 0:
        raise ValueError(f"Not valid wgrad shape N={N} H={H} W={W} Cin={Cin} Cout={Cout} taps={n}")
    ws = _workspace("wgrad", nbytes, x.device)
    dw = torch.empty(Cout, n, Cin, dtype=torch.float32, device=x.device)
    db = torch.empty(Cout, dtype=torch.float32, device=x.device) if want_bias else None
    ev = _prof_begin()
    _lib.call("mcgm_conv_wgrad_taps", x.data_ptr(), dy.data_ptr(), dw.data_ptr(), _ptr(db), N, H, W, Cin, Cout, n, dh, dw_,
              ws.data_ptr(), _stream())
    _lib.launch_count += 1
    _prof_end(ev, "wgrad", 2.0 * N * H * W * Cout * Cin * n, f"wgrad_taps {N}x{H}x{W} {Cin}->{Cout} t{n}")
    return dw, db


class ConvTapsFn(torch.autograd.Function):
    """nn.Conv2d(C, K, (R, S), 1, (ph, pw)) cropped to the input's H x W — the masked / causal stacks of
    MCGatedMaskedConv2d (src/models/mcpixelcnn.py:31-36, 48-52) — as an explicit-tap tcgen05 conv.
    w: fp32 OIHW master weight; tap (r, s) reads input pixel (i + r - ph, j + s - pw).
    res: optional same-shape tensor added in the epilogue (``v2h + h_horiz``)."""

    @staticmethod
    def forward(ctx, x, w, bias, pad, res):
        K, C, R, S = w.shape
        taps = [(r - pad[0], s - pad[1]) for r in range(R) for s in range(S)]
        wf, wflip = pack_weight(w.detach())                       # [K,R,S,C] and flipped [C,R,S,K]
        out = conv_fwd_taps(x, wf.view(K, R * S, C), taps, torch.bfloat16, bias, res)
        ctx.save_for_backward(x)
        ctx.wflip, ctx.taps, ctx.wshape = wflip, taps, (K, C, R, S)
        ctx.has_bias, ctx.has_res = bias is not None, res is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        K, C, R, S = ctx.wshape
        dout = dout.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            # dx[q] = sum_t dout[q - off_t] . w[:, t, :]^T ; the flipped pack holds tap T-1-t at index t
            rtaps = [(-dh, -dw_) for (dh, dw_) in reversed(ctx.taps)]
            dx = conv_fwd_taps(dout, ctx.wflip.view(C, R * S, K), rtaps, torch.bfloat16)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dwk, db = conv_wgrad_taps(x, dout, ctx.taps, want_bias=ctx.has_bias)       # [K, T, C]
            dw = dwk.view(K, R, S, C).permute(0, 3, 1, 2)
        return dx, dw, db, None, (dout if ctx.has_res and ctx.needs_input_grad[4] else None)


class SGateFn(torch.autograd.Function):
    """MCGatedActivation (src/models/mcpixelcnn.py:9-20) on NHWC: h [N,H,W,2C] bf16 ->
    mask * relu(bn(h[..., :C])) * sigmoid(h[..., C:]) as one kernel (plus the statistics pass in training)."""

    @staticmethod
    def forward(ctx, h, gamma, beta, labels, packed, running_mean, running_var, momentum, eps, training):
        _check_cuda(h, gamma, beta, labels)
        if h.dtype != torch.bfloat16:
            raise ValueError("Not valid dtype: the gated activation takes bfloat16")
        N, H, W, C2 = h.shape
        C = C2 // 2
        ws = _workspace("sgate", _lib.query("mcgm_sgate_workspace_bytes", C), h.device)
        if training:
            mean = torch.empty(C, dtype=torch.float32, device=h.device)
            rstd = torch.empty(C, dtype=torch.float32, device=h.device)
            _lib.call("mcgm_sgate_stats", h.data_ptr(), N * H * W, C, mean.data_ptr(), rstd.data_ptr(), _ptr(running_mean),
                      _ptr(running_var), float(momentum), float(eps), ws.data_ptr(), _stream())
            _lib.launch_count += 1
        else:
            mean, rstd = running_mean, torch.rsqrt(running_var + eps)
        out = torch.empty(N, H, W, C, dtype=torch.bfloat16, device=h.device)
        ev = _prof_begin()
        _lib.call("mcgm_sgate_fwd", h.data_ptr(), out.data_ptr(), N, H * W, C, labels.data_ptr(), packed.bits.data_ptr(),
                  packed.M, mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _stream())
        _prof_end(ev, "gate", (h.numel() + out.numel()) * 2, f"sgate_fwd {N}x{H}x{W}x{C}")
        ctx.save_for_backward(h, gamma, beta, mean, rstd)
        ctx.labels, ctx.packed, ctx.training = labels, packed, training
        return out

    @staticmethod
    def backward(ctx, dout):
        h, gamma, beta, mean, rstd = ctx.saved_tensors
        N, H, W, C2 = h.shape
        C = C2 // 2
        dout = dout.contiguous()
        dh = torch.empty_like(h)
        dgamma = torch.empty(C, dtype=torch.float32, device=h.device)
        dbeta = torch.empty(C, dtype=torch.float32, device=h.device)
        ws = _workspace("sgate", _lib.query("mcgm_sgate_workspace_bytes", C), h.device)
        ev = _prof_begin()
        _lib.call("mcgm_sgate_bwd", dout.data_ptr(), h.data_ptr(), dh.data_ptr(), N, H * W, C, ctx.labels.data_ptr(),
                  ctx.packed.bits.data_ptr(), ctx.packed.M, mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                  beta.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), int(ctx.training), ws.data_ptr(), _stream())
        _lib.launch_count += 2
        _prof_end(ev, "gate", (2 * (dout.numel() + h.numel()) + dh.numel()) * 2, f"sgate_bwd {N}x{H}x{W}x{C}")
        return dh, dgamma, dbeta, None, None, None, None, None, None, None
