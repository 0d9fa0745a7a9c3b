"""``MultimodalController`` and ``Wrapper`` — the operator of the reference
(src/modules/modules.py:49-85) behind the same constructor, buffer and list protocol.

Drop-in constraints kept:
  * the class is named ``MultimodalController`` (src/models/utils.py:51,116 selects
    modules by ``__class__.__name__``);
  * ``codebook`` is a live fp32 ``[num_mode, input_size]`` registered buffer that
    outside code may replace with a different row count (models/utils.py:43,53,117-119);
    the bit-packed device copy used by the kernels is re-derived whenever the buffer
    object, its version counter, shape or device changes;
  * ``forward([x, indicator, *rest]) -> [x * code, indicator, *rest]`` with
    ``code = indicator @ codebook`` detached — no gradient reaches the indicator.
The multiply itself runs in libmcgm.so (mcgm_mc_forward); there is no CPU fallback.
"""
import torch
import torch.nn as nn

from . import ops


class MultimodalController(nn.Module):
    def __init__(self, input_size, num_mode, controller_rate=0.5):
        super().__init__()
        self.input_size = input_size
        self.num_mode = num_mode
        self.controller_rate = controller_rate
        self.register_buffer('codebook', self.make_codebook())
        self._packed = None
        self._packed_key = None
        self._packed_src = None

    def make_codebook(self):
        """Fixed binary codebook, bit-identical to the reference for the same RNG state.

        src/modules/modules.py:58-69: rate 1 -> all ones; otherwise rounds of
        Bernoulli(rate) draws of a whole [num_mode, input_size] block from torch's global
        CPU generator are merged into a Python ``set`` of row tuples until it holds at
        least ``num_mode`` distinct rows, and the first ``num_mode`` rows in the set's
        iteration order are kept.  Bit-exactness depends on the torch CPU Bernoulli stream
        AND on CPython's set ordering, so this stays on the host and is never re-derived
        on the device.
        """
        M, C = self.num_mode, self.input_size
        if self.controller_rate == 1:
            return torch.ones(M, C, dtype=torch.float)
        sampler = torch.distributions.bernoulli.Bernoulli(probs=self.controller_rate)
        rows = set()
        while len(rows) < M:
            rows.update(map(tuple, sampler.sample((M, C)).tolist()))
        return torch.tensor(list(rows)[:M], dtype=torch.float)

    # ---- device-side packed copy -------------------------------------------------
    def packed(self):
        """Bit-packed codebook on the buffer's device, rebuilt if the buffer changed."""
        cb = self.codebook
        key = (cb._version, tuple(cb.shape), cb.device)
        # the source tensor itself is part of the key (compared with `is`): create() / transit() replace the
        # buffer object, and the id() of a freed tensor can be handed to its successor
        if self._packed is None or self._packed_src is not cb or self._packed_key != key:
            if cb.dtype != torch.float32:
                raise ValueError('Not valid codebook dtype')
            if not bool(((cb == 0) | (cb == 1)).all()):
                self._packed = False          # non-binary codebook (e.g. models/utils.py transit): dense path
            else:
                self._packed = ops.PackedCodebook(cb.contiguous())
            self._packed_key = key
            self._packed_src = cb
        return self._packed

    def packed_bits(self):
        """``packed()`` for the fused blocks, which gather mask rows by label from the bit-packed copy only."""
        packed = self.packed()
        if packed is False:
            raise ValueError('Not valid codebook: the fused blocks need a binary {0,1} codebook (a non-binary one '
                             'only works through MultimodalController.forward, which multiplies by the dense code)')
        return packed

    def forward(self, input):
        x, indicator = input[0], input[1]
        if not x.is_cuda:
            raise RuntimeError('MultimodalController runs on CUDA tensors only (no CPU fallback); '
                               'got a tensor on ' + str(x.device))
        if x.dtype not in (torch.float32, torch.bfloat16):
            raise ValueError('Not valid dtype')
        if x.dim() < 2 or x.size(1) != self.input_size:
            raise ValueError('Not valid input size')
        xc = x.contiguous()
        if indicator.dim() == 1 and not indicator.is_floating_point():
            # fast path: integer mode labels -> row gather from the packed codebook
            packed = self.packed()
            if packed is False:
                code = self.codebook.index_select(0, indicator.long()).contiguous()
                out = ops.MCMaskFn.apply(xc, None, None, code, x.dim() > 2)
            else:
                out = ops.MCMaskFn.apply(xc, ops.as_labels(indicator), packed, None, x.dim() > 2)
        else:
            # reference semantics for any (not necessarily one-hot) indicator: a tiny [N,M]x[M,C] product
            code = indicator.detach().to(torch.float32).matmul(self.codebook).contiguous()
            out = ops.MCMaskFn.apply(xc, None, None, code, x.dim() > 2)
        return [out, *input[1:]]


class Wrapper(nn.Module):
    """Lifts an ``nn.Module`` to the list protocol (src/modules/modules.py:79-85)."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, input):
        return [self.module(input[0]), *input[1:]]
