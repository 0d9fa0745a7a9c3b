"""GPU parity checks: every kernel family of libmcgm.so against the oracle / plain torch fp32
on identical seeded inputs.  Each check returns a list of (label, error, tolerance) rows;
tests/test_gpu_kernels.py asserts on them and tools/gpu_bringup.py prints them all.

Tolerances (written here, used everywhere):
  EXACT      0          integer / bit work: packed codebook, masks
  F32        2e-5       fp32 storage kernels, and bf16-input kernels compared on the SAME
                        bf16-rounded inputs with fp32 output (only the summation order differs)
  BF16_OUT   6e-3       same, but the result is stored in bf16 (one rounding: 2^-8 relative)
"""
import torch
import torch.nn.functional as F

from mcgm_b200 import _lib, ops
from mcgm_b200._lib import (EPI_MASK, EPI_POSMASK, EPI_RELU, EPI_RES, EPI_RES_UP2, EPI_TANH, GATE_BN, GATE_MASK,
                            GATE_RELU, GATE_UP2)
from oracle import mcgan_oracle as O

EXACT, F32_TOL, BF16_OUT = 0.0, 2e-5, 6e-3
DEV = "cuda"


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def nhwc(x):   # NCHW -> NHWC contiguous
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def _codebook(M, C, seed):
    torch.manual_seed(seed)
    return O.make_codebook(M, C, 0.5)


# ----------------------------------------------------------------------------- codebook / mask
def check_codebook_and_mask():
    rows = []
    for (M, C, N, HW) in [(10, 256, 8, 16), (100, 128, 5, 4), (1623, 512, 6, 4), (7, 64, 3, 1)]:
        cb = _codebook(M, C, 11 + M)
        packed = ops.PackedCodebook(cb.to(DEV))
        bits = packed.bits.cpu().numpy().view("uint32")
        import numpy as np
        want = np.zeros((M, (C + 31) // 32), dtype=np.uint32)
        cbn = cb.numpy().astype(np.uint32)
        for c in range(C):
            want[:, c // 32] |= cbn[:, c] << np.uint32(c % 32)
        rows.append((f"pack_codebook M={M} C={C}", float((bits != want).sum()), EXACT))
        g = torch.Generator().manual_seed(M)
        labels = torch.randint(0, M, (N,), generator=g)
        ind = O.one_hot(labels, M)
        side = int(HW ** 0.5)
        x = torch.randn(N, C, side, side, generator=g) if HW > 1 else torch.randn(N, C, generator=g)
        ref = O.mc_mask(x, ind, cb)
        xd = x.to(DEV)
        lab = labels.to(DEV).to(torch.int32)
        # NCHW fp32 (the reference layout), label path and dense-code path
        y = ops.mc_mask(xd, lab, packed, None, nchw=x.dim() > 2)
        rows.append((f"mc_mask nchw f32 labels M={M} C={C}", float((y.cpu() != ref).sum()), EXACT))
        code = ind.matmul(cb).to(DEV).contiguous()
        y = ops.mc_mask(xd, None, None, code, nchw=x.dim() > 2)
        rows.append((f"mc_mask nchw f32 code M={M} C={C}", float((y.cpu() != ref).sum()), EXACT))
        if x.dim() > 2:
            xb = nhwc(x).to(DEV).to(torch.bfloat16)
            y = ops.mc_mask(xb, lab, packed, None, nchw=False)
            refb = nhwc(O.mc_mask(nchw(xb.float().cpu()), ind, cb)).to(torch.bfloat16)
            rows.append((f"mc_mask nhwc bf16 labels M={M} C={C}", float((y.cpu() != refb).sum()), EXACT))
    return rows


# ----------------------------------------------------------------------------- BN stats + gate
def _gate_ref(x_nchw, gamma, beta, ind, cb, bn, relu, up, rm=None, rv=None):
    h = x_nchw
    if bn:
        h = F.batch_norm(h, rm, rv, gamma, beta, True, 0.1, 1e-5)
    if relu:
        h = F.relu(h)
    if up:
        h = F.interpolate(h, scale_factor=2, mode="nearest")
    return O.mc_mask(h, ind, cb)


def check_gate():
    rows = []
    cases = [(4, 8, 8, 256, 10, True, True, True), (3, 16, 16, 128, 10, True, True, False),
             (2, 4, 4, 64, 5, False, True, False), (5, 8, 8, 512, 1623, True, True, True),
             (2, 32, 32, 256, 10, True, True, False)]
    for dtype in (torch.float32, torch.bfloat16):
        for (N, H, W, C, M, bn, relu, up) in cases:
            g = torch.Generator().manual_seed(N * 131 + C)
            cb = _codebook(M, C, C + M)
            labels = torch.randint(0, M, (N,), generator=g)
            ind = O.one_hot(labels, M)
            x = (torch.randn(N, C, H, W, generator=g) * 1.5 + 0.3).to(dtype).float()   # representable in dtype
            gamma = (1 + 0.1 * torch.randn(C, generator=g)).requires_grad_(True)
            beta = (0.1 * torch.randn(C, generator=g)).requires_grad_(True)
            rm, rv = torch.zeros(C), torch.ones(C)
            xr = x.clone().requires_grad_(True)
            y_ref = _gate_ref(xr, gamma, beta, ind, cb, bn, relu, up, rm, rv)
            raw_ref = O.mc_mask(xr, ind, cb)
            up_f = 2 if up else 1
            dy = torch.randn(N, C, H * up_f, W * up_f, generator=g).to(dtype).float()
            draw = torch.randn(N, C, H, W, generator=g).to(dtype).float()
            ((y_ref * dy).sum() + (raw_ref * draw).sum()).backward()

            flags = (GATE_BN if bn else 0) | (GATE_RELU if relu else 0) | (GATE_UP2 if up else 0) | GATE_MASK
            packed = ops.PackedCodebook(cb.to(DEV))
            lab = labels.to(DEV).to(torch.int32)
            xd = nhwc(x).to(DEV).to(dtype)
            gd, bd = gamma.detach().to(DEV), beta.detach().to(DEV)
            rmd, rvd = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
            mean = rstd = None
            tag = f"{str(dtype)[6:]} N{N} {H}x{W} C{C} M{M} bn{int(bn)} up{int(up)}"
            tol = F32_TOL if dtype == torch.float32 else BF16_OUT
            if bn:
                mean, rstd = ops.bn_stats(xd, rmd, rvd, 0.1, 1e-5)
                rows.append((f"bn_stats mean {tag}", rel(mean, x.mean(dim=(0, 2, 3))), 1e-4))
                rows.append((f"bn_stats rstd {tag}", rel(rstd, (x.var(dim=(0, 2, 3), unbiased=False) + 1e-5).rsqrt()), F32_TOL))
                rows.append((f"bn running_mean {tag}", rel(rmd, rm), 1e-4))
                rows.append((f"bn running_var {tag}", rel(rvd, rv), F32_TOL))
            y, raw = ops.gate_fwd(xd, flags, lab, packed, mean, rstd, gd if bn else None, bd if bn else None, True)
            rows.append((f"gate_fwd y {tag}", rel(nchw(y.float()), y_ref), tol))
            rows.append((f"gate_fwd raw {tag}", float((nchw(raw.float()).cpu() != raw_ref.detach()).sum()), EXACT))
            dx, dg, db = ops.gate_bwd(nhwc(dy).to(DEV).to(dtype), nhwc(draw).to(DEV).to(dtype), xd, flags, lab, packed,
                                      mean, rstd, gd if bn else None, bd if bn else None)
            rows.append((f"gate_bwd dx {tag}", rel(nchw(dx.float()), xr.grad), tol * (4 if bn else 1)))
            if bn:
                rows.append((f"gate_bwd dgamma {tag}", rel(dg, gamma.grad), 1e-4))
                rows.append((f"gate_bwd dbeta {tag}", rel(db, beta.grad), 1e-4))
    return rows


def check_pool():
    rows = []
    g = torch.Generator().manual_seed(5)
    for dtype in (torch.float32, torch.bfloat16):
        x = torch.randn(3, 64, 8, 8, generator=g).to(dtype).float()
        y = ops.pool2(nhwc(x).to(DEV).to(dtype), 0.25)
        tol = F32_TOL if dtype == torch.float32 else BF16_OUT
        rows.append((f"pool2 {dtype}", rel(nchw(y.float()), F.avg_pool2d(x, 2)), tol))
        d = ops.unpool2(nhwc(x).to(DEV).to(dtype), 1.0)
        rows.append((f"unpool2 {dtype}", rel(nchw(d.float()), F.interpolate(x, scale_factor=2, mode="nearest")), EXACT + 1e-12))
    return rows


# ----------------------------------------------------------------------------- tensor-core conv
CONV_SHAPES = [  # N, H, W, Cin, Cout, R
    (2, 32, 32, 64, 64, 3), (4, 8, 8, 128, 128, 3), (2, 16, 16, 256, 256, 3), (8, 4, 4, 64, 128, 1),
    (3, 16, 16, 128, 256, 1), (2, 8, 8, 64, 3, 3), (1, 32, 32, 256, 128, 3), (5, 8, 8, 64, 64, 3),
    # enough M tiles that the two-tiles-per-CTA (MT=2) forward and two-taps-per-CTA (NT=2) wgrad paths run,
    # the second with an ODD tile count (ragged last pair)
    (40, 32, 32, 64, 128, 3), (593, 8, 8, 64, 64, 3), (40, 32, 32, 64, 256, 1),
]


def _conv_inputs(N, H, W, Cin, Cout, R, seed=0):
    g = torch.Generator().manual_seed(seed + N * 7 + Cin)
    x = torch.randn(N, Cin, H, W, generator=g).to(torch.bfloat16).float()
    w = (torch.randn(Cout, Cin, R, R, generator=g) / (Cin * R * R) ** 0.5).to(torch.bfloat16).float()
    b = torch.randn(Cout, generator=g)
    return g, x, w, b


def check_conv_fwd(shapes=CONV_SHAPES):
    rows = []
    for (N, H, W, Cin, Cout, R) in shapes:
        g, x, w, b = _conv_inputs(N, H, W, Cin, Cout, R)
        ref = F.conv2d(x, w, b, padding=(R - 1) // 2)
        xd = nhwc(x).to(DEV).to(torch.bfloat16)
        wf, wflip = ops.pack_weight(w.to(DEV))
        rows.append((f"pack_weight krsc {Cout}x{Cin}x{R}", float((wf.float().cpu() != w.permute(0, 2, 3, 1)).sum()), EXACT))
        tag = f"N{N} {H}x{W} {Cin}->{Cout} k{R}"
        y = ops.conv_fwd(xd, wf, torch.float32, b.to(DEV))
        rows.append((f"conv_fwd f32 {tag}", rel(nchw(y), ref), F32_TOL))
        y = ops.conv_fwd(xd, wf, torch.bfloat16, b.to(DEV))
        rows.append((f"conv_fwd bf16 {tag}", rel(nchw(y.float()), ref), BF16_OUT))
        if Cout % 64 == 0:
            # dgrad == forward conv with the flipped / transposed weight
            dy = torch.randn(N, Cout, H, W, generator=g).to(torch.bfloat16).float()
            xr = x.clone().requires_grad_(True)
            F.conv2d(xr, w, None, padding=(R - 1) // 2).backward(dy)
            dx = ops.conv_fwd(nhwc(dy).to(DEV).to(torch.bfloat16), wflip, torch.float32)
            rows.append((f"conv dgrad f32 {tag}", rel(nchw(dx), xr.grad), F32_TOL))
    return rows


def check_conv_epilogues():
    rows = []
    N, H, W, Cin, Cout, R, M = 4, 8, 8, 64, 128, 3, 10
    g, x, w, b = _conv_inputs(N, H, W, Cin, Cout, R, seed=3)
    cb = _codebook(M, Cout, 77)
    labels = torch.randint(0, M, (N,), generator=g)
    ind = O.one_hot(labels, M)
    packed = ops.PackedCodebook(cb.to(DEV))
    lab = labels.to(DEV).to(torch.int32)
    xd = nhwc(x).to(DEV).to(torch.bfloat16)
    wf, _ = ops.pack_weight(w.to(DEV))
    base = F.conv2d(x, w, b, padding=1)
    res = torch.randn(N, Cout, H, W, generator=g).to(torch.bfloat16).float()
    res_lo = torch.randn(N, Cout, H // 2, W // 2, generator=g).to(torch.bfloat16).float()
    aux = torch.randn(N, Cout, H, W, generator=g).to(torch.bfloat16).float()
    bd = b.to(DEV)
    y = ops.conv_fwd(xd, wf, torch.float32, bd, nhwc(res).to(DEV).to(torch.bfloat16), None, None, None, EPI_RES)
    rows.append(("conv epi +res", rel(nchw(y), base + res), F32_TOL))
    y = ops.conv_fwd(xd, wf, torch.float32, bd, nhwc(res_lo).to(DEV).to(torch.bfloat16), None, None, None, EPI_RES_UP2)
    rows.append(("conv epi +up2(res)", rel(nchw(y), base + F.interpolate(res_lo, scale_factor=2, mode="nearest")), F32_TOL))
    y = ops.conv_fwd(xd, wf, torch.float32, bd, None, None, lab, packed, EPI_RELU | EPI_MASK)
    rows.append(("conv epi relu*mask", rel(nchw(y), O.mc_mask(F.relu(base), ind, cb)), F32_TOL))
    y = ops.conv_fwd(xd, wf, torch.float32, bd, None, nhwc(aux).to(DEV).to(torch.bfloat16), None, None, EPI_POSMASK)
    rows.append(("conv epi posmask", rel(nchw(y), base * (aux > 0)), F32_TOL))
    y = ops.conv_fwd(xd, wf, torch.float32, bd, None, None, None, None, EPI_TANH)
    rows.append(("conv epi tanh", rel(nchw(y), torch.tanh(base)), F32_TOL))
    return rows


def check_conv_wgrad(shapes=CONV_SHAPES):
    rows = []
    for (N, H, W, Cin, Cout, R) in shapes:
        if Cout % 64:
            continue
        g, x, w, b = _conv_inputs(N, H, W, Cin, Cout, R, seed=9)
        dy = torch.randn(N, Cout, H, W, generator=g).to(torch.bfloat16).float()
        wr = w.clone().requires_grad_(True)
        br = b.clone().requires_grad_(True)
        F.conv2d(x, wr, br, padding=(R - 1) // 2).backward(dy)
        dw, db = ops.conv_wgrad(nhwc(x).to(DEV).to(torch.bfloat16), nhwc(dy).to(DEV).to(torch.bfloat16), R, R)
        tag = f"N{N} {H}x{W} {Cin}->{Cout} k{R}"
        rows.append((f"conv wgrad {tag}", rel(dw.permute(0, 3, 1, 2), wr.grad), F32_TOL))
        rows.append((f"conv dbias {tag}", rel(db, br.grad), F32_TOL))
    return rows


# ----------------------------------------------------------------------------- thin convs
def check_thin():
    rows = []
    for (N, H, W, Ct, Cw, R) in [(4, 8, 8, 3, 128, 3), (2, 32, 32, 3, 64, 3), (4, 16, 16, 3, 128, 1),
                                 (2, 8, 8, 1, 64, 3), (2, 16, 16, 3, 256, 3)]:
        g = torch.Generator().manual_seed(N + Cw)
        M = 10
        cb = _codebook(M, Cw, 5)
        labels = torch.randint(0, M, (N,), generator=g)
        ind = O.one_hot(labels, M)
        packed = ops.PackedCodebook(cb.to(DEV))
        lab = labels.to(DEV).to(torch.int32)
        tag = f"N{N} {H}x{W} Ct{Ct} Cw{Cw} k{R}"
        pad = (R - 1) // 2
        # thin -> wide  (Conv2d(3, Cw))
        # the GEMM path stores the im2col'd image and the weights in bf16: compare on representable inputs
        img = torch.randn(N, Ct, H, W, generator=g).to(torch.bfloat16).float()
        w = (torch.randn(Cw, Ct, R, R, generator=g) * 0.3).to(torch.bfloat16).float()
        b = torch.randn(Cw, generator=g)
        imr, wr = img.clone().requires_grad_(True), w.clone().requires_grad_(True)
        out_ref = O.mc_mask(F.relu(F.conv2d(imr, wr, b, padding=pad)), ind, cb)
        dy = torch.randn(N, Cw, H, W, generator=g).to(torch.bfloat16).float()
        out_ref.backward(dy)
        imd = nhwc(img).to(DEV).requires_grad_(True)
        wd = w.to(DEV).requires_grad_(True)
        out = ops.WideFromThinFn.apply(imd, wd, b.to(DEV), (lab, packed), False, None)
        rows.append((f"wide_from_thin fwd {tag}", rel(nchw(out.float()), out_ref), BF16_OUT))
        out.backward(nhwc(dy).to(DEV).to(torch.bfloat16))
        # the GPU masks dy with its own (bf16-rounded) output sign; identical except at exact zeros
        rows.append((f"wide_from_thin dthin {tag}", rel(nchw(imd.grad), imr.grad), 2e-3))
        rows.append((f"wide_from_thin dw {tag}", rel(wd.grad, wr.grad), 2e-3))
        # wide -> thin  (Conv2d(Cw, 3) + tanh)
        xw = torch.randn(N, Cw, H, W, generator=g).to(torch.bfloat16).float()
        w2 = (torch.randn(Ct, Cw, R, R, generator=g) / (Cw * R * R) ** 0.5).to(torch.bfloat16).float()
        b2 = torch.randn(Ct, generator=g) * 0.1
        xr, w2r, b2r = xw.clone().requires_grad_(True), w2.clone().requires_grad_(True), b2.clone().requires_grad_(True)
        o_ref = torch.tanh(F.conv2d(xr, w2r, b2r, padding=pad))
        dth = torch.randn(N, Ct, H, W, generator=g)
        o_ref.backward(dth)   # (dth * tanh' is rounded to bf16 inside the GEMM path: BF16_OUT-level error)
        xwd = nhwc(xw).to(DEV).to(torch.bfloat16).requires_grad_(True)
        w2d, b2d = w2.to(DEV).requires_grad_(True), b2.to(DEV).requires_grad_(True)
        o = ops.ThinFromWideFn.apply(xwd, w2d, b2d, True)
        rows.append((f"thin_from_wide fwd {tag}", rel(nchw(o), o_ref), F32_TOL))
        o.backward(nhwc(dth).to(DEV))
        rows.append((f"thin_from_wide dwide {tag}", rel(nchw(xwd.grad.float()), xr.grad), BF16_OUT))
        rows.append((f"thin_from_wide dw {tag}", rel(w2d.grad, w2r.grad), BF16_OUT))
        rows.append((f"thin_from_wide db {tag}", rel(b2d.grad, b2r.grad), 1e-4))
    return rows


def check_conv_ex():
    """Generalised geometries: nearest-up2+conv3x3 as a phase conv, 4x4 stride-2 conv, ConvTranspose2d(4,2,1)."""
    from mcgm_b200._lib import CONV_PHASE2, CONV_S2K4
    rows = []
    for (N, H, Cin, Cout) in [(4, 8, 64, 64), (40, 16, 64, 128), (3, 4, 128, 256), (2, 16, 256, 64)]:
        g = torch.Generator().manual_seed(N + H + Cin)
        x = torch.randn(N, Cin, H, H, generator=g).to(torch.bfloat16).float()
        tag = f"N{N} {H}x{H} {Cin}->{Cout}"
        # --- up2 + conv3x3 (mcgan.py:17-19).  Integer-valued weights: their phase sums stay bf16-exact.
        w = torch.randint(-2, 3, (Cout, Cin, 3, 3), generator=g).float() * 0.125
        b = torch.randn(Cout, generator=g)
        xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        ref = F.conv2d(F.interpolate(xr, scale_factor=2, mode="nearest"), wr, br, padding=1)
        dy = torch.randn(N, Cout, 2 * H, 2 * H, generator=g).to(torch.bfloat16).float()
        ref.backward(dy)
        xd = nhwc(x).to(DEV).to(torch.bfloat16).requires_grad_(True)
        wd, bd = w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
        wp, _ = ops.up2_conv_packs(wd.detach())
        out32 = ops.conv_fwd_ex(CONV_PHASE2, xd.detach(), wp, torch.float32, bd.detach())
        rows.append((f"up2conv fwd f32 {tag}", rel(nchw(out32), ref), F32_TOL))
        out = ops.ConvUp2Fn.apply(xd, wd, bd, None)
        rows.append((f"up2conv fwd bf16 {tag}", rel(nchw(out.float()), ref), BF16_OUT))
        out.backward(nhwc(dy).to(DEV).to(torch.bfloat16))
        rows.append((f"up2conv dgrad {tag}", rel(nchw(xd.grad.float()), xr.grad), BF16_OUT))
        rows.append((f"up2conv wgrad {tag}", rel(wd.grad, wr.grad), 5e-5))
        rows.append((f"up2conv dbias {tag}", rel(bd.grad, br.grad), F32_TOL))
        # --- Conv2d(Cin, Cout, 4, 2, 1) (mcvae.py:41-49) and its gradients
        w4 = (torch.randn(Cout, Cin, 4, 4, generator=g) / (Cin * 16) ** 0.5).to(torch.bfloat16).float()
        xr, w4r = x.clone().requires_grad_(True), w4.clone().requires_grad_(True)
        ref = F.conv2d(xr, w4r, None, stride=2, padding=1)
        dy = torch.randn(N, Cout, H // 2, H // 2, generator=g).to(torch.bfloat16).float()
        ref.backward(dy)
        w4k = w4.permute(0, 2, 3, 1).contiguous().to(DEV).to(torch.bfloat16)            # KRSC
        xdd = nhwc(x).to(DEV).to(torch.bfloat16)
        dyd = nhwc(dy).to(DEV).to(torch.bfloat16)
        if H >= 16 or (H // 2) * (H // 2) * N >= 128:
            y = ops.conv_fwd_ex(CONV_S2K4, xdd, w4k, torch.float32)
            rows.append((f"conv4x4s2 fwd {tag}", rel(nchw(y), ref), F32_TOL))
            dw4, _ = ops.conv_wgrad_ex(CONV_S2K4, xdd, dyd, want_bias=False)             # [Cout,4,4,Cin]
            rows.append((f"conv4x4s2 wgrad {tag}", rel(dw4.permute(0, 3, 1, 2), w4r.grad), 5e-5))
            # data gradient = phase conv with W'[ph,pw][cin][a,b][cout] = w4[cout, cin, 3-2a-ph, 3-2b-pw]
            idx = torch.tensor([[3, 1], [2, 0]])
            wt = w4.permute(1, 0, 2, 3)                                                   # [Cin, Cout, 4, 4]
            wph = torch.stack([wt[:, :, idx[ph]][:, :, :, idx[pw]] for ph in range(2) for pw in range(2)])
            wph = wph.permute(0, 1, 3, 4, 2).contiguous().to(DEV).to(torch.bfloat16)      # [4, Cin, 2, 2, Cout]
            dx = ops.conv_fwd_ex(CONV_PHASE2, dyd, wph, torch.float32)
            rows.append((f"conv4x4s2 dgrad (phase conv) {tag}", rel(nchw(dx), xr.grad), F32_TOL))
        # --- ConvTranspose2d(Cin, Cout, 4, 2, 1) forward (mcvae.py:87-92) as a phase conv
        wt4 = (torch.randn(Cin, Cout, 4, 4, generator=g) / (Cin * 4) ** 0.5).to(torch.bfloat16).float()
        ref = F.conv_transpose2d(x, wt4, None, stride=2, padding=1)
        idx = torch.tensor([[3, 1], [2, 0]])
        wph = torch.stack([wt4[:, :, idx[ph]][:, :, :, idx[pw]] for ph in range(2) for pw in range(2)])  # [4,Cin,Cout,2,2]
        wph = wph.permute(0, 2, 3, 4, 1).contiguous().to(DEV).to(torch.bfloat16)          # [4, Cout, 2, 2, Cin]
        y = ops.conv_fwd_ex(CONV_PHASE2, xdd, wph, torch.float32)
        rows.append((f"convT4x4s2 fwd (phase conv) {tag}", rel(nchw(y), ref), F32_TOL))
    return rows


def _tf32(t):
    """Round-toward-zero to the 10-bit TF32 mantissa (exactly representable for the tensor cores)."""
    return (t.contiguous().view(torch.int32) & -8192).view(torch.float32)


def check_conv_tf32():
    """fp32-operand (TF32 tensor core) variants of the conv kernels — the fp32 parity pipeline.
    On TF32-representable inputs only the summation order differs (2e-5); on arbitrary fp32 inputs the
    operands lose 13 mantissa bits (1e-3, the north-star fp32/tf32 tolerance)."""
    from mcgm_b200._lib import CONV_PHASE2, CONV_S1, CONV_S2K4
    rows = []
    for (N, H, Cin, Cout, R) in [(4, 8, 64, 64, 3), (2, 16, 128, 128, 3), (8, 4, 64, 256, 1), (40, 16, 64, 128, 3)]:
        g = torch.Generator().manual_seed(N + Cin + R)
        for exact in (True, False):
            x = torch.randn(N, Cin, H, H, generator=g)
            w = torch.randn(Cout, Cin, R, R, generator=g) / (Cin * R * R) ** 0.5
            dy = torch.randn(N, Cout, H, H, generator=g)
            if exact:
                x, w, dy = _tf32(x), _tf32(w), _tf32(dy)
            tol = F32_TOL if exact else 1e-3
            b = torch.randn(Cout, generator=g)
            xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
            ref = F.conv2d(xr, wr, br, padding=(R - 1) // 2)
            ref.backward(dy)
            tag = f"{'tf32-exact' if exact else 'fp32'} N{N} {H}x{H} {Cin}->{Cout} k{R}"
            xd, dyd = nhwc(x).to(DEV), nhwc(dy).to(DEV)
            wk = w.permute(0, 2, 3, 1).contiguous().to(DEV)
            y = ops.conv_fwd_tf32(CONV_S1, xd, wk, R, R, b.to(DEV))
            rows.append((f"tf32 conv fwd {tag}", rel(nchw(y), ref), tol))
            wflip = w.flip(2, 3).permute(1, 2, 3, 0).contiguous().to(DEV)
            dx = ops.conv_fwd_tf32(CONV_S1, dyd, wflip, R, R)
            rows.append((f"tf32 conv dgrad {tag}", rel(nchw(dx), xr.grad), tol))
            if not exact:
                # fp32-accurate variants of the parity pipeline: 3xTF32 forward, bf16x3 weight gradient
                y3 = ops.conv_f32(CONV_S1, xd, wk, R, R, b.to(DEV))
                rows.append((f"3xTF32 conv fwd {tag}", rel(nchw(y3), ref), F32_TOL))
                dw, db = ops.wgrad_f32(CONV_S1, xd, dyd, R, R)
                rows.append((f"bf16x3 conv wgrad {tag}", rel(dw.permute(0, 3, 1, 2), wr.grad), F32_TOL))
                rows.append((f"bf16x3 conv dbias {tag}", rel(db, br.grad), F32_TOL))
    # 4x4 stride-2 and its adjoint in TF32
    g = torch.Generator().manual_seed(77)
    N, H, Cin, Cout = 8, 16, 64, 128
    x, w4 = _tf32(torch.randn(N, Cin, H, H, generator=g)), _tf32(torch.randn(Cout, Cin, 4, 4, generator=g) * 0.05)
    dy = _tf32(torch.randn(N, Cout, H // 2, H // 2, generator=g))
    xr, wr = x.clone().requires_grad_(True), w4.clone().requires_grad_(True)
    ref = F.conv2d(xr, wr, None, stride=2, padding=1)
    ref.backward(dy)
    y = ops.conv_fwd_tf32(CONV_S2K4, nhwc(x).to(DEV), w4.permute(0, 2, 3, 1).contiguous().to(DEV))
    rows.append(("tf32 conv4x4s2 fwd", rel(nchw(y), ref), F32_TOL))
    wph = ops._phase_weights(w4.permute(1, 0, 2, 3).to(DEV))
    dx = ops.conv_fwd_tf32(CONV_PHASE2, nhwc(dy).to(DEV), wph)
    rows.append(("tf32 conv4x4s2 dgrad (phase conv)", rel(nchw(dx), xr.grad), F32_TOL))
    return rows


class _SNLayer:
    def __init__(self, w, u, v):
        self.weight_orig, self.weight_u, self.weight_v = w, u, v


def check_sn():
    """Batched spectral-norm power iteration / W/sigma packing / backward vs the oracle (torch CPU)."""
    rows = []
    g = torch.Generator().manual_seed(3)
    shapes = [(128, 128, 3, 3), (128, 3, 3, 3), (128, 128, 1, 1), (1, 128, 1, 1), (64, 64, 3, 3)]
    ws = [torch.randn(*sh, generator=g) * 0.1 for sh in shapes]
    us = [F.normalize(torch.randn(sh[0], generator=g), dim=0) for sh in shapes]
    vs = [F.normalize(torch.randn(sh[1] * sh[2] * sh[3], generator=g), dim=0) for sh in shapes]
    layers = [_SNLayer(w.to(DEV), u.to(DEV), v.to(DEV)) for w, u, v in zip(ws, us, vs)]
    batch = ops.SNBatch(layers)
    for it in range(2):                      # two consecutive training forwards (u, v carried in place)
        states = batch.run(True)
        for i, (sh, L) in enumerate(zip(shapes, layers)):
            sd = {"weight_orig": ws[i].clone().requires_grad_(True), "weight_u": us[i], "weight_v": vs[i]}
            wbar = O.sn_weight(sd, "", True)          # updates us[i], vs[i] in place like the reference
            st = states[L]
            K, CRS = sh[0], sh[1] * sh[2] * sh[3]
            tag = f"iter{it} {sh}"
            rows.append((f"sn u {tag}", rel(st[1:1 + K], us[i]), F32_TOL))
            rows.append((f"sn v {tag}", rel(st[1 + K:1 + K + CRS], vs[i]), F32_TOL))
            rows.append((f"sn buffers in place {tag}", rel(L.weight_u, us[i]) + rel(L.weight_v, vs[i]), 2 * F32_TOL))
            rows.append((f"sn wbar {tag}", rel(L.weight_orig * st[0], wbar.detach()), F32_TOL))
            dwbar = torch.randn(*sh, generator=g)
            wbar.backward(dwbar)
            # backward from a KRSC-ordered gradient (conv_wgrad layout) and from an OIHW-ordered one
            dk = dwbar.permute(0, 2, 3, 1).contiguous().to(DEV)
            got = ops.sn_weight_bwd(dk, CRS, 1, L.weight_orig, st)
            rows.append((f"sn bwd krsc {tag}", rel(got, sd["weight_orig"].grad), 1e-4))
            got = ops.sn_weight_bwd(dwbar.to(DEV).contiguous(), CRS, 0, L.weight_orig, st)
            rows.append((f"sn bwd oihw {tag}", rel(got, sd["weight_orig"].grad), 1e-4))
            if sh[1] <= 4:
                pk = ops.pack_weight_padded(L.weight_orig, st[0:1], 64).float().cpu().reshape(K, 64)
                want = torch.zeros(K, 64)
                want[:, :CRS] = wbar.detach().permute(0, 2, 3, 1).reshape(K, CRS)
                rows.append((f"pack_weight_padded {tag}", rel(pk, want.to(torch.bfloat16).float()), 1e-6))
    # eval mode: no power iteration, sigma from the stored u, v
    states = batch.run(False)
    sd = {"weight_orig": ws[0], "weight_u": us[0], "weight_v": vs[0]}
    rows.append(("sn eval wbar", rel(layers[0].weight_orig * states[layers[0]][0], O.sn_weight(sd, "", False)), F32_TOL))
    return rows


def check_conv_variants():
    """The same parity rows with the alternative tilings forced on / off through mcgm_debug_set."""
    rows = []
    big = [sh for sh in CONV_SHAPES if sh[0] >= 40]
    for key, val, tag in [(7, 1, "BN256/MT2"), (6, 0, "MT1 only"), (8, 0, "NT1 only"), (17, 1, "wgrad blocked 5-D maps"),
                          (16, 0, "no CTA pairs")]:
        _lib.call("mcgm_debug_set", key, val)
        try:
            if key != 17:
                rows += [(f"[{tag}] {n}", e, t) for (n, e, t) in check_conv_fwd(big) if n.startswith("conv")]
            if key != 16:
                rows += [(f"[{tag}] {n}", e, t) for (n, e, t) in check_conv_wgrad(big)]
        finally:
            _lib.call("mcgm_debug_set", key, {7: 0, 6: 1, 8: 1, 17: 0, 16: 3}[key])
    return rows


def check_conv_cluster():
    """CTA pairs (cta_group::2 MMAs, mcgm_debug_set key 16 — the default for the 256-wide and the 128-wide pair tiles) and
    thread-block clusters with multicast weights (key 11): large-M forward convs — the shapes whose launch takes those
    paths — must be BIT-IDENTICAL with one CTA per tile, CTA pairs, and clusters of 2 and 4 CTAs (the arithmetic is the
    same, only the operand delivery and the MMA's M extent differ), including odd tile counts (a dummy tile for the last cluster), every
    fused epilogue and the epilogue statistics; the cluster-of-2 result is also held against torch's fp32 conv."""
    from mcgm_b200._lib import CONV_PHASE2, CONV_S1, EPI_BIAS
    rows = []
    M = 10
    cases = [("s1", 80, 32, 64, 256, 3, 0), ("s1", 75, 32, 64, 256, 3, EPI_RES), ("s1", 77, 32, 128, 256, 1, EPI_RES_UP2),
             ("s1", 160, 32, 64, 128, 3, EPI_RES), ("s1", 151, 32, 64, 128, 3, EPI_POSMASK), ("s1", 160, 32, 128, 128, 3, EPI_RELU | EPI_MASK),
             ("ph", 150, 16, 64, 256, 0, 0), ("ph", 321, 8, 128, 128, 0, 0)]
    for (kind, N, H, Cin, Cout, R, fl) in cases:
        g = torch.Generator().manual_seed(100 + N)
        x = torch.randn(N, H, H, Cin, generator=g).to(DEV).to(torch.bfloat16)
        bias = torch.randn(Cout, generator=g).to(DEV)
        Ho = H if kind == "s1" else 2 * H
        res = aux = labels = packed = None
        if fl & EPI_RES:
            res = torch.randn(N, Ho, Ho, Cout, generator=g).to(DEV).to(torch.bfloat16)
        if fl & EPI_RES_UP2:
            res = torch.randn(N, Ho // 2, Ho // 2, Cout, generator=g).to(DEV).to(torch.bfloat16)
        if fl & EPI_POSMASK:
            aux = torch.randn(N, Ho, Ho, Cout, generator=g).to(DEV).to(torch.bfloat16)
        if fl & EPI_MASK:
            packed = ops.PackedCodebook(_codebook(M, Cout, 3).to(DEV))
            labels = torch.randint(0, M, (N,), generator=g).to(DEV).to(torch.int32)
        if kind == "s1":
            w = (torch.randn(Cout, R, R, Cin, generator=g) * 0.05).to(DEV).to(torch.bfloat16)
            run = lambda: ops.conv_fwd(x, w, torch.bfloat16, bias, res, aux, labels, packed, fl)       # noqa: E731
            run_st = (lambda: ops.conv_fwd_bnstats(CONV_S1, x, w, bias, res, fl)) if not (fl & (EPI_POSMASK | EPI_MASK)) else None
        else:
            w = (torch.randn(4, Cout, 2, 2, Cin, generator=g) * 0.05).to(DEV).to(torch.bfloat16)
            run = lambda: ops.conv_fwd_ex(CONV_PHASE2, x, w, torch.bfloat16, bias)                     # noqa: E731
            run_st = lambda: ops.conv_fwd_bnstats(CONV_PHASE2, x, w, bias)                             # noqa: E731
        outs, stats = {}, {}
        # 1 = one CTA per tile (CTA pairs off), "pair" = cta_group::2 MMAs (key 16, the default), 2 / 4 = multicast clusters
        for cl in (1, "pair", 2, 4):
            _lib.call("mcgm_debug_set", 11, cl if isinstance(cl, int) else 1)
            _lib.call("mcgm_debug_set", 16, 3 if cl == "pair" else 0)
            try:
                outs[cl] = run()
                if run_st is not None:
                    o2, st = run_st()
                    stats[cl] = ops.bn_finalize(st, Cout)
                    rows.append((f"cluster{cl} {kind} N={N} {Cin}->{Cout} f{fl:#x}: bnstats output identical", float((o2.float() != outs[cl].float()).sum()), EXACT))
            finally:
                _lib.call("mcgm_debug_set", 11, 1)
                _lib.call("mcgm_debug_set", 16, 3)
        tag = f"{kind} N={N} H={H} {Cin}->{Cout} k{R} f{fl:#x}"
        for cl in ("pair", 2, 4):
            rows.append((f"cluster{cl} == cluster1 {tag}", float((outs[cl].float() != outs[1].float()).sum()), EXACT))
            if stats:
                rows.append((f"cluster{cl} stats mean {tag}", rel(stats[cl][0], stats[1][0]), F32_TOL))
                rows.append((f"cluster{cl} stats rstd {tag}", rel(stats[cl][1], stats[1][1]), F32_TOL))
        if kind == "s1" and not (fl & (EPI_POSMASK | EPI_MASK | EPI_RES_UP2)):
            ref = F.conv2d(nchw(x.float()), w.float().permute(0, 3, 1, 2), bias, padding=(R - 1) // 2)
            if res is not None:
                ref = ref + nchw(res.float())
            rows.append((f"cluster2 vs torch fp32 conv {tag}", rel(nchw(outs[2].float()), ref), BF16_OUT))
        if stats:
            mref, rref = ops.bn_stats(outs[2])
            rows.append((f"cluster2 stats vs statistics pass {tag}", max(rel(stats[2][0], mref), rel(stats[2][1], rref)), F32_TOL))
    return rows


ALL_KERNEL_CHECKS = [check_codebook_and_mask, check_gate, check_pool, check_conv_fwd, check_conv_epilogues,
                     check_conv_wgrad, check_thin, check_conv_variants, check_sn, check_conv_ex, check_conv_tf32]


# ----------------------------------------------------------------------------- optimiser
def check_optim():
    """optim.FusedAdam (mcgm_adam_step / mcgm_grad_clip_coef) against torch.optim.Adam
    (+ clip_grad_norm_) on identical parameters and gradients, several steps.  fp32 elementwise
    arithmetic in a different association order: agreement to a few ulp of the update."""
    from mcgm_b200.optim import FusedAdam
    rows = []
    for (clip, wd, betas) in [(None, 0.0, (0.0, 0.9)), (1.0, 0.0, (0.9, 0.999)), (0.05, 1e-2, (0.5, 0.99))]:
        g = torch.Generator().manual_seed(7)
        shapes = [(64, 32, 3, 3), (64,), (130, 7), (5,), (256, 128, 1, 1)]
        ref = [torch.nn.Parameter(torch.randn(*s, generator=g).to(DEV)) for s in shapes]
        got = [torch.nn.Parameter(p.detach().clone()) for p in ref]
        o_ref = torch.optim.Adam(ref, lr=2e-3, betas=betas, weight_decay=wd)
        o_got = FusedAdam(got, lr=2e-3, betas=betas, weight_decay=wd, clip_norm=clip)
        for it in range(5):
            grads = [torch.randn(*s, generator=g).to(DEV) * (10.0 if it == 2 else 0.3) for s in shapes]
            o_ref.zero_grad()
            o_got.zero_grad()
            for p, q, gr in zip(ref, got, grads):
                p.grad = gr.clone()
                q.grad.add_(gr)                       # accumulates into the flat-buffer view, like autograd
            if clip is not None:
                tn = torch.nn.utils.clip_grad_norm_(ref, clip)
            o_ref.step()
            o_got.step()
            if clip is not None:
                rows.append((f"adam clip={clip} step {it}: total_norm", rel(o_got.total_norm, tn), F32_TOL))
        for i, (p, q) in enumerate(zip(ref, got)):
            rows.append((f"adam clip={clip} wd={wd} betas={betas}: param {i} after 5 steps", rel(q, p), F32_TOL))
        rows.append((f"adam clip={clip}: device step counter", float(abs(int(o_got.step_dev.item()) - 5)), EXACT))
        rows.append((f"adam clip={clip}: params are views of the flat buffer",
                     float(got[0].data_ptr() != o_got.flat.param.data_ptr()), EXACT))
    return rows


# ----------------------------------------------------------------------------- Glow flow kernels
def check_flow():
    """csrc/flow.cu through ops.FlowFrontFn / ops.FlowTailFn against the reference formulation
    (ActNorm -> 1x1 conv; ZeroConv scale -> sigmoid -> affine coupling -> logdet) in torch fp32 on the
    CPU, values and every gradient.  fp32 arithmetic, different summation order: F32 tolerance; the bf16
    network input is compared after the same rounding."""
    rows = []
    for (N, H, C, affine) in [(5, 16, 12, True), (3, 8, 24, True), (4, 4, 48, True), (2, 8, 4, True), (3, 4, 24, False)]:
        g = torch.Generator().manual_seed(100 + C + H)
        x = torch.randn(N, H, H, C, generator=g)
        loc, scale = torch.randn(1, C, 1, 1, generator=g) * 0.3, torch.rand(1, C, 1, 1, generator=g) + 0.5
        Wm = torch.randn(C, C, generator=g) / C ** 0.5 + torch.eye(C)
        dy, dxa = torch.randn(N, H, H, C, generator=g), torch.randn(N, H, H, 64, generator=g).to(torch.bfloat16)
        # reference (NCHW, as the modules compute it)
        xr, lr_, sr, Wr = [t.clone().requires_grad_(True) for t in (x, loc, scale, Wm)]
        a = sr * (xr.permute(0, 3, 1, 2) + lr_)
        yr = F.conv2d(a, Wr[:, :, None, None]).permute(0, 2, 3, 1)
        (yr * dy).sum().backward(retain_graph=True)
        (yr[..., :C // 2] * dxa[..., :C // 2].float()).sum().backward()
        xd, ld, sd_, Wd = [t.to(DEV).requires_grad_(True) for t in (x, loc, scale, Wm)]
        y, xa = ops.FlowFrontFn.apply(xd, ld, sd_, Wd)
        torch.autograd.backward([y, xa], [dy.to(DEV), dxa.to(DEV)])
        tag = f"flow front N={N} H={H} C={C}"
        rows.append((tag + " y", rel(y, yr), F32_TOL))
        want_xa = torch.zeros(N, H, H, 64)
        want_xa[..., :C // 2] = yr[..., :C // 2].detach()
        rows.append((tag + " xa (bf16, zero padded)", rel(xa.float(), want_xa.to(torch.bfloat16).float()), BF16_OUT))
        rows.append((tag + " xa padding is zero", float(xa[..., C // 2:].float().abs().max()), EXACT))
        for nm, got, ref in (("dx", xd.grad, xr.grad), ("dloc", ld.grad, lr_.grad), ("dscale", sd_.grad, sr.grad),
                             ("dW", Wd.grad, Wr.grad)):
            rows.append((f"{tag} {nm}", rel(got, ref), F32_TOL))
        # tail
        yv = torch.randn(N, H, H, C, generator=g)
        nout = C if affine else C // 2
        net = torch.zeros(N, H, H, 64)
        net[..., :nout] = torch.randn(N, H, H, nout, generator=g)
        zs = torch.randn(1, nout, 1, 1, generator=g) * 0.2
        dout, dld = torch.randn(N, H, H, C, generator=g), torch.randn(N, generator=g)
        yr2, nr, zr = [t.clone().requires_grad_(True) for t in (yv, net, zs)]
        hcl = nr[..., :nout] * torch.exp(zr.reshape(nout) * 3)
        in_a, in_b = yr2[..., :C // 2], yr2[..., C // 2:]
        if affine:
            s = torch.sigmoid(hcl[..., :C // 2] + 2)
            outr = torch.cat([in_a, (in_b + hcl[..., C // 2:]) * s], -1)
            ldr = torch.log(s).reshape(N, -1).sum(1)
            ((outr * dout).sum() + (ldr * dld).sum()).backward()
        else:
            outr = torch.cat([in_a, in_b + hcl], -1)
            (outr * dout).sum().backward()
        yd2, nd, zd = [t.to(DEV).requires_grad_(True) for t in (yv, net, zs)]
        out, ldg = ops.FlowTailFn.apply(yd2, nd, zd, affine)
        if affine:
            torch.autograd.backward([out, ldg], [dout.to(DEV), dld.to(DEV)])
            rows.append((f"flow tail C={C} logdet", rel(ldg, ldr), F32_TOL))
        else:
            out.backward(dout.to(DEV))
        tag = f"flow tail N={N} H={H} C={C} affine={affine}"
        rows.append((tag + " out", rel(out, outr), F32_TOL))
        rows.append((tag + " dy", rel(yd2.grad, yr2.grad), F32_TOL))
        rows.append((tag + " dnet", rel(nd.grad, nr.grad), F32_TOL))
        rows.append((tag + " dzscale", rel(zd.grad, zr.grad), F32_TOL))
    return rows


# ----------------------------------------------------------------------------- BN statistics in the conv epilogue
def check_conv_bnstats():
    """mcgm_conv_fwd_bnstats: the output must be bit-identical to the plain conv's, and the finalised
    epilogue partials must equal the statistics pass over that output (mcgm_bn_stats) — same values
    (the bf16-rounded stores), different summation order."""
    from mcgm_b200._lib import CONV_PHASE2, CONV_S1, EPI_BIAS
    rows = []
    for (kind, N, H, Cin, res_up, Co) in [(CONV_S1, 4, 16, 256, True, 256), (CONV_S1, 5, 8, 256, False, 256), (CONV_S1, 3, 32, 128, True, 256),
                                          (CONV_PHASE2, 4, 8, 256, False, 256), (CONV_PHASE2, 5, 4, 256, False, 256),
                                          (CONV_S1, 37, 16, 128, False, 128), (CONV_S1, 150, 8, 64, True, 64), (CONV_PHASE2, 80, 8, 128, False, 128),
                                          (CONV_S1, 300, 8, 128, False, 128), (CONV_S1, 3, 4, 64, False, 64)]:
        g = torch.Generator().manual_seed(31 + H + N)
        x = torch.randn(N, H, H, Cin, generator=g).to(DEV).to(torch.bfloat16)
        bias = torch.randn(Co, generator=g).to(DEV)
        if kind == CONV_S1:
            w = (torch.randn(Co, 3, 3, Cin, generator=g) * 0.05).to(DEV).to(torch.bfloat16)
            res = torch.randn(N, H // 2, H // 2, Co, generator=g).to(DEV).to(torch.bfloat16) if res_up else None
            fl = EPI_RES_UP2 if res_up else 0
            ref = ops.conv_fwd(x, w, torch.bfloat16, bias, res, None, None, None, fl)
            out, st = ops.conv_fwd_bnstats(kind, x, w, bias, res, fl)
        else:
            w = (torch.randn(4, Co, 2, 2, Cin, generator=g) * 0.05).to(DEV).to(torch.bfloat16)
            ref = ops.conv_fwd_ex(kind, x, w, torch.bfloat16, bias)
            out, st = ops.conv_fwd_bnstats(kind, x, w, bias)
        tag = f"conv+bnstats kind={kind} N={N} H={H} Cin={Cin} Cout={Co}"
        rows.append((tag + " output identical", float((out.float() != ref.float()).sum()), EXACT))
        rm0, rv0 = torch.zeros(Co, device=DEV), torch.ones(Co, device=DEV)
        rm1, rv1 = rm0.clone(), rv0.clone()
        mean, rstd = ops.bn_finalize(st, Co, rm0, rv0, 0.1, 1e-5)
        mref, rref = ops.bn_stats(ref, rm1, rv1, 0.1, 1e-5)
        rows.append((tag + " count", float(abs(st.count - ref.numel() // Co)), EXACT))
        rows.append((tag + " mean", rel(mean, mref), F32_TOL))
        rows.append((tag + " rstd", rel(rstd, rref), F32_TOL))
        rows.append((tag + " running_mean", rel(rm0, rm1), F32_TOL))
        rows.append((tag + " running_var", rel(rv0, rv1), F32_TOL))
    return rows


# ----------------------------------------------------------------------------- batched / grouped launch forms
def check_batched_paths():
    """The batched forms against their per-layer forms on identical inputs:
    mcgm_conv_wgrad_grouped vs mcgm_conv_wgrad (different split-K factor: F32 tolerance),
    mcgm_sn_weight_bwd_batched vs mcgm_sn_weight_bwd, mcgm_pack_weight_batched vs mcgm_pack_weight(_padded)."""
    import ctypes as c
    rows = []
    g = torch.Generator().manual_seed(5)
    # grouped weight gradients: 3 same-shape problems
    for (N, H, Cin, Cout, R) in [(32, 8, 128, 128, 3), (16, 16, 128, 128, 1)]:
        xs = [torch.randn(N, H, H, Cin, generator=g).to(DEV).to(torch.bfloat16) for _ in range(3)]
        dys = [torch.randn(N, H, H, Cout, generator=g).to(DEV).to(torch.bfloat16) for _ in range(3)]
        ref = [ops.conv_wgrad(x, dy, R, R, want_bias=True) for x, dy in zip(xs, dys)]
        n = 3
        dw = torch.empty(n, Cout, R, R, Cin, device=DEV)
        db = torch.empty(n, Cout, device=DEV)
        ws = torch.empty(_lib.query("mcgm_conv_wgrad_grouped_workspace_bytes", n, N, H, H, Cin, Cout, R, R) // 4, device=DEV)
        _lib.call("mcgm_conv_wgrad_grouped", n, (c.c_void_p * n)(*[t.data_ptr() for t in xs]),
                  (c.c_void_p * n)(*[t.data_ptr() for t in dys]), (c.c_void_p * n)(*[dw[i].data_ptr() for i in range(n)]),
                  (c.c_void_p * n)(*[db[i].data_ptr() for i in range(n)]), N, H, H, Cin, Cout, R, R, ws.data_ptr(),
                  torch.cuda.current_stream().cuda_stream)
        for i in range(n):
            rows.append((f"wgrad_grouped {N}x{H}x{H} {Cin}->{Cout} k{R} group {i}: dw", rel(dw[i], ref[i][0]), F32_TOL))
            rows.append((f"wgrad_grouped {N}x{H}x{H} {Cin}->{Cout} k{R} group {i}: dbias", rel(db[i], ref[i][1]), F32_TOL))
    # spectral-norm backward, batched (two items share one gradient) vs per layer
    class _M:            # the minimum of a spectral-normed module that ops.SNBatch / sn_pack_all look at
        pass
    layers = []
    for (K, C, R) in [(128, 128, 3), (128, 128, 1), (128, 3, 3)]:
        w = torch.nn.Parameter((torch.randn(K, C, R, R, generator=g) * 0.1).to(DEV))
        state = torch.cat([torch.rand(1, generator=g) + 0.5, torch.randn(K, generator=g), torch.randn(C * R * R, generator=g)]).to(DEV)
        layers.append((w, state))
    ops.SN_DEFER = True
    try:
        want = []
        items = []
        for (w, state) in layers:
            K, C, R, _ = w.shape
            for rep in range(2):
                dwbar = torch.randn(K, R, R, C, generator=g).to(DEV)
                items.append((dwbar, w, state))
        ops.SN_DEFER = False
        for (w, state) in layers:
            acc = torch.zeros_like(w)
            for (dwbar, w2, st) in items:
                if w2 is w:
                    acc += ops.sn_weight_bwd(dwbar, dwbar.shape[1] * dwbar.shape[2] * dwbar.shape[3], 1, w, st)
            want.append(acc)
        for (w, _) in layers:
            w.grad = None
        for (dwbar, w, st) in items:
            ops._sn_pending.append((dwbar, dwbar.shape[1] * dwbar.shape[2] * dwbar.shape[3], 1, w, st))
        ops._sn_flush()
        for (w, _), ref_g in zip(layers, want):
            rows.append((f"sn_weight_bwd_batched {tuple(w.shape)} (2 items, grad +=)", rel(w.grad, ref_g), F32_TOL))
    finally:
        ops.SN_DEFER = False
    # batched packing
    mods, states = [], {}
    for (w, state) in layers:
        m = _M()
        m.weight_orig = w
        mods.append(m)
        states[m] = state
    packs = ops.sn_pack_all(states)
    for m in mods:
        w, st = m.weight_orig, states[m]
        if w.shape[1] <= 4:
            ref_p = ops.pack_weight_padded(w.detach(), st[0:1], packs[m].shape[-1])
            rows.append((f"pack_weight_batched padded {tuple(w.shape)}", float((packs[m].float() != ref_p.float()).sum()), EXACT))
        else:
            rf, rb = ops.pack_weight(w.detach(), st[0:1])
            rows.append((f"pack_weight_batched KRSC {tuple(w.shape)}", float((packs[m][0].float() != rf.float()).sum()), EXACT))
            rows.append((f"pack_weight_batched flipped {tuple(w.shape)}", float((packs[m][1].float() != rb.float()).sum()), EXACT))
    return rows
