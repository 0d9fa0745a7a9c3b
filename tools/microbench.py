"""Kernel microbenchmarks (CUDA events, L2 flushed between launches): conv fwd / wgrad / gate."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402
entry.build()
from mcgm_b200 import ops  # noqa: E402
from mcgm_b200._lib import GATE_BN, GATE_MASK, GATE_RELU, GATE_UP2  # noqa: E402

from mcgm_b200 import _lib  # noqa: E402
_lib.load()                  # applies MCGM_DEBUG_SET
print("debug_set", os.environ.get("MCGM_DEBUG_SET", ""))
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


which = sys.argv[1:] or ["conv", "wgrad", "gate"]
N = 512
if "conv" in which:
    for (H, Cin, Cout, R) in [(32, 256, 256, 3), (16, 256, 256, 3), (8, 256, 256, 3), (32, 128, 128, 3), (16, 128, 128, 3),
                              (8, 128, 128, 3), (16, 256, 256, 1), (16, 128, 128, 1), (32, 64, 128, 1), (32, 256, 32, 1)]:
        x = torch.randn(N, H, H, Cin, device=dev).to(torch.bfloat16)
        w = torch.randn(Cout, R, R, Cin, device=dev).to(torch.bfloat16)
        t = timeit(lambda: ops.conv_fwd(x, w, torch.bfloat16 if Cout != 32 else torch.float32))
        fl = 2.0 * N * H * H * Cin * Cout * R * R
        print(f"conv_fwd N{N} {H}x{H} {Cin}->{Cout} k{R}: {t * 1e6:8.1f} us  {fl / t / 1e12:7.1f} TFLOP/s", flush=True)
if "epi" in which:
    from mcgm_b200._lib import EPI_BIAS, EPI_MASK, EPI_POSMASK, EPI_RELU, EPI_RES, EPI_RES_UP2
    M = 10
    labels = torch.randint(0, M, (N,), device=dev).to(torch.int32)
    for (H, Cin, Cout, R, fl) in [(32, 128, 128, 3, 0x3), (32, 128, 128, 3, 0x40), (32, 64, 128, 1, 0x19), (32, 64, 128, 1, 0x1),
                                  (16, 128, 128, 3, 0x3), (16, 128, 128, 3, 0x19), (8, 128, 128, 3, 0x3), (32, 256, 256, 3, 0x5),
                                  (16, 128, 128, 1, 0x1), (16, 256, 256, 1, 0x1)]:
        x = torch.randn(N, H, H, Cin, device=dev).to(torch.bfloat16)
        w = torch.randn(Cout, R, R, Cin, device=dev).to(torch.bfloat16)
        bias = torch.randn(Cout, device=dev) if fl & EPI_BIAS else None
        res = torch.randn(N, H // 2 if fl & EPI_RES_UP2 else H, H // 2 if fl & EPI_RES_UP2 else H, Cout, device=dev).to(torch.bfloat16) \
            if fl & (EPI_RES | EPI_RES_UP2) else None
        aux = torch.randn(N, H, H, Cout, device=dev).to(torch.bfloat16) if fl & EPI_POSMASK else None
        packed = ops.PackedCodebook((torch.rand(M, Cout) > 0.5).float().to(dev)) if fl & EPI_MASK else None
        t = timeit(lambda: ops.conv_fwd(x, w, torch.bfloat16, bias, res, aux, labels if packed else None, packed, fl))
        fl_ = 2.0 * N * H * H * Cin * Cout * R * R
        by = (x.numel() + N * H * H * Cout + (res.numel() if res is not None else 0) + (aux.numel() if aux is not None else 0)) * 2
        print(f"conv_epi N{N} {H}x{H} {Cin}->{Cout} k{R} f{fl:#x}: {t * 1e6:8.1f} us  {fl_ / t / 1e12:7.1f} TFLOP/s  {by / t / 1e9:7.0f} GB/s", flush=True)
if "thin" in which:
    M = 10
    labels = torch.randint(0, M, (N,), device=dev).to(torch.int32)
    packed = ops.PackedCodebook((torch.rand(M, 128) > 0.5).float().to(dev))
    img = torch.randn(N, 32, 32, 3, device=dev)
    for R in (3, 1):
        w = torch.nn.Parameter(torch.randn(128, 3, R, R, device=dev) * 0.1)
        b = torch.nn.Parameter(torch.zeros(128, device=dev))
        gate = (labels, packed) if R == 3 else None
        t = timeit(lambda: ops.WideFromThinFn.apply(img, w, b, gate, True, None))
        print(f"thin_conv fwd k{R}: {t * 1e6:7.1f} us  {(img.numel() * 4 + N * 1024 * 128 * 2) / t / 1e9:7.0f} GB/s", flush=True)
        out = ops.WideFromThinFn.apply(img, w, b, gate, True, None)
        dout = torch.randn_like(out)
        def bw():
            w.grad = None; b.grad = None
            out.backward(dout, retain_graph=True)
        t = timeit(bw)
        print(f"thin_conv wgrad k{R}: {t * 1e6:7.1f} us  {(img.numel() * 4 + N * 1024 * 128 * 2) / t / 1e9:7.0f} GB/s", flush=True)
if "wgrad" in which:
    for (H, Cin, Cout, R) in [(32, 256, 256, 3), (16, 256, 256, 3), (8, 256, 256, 3), (32, 128, 128, 3), (16, 128, 128, 3),
                              (8, 128, 128, 3), (16, 256, 256, 1), (32, 64, 128, 1), (32, 256, 64, 1)]:
        x = torch.randn(N, H, H, Cin, device=dev).to(torch.bfloat16)
        dy = torch.randn(N, H, H, Cout, device=dev).to(torch.bfloat16)
        t = timeit(lambda: ops.conv_wgrad(x, dy, R, R, want_bias=False))
        fl = 2.0 * N * H * H * Cin * Cout * R * R
        print(f"conv_wgrad N{N} {H}x{H} {Cin}->{Cout} k{R}: {t * 1e6:8.1f} us  {fl / t / 1e12:7.1f} TFLOP/s", flush=True)
if "gate" in which:
    M = 10
    cb = (torch.rand(M, 256) > 0.5).float().to(dev)
    packed = ops.PackedCodebook(cb)
    labels = torch.randint(0, M, (N,), device=dev).to(torch.int32)
    gamma, beta = torch.ones(256, device=dev), torch.zeros(256, device=dev)
    for (H, up) in [(32, False), (16, True), (16, False), (8, True)]:
        x = torch.randn(N, H, H, 256, device=dev).to(torch.bfloat16)
        T = x.numel() * 2
        t = timeit(lambda: ops.bn_stats(x))
        print(f"bn_stats {H}x{H}: {t * 1e6:7.1f} us {T / t / 1e9:7.0f} GB/s", flush=True)
        mean, rstd = ops.bn_stats(x)
        fl = GATE_BN | GATE_RELU | GATE_MASK | (GATE_UP2 if up else 0)
        t = timeit(lambda: ops.gate_fwd(x, fl, labels, packed, mean, rstd, gamma, beta, up))
        by = T * ((1 + 4 + 1) if up else 2)
        print(f"gate_fwd {H}x{H} up{int(up)} raw{int(up)}: {t * 1e6:7.1f} us {by / t / 1e9:7.0f} GB/s", flush=True)
        f = 2 if up else 1
        dy = torch.randn(N, H * f, H * f, 256, device=dev).to(torch.bfloat16)
        draw = torch.randn_like(x) if up else None
        t = timeit(lambda: ops.gate_bwd(dy, draw, x, fl, labels, packed, mean, rstd, gamma, beta))
        by = 2 * (dy.numel() * 2 + T) + T + (T if up else 0)
        print(f"gate_bwd {H}x{H} up{int(up)} (reduce+apply): {t * 1e6:7.1f} us {by / t / 1e9:7.0f} GB/s", flush=True)
    x = torch.randn(N, 16, 16, 128, device=dev).to(torch.bfloat16)
    cb = (torch.rand(M, 128) > 0.5).float().to(dev)
    packed = ops.PackedCodebook(cb)
    t = timeit(lambda: ops.gate_fwd(x, GATE_RELU | GATE_MASK, labels, packed, None, None, None, None, True))
    print(f"relu_gate 16x16 C128 +raw: {t * 1e6:7.1f} us {x.numel() * 2 * 3 / t / 1e9:7.0f} GB/s", flush=True)
if "legs" in which:
    # the forward conv's pipeline legs in isolation (debug knob 15; results are garbage, only the time matters)
    from mcgm_b200._lib import EPI_BIAS, EPI_RES
    LEGS = [(0, "all"), (8, "loads + MMA (no epilogue)"), (9, "loads only"), (9 | 4, "activation loads only"),
            (9 | 2, "weight loads only"), (2 | 4 | 8, "MMA only"), (2 | 4, "MMA + epilogue"), (1 | 2 | 4, "epilogue only")]
    for (H, C, fl) in [(32, 128, 0), (32, 128, 3), (32, 256, 0), (8, 128, 0)]:
        x = torch.randn(N, H, H, C, device=dev).to(torch.bfloat16)
        w = torch.randn(C, 3, 3, C, device=dev).to(torch.bfloat16)
        bias = torch.randn(C, device=dev) if fl & EPI_BIAS else None
        res = torch.randn(N, H, H, C, device=dev).to(torch.bfloat16) if fl & EPI_RES else None
        for v, name in LEGS:
            _lib.call("mcgm_debug_set", 15, v)
            t = timeit(lambda: ops.conv_fwd(x, w, torch.bfloat16, bias, res, None, None, None, fl))
            print(f"legs {H}x{H} {C}->{C} f{fl:#x} {name:28s}: {t * 1e6:8.1f} us", flush=True)
        _lib.call("mcgm_debug_set", 15, 0)
if "wgbias" in which:
    # weight gradient with / without the fused bias gradient
    for (H, C) in [(32, 128), (16, 128), (32, 256)]:
        x = torch.randn(N, H, H, C, device=dev).to(torch.bfloat16)
        dy = torch.randn(N, H, H, C, device=dev).to(torch.bfloat16)
        for wb in (False, True):
            t = timeit(lambda: ops.conv_wgrad(x, dy, 3, 3, want_bias=wb))
            print(f"conv_wgrad {H}x{H} {C}->{C} k3 want_bias={int(wb)}: {t * 1e6:8.1f} us", flush=True)
