"""Run every GPU parity check and print one line per row (never stops at the first failure).
Usage on the GPU box:  python tools/gpu_bringup.py [check_name ...]"""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402

g.build()
import gpu_checks  # noqa: E402

def diag_conv():
    """Structured 1x1 / 3x3 probes that localise a wrong descriptor / swizzle / K-advance."""
    import torch.nn.functional as F
    from mcgm_b200 import ops
    torch.manual_seed(0)
    N, H, W, Cin, Cout = 2, 8, 8, 64, 64
    w = (torch.randn(Cout, Cin, 1, 1) * 0.1).to(torch.bfloat16).float()
    wf, _ = ops.pack_weight(w.cuda())
    for chunk in range(4):
        x = torch.zeros(N, Cin, H, W)
        x[:, chunk * 16:(chunk + 1) * 16] = torch.randn(N, 16, H, W).to(torch.bfloat16).float()
        ref = F.conv2d(x, w)
        y = ops.conv_fwd(x.permute(0, 2, 3, 1).contiguous().cuda().to(torch.bfloat16), wf, torch.float32)
        y = y.permute(0, 3, 1, 2).cpu()
        err = (y - ref).abs()
        print(f"[diag] 1x1 k-chunk {chunk}: max|err|={float(err.max()):.3e} max|ref|={float(ref.abs().max()):.3e} "
              f"max|y|={float(y.abs().max()):.3e} rows_bad={int((err.amax(dim=1) > 1e-3).sum())}/128 "
              f"cols_bad={int((err.amax(dim=(0, 2, 3)) > 1e-3).sum())}/64", flush=True)
    x = torch.randn(N, Cin, H, W).to(torch.bfloat16).float()
    ref = F.conv2d(x, w)
    y = ops.conv_fwd(x.permute(0, 2, 3, 1).contiguous().cuda().to(torch.bfloat16), wf, torch.float32).permute(0, 3, 1, 2).cpu()
    e = (y - ref).abs().amax(dim=1).reshape(-1)
    print("[diag] per-pixel max err (128 pixels):", " ".join(f"{float(v):.0e}" for v in e), flush=True)
    print("[diag] y[0,:8,0,0] =", y[0, :8, 0, 0].tolist(), " ref =", ref[0, :8, 0, 0].tolist(), flush=True)
    return []


if os.environ.get("MCGM_DEBUG_SET"):
    from mcgm_b200 import _lib as _l
    for kv in os.environ["MCGM_DEBUG_SET"].split(","):
        k, v = kv.split("=")
        _l.call("mcgm_debug_set", int(k), int(v))
        print("debug_set", k, v, flush=True)
gpu_checks.diag_conv = diag_conv
names = sys.argv[1:] or [f.__name__ for f in gpu_checks.ALL_KERNEL_CHECKS]
print("device:", torch.cuda.get_device_name(0), flush=True)
n_bad = 0
for name in names:
    t0 = time.time()
    try:
        rows = getattr(gpu_checks, name)()
        torch.cuda.synchronize()
    except Exception:
        print(f"[{name}] EXCEPTION\n{traceback.format_exc()}", flush=True)
        n_bad += 1
        # a trapped kernel poisons the context; stop here so later output is not misleading
        if "CUDA" in traceback.format_exc() or "cuda" in traceback.format_exc():
            break
        continue
    for (label, err, tol) in rows:
        ok = err <= tol
        n_bad += (not ok)
        print(f"[{name}] {'ok  ' if ok else 'FAIL'} {label}: err={err:.3e} tol={tol:.1e}", flush=True)
    print(f"[{name}] done in {time.time() - t0:.1f}s", flush=True)
print("FAILED rows:", n_bad)
sys.exit(1 if n_bad else 0)
