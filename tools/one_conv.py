"""One forward conv launch loop for ncu (debug knobs from MCGM_DEBUG_SET): python tools/one_conv.py H C flags [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as entry
entry.build()
from mcgm_b200 import ops, _lib
_lib.load()
H, C, fl = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3], 0)
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 4
N, dev = 512, "cuda"
x = torch.randn(N, H, H, C, device=dev).to(torch.bfloat16)
w = torch.randn(C, 3, 3, C, device=dev).to(torch.bfloat16)
bias = torch.randn(C, device=dev) if fl & 1 else None
res = torch.randn(N, H, H, C, device=dev).to(torch.bfloat16) if fl & 2 else None
aux = torch.randn(N, H, H, C, device=dev).to(torch.bfloat16) if fl & 64 else None
for _ in range(iters):
    ops.conv_fwd(x, w, torch.bfloat16, bias, res, aux, None, None, fl)
torch.cuda.synchronize()
print("ok")
