"""Per-shape time breakdown of one eager MCGAN iteration (5 D + 1 G, batch 512): CUDA events around
every tagged native call (ops.PROFILE), grouped by (tag, shape descriptor), sorted by total time.
Everything not bracketed (torch launches, un-tagged native calls, gaps) shows up as 'other'."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402
entry.build()
from mcgm_b200 import cfg, models, ops, process_control, train_gan  # noqa: E402

cfg.update(device="cuda", data_name="CIFAR10", model_name="mcgan", control={"controller_rate": "0.5"})
process_control()
cfg["classes_size"] = 10
torch.manual_seed(0)
model = models.mcgan().cuda()
model.train(True)
opt = train_gan.make_optimizers(model)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
g = torch.Generator().manual_seed(1)
batch = {"img": (torch.rand(B, 3, 32, 32, generator=g) * 2 - 1).cuda(), "label": torch.randint(0, 10, (B,), generator=g).cuda()}
for _ in range(2):
    train_gan.train_step(model, opt, batch)
torch.cuda.synchronize()
ops.PROFILE = {}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda._sleep(int(0.12 * 1.9e9))       # let the host run ahead: no launch gaps inside the event pairs
e0.record()
train_gan.train_step(model, opt, batch)
e1.record()
torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
total = e0.elapsed_time(e1)
rows = {}
for tag, recs in prof.items():
    for a, b, work, desc, _ref in recs:
        r = rows.setdefault((tag, desc), [0, 0.0, 0.0])
        r[0] += 1; r[1] += a.elapsed_time(b); r[2] += work
acc = sum(r[1] for r in rows.values())
print(f"eager iteration {total:.2f} ms, bracketed native calls {acc:.2f} ms")
print(f"{'tag':6s} {'shape':58s} {'n':>4s} {'ms':>8s} {'%':>6s} {'us/call':>8s} {'rate':>10s}")
for (tag, desc), (n, ms, work) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    hbm = tag in ("gate", "thin")
    rate = work / (ms * 1e-3) / (1e9 if hbm else 1e12)
    print(f"{tag:6s} {desc:58s} {n:4d} {ms:8.3f} {100 * ms / acc:6.2f} {ms / n * 1e3:8.1f} {rate:8.1f} {'GB/s' if hbm else 'TF/s'}")
