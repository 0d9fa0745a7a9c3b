"""Short, deterministic slice of the bench workload for ncu (MCGAN CIFAR10-shaped, batch 512):
--mode d : one D update (D(real), G(z), D(G.detach()), backward, Adam)
--mode g : one G update (G(z), D(G), backward, Adam)
--mode step : one full iteration (5 D + 1 G)"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="d")
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--warm", type=int, default=0)
a = ap.parse_args()

import __graft_entry__ as entry  # noqa: E402
entry.build()
from mcgm_b200 import cfg, models, process_control, train_gan  # noqa: E402
from mcgm_b200.config import GAN_HYPER  # noqa: E402

cfg.update(device="cuda", data_name="CIFAR10", model_name="mcgan", control={"controller_rate": "0.5"})
process_control()
cfg["classes_size"] = 10
torch.manual_seed(0)
model = models.mcgan().cuda()
model.train(True)
opt = train_gan.make_optimizers(model)
g = torch.Generator().manual_seed(1)
batch = {"img": (torch.rand(a.batch, 3, 32, 32, generator=g) * 2 - 1).cuda(),
         "label": torch.randint(0, 10, (a.batch,), generator=g).cuda()}
if a.mode == "d":
    GAN_HYPER["iter"] = {"generator": 0, "discriminator": 1}
elif a.mode == "g":
    GAN_HYPER["iter"] = {"generator": 1, "discriminator": 0}
for _ in range(a.warm):
    train_gan.train_step(model, opt, batch)
torch.cuda.synchronize()
torch.cuda.profiler.start()          # ncu --profile-from-start off: only the update below is captured
out = train_gan.train_step(model, opt, batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", {k: (float(v) if v is not None else None) for k, v in out.items() if k != "loss"})
