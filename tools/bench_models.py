"""Informational throughput of the other two model families on the B200 kernels (not the bench.py headline):
MCVAE COIL100-shaped (100 modes, widths 64/128/256, 2 ResBlocks) and MCGlow CIFAR10-shaped (K=16, L=3,
hidden 512, affine, LU) — BASELINE.json configs[3] and [4].  One step = forward + backward + clip + Adam
(src/train_vae.py:106-111, src/train_glow.py:116-121), eager issue, CUDA events, synthetic data."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402
entry.build()
from mcgm_b200 import cfg, models, process_control, train_vae  # noqa: E402


def run(name, batch, steps=10, warm=4):
    classes = {"mcvae": 100, "mcglow": 10}[name]
    cfg.update(device="cuda", data_name="COIL100" if name == "mcvae" else "CIFAR10", model_name=name,
               control={"controller_rate": "0.5"})
    process_control()
    cfg["classes_size"] = classes
    torch.manual_seed(0)
    np.random.seed(0)
    model = getattr(models, name)().cuda()
    model.train(True)
    g = torch.Generator().manual_seed(1)
    inp = {"img": (torch.rand(batch, *cfg["data_shape"], generator=g) * 2 - 1).cuda(),
           "label": torch.randint(0, classes, (batch,), generator=g).cuda()}
    if name == "mcglow":
        with torch.no_grad():
            model(inp)                     # data-dependent ActNorm initialisation (train_glow.py:60-67)
    opt = train_vae.make_optimizer(model)
    for _ in range(warm):
        train_vae.train_step(model, opt, inp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = train_vae.train_step(model, opt, inp)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{name} eager : batch {batch}, {ms:.2f} ms/step, {batch / ms * 1e3:.0f} img/s, loss {float(out['loss']):.4f}", flush=True)
    del out                                   # frees the eager autograd graph (its AccumulateGrad nodes are bound to this stream)
    graphed = train_vae.GraphedTrainStep(model, opt, inp)
    for _ in range(2):
        graphed(inp)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        out = graphed(inp)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{name} graph : batch {batch}, {ms:.2f} ms/step, {batch / ms * 1e3:.0f} img/s, loss {float(out['loss']):.4f}, "
          f"{graphed.native_launches} native launches / step", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["mcvae", "mcglow"]
    if "mcvae" in which:
        run("mcvae", 512)
    if "mcglow" in which:
        run("mcglow", 64)
