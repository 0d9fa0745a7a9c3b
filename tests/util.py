"""Shared helpers for the tests (host side)."""
import hashlib
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def sd_hash(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def build_model(case, device="cpu"):
    """Build mcgm_b200's MCGAN for a golden case the way a user of the reference would:
    fill cfg, process_control(), optional width override, seed, models.mcgan()."""
    import mcgm_b200
    from mcgm_b200 import cfg, models, process_control
    cfg.update(device=device, data_name=case["data_name"], model_name="mcgan",
               control={"controller_rate": "0.5"})
    process_control()
    cfg["classes_size"] = case["classes"]
    if case.get("g_widths"):
        cfg["gan"]["generator_hidden_size"] = list(case["g_widths"])
        cfg["gan"]["discriminator_hidden_size"] = list(case["d_widths"])
    torch.manual_seed(case["seed"])
    return models.mcgan()


def load_arrays(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files if z[k].dtype.kind != "U"}      # string arrays: read with numpy


def d_strides(case):
    n = len(case["d_widths"]) if case.get("d_widths") else 4
    n_down = n - 3 if case["data_name"] in ("CIFAR10", "CIFAR100") else n - 2
    return [2] * n_down + [1] * (n - 1 - n_down)


def rel_err(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))
