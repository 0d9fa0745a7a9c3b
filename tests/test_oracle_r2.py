"""The round-2 oracle restatements (Glow reverse / eval-mode sampling paths, MC-gated PixelCNN) pinned to fixtures
generated from the unmodified reference (oracle/make_golden_r2.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from util import GOLDEN, load_arrays, rel_err, sd_hash


@pytest.fixture(scope="module")
def r2():
    with open(os.path.join(GOLDEN, "r2_golden.json")) as f:
        return json.load(f)


def fill_glow_zero_layers(named_parameters):
    """The deterministic non-zero fill of Glow's zero-initialised layers that make_golden_r2.py applied to the
    reference model (a trained model has them non-zero; with zeros the coupling is the identity)."""
    gw = torch.Generator().manual_seed(14)
    with torch.no_grad():
        for n, p in named_parameters:
            if n.endswith('net.8.module.conv.weight') or n.endswith('prior.conv.weight'):
                p.copy_((torch.randn(p.shape, generator=gw) * 0.01).to(p.device))
            elif n.endswith('net.8.module.scale') or n.endswith('prior.scale') or (n.endswith('conv.bias') and 'prior' in n):
                p.copy_((torch.randn(p.shape, generator=gw) * 0.05).to(p.device))


def build(model_name, data_name, classes, **over):
    from mcgm_b200 import cfg, models, process_control
    cfg.update(device="cpu", data_name=data_name, model_name=model_name, control={"controller_rate": "0.5"})
    process_control()
    cfg["classes_size"] = classes
    for k, v in over.items():
        sec, key = k.split("__")
        cfg[sec][key] = v
    torch.manual_seed(0)
    np.random.seed(0)
    return getattr(models, model_name)()


def test_gan_eval_generate_oracle_matches_reference():
    from oracle import mcgan_oracle as O
    arr = load_arrays("eval_generate")
    m = build("mcgan", "CIFAR10", 10, gan__generator_hidden_size=[64] * 4, gan__discriminator_hidden_size=[64] * 4)
    sd = O.make_trainable(m.state_dict())
    ind = O.one_hot(arr["gan/label"], 10)
    with torch.no_grad():
        O.generator(arr["gan/zs"][0], ind, sd)
        O.generator(arr["gan/zs"][1], ind, sd)
        out = O.generator(arr["gan/zs"][2], ind, sd, training=False)
        d = O.discriminator(out, ind, sd, [2, 1, 1], training=False)
    assert rel_err(out, arr["gan/generated"]) < 1e-5
    assert rel_err(d, arr["gan/d_eval"]) < 1e-5


def test_vae_eval_generate_oracle_matches_reference(r2):
    from oracle import mcvae_oracle as V
    arr = load_arrays("eval_generate")
    m = build("mcvae", "COIL100", 100)
    sd = V.make_trainable(m.state_dict())
    with torch.no_grad():
        V.forward(arr["vae/img"], arr["vae/label"], sd, 100, arr["vae/eps"])
        gen = V.decoder(arr["vae/z"], V.one_hot(arr["vae/label"], 100), sd, False) * 2 - 1
        ev = V.forward(arr["vae/img"], arr["vae/label"], sd, 100, arr["vae/eps"], training=False)
    assert rel_err(gen, arr["vae/generated"]) < 1e-5
    assert rel_err(ev["img"], arr["vae/eval_recon"]) < 1e-5
    assert abs(float(ev["loss"]) - r2["vae_eval_loss"]) < 1e-5


def test_glow_reverse_oracle_matches_reference():
    from oracle import mcglow_oracle as G
    arr = load_arrays("eval_generate")
    m = build("mcglow", "CIFAR10", 10, glow__K=2, glow__hidden_size=128)
    sd = G.make_trainable(m.state_dict())
    with torch.no_grad():
        G.forward(arr["glow/img"], arr["glow/label"], sd, 10, arr["glow/noise"])            # init pass
    fill_glow_zero_layers([(k, v) for k, v in sd.items() if v.requires_grad])
    with torch.no_grad():
        gen = G.reverse([arr[f"glow/z{i}"] for i in range(3)], arr["glow/label"], sd, 10, False)
        fz = G.forward(arr["glow/img"], arr["glow/label"], sd, 10, arr["glow/noise"], training=False)["z"]
        rec = G.reverse(fz, arr["glow/label"], sd, 10, True)
    assert rel_err(gen, arr["glow/generated"]) < 1e-4
    assert rel_err(rec, arr["glow/reconstructed"]) < 1e-4


def test_pixelcnn_oracle_matches_reference(r2):
    from oracle import mcpixelcnn_oracle as P
    case, arr = r2["pixelcnn_small"], load_arrays("pixelcnn_small")
    m = build("mcpixelcnn", "CIFAR10", 10, pixelcnn__num_layer=3, pixelcnn__hidden_size=64)
    sd0 = m.state_dict()
    assert [[k, list(v.shape)] for k, v in sd0.items()] == case["keys"]
    assert sd_hash(sd0) == case["state_dict_sha256"]
    sd = P.make_trainable(sd0)
    out = P.forward(arr["img"], arr["label"], sd, 10)
    out["loss"].backward()
    assert abs(float(out["loss"].detach()) - case["loss_fwd"]) < 1e-5
    assert rel_err(out["logits"].detach()[:4], arr["logits"]) < 1e-5
    for k, ref in case["grad_norms"].items():
        assert abs(float(sd[k].grad.norm()) - ref) <= 2e-4 * max(ref, 1e-6) + 1e-9, k
    with torch.no_grad():
        ev = P.forward(arr["img"], arr["label"], sd, 10, training=False)
    assert abs(float(ev["loss"]) - case["loss_eval"]) < 1e-5
    assert rel_err(ev["logits"][:4], arr["logits_eval"]) < 1e-5
