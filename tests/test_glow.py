"""MCGlow (BASELINE.json configs[4]): host parity + oracle pinning on the CPU,
model / gradient / train-step parity on the GPU.  The fixture model has K=2 flows per level and
hidden width 128 (the full K=16 / 512 model runs the same code path; its size only scales the loops)."""
import numpy as np
import pytest
import torch

from util import load_arrays, rel_err, sd_hash


def _build(device="cpu"):
    from mcgm_b200 import cfg, models, process_control
    cfg.update(device=device, data_name="CIFAR10", model_name="mcglow", control={"controller_rate": "0.5"})
    process_control()
    cfg["classes_size"] = 10
    cfg["glow"]["K"] = 2
    cfg["glow"]["hidden_size"] = 128
    torch.manual_seed(0)
    np.random.seed(0)
    return models.mcglow()


def test_glow_state_dict_bit_identical_to_reference(golden):
    case = golden["cases"]["glow_small"]
    sd = _build().state_dict()
    assert [[k, list(v.shape)] for k, v in sd.items()] == case["keys"]
    assert sd_hash(sd) == case["state_dict_sha256"]


def test_glow_oracle_matches_reference(golden):
    from oracle import mcglow_oracle as G
    case, arr = golden["cases"]["glow_small"], load_arrays("glow_small")
    sd = G.make_trainable(_build().state_dict())
    with torch.no_grad():
        out = G.forward(arr["img"], arr["label"], sd, 10, arr["noise"])          # data-dependent init pass
    assert abs(float(out["loss"]) - case["loss_init"]) < 1e-5
    for k, v in arr.items():
        if k.startswith("init/"):
            assert rel_err(sd[k[5:]].detach(), v) < 1e-5, k
    out = G.forward(arr["img"], arr["label"], sd, 10, arr["noise"])
    out["loss"].backward()
    assert abs(float(out["loss"].detach()) - case["loss_fwd"]) < 1e-5
    for i, z in enumerate(out["z"]):
        assert rel_err(z.detach(), arr[f"z{i}"]) < 1e-5
    for k, ref in case["grad_norms"].items():
        assert abs(float(sd[k].grad.norm()) - ref) <= 2e-4 * max(ref, 1e-6) + 1e-9, k
    opt = G.make_adam(sd)
    G.train_step(sd, opt, arr["img"], arr["label"], 10, arr["noise"])
    out2 = G.forward(arr["img"], arr["label"], sd, 10, arr["noise"])
    assert abs(float(out2["loss"].detach()) - case["loss_step2"]) < 1e-4


@pytest.mark.gpu
def test_glow_forward_gradients_and_step_vs_reference(golden):
    """bf16 storage inside the coupling networks; the flow variables themselves stay fp32."""
    from mcgm_b200 import train_vae
    from oracle import mcglow_oracle as G
    case, arr = golden["cases"]["glow_small"], load_arrays("glow_small")
    model = _build("cuda")
    sd = G.make_trainable(model.state_dict())
    model = model.to("cuda")
    model.train(True)
    inp = {"img": arr["img"].cuda(), "label": arr["label"].cuda()}
    noise = arr["noise"].cuda()
    with torch.no_grad():
        out = model(inp, noise)                                                   # ActNorm init pass
    print("glow init loss", float(out["loss"]), case["loss_init"])
    assert abs(float(out["loss"]) - case["loss_init"]) < 2e-2
    assert int(model.blocks[0].flows[0].coupling.net[1].module.initialized) == 1
    with torch.no_grad():
        G.forward(arr["img"], arr["label"], sd, 10, arr["noise"])
    out = model(inp, noise)
    out["loss"].backward()
    ref = G.forward(arr["img"], arr["label"], sd, 10, arr["noise"])
    ref["loss"].backward()
    print("glow loss", float(out["loss"].detach()), float(ref["loss"].detach()))
    assert abs(float(out["loss"].detach()) - case["loss_fwd"]) < 2e-2
    for i, z in enumerate(out["z"]):
        assert rel_err(z.detach().float().cpu(), arr[f"z{i}"]) < 3e-2
    errs = sorted(((rel_err(p.grad.float().cpu(), sd[k].grad), k) for k, p in model.named_parameters()
                   if float(sd[k].grad.norm()) > 1e-6), reverse=True)
    print("glow worst gradient errors:", errs[:4], "median", errs[len(errs) // 2])
    assert errs[0][0] < 1.2e-1, errs[:4]
    assert errs[len(errs) // 2][0] < 3e-2
    # one optimiser step (clip + Adam 3e-4), then the next loss
    opt = train_vae.make_optimizer(model)            # FusedAdam: clip_grad_norm_(1) folded into the step
    opt.zero_grad()
    out = model(inp, noise)
    out["loss"].backward()
    opt.step()
    out2 = model(inp, noise)
    print("glow step loss", float(out2["loss"].detach()), case["loss_step2"])
    assert abs(float(out2["loss"].detach()) - case["loss_step2"]) < 2e-2
    # reverse() inverts forward(): reconstruct the dequantised image from z
    model.train(False)
    with torch.no_grad():
        z = model({"img": inp["img"], "label": inp["label"]}, noise)["z"]
        rec = model.reverse({"z": z, "label": inp["label"], "reconstruct": True})["img"]
    want = torch.clamp(inp["img"] * 0.5 + noise / 256, -.5, .5) * 2
    assert rel_err(rec.cpu(), want.cpu()) < 3e-2
