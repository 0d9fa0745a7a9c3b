"""Pin the oracle (oracle/mcgan_oracle.py) against fixtures generated FROM THE REFERENCE
(oracle/make_golden.py): forward outputs, gradients, one full 5D+1G iteration."""
import pytest
import torch

from oracle import mcgan_oracle as O
from util import build_model, d_strides, load_arrays, rel_err


@pytest.mark.parametrize("name", ["tiny_cifar", "tiny_omni"])
def test_oracle_forward_matches_reference(golden, name):
    case = golden["cases"][name]
    arr = load_arrays(name)
    sd = O.make_trainable(build_model(case).state_dict())
    ind = O.one_hot(arr["label"], case["classes"])
    with torch.no_grad():
        gen = O.generator(arr["zs"][0], ind, sd)
    assert rel_err(gen, arr["gen_out"]) < 1e-5
    sd = O.make_trainable(build_model(case).state_dict())
    with torch.no_grad():
        d = O.discriminator(arr["img"], ind, sd, d_strides(case))
    assert rel_err(d, arr["d_real"]) < 1e-5


@pytest.mark.parametrize("name", ["tiny_cifar", "tiny_omni"])
def test_oracle_gradients_match_reference(golden, name):
    case = golden["cases"][name]
    arr = load_arrays(name)
    sd = O.make_trainable(build_model(case).state_dict())
    ind = O.one_hot(arr["label"], case["classes"])
    loss = -O.discriminator(O.generator(arr["zs"][0], ind, sd), ind, sd, d_strides(case)).mean()
    loss.backward()
    assert abs(float(loss) - case["g_loss_fwd"]) < 1e-5 * max(1.0, abs(case["g_loss_fwd"]))
    for k, ref in case["grad_norms"].items():
        got = float(sd[k].grad.norm())
        assert abs(got - ref) <= 1e-4 * max(ref, 1e-6), k


@pytest.mark.parametrize("name", ["tiny_cifar", "tiny_omni"])
def test_oracle_full_iteration_matches_reference(golden, name):
    case = golden["cases"][name]
    arr = load_arrays(name)
    sd0 = build_model(case).state_dict()
    sd = O.make_trainable(sd0)
    opt = O.make_adams(sd)
    d_loss, g_loss = O.train_iteration(sd, opt, arr["img"], arr["label"], list(arr["zs"]), case["classes"],
                                       d_strides(case))
    assert abs(float(d_loss) - case["loss_D"]) < 1e-4
    assert abs(float(g_loss) - case["loss_G"]) < 1e-4
    for k, ref in case["delta_norms"].items():
        got = float((sd[k].detach().float() - sd0[k].float()).norm())
        assert abs(got - ref) <= 2e-3 * max(ref, 1e-7) + 1e-7, (k, got, ref)
    for k, v in arr.items():
        if k.startswith("after/"):
            assert rel_err(sd[k[6:]].detach(), v) < 1e-4, k
