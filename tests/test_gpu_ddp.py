"""-m gpu: parity of the data-parallel paths that bench.py measures.

1 GPU : the three graph-replay forms of the iteration (GraphedTrainStep, GraphedDDPIngraphTrainStep and
        GraphedDDPTrainStep with world-size-1 reducers) are the same computation as the eager train_step.
2 GPUs: two NCCL ranks on distinct shards (torchrun, skipped when the box has one GPU) against
        (a) the ORACLE run on the same two shards with the gradients averaged (tests/ddp_oracle.py), fp32 and with
            simulated bf16 storage, and (b) the same two-shard iteration computed in ONE process with our own kernels
            and the gradients averaged by hand — which isolates the collective path (must agree to 1e-4)."""
import copy
import os
import socket
import subprocess
import sys
import tempfile

import pytest
import torch

from util import ROOT, build_model, d_strides, load_arrays, rel_err

pytestmark = pytest.mark.gpu


def _state_worst(sd_a, sd_b):
    return max((rel_err(sd_a[k].float().cpu(), sd_b[k].float().cpu()), k) for k in sd_a if sd_a[k].is_floating_point())


@pytest.mark.parametrize("form", ["single", "ingraph", "segmented"])
def test_graph_forms_equal_eager_step(golden, form):
    from mcgm_b200 import train_gan
    case = golden["cases"]["tiny_cifar"]
    arr = load_arrays("tiny_cifar")
    m_e = build_model(case, device="cuda").to("cuda")
    m_g = copy.deepcopy(m_e)
    m_e.train(True)
    m_g.train(True)
    sd_start = {k: v.clone() for k, v in m_g.state_dict().items()}
    batch = {"img": arr["img"].cuda(), "label": arr["label"].cuda(), "z": [z.cuda() for z in arr["zs"]]}
    opt_g = train_gan.make_optimizers(m_g, capturable=True)
    if form == "single":
        graphed = train_gan.GraphedTrainStep(m_g, opt_g, case["batch"], warmup=1, inject_noise=True)
    else:
        red = {k: train_gan.GradAllReduce(getattr(m_g, k), overlap=False) for k in ("generator", "discriminator")}
        graphed = train_gan.make_ddp_step(m_g, opt_g, red, case["batch"], mode=form, inject_noise=True)
    # the warm-up / capture iterations must leave the caller's model and optimiser state untouched (StateGuard)
    for k, v in m_g.state_dict().items():
        assert torch.equal(v, sd_start[k]), f"graph warm-up changed {k}"
    assert all(int(o.step_dev) == 0 and float(o.exp_avg.abs().max()) == 0.0 for o in opt_g.values())
    opt_e = train_gan.make_optimizers(m_e, capturable=True)
    for _ in range(2):
        out_e = train_gan.train_step(m_e, opt_e, batch, z=batch["z"])
        out_g = graphed(batch)
    assert abs(float(out_e["loss_D"]) - float(out_g["loss_D"])) < 1e-4
    assert abs(float(out_e["loss_G"]) - float(out_g["loss_G"])) < 1e-4
    worst = _state_worst(m_g.state_dict(), m_e.state_dict())
    assert worst[0] < 1e-4, worst


def test_graph_replay_follows_lr_changes(golden):
    """A captured step reads lr from a staged device buffer: changing param_groups[0]['lr'] (what
    ReduceLROnPlateau does, src/train_vae.py:35,78-81) changes the replayed update like the eager one."""
    from mcgm_b200 import train_gan
    case = golden["cases"]["tiny_cifar"]
    arr = load_arrays("tiny_cifar")
    m_e = build_model(case, device="cuda").to("cuda")
    m_g = copy.deepcopy(m_e)
    m_e.train(True)
    m_g.train(True)
    batch = {"img": arr["img"].cuda(), "label": arr["label"].cuda(), "z": [z.cuda() for z in arr["zs"]]}
    opt_g = train_gan.make_optimizers(m_g, capturable=True)
    graphed = train_gan.GraphedTrainStep(m_g, opt_g, case["batch"], warmup=1, inject_noise=True)
    opt_e = train_gan.make_optimizers(m_e, capturable=True)
    for lr in (2e-4, 5e-5):
        for o in list(opt_g.values()) + list(opt_e.values()):
            o.param_groups[0]["lr"] = lr
        train_gan.train_step(m_e, opt_e, batch, z=batch["z"])
        graphed(batch)
    worst = _state_worst(m_g.state_dict(), m_e.state_dict())
    assert worst[0] < 1e-4, worst


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run with gpurun --gpus 2)")
@pytest.mark.parametrize("mode", ["ingraph", "segmented", "eager"])
def test_two_rank_nccl_iteration_vs_oracle_shards(golden, mode):
    from ddp_oracle import ddp_train_iteration, shard_inputs
    from mcgm_b200 import train_gan
    from oracle import bf16_storage_sim as S
    from oracle import mcgan_oracle as O
    name, world = "tiny_cifar", 2
    case, arr = golden["cases"][name], load_arrays(name)
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "res.pt")
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                            "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                            os.path.join(ROOT, "tests", "ddp_worker.py"), out, mode, name],
                           capture_output=True, text=True, timeout=300, cwd=ROOT)
        assert r.returncode == 0, r.stderr[-3000:]
        res = [torch.load(f"{out}.{k}") for k in range(world)]
    print("launch form:", res[0]["form"], res[0].get("diag"))
    # every rank ends with identical parameters (BN running statistics are per-rank by design)
    for k, v in res[0]["sd"].items():
        if v.is_floating_point() and "running_" not in k:
            assert torch.equal(v, res[1]["sd"][k]), f"ranks diverged on {k}"
    shards = [shard_inputs(arr, world, r_) for r_ in range(world)]
    sd0 = build_model(case, device="cpu").state_dict()

    def oracle_run(simulate):
        sds = [O.make_trainable(sd0) for _ in range(world)]
        opts = [O.make_adams(s) for s in sds]
        args = (O, sds, opts, [s[0] for s in shards], [s[1] for s in shards], [s[2] for s in shards], case["classes"],
                d_strides(case))
        if simulate:
            with S._Patched():
                losses = ddp_train_iteration(*args)
        else:
            losses = ddp_train_iteration(*args)
        return losses, sds

    (l32, sds32), (lsim, sdssim) = oracle_run(False), oracle_run(True)
    for r_ in range(world):
        print(f"rank {r_}: gpu D {res[r_]['loss_D']:.4f} G {res[r_]['loss_G']:.4f} | oracle fp32 {l32[r_]} | bf16-storage oracle {lsim[r_]}")
        assert abs(res[r_]["loss_D"] - l32[r_][0]) < 0.2 and abs(res[r_]["loss_G"] - l32[r_][1]) < 0.2
        assert abs(res[r_]["loss_D"] - lsim[r_][0]) < 3e-2 and abs(res[r_]["loss_G"] - lsim[r_][1]) < 3e-2
    # parameter updates against the fp32 oracle: update norms (Adam's first steps are sign-like) as in the N=1 test
    bad = []
    from test_gpu_parity_r2 import _feeds_batchnorm
    for k, v0 in sd0.items():
        if not v0.is_floating_point() or not sds32[0][k].requires_grad or _feeds_batchnorm(k):
            continue           # (generator biases in front of a BatchNorm: analytically zero gradient, pure noise)
        ref = float((sds32[0][k].detach() - v0).norm())
        got = float((res[0]["sd"][k] - v0).norm())
        if ref > 1e-4 and abs(got - ref) > 0.15 * ref + 2.5e-4:
            bad.append((k, got, ref))
    assert not bad, bad[:5]
    # per-rank BatchNorm statistics follow each rank's own shard
    for r_ in range(world):
        for k, v in res[r_]["sd"].items():
            if k.endswith("running_mean") or k.endswith("running_var"):
                assert rel_err(v, sds32[r_][k]) < 3e-2, (r_, k)

    # (b) the same two-shard iteration in ONE process with our kernels, gradients averaged by hand
    models_ = [build_model(case, device="cuda").to("cuda") for _ in range(world)]
    for m in models_:
        m.train(True)
    opts = [train_gan.make_optimizers(m) for m in models_]
    dev_sh = [(s[0].cuda(), s[1].cuda(), [z.cuda() for z in s[2]]) for s in shards]

    def avg(net):
        flats = [o[net].flat.grad for o in opts]
        mean = sum(flats) / world
        for f in flats:
            f.copy_(mean)

    for i in range(5):
        for r_ in range(world):
            opts[r_]["discriminator"].zero_grad(); opts[r_]["generator"].zero_grad()
            d_last = train_gan.d_forward_backward(models_[r_], dev_sh[r_][0], dev_sh[r_][1], dev_sh[r_][2][i])
        avg("discriminator")
        for r_ in range(world):
            opts[r_]["discriminator"].step()
    for r_ in range(world):
        opts[r_]["discriminator"].zero_grad(); opts[r_]["generator"].zero_grad()
        train_gan.g_forward_backward(models_[r_], dev_sh[r_][1], dev_sh[r_][2][5])
    avg("generator")
    for r_ in range(world):
        opts[r_]["generator"].step()
    worst = _state_worst(res[0]["sd"], models_[0].state_dict())
    print("2-rank NCCL vs single-process hand-averaged:", worst)
    assert worst[0] < 1e-4, worst
