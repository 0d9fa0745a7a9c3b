"""torchrun worker of tests/test_gpu_ddp.py: one MCGAN data-parallel iteration on this rank's shard over NCCL.
argv: <out.pt> <mode: ingraph|segmented|eager> <case name>"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    out_path, mode, name = sys.argv[1], sys.argv[2], sys.argv[3]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from ddp_oracle import shard_inputs
    from mcgm_b200 import train_gan
    from util import build_model, load_arrays
    with open(os.path.join(ROOT, "tests", "golden", "mcgan_golden.json")) as f:
        case = json.load(f)["cases"][name]
    arr = load_arrays(name)
    model = build_model(case, device=str(dev)).to(dev)
    model.train(True)
    img, label, zs = shard_inputs(arr, world, rank)
    batch = {"img": img.to(dev), "label": label.to(dev), "z": [z.to(dev) for z in zs]}
    graph = mode != "eager"
    opt = train_gan.make_optimizers(model, capturable=graph)
    red = {k: train_gan.GradAllReduce(getattr(model, k), overlap=not graph) for k in ("generator", "discriminator")}
    step = None
    if graph:
        step = train_gan.make_ddp_step(model, opt, red, case["batch"], mode=mode, inject_noise=True)
        out = step(batch)
        form = step.launch_form
    else:
        out = train_gan.train_step(model, opt, batch, z=batch["z"], reducers=red)
        form = "eager"
    torch.cuda.synchronize()
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    res = {"rank": rank, "loss_D": float(out["loss_D"]), "loss_G": float(out["loss_G"]), "form": form, "sd": sd}
    if not graph:
        # diagnostic: one more D backward; after finish() the flat gradient buffer must be bit-identical on all ranks
        red["discriminator"].zero_grad(); red["generator"].zero_grad()
        red["generator"].enabled = False
        train_gan.d_forward_backward(model, batch["img"], batch["label"], batch["z"][0])
        local = red["discriminator"].flat.clone()
        red["discriminator"].finish()
        red["generator"].enabled = True
        both = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(both, red["discriminator"].flat)
        diff = (both[0] - both[1]).abs()
        exp = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(exp, local)
        want = (exp[0] + exp[1]) / world
        res["diag"] = {"ranks_max_diff": float(diff.max()), "n_diff": int((diff > 0).sum()), "numel": diff.numel(),
                       "first_diff": int((diff > 0).nonzero()[0]) if bool((diff > 0).any()) else -1,
                       "vs_expected_max": float((both[0] - want).abs().max()),
                       "offsets": list(opt["discriminator"].flat.offsets)[:6], "buckets": red["discriminator"].buckets,
                       "pending": list(red["discriminator"].pending)}
    torch.save(res, f"{out_path}.{rank}")
    train_gan.shutdown_process_group(step)


if __name__ == "__main__":
    main()
