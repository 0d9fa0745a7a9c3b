"""CPU: the host side of the input pipeline / metric logging and the DDP oracle helper."""
import torch

from util import build_model, d_strides, load_arrays


def test_collate_matches_reference_semantics():
    from mcgm_b200 import data
    ds = data.SyntheticImages(10, [3, 32, 32], 7, seed=1)
    batch = data.collate(data.input_collate([ds[i] for i in range(4)]))          # src/data.py:65-73, src/utils.py:259-262
    assert batch["img"].shape == (4, 3, 32, 32) and batch["label"].shape == (4,)
    assert float(batch["img"].min()) >= -1.0 and float(batch["img"].max()) <= 1.0
    assert torch.equal(batch["img"], ds.batch(torch.arange(4))["img"])
    assert data.input_collate([torch.zeros(2), torch.ones(2)]).shape == (2, 2)       # non-dict samples: default collate


def test_device_meter_running_means_on_cpu():
    from mcgm_b200 import data
    meter = data.DeviceMeter("cpu", flush_every=0)
    ref_mean, ref_count = 0.0, 0
    for v, n in ((0.3, 128), (0.7, 128), (0.1, 64)):
        meter.append({"loss": torch.tensor(v)}, "train", n)
        ref_count += n                                                             # src/logger.py:34-50
        ref_mean = ((ref_count - n) * ref_mean + n * v) / ref_count
    assert abs(meter.mean["train/loss"] - ref_mean) < 1e-6      # values enter as float32 tensors


def test_ddp_oracle_world1_equals_train_iteration(golden):
    from ddp_oracle import ddp_train_iteration
    from oracle import mcgan_oracle as O
    case, arr = golden["cases"]["tiny_cifar"], load_arrays("tiny_cifar")
    sd0 = build_model(case).state_dict()
    a, b = O.make_trainable(sd0), O.make_trainable(sd0)
    la = O.train_iteration(a, O.make_adams(a), arr["img"], arr["label"], list(arr["zs"]), case["classes"], d_strides(case))
    lb = ddp_train_iteration(O, [b], [O.make_adams(b)], [arr["img"]], [arr["label"]], [list(arr["zs"])], case["classes"],
                             d_strides(case))
    assert abs(float(la[0]) - lb[0][0]) < 1e-6 and abs(float(la[1]) - lb[0][1]) < 1e-6
    for k in a:
        assert torch.equal(a[k].detach(), b[k].detach()), k
