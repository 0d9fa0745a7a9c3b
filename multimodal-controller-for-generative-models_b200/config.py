"""Global ``cfg`` dict and ``process_control`` — the subset of the reference's
configuration layer that the hot-path models read at construction time
(src/config.py:1-6, src/config.yml, src/utils.py:104-192).

The reference loads ``config.yml`` from the CWD into a module-level dict and its
models read ``cfg[...]`` when they are built (src/models/mcgan.py:212-221) and in
``forward`` (``cfg['classes_size']``, ``cfg['device']``; mcgan.py:195-196).  The
same contract is kept here so `models.mcgan()` is a drop-in: fill ``cfg`` and call
``process_control()``.
"""

cfg = {
    "control": {"controller_rate": "0.5"},
    "data_name": "CIFAR10",
    "model_name": "mcgan",
    "subset": "label",
    "device": "cuda",
    "world_size": 1,
    "init_seed": 0,
    # B200 build: storage type of activations inside the fused blocks ('bf16'; fp32 accumulate)
    "compute_dtype": "bf16",
    # B200 build: BatchNorm batch statistics accumulated by the producing conv's (staged) epilogue instead of a pass
    # over the tensor (MCGM_BN_STATS_IN_CONV=0 switches back to the separate statistics kernels)
    "bn_stats_in_conv": __import__("os").environ.get("MCGM_BN_STATS_IN_CONV", "1") == "1",
}

# data_name -> (data_shape, generate_per_mode)                      src/utils.py:107-127
_DATA = {
    "MNIST": ([1, 32, 32], 1000), "FashionMNIST": ([1, 32, 32], 1000),
    "Omniglot": ([1, 32, 32], 20),
    "SVHN": ([3, 32, 32], 1000), "CIFAR10": ([3, 32, 32], 1000),
    "COIL100": ([3, 32, 32], 100),
    "ImageNet32": ([3, 32, 32], 20),
    "CelebA-HQ": ([3, 128, 128], 20), "ImageNet": ([3, 128, 128], 20),
}


def process_control():
    """Derive shapes / widths / batch sizes from ``cfg`` (src/utils.py:104-192)."""
    if "controller_rate" in cfg["control"]:
        cfg["controller_rate"] = float(cfg["control"]["controller_rate"])
    if cfg["data_name"] not in _DATA:
        raise ValueError("Not valid dataset")
    shape, per_mode = _DATA[cfg["data_name"]]
    cfg["data_shape"] = list(shape)
    cfg["generate_per_mode"] = per_mode
    side = cfg["data_shape"][1]
    if side not in (32, 128):
        raise ValueError("Not valid data shape")
    name = cfg["model_name"]
    if name in ("cgan", "mcgan"):
        if side == 32:
            if cfg["data_name"] == "CIFAR10":
                g, d = [256] * 4, [128] * 4
            else:
                g, d = [512, 256, 128, 64], [64, 128, 256, 512]
        else:
            g, d = [1024, 512, 256, 128, 64], [64, 128, 256, 512, 1024]
        cfg["gan"] = {"latent_size": 128, "generator_hidden_size": g,
                      "discriminator_hidden_size": d, "embedding_size": 32}
    elif name in ("cvae", "mcvae"):
        cfg["vae"] = ({"hidden_size": [64, 128, 256], "latent_size": 128} if side == 32 else
                      {"hidden_size": [64, 128, 256, 512, 512], "latent_size": 256})
        cfg["vae"].update(num_res_block=2, embedding_size=32)
    elif name in ("cpixelcnn", "mcpixelcnn"):                    # src/utils.py:139-143
        cfg["pixelcnn"] = {"num_layer": 15, "hidden_size": 128, "num_embedding": 512}
    elif name in ("cglow", "mcglow"):
        cfg["glow"] = {"hidden_size": 512, "K": 16, "L": 3 if side == 32 else 5,
                       "affine": True, "conv_lu": True}
    cfg["batch_size"] = {"train": 128, "test": 512} if side == 32 else {"train": 32, "test": 128}
    return


# train_gan.py:33-55 — the hyper-parameters the script hard-codes for mcgan
GAN_HYPER = {
    "lr": {"generator": 2e-4, "discriminator": 2e-4},
    "iter": {"generator": 1, "discriminator": 5},
    "betas": {"generator": (0.5, 0.999), "discriminator": (0.5, 0.999)},
    "weight_decay": 0,
    "loss_type": "Hinge",
    "optimizer_name": "Adam",
}
