"""mcgm_b200 — B200-native (sm_100a) hot path of the Multimodal-Controller
generative models: the MC-gated conv block behind the reference's own module API.

Import as ``mcgm_b200`` (see ``mcgm_b200/__init__.py`` at the repo root).

Public surface (mirrors the reference, src/modules/modules.py and src/models/mcgan.py):
    cfg, process_control                      — config dict (src/config.py, src/utils.py:104-192)
    MultimodalController, Wrapper             — src/modules/modules.py:49-85
    models.mcgan()                            — src/models/mcgan.py:212-221
    train_gan.train_step / make_optimizer     — src/train_gan.py:128-194, 223-236
"""
from .config import cfg, process_control  # noqa: F401
from .modules import MultimodalController, Wrapper  # noqa: F401
from . import models  # noqa: F401

__all__ = ["cfg", "process_control", "MultimodalController", "Wrapper", "models"]
