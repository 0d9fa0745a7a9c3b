"""Model factories of the hot path (src/models/__init__.py): ``mcgan()`` reads ``cfg``."""
from .mcgan import MCGAN, Generator, Discriminator, mcgan  # noqa: F401
from .utils import init_param, make_SpectralNormalization  # noqa: F401
