"""Model factories of the hot path (src/models/__init__.py): ``mcgan()`` / ``mcvae()`` / ``mcglow()`` / ``mcpixelcnn()`` read ``cfg``."""
from .mcgan import MCGAN, Generator, Discriminator, mcgan  # noqa: F401
from .mcvae import MCVAE, mcvae  # noqa: F401
from .mcglow import MCGlow, mcglow  # noqa: F401
from .mcpixelcnn import MCGatedPixelCNN, mcpixelcnn  # noqa: F401
from .utils import init_param, make_SpectralNormalization  # noqa: F401
