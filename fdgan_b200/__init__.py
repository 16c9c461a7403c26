"""fdgan_b200 -- B200-native FD-GAN hot path (generator, Fusion-discriminator, frequency decomposition,
VGG16 features) behind the reference's torch.nn.Module surface.  Importing this package loads the in-tree
CUDA library and fails loudly if it has not been built."""
from . import _lib  # noqa: F401  (raises ImportError when libfdgan_b200.so is missing)
from .dehaze1113 import D, FDGAN  # noqa: F401
from .loss import Blur, Laplacian, blur, freq_concat, laplace_filter  # noqa: F401
from .vgg16 import Vgg16  # noqa: F401
from . import compat, pytorch_ssim  # noqa: F401  (compat.install(): the reference's import lines, unchanged)

__all__ = ["FDGAN", "D", "Vgg16", "Blur", "Laplacian", "blur", "laplace_filter", "freq_concat"]
