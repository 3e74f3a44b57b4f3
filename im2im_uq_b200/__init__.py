"""im2im_uq_b200 - B200 (sm_100a) hot path for aangelopoulos/im2im-uq behind the reference's own Python surface.

Host code is Python/PyTorch (device memory, streams, torch.distributed); everything that touches a pixel is a
hand-written CUDA kernel in ``csrc/`` reached through the C ABI declared in ``include/im2im_uq.h``.
There is no CPU fallback: entry points raise if the CUDA library or a CUDA device is missing.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"
