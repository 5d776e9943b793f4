"""az_b200 - host side of the B200-native self-play engine.

Python here only owns plumbing: device memory (torch), streams / CUDA graphs, torch.distributed.
All search, environment and encoding work happens in the sm_100a kernels of libaz_b200.so, reached
through the C ABI declared in include/az_b200.h.  There is no CPU fallback: importing
az_b200.native without the built library, or creating an engine without a CUDA device, raises.
"""
from .native import AzConfig, AzLayout, NativeError, lib  # noqa: F401
