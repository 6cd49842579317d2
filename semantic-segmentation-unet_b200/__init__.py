"""unetb200 -- B200-native U-Net hot path (train step + tiled inference) behind the reference's Python surface.

Host code is Python; torch tensors are the allocation shell; all arithmetic runs in hand-written sm_100a CUDA
(libunetb200.so, C ABI declared in include/unetb200.h) reached through ctypes.  There is no CPU or library
fallback: importing `unetb200._C` fails loudly when the shared library is missing.
"""
__version__ = "0.1.0"
