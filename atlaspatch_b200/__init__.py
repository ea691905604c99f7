"""atlaspatch_b200 -- B200-native drop-in for AtlasPatch's per-slide hot path.

Host side is Python mirroring the reference's plug-in interfaces; all arithmetic on the path runs in
hand-written sm_100a CUDA behind the C ABI of libatlaspatch_b200.so (include/atlaspatch_b200.h).
Importing the package does not need a GPU; calling into it does (no CPU fallback).
"""
__version__ = "0.1.0"

from atlaspatch_b200._lib import AtlasB200Error, Context, load_library  # noqa: F401
