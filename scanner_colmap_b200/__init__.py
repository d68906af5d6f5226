"""B200-native (sm_100a) replacement for the feature-matching hot path of garyjyzhang/scanner-colmap.

Scope: the SIFT descriptor matcher behind ``integration/feature_matching.py`` ->
``SequentialMatchingCPU`` -> ``colmap::MatchSiftFeaturesCPU``
(``/root/reference/integration/op_cpp/sequential_matching.cc:154``) and nothing else.

* ``csrc/``      CUDA kernels + the C ABI (``include/smb.h``) -> ``libsmb.so``
* ``matcher``    ctypes mirror of the C ABI
* ``synth``      synthetic SIFT-like descriptors
"""
from .matcher import SiftMatcher, SmbError, load_library, sequential_pairs  # noqa: F401

__all__ = ["SiftMatcher", "SmbError", "load_library", "sequential_pairs"]
