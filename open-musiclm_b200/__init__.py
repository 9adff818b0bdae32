"""open-musiclm hot path, B200-native (sm_100a): package root.

Only what the TokenConditionedTransformer training path needs lives here:
  csrc/     hand-written CUDA kernels + the C ABI (libomlm_b200.so)
  lib.py    ctypes binding of that ABI (no fallback)
"""
__version__ = "0.1.0"
