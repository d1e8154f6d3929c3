"""apophenia_b200 -- B200-native likelihood/score hot path behind Apophenia's
apop_model / apop_vtable interface.

The package holds only what the path needs: `csrc/` (hand-written sm_100a CUDA
kernels and the C-ABI, built into libapop_b200.so), `abi.py` (the reference's
struct layouts for ctypes) and `api.py` (thin host-side mirror of the
reference's operators for tests and the benchmark).
"""
from . import abi  # noqa: F401
from .api import *  # noqa: F401,F403
