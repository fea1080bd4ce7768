"""egaze: host-side mirror of the reference's hot-path operator interface, executed by hand-written sm_100a kernels
(libegaze.so, C-ABI in include/egaze.h).  No CPU fallback, no cuDNN/cuBLAS on the path."""
from . import _lib, ops, engine  # noqa: F401
from .modules import TrunkSequential, DecoderSequential  # noqa: F401
