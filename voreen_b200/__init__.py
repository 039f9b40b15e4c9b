"""voreen_b200 -- B200-native direct volume ray casting behind Voreen's raycaster interface.

The package holds only what the hot path needs: csrc/ (CUDA kernels + the C ABI of include/b200rc.h), host/ (C++ mirror
of the reference's processor / transfer-function / camera interface) and thin ctypes bindings used by tests and
bench.py.  PyTorch is used for device memory, streams and torch.distributed only.
"""
from . import abi  # noqa: F401
from . import host  # noqa: F401
from .raycaster import Context, default_params, default_eep, make_volume_desc  # noqa: F401
