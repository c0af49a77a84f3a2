"""lintransunet_b200 -- B200-native (sm_100a) forward hot path of LinTransUNet's MaskTransUnet.

    from lintransunet_b200 import MaskTransUnet, get_model_dict

The package contains the C-ABI kernel library (csrc/, include/ltu_b200.h), its ctypes binding
(_native.py, ops.py), the drop-in nn.Module (unet.py) and the patch-sharded sliding-window
inference driver (sliding_window.py).  Importing the package does not need a GPU; running it
does, and the native library must have been built (python -m lintransunet_b200.build).
"""
from .unet import MaskTransUnet, Model_Dict, get_model_dict  # noqa: F401

__version__ = "0.1.0"
