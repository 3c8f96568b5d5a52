"""multi_view_sort_b200 - B200-native (sm_100a) VMM / Neural-EM latent-view decomposition.

A drop-in for the hot path of hjjpku/multi_view_sort (train_nem.py `NEM`,
`Multi_View_Net`, tools/NEM_utilities.py): same constructor options and state_dict, with
every EM iteration, the E-step, the outer loss, the head and all their gradients computed by
hand-written CUDA (TMA + tcgen05) behind the C-ABI in `include/vmm_b200.h`.

There is no CPU fallback: any compute call raises if `libvmm_b200.so` is not built
(`python -m multi_view_sort_b200.build`) or no sm_100 GPU is present.
"""
from .model import Model
from .nem import NEM, Multi_View_Net
from . import nem_utilities

__all__ = ["Model", "NEM", "Multi_View_Net", "nem_utilities"]
__version__ = "0.1.0"
