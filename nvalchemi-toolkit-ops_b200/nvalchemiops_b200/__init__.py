"""nvalchemiops_b200 — B200-native (sm_100a) cell-list neighbor lists.

Drop-in for ONE path of NVIDIA/nvalchemi-toolkit-ops: ``nvalchemiops.neighborlist.neighbor_list`` ->
``cell_list`` / ``batch_cell_list`` (reference: nvalchemiops/neighborlist/neighborlist.py:41-310).
Hand-written CUDA behind a C ABI (``include/nvalchemi_nl_b200.h``); PyTorch is used only for device
memory, streams and ``torch.distributed``.  No Warp, no Triton, no CPU fallback: every entry point
raises if the CUDA library or a CUDA device is missing.
"""
from . import config  # noqa: F401
from ._lib import library_path, launch_count  # noqa: F401

__version__ = "0.1.0"
