"""Neighbor-list API of the B200 path — same public names as ``nvalchemiops.neighborlist`` for the
accelerated cell-list route (reference nvalchemiops/neighborlist/__init__.py:16-74)."""
from .batch_cell_list import batch_cell_list, estimate_batch_cell_list_sizes
from .cell_list import cell_list, estimate_cell_list_sizes
from .neighbor_utils import (
    NeighborOverflowError,
    allocate_cell_list,
    estimate_max_neighbors,
    get_neighbor_list_from_neighbor_matrix,
)
from .neighborlist import neighbor_list

__all__ = [
    "NeighborOverflowError",
    "allocate_cell_list",
    "batch_cell_list",
    "cell_list",
    "estimate_batch_cell_list_sizes",
    "estimate_cell_list_sizes",
    "estimate_max_neighbors",
    "get_neighbor_list_from_neighbor_matrix",
    "neighbor_list",
]
