"""Neighbor-list API of the B200 path — same public names as ``nvalchemiops.neighborlist`` for the
accelerated cell-list route (reference nvalchemiops/neighborlist/__init__.py:16-74)."""
from .batch_cell_list import (batch_build_cell_list, batch_cell_list, batch_query_cell_list,
                              estimate_batch_cell_list_sizes)
from .cell_list import build_cell_list, cell_list, estimate_cell_list_sizes, query_cell_list
from .neighbor_utils import (
    NeighborOverflowError,
    allocate_cell_list,
    estimate_max_neighbors,
    get_neighbor_list_from_neighbor_matrix,
)
from . import ops  # noqa: F401  (registers the torch custom ops)
from .neighborlist import neighbor_list
from .rebuild_detection import (cell_list_needs_rebuild, check_cell_list_rebuild_needed,
                                check_neighbor_list_rebuild_needed, neighbor_list_needs_rebuild)

__all__ = [
    "NeighborOverflowError",
    "allocate_cell_list",
    "batch_build_cell_list",
    "batch_cell_list",
    "batch_query_cell_list",
    "build_cell_list",
    "cell_list",
    "cell_list_needs_rebuild",
    "check_cell_list_rebuild_needed",
    "check_neighbor_list_rebuild_needed",
    "estimate_batch_cell_list_sizes",
    "estimate_cell_list_sizes",
    "estimate_max_neighbors",
    "get_neighbor_list_from_neighbor_matrix",
    "neighbor_list",
    "neighbor_list_needs_rebuild",
    "query_cell_list",
]
