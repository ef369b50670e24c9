"""Neighbor-list API of the B200 path — same public names as ``nvalchemiops.neighborlist`` for the
accelerated cell-list route (reference nvalchemiops/neighborlist/__init__.py:16-74): every name of the reference's
``__all__`` is here, with the reference's signature."""
from .batch_cell_list import (batch_build_cell_list, batch_cell_list, batch_query_cell_list,
                              estimate_batch_cell_list_sizes)
from .batch_naive import batch_naive_neighbor_list
from .batch_naive_dual_cutoff import batch_naive_neighbor_list_dual_cutoff
from .cell_list import build_cell_list, cell_list, estimate_cell_list_sizes, query_cell_list
from .naive import naive_neighbor_list
from .naive_dual_cutoff import naive_neighbor_list_dual_cutoff
from .neighbor_utils import (
    NeighborOverflowError,
    allocate_cell_list,
    compute_naive_num_shifts,
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
    "batch_naive_neighbor_list",
    "batch_naive_neighbor_list_dual_cutoff",
    "batch_query_cell_list",
    "build_cell_list",
    "cell_list",
    "cell_list_needs_rebuild",
    "check_cell_list_rebuild_needed",
    "check_neighbor_list_rebuild_needed",
    "compute_naive_num_shifts",
    "estimate_batch_cell_list_sizes",
    "estimate_cell_list_sizes",
    "estimate_max_neighbors",
    "get_neighbor_list_from_neighbor_matrix",
    "naive_neighbor_list",
    "naive_neighbor_list_dual_cutoff",
    "neighbor_list",
    "neighbor_list_needs_rebuild",
    "query_cell_list",
]
