"""``batch_cell_list`` — many independent systems in one launch sequence.

Mirrors ``nvalchemiops/neighborlist/batch_cell_list.py:1229-1468``.
"""
from __future__ import annotations

import torch

from . import _engine
from .cell_list import _CACHE_KEYS, _estimate_grid, _run


def estimate_batch_cell_list_sizes(cell: torch.Tensor, pbc: torch.Tensor, cutoff: float, max_nbins: int = 1000):
    """``(max_total_cells, neighbor_search_radius [S,3] int32)`` — batch_cell_list.py:659-736: every system's grid is
    halved until it has at most ``max_nbins`` cells, the counts are summed.  One ``.item()`` sync for the whole batch
    (no per-system sync).  See ``estimate_cell_list_sizes`` for how the result is used."""
    ns = cell.shape[0]
    if ns == 0 or cutoff <= 0:
        return 1, torch.zeros((ns, 3), device=cell.device, dtype=torch.int32)
    _engine._dtype_code(cell.dtype)
    cells, radius = _estimate_grid(cell, pbc, cutoff, max_nbins)
    return int(cells.sum().item()), radius


def batch_cell_list(
    positions: torch.Tensor,
    cutoff: float,
    cell: torch.Tensor,
    pbc: torch.Tensor,
    batch_idx: torch.Tensor,
    max_neighbors: int | None = None,
    half_fill: bool = False,
    fill_value: int | None = None,
    return_neighbor_list: bool = False,
    neighbor_matrix: torch.Tensor | None = None,
    neighbor_matrix_shifts: torch.Tensor | None = None,
    num_neighbors: torch.Tensor | None = None,
    cells_per_dimension: torch.Tensor | None = None,
    neighbor_search_radius: torch.Tensor | None = None,
    atom_periodic_shifts: torch.Tensor | None = None,
    atom_to_cell_mapping: torch.Tensor | None = None,
    atoms_per_cell_count: torch.Tensor | None = None,
    cell_atom_start_indices: torch.Tensor | None = None,
    cell_atom_list: torch.Tensor | None = None,
    batch_ptr: torch.Tensor | None = None,
):
    """Neighbor lists of a batch of systems (reference batch_cell_list.py:1229-1468).

    ``cell`` [S,3,3], ``pbc`` [S,3], ``batch_idx`` [N] (atoms of a system need not be contiguous).  ``batch_ptr``
    is an optional extra (the dispatcher has it): it saves one counting pass.  Atoms only pair within their own
    system.  Returns the same tuples as ``cell_list``.
    """
    total_atoms = positions.shape[0]
    empty_fill = -1  # batch_cell_list.py:1369
    if fill_value is None:
        fill_value = total_atoms
    cache = dict(zip(_CACHE_KEYS, (cells_per_dimension, neighbor_search_radius, atom_periodic_shifts, atom_to_cell_mapping,
                                   atoms_per_cell_count, cell_atom_start_indices, cell_atom_list)))
    return _run(positions, cutoff, cell.reshape(-1, 3, 3), pbc.reshape(-1, 3), batch_idx, batch_ptr, max_neighbors,
                half_fill, fill_value, return_neighbor_list, neighbor_matrix, neighbor_matrix_shifts, num_neighbors,
                cache, empty_fill=empty_fill)


def batch_build_cell_list(positions, cutoff, cell, pbc, batch_idx, cells_per_dimension, neighbor_search_radius,
                          atom_periodic_shifts, atom_to_cell_mapping, atoms_per_cell_count, cell_atom_start_indices,
                          cell_atom_list) -> None:
    """Batched ``build_cell_list`` (reference batch_cell_list.py:1070-1135): every system's grid has at most
    ``atoms_per_cell_count.numel() // num_systems`` cells (batch_cell_list.py:162-176)."""
    from .ops import batch_build_cell_list_op

    _engine._require_cuda(positions, "positions")
    batch_build_cell_list_op(positions, float(cutoff), cell, pbc, batch_idx, cells_per_dimension, neighbor_search_radius,
                             atom_periodic_shifts, atom_to_cell_mapping, atoms_per_cell_count, cell_atom_start_indices,
                             cell_atom_list)


def batch_query_cell_list(positions, cell, pbc, cutoff, batch_idx, cells_per_dimension, neighbor_search_radius,
                          atom_periodic_shifts, atom_to_cell_mapping, atoms_per_cell_count, cell_atom_start_indices,
                          cell_atom_list, neighbor_matrix, neighbor_matrix_shifts, num_neighbors, half_fill=False) -> None:
    """Batched ``query_cell_list`` — note the reference's argument order ``(positions, cell, pbc, cutoff, batch_idx, ...)``
    (batch_cell_list.py:1139-1144)."""
    from .ops import batch_query_cell_list_op

    _engine._require_cuda(positions, "positions")
    batch_query_cell_list_op(positions, cell, pbc, float(cutoff), batch_idx, cells_per_dimension, neighbor_search_radius,
                             atom_periodic_shifts, atom_to_cell_mapping, atoms_per_cell_count, cell_atom_start_indices,
                             cell_atom_list, neighbor_matrix, neighbor_matrix_shifts, num_neighbors, bool(half_fill))
