"""``batch_cell_list`` — many independent systems in one launch sequence.

Mirrors ``nvalchemiops/neighborlist/batch_cell_list.py:1229-1468``.
"""
from __future__ import annotations

import torch

from . import _engine
from .cell_list import _attach, _find_handle, _run


def estimate_batch_cell_list_sizes(cell: torch.Tensor, pbc: torch.Tensor, cutoff: float, max_nbins: int = 1000):
    """Signature of batch_cell_list.py:659-736; see ``estimate_cell_list_sizes`` for what it means here."""
    from .cell_list import estimate_cell_list_sizes

    ns = cell.shape[0]
    if ns == 0 or cutoff <= 0:
        return 1, torch.zeros((ns, 3), device=cell.device, dtype=torch.int32)
    total, radii = 0, []
    for s in range(ns):
        n, r = estimate_cell_list_sizes(cell[s], pbc[s], cutoff, max_nbins)
        total += min(n, max_nbins)
        radii.append(r)
    return total, torch.stack(radii).to(cell.device)


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
    cache = {"cells_per_dimension": cells_per_dimension, "neighbor_search_radius": neighbor_search_radius}
    return _run(positions, cutoff, cell.reshape(-1, 3, 3), pbc.reshape(-1, 3), batch_idx, batch_ptr, max_neighbors,
                half_fill, fill_value, return_neighbor_list, neighbor_matrix, neighbor_matrix_shifts, num_neighbors,
                cache, empty_fill=empty_fill)


def batch_build_cell_list(positions, cutoff, cell, pbc, batch_idx, cells_per_dimension, neighbor_search_radius,
                          atom_periodic_shifts, atom_to_cell_mapping, atoms_per_cell_count, cell_atom_start_indices,
                          cell_atom_list) -> None:
    """Batched ``build_cell_list`` (reference batch_cell_list.py:1070-1135); see ``cell_list.build_cell_list``."""
    if positions.shape[0] == 0 or cutoff <= 0:
        return
    h = _engine.build(positions, cutoff, cell.reshape(-1, 3, 3), pbc.reshape(-1, 3), batch_idx=batch_idx)
    _engine.export_cache(h, cells_per_dimension, neighbor_search_radius, atom_periodic_shifts, atom_to_cell_mapping,
                         atoms_per_cell_count, cell_atom_start_indices, cell_atom_list)
    _attach(h, cells_per_dimension, atom_periodic_shifts, atom_to_cell_mapping, atoms_per_cell_count,
            cell_atom_start_indices, cell_atom_list)


def batch_query_cell_list(positions, cell, pbc, cutoff, batch_idx, cells_per_dimension, neighbor_search_radius,
                          atom_periodic_shifts, atom_to_cell_mapping, atoms_per_cell_count, cell_atom_start_indices,
                          cell_atom_list, neighbor_matrix, neighbor_matrix_shifts, num_neighbors, half_fill=False) -> None:
    """Batched ``query_cell_list`` — note the reference's argument order ``(positions, cell, pbc, cutoff, batch_idx, ...)``
    (batch_cell_list.py:1139-1144)."""
    if positions.shape[0] == 0 or cutoff <= 0:
        return
    h = _find_handle(cell_atom_list, atom_to_cell_mapping, atoms_per_cell_count, cell_atom_start_indices,
                     atom_periodic_shifts, cells_per_dimension)
    if cutoff > h.cutoff * (1.0 + 1e-12):
        raise ValueError(f"query cutoff {cutoff} exceeds the cutoff {h.cutoff} the cell list was built for")
    _engine.refresh_positions(h, positions)
    _engine.query_matrix(h, _engine.cutoff_sq_in_dtype(cutoff, positions.dtype), neighbor_matrix, neighbor_matrix_shifts,
                         num_neighbors, 0, half_fill, pad_rows=False)
