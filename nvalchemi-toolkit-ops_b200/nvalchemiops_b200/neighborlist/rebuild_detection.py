"""Rebuild detection (reference nvalchemiops/neighborlist/rebuild_detection.py:258-625).

Both checks are stateless functions of their tensor arguments, like the reference's: ``neighbor_list_needs_rebuild``
compares two position arrays, ``cell_list_needs_rebuild`` re-hashes the current positions on the grid
``(cell, pbc, cells_per_dimension)`` and compares with the ``atom_to_cell_mapping`` that ``build_cell_list`` wrote.
"""
from __future__ import annotations

import torch

from . import _engine


def cell_list_needs_rebuild(current_positions: torch.Tensor, atom_to_cell_mapping: torch.Tensor,
                            cells_per_dimension: torch.Tensor, cell: torch.Tensor, pbc: torch.Tensor,
                            batch_idx: torch.Tensor | None = None) -> torch.Tensor:
    """bool tensor [1]: True if any atom moved to a different cell of the grid ``build_cell_list`` used
    (reference :336-383; same signature, plus an optional ``batch_idx`` for caches of ``batch_build_cell_list``)."""
    device = current_positions.device
    if current_positions.shape[0] == 0:
        return torch.tensor([False], device=device, dtype=torch.bool)
    cell = cell if cell.ndim == 3 else cell.unsqueeze(0)
    return _engine.cells_changed_cache(current_positions, cell, pbc, batch_idx, cells_per_dimension,
                                       atom_to_cell_mapping).to(torch.bool)


def neighbor_list_needs_rebuild(reference_positions: torch.Tensor, current_positions: torch.Tensor,
                                skin_distance_threshold: float) -> torch.Tensor:
    """bool tensor [1]: True if any atom moved farther than ``skin_distance_threshold`` from its reference position
    (reference :457-503)."""
    device = current_positions.device
    if reference_positions.shape != current_positions.shape:
        return torch.tensor([True], device=device, dtype=torch.bool)  # reference :405-407
    if current_positions.shape[0] == 0:
        return torch.tensor([False], device=device, dtype=torch.bool)
    return _engine.moved_beyond(reference_positions, current_positions, skin_distance_threshold).to(torch.bool)


def check_cell_list_rebuild_needed(
    cells_per_dimension: torch.Tensor,
    neighbor_search_radius: torch.Tensor,
    atom_periodic_shifts: torch.Tensor,
    atom_to_cell_mapping: torch.Tensor,
    atoms_per_cell_count: torch.Tensor,
    cell_atom_start_indices: torch.Tensor,
    cell_atom_list: torch.Tensor,
    current_positions: torch.Tensor,
    current_cell: torch.Tensor,
    current_pbc: torch.Tensor,
    cutoff: float,
) -> bool:
    """Python-bool convenience wrapper with the reference's argument list (:505-576): the seven cache tensors of
    ``build_cell_list`` in their usual order, then the current positions / cell / pbc and the (unused) cutoff.  Only
    ``atom_to_cell_mapping`` and ``cells_per_dimension`` enter the check, as in the reference."""
    return bool(cell_list_needs_rebuild(current_positions, atom_to_cell_mapping, cells_per_dimension, current_cell,
                                        current_pbc).item())


def check_neighbor_list_rebuild_needed(reference_positions, current_positions, skin_distance_threshold) -> bool:
    """Python-bool convenience wrapper (reference :579-625)."""
    return bool(neighbor_list_needs_rebuild(reference_positions, current_positions, skin_distance_threshold).item())
