"""Rebuild detection (reference nvalchemiops/neighborlist/rebuild_detection.py:258-625).

``neighbor_list_needs_rebuild`` is a stateless comparison of two position arrays.  ``cell_list_needs_rebuild`` asks
whether any atom left the cell it was binned into; it needs the cell list of the last ``build_cell_list`` call, which
this package keeps in an opaque workspace attached to the cache tensors that call filled in.
"""
from __future__ import annotations

import torch

from . import _engine
from .cell_list import _find_handle


def cell_list_needs_rebuild(current_positions: torch.Tensor, atom_to_cell_mapping: torch.Tensor,
                            cells_per_dimension: torch.Tensor, cell: torch.Tensor, pbc: torch.Tensor) -> torch.Tensor:
    """bool tensor [1]: True if any atom moved to a different cell of the grid used by the last build
    (reference :336-383; same signature).  ``atom_to_cell_mapping`` must be the tensor ``build_cell_list`` filled."""
    device = current_positions.device
    if current_positions.shape[0] == 0:
        return torch.tensor([False], device=device, dtype=torch.bool)
    h = _find_handle(atom_to_cell_mapping, cells_per_dimension)
    return _engine.cells_changed(h, current_positions).to(torch.bool)


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


def check_cell_list_rebuild_needed(current_positions, atom_to_cell_mapping, cells_per_dimension, cell, pbc) -> bool:
    """Python-bool convenience wrapper (reference :505-576)."""
    return bool(cell_list_needs_rebuild(current_positions, atom_to_cell_mapping, cells_per_dimension, cell, pbc).item())


def check_neighbor_list_rebuild_needed(reference_positions, current_positions, skin_distance_threshold) -> bool:
    """Python-bool convenience wrapper (reference :579-625)."""
    return bool(neighbor_list_needs_rebuild(reference_positions, current_positions, skin_distance_threshold).item())
