"""``batch_naive_neighbor_list`` — the reference's batched O(N^2) entry point
(nvalchemiops/neighborlist/batch_naive.py:480-763), served by the B200 cell-list engine (see naive.py)."""
from __future__ import annotations

import torch

from .naive import _check_cell_pbc, _naive_route
from .neighbor_utils import _prepare_batch_idx_ptr


def batch_naive_neighbor_list(
    positions: torch.Tensor,
    cutoff: float,
    batch_idx: torch.Tensor | None = None,
    batch_ptr: torch.Tensor | None = None,
    pbc: torch.Tensor | None = None,
    cell: torch.Tensor | None = None,
    max_neighbors: int | None = None,
    half_fill: bool = False,
    fill_value: int | None = None,
    return_neighbor_list: bool = False,
    neighbor_matrix: torch.Tensor | None = None,
    neighbor_matrix_shifts: torch.Tensor | None = None,
    num_neighbors: torch.Tensor | None = None,
    shift_range_per_dimension: torch.Tensor | None = None,
    shift_offset: torch.Tensor | None = None,
    total_shifts: int | None = None,
    max_atoms_per_system: int | None = None,
):
    """Neighbors within each system of a batch; pairs never cross systems.  Same contract as the reference: one of
    ``batch_idx`` / ``batch_ptr`` is enough (batch_naive.py:702-707), 2-tuples without PBC, 3-tuples with PBC, in-place
    reuse of pre-allocated tensors.  ``shift_*`` and ``max_atoms_per_system`` tune the reference's naive kernel launch
    and are accepted for signature compatibility."""
    _check_cell_pbc(cell, pbc)
    batch_idx, batch_ptr = _prepare_batch_idx_ptr(batch_idx=batch_idx, batch_ptr=batch_ptr, num_atoms=positions.shape[0],
                                                  device=positions.device)
    if pbc is not None:
        ns = int(batch_ptr.shape[0] - 1)
        cell = cell if cell.ndim == 3 else cell.unsqueeze(0)
        pbc = pbc if pbc.ndim == 2 else pbc.unsqueeze(0)
        if cell.shape[0] == 1 and ns > 1:      # one cell shared by every system
            cell = cell.repeat(ns, 1, 1)
        if pbc.shape[0] == 1 and ns > 1:
            pbc = pbc.repeat(ns, 1)
    return _naive_route(positions, cutoff, cell, pbc, batch_idx, batch_ptr, max_neighbors, half_fill, fill_value,
                        return_neighbor_list, neighbor_matrix, neighbor_matrix_shifts, num_neighbors)
