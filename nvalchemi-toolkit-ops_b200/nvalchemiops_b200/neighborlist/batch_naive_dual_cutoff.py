"""``batch_naive_neighbor_list_dual_cutoff`` — the batched dual-cutoff entry point
(reference nvalchemiops/neighborlist/batch_naive_dual_cutoff.py:592-1000), served by the B200 cell-list engine:
one batched build with the larger cutoff, one sweep per cutoff (see naive_dual_cutoff.py)."""
from __future__ import annotations

import torch

from .naive import _check_cell_pbc
from .naive_dual_cutoff import _dual_route
from .neighbor_utils import _prepare_batch_idx_ptr


def batch_naive_neighbor_list_dual_cutoff(
    positions: torch.Tensor,
    cutoff1: float,
    cutoff2: float,
    batch_idx: torch.Tensor | None = None,
    batch_ptr: torch.Tensor | None = None,
    pbc: torch.Tensor | None = None,
    cell: torch.Tensor | None = None,
    max_neighbors1: int | None = None,
    max_neighbors2: int | None = None,
    half_fill: bool = False,
    fill_value: int | None = None,
    return_neighbor_list: bool = False,
    neighbor_matrix1: torch.Tensor | None = None,
    neighbor_matrix2: torch.Tensor | None = None,
    neighbor_matrix_shifts1: torch.Tensor | None = None,
    neighbor_matrix_shifts2: torch.Tensor | None = None,
    num_neighbors1: torch.Tensor | None = None,
    num_neighbors2: torch.Tensor | None = None,
    shift_range_per_dimension: torch.Tensor | None = None,
    shift_offset: torch.Tensor | None = None,
    total_shifts: int | None = None,
    max_atoms_per_system: int | None = None,
):
    """Per-system neighbors within ``cutoff1`` and ``cutoff2`` for a batch; same contract as the reference
    (4-tuple without PBC, 6-tuple with PBC; one of ``batch_idx`` / ``batch_ptr`` is enough)."""
    _check_cell_pbc(cell, pbc)
    batch_idx, batch_ptr = _prepare_batch_idx_ptr(batch_idx=batch_idx, batch_ptr=batch_ptr, num_atoms=positions.shape[0],
                                                  device=positions.device)
    if pbc is not None:
        ns = int(batch_ptr.shape[0] - 1)
        cell = cell if cell.ndim == 3 else cell.unsqueeze(0)
        pbc = pbc if pbc.ndim == 2 else pbc.unsqueeze(0)
        if cell.shape[0] == 1 and ns > 1:
            cell = cell.repeat(ns, 1, 1)
        if pbc.shape[0] == 1 and ns > 1:
            pbc = pbc.repeat(ns, 1)
    return _dual_route(positions, cutoff1, cutoff2, cell, pbc, batch_idx, batch_ptr, max_neighbors1, max_neighbors2,
                       half_fill, fill_value, return_neighbor_list, neighbor_matrix1, neighbor_matrix2,
                       neighbor_matrix_shifts1, neighbor_matrix_shifts2, num_neighbors1, num_neighbors2)
