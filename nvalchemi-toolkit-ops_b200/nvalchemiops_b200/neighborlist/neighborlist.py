"""``neighbor_list`` — the unified entry point (reference nvalchemiops/neighborlist/neighborlist.py:41-310).

Same signature, method auto-selection, kwargs forwarding and return arity as the reference.  Every method name
that the reference routes to an O(N^2) "naive" Warp kernel goes to the public function of the same name
(naive.py, batch_naive.py, naive_dual_cutoff.py, batch_naive_dual_cutoff.py), all served by the same B200 cell-list
engine (the neighbor set does not depend on the search algorithm); only the return arity differs (2-tuples without
PBC).  The dual-cutoff methods build one cell list with the larger cutoff and query it twice.
"""
from __future__ import annotations

import torch

from .batch_cell_list import batch_cell_list
from .batch_naive import batch_naive_neighbor_list
from .batch_naive_dual_cutoff import batch_naive_neighbor_list_dual_cutoff
from .cell_list import cell_list
from .naive import naive_neighbor_list
from .naive_dual_cutoff import naive_neighbor_list_dual_cutoff
from .neighbor_utils import _prepare_batch_idx_ptr


def neighbor_list(
    positions: torch.Tensor,
    cutoff: float,
    cell: torch.Tensor | None = None,
    pbc: torch.Tensor | None = None,
    batch_idx: torch.Tensor | None = None,
    batch_ptr: torch.Tensor | None = None,
    cutoff2: float | None = None,
    half_fill: bool = False,
    fill_value: int | None = None,
    return_neighbor_list: bool = False,
    method: str | None = None,
    **kwargs,
):
    """Compute neighbor lists; see the reference docstring (neighborlist.py:55-211) for the full contract.

    Returns
    -------
    * no PBC (naive methods): ``(neighbor_matrix, num_neighbors)`` or ``(neighbor_list, neighbor_ptr)``
    * with PBC / cell-list methods: ``(neighbor_matrix, num_neighbors, neighbor_matrix_shifts)`` or
      ``(neighbor_list, neighbor_ptr, neighbor_list_shifts)``
    """
    if method is None:
        total_atoms = positions.shape[0]
        if cutoff2 is not None:
            method = "naive_dual_cutoff"
        elif total_atoms >= 5000:
            method = "cell_list"
            if cell is None or pbc is None:
                cell = torch.eye(3, dtype=positions.dtype, device=positions.device).reshape(1, 3, 3)
                pbc = torch.tensor([False, False, False], dtype=torch.bool, device=positions.device)
        else:
            method = "naive"
        if batch_idx is not None or batch_ptr is not None:
            method = "batch_" + method
            batch_idx, batch_ptr = _prepare_batch_idx_ptr(batch_idx, batch_ptr, total_atoms, positions.device)
    match method:
        case "naive":
            return naive_neighbor_list(positions, cutoff, pbc=pbc, cell=cell, half_fill=half_fill, fill_value=fill_value,
                                       return_neighbor_list=return_neighbor_list, **kwargs)
        case "cell_list":
            return cell_list(positions, cutoff, cell, pbc, half_fill=half_fill, fill_value=fill_value,
                             return_neighbor_list=return_neighbor_list, **kwargs)
        case "batch_naive":
            return batch_naive_neighbor_list(positions, cutoff, pbc=pbc, cell=cell, batch_idx=batch_idx,
                                             batch_ptr=batch_ptr, half_fill=half_fill, fill_value=fill_value,
                                             return_neighbor_list=return_neighbor_list, **kwargs)
        case "batch_cell_list":
            if cell is None or pbc is None:
                # auto-selected batch_cell_list without a cell (>= 5000 atoms): one open unit cell per system
                ns = int(batch_ptr.shape[0] - 1)
                cell = torch.eye(3, dtype=positions.dtype, device=positions.device).reshape(1, 3, 3).repeat(ns, 1, 1)
                pbc = torch.zeros((ns, 3), dtype=torch.bool, device=positions.device)
            elif cell.ndim == 3 and cell.shape[0] == 1 and batch_ptr is not None and batch_ptr.shape[0] - 1 > 1:
                ns = int(batch_ptr.shape[0] - 1)
                cell = cell.repeat(ns, 1, 1)
                pbc = pbc.reshape(1, 3).repeat(ns, 1)
            return batch_cell_list(positions, cutoff, cell, pbc, batch_idx, half_fill=half_fill,
                                   fill_value=fill_value, return_neighbor_list=return_neighbor_list,
                                   batch_ptr=batch_ptr, **kwargs)
        case "naive_dual_cutoff":
            return naive_neighbor_list_dual_cutoff(positions, cutoff, cutoff2, pbc=pbc, cell=cell, half_fill=half_fill,
                                                   fill_value=fill_value, return_neighbor_list=return_neighbor_list,
                                                   **kwargs)
        case "batch_naive_dual_cutoff":
            return batch_naive_neighbor_list_dual_cutoff(positions, cutoff, cutoff2, pbc=pbc, cell=cell,
                                                         batch_idx=batch_idx, batch_ptr=batch_ptr, half_fill=half_fill,
                                                         fill_value=fill_value,
                                                         return_neighbor_list=return_neighbor_list, **kwargs)
        case _:
            raise ValueError(f"Invalid method: {method}")
