"""torch custom op of the padded-matrix path — the counterpart of the reference's ``nvalchemiops::build_cell_list`` +
``nvalchemiops::query_cell_list`` (+ ``batch_`` variants) custom ops (cell_list.py:725-749, 892-912;
batch_cell_list.py:739-763, 915-936): mutation-only, returns None, so ``torch.compile`` can keep it in the graph with
pre-allocated outputs (reference test_cell_list.py:598-844).  Registered under this package's own namespace — the
reference registers ``nvalchemiops::*`` at import and double registration raises.

    nvalchemiops_b200::neighbor_matrix(Tensor positions, float cutoff, Tensor cell, Tensor pbc, Tensor? batch_idx,
        Tensor? batch_ptr, Tensor(a!) neighbor_matrix, Tensor(b!) neighbor_matrix_shifts, Tensor(c!) num_neighbors,
        int fill_value, bool half_fill, float cutoff_sq) -> ()

One call = grid + hash + counting sort + fused sweep/fill (no host sync, every output slot written once).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _engine


@torch.library.custom_op(
    "nvalchemiops_b200::neighbor_matrix",
    mutates_args=("neighbor_matrix", "neighbor_matrix_shifts", "num_neighbors"),
)
def neighbor_matrix_op(
    positions: torch.Tensor,
    cutoff: float,
    cell: torch.Tensor,
    pbc: torch.Tensor,
    batch_idx: Optional[torch.Tensor],
    batch_ptr: Optional[torch.Tensor],
    neighbor_matrix: torch.Tensor,
    neighbor_matrix_shifts: torch.Tensor,
    num_neighbors: torch.Tensor,
    fill_value: int,
    half_fill: bool,
    cutoff_sq: float,
) -> None:
    h = _engine.build(positions, cutoff, cell, pbc, batch_idx=batch_idx, batch_ptr=batch_ptr)
    _engine.query_matrix(h, cutoff_sq, neighbor_matrix, neighbor_matrix_shifts, num_neighbors, fill_value, half_fill)
