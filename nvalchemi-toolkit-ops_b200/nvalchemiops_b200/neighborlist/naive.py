"""``naive_neighbor_list`` — the reference's O(N^2) entry point (nvalchemiops/neighborlist/naive.py:400-706), served by
the B200 cell-list engine.

The neighbor set ``{(i, j, s) : |r_j - r_i + s.cell| < rc}`` does not depend on the search algorithm, so nothing
quadratic is built here: the same build + sweep kernels run, and this module reproduces what makes the reference's
naive route different for a caller — its signature (``pbc`` / ``cell`` optional), its return arity (2-tuples without
PBC), its cutoff rule (``cutoff**2`` in Python double, cast to the input precision; naive.py:290), its handling of
``cutoff <= 0`` and of pre-allocated buffers.  The ``shift_*`` arguments describe the naive kernel's image enumeration
and are accepted for signature compatibility; the engine derives the images from its grid.
"""
from __future__ import annotations

import torch

from . import _engine
from .cell_list import _run
from .neighbor_utils import estimate_max_neighbors


def _check_cell_pbc(cell, pbc):
    if pbc is None and cell is not None:
        raise ValueError("If cell is provided, pbc must also be provided")
    if pbc is not None and cell is None:
        raise ValueError("If pbc is provided, cell must also be provided")


def _open_or_given_cell(positions, cell, pbc, num_systems):
    """[S,3,3] cell and [S,3] pbc for the engine: the caller's, or one open unit cell per system."""
    dev = positions.device
    if pbc is None:
        cell_ = torch.eye(3, dtype=positions.dtype, device=dev).reshape(1, 3, 3).repeat(num_systems, 1, 1)
        pbc_ = torch.zeros((num_systems, 3), dtype=torch.bool, device=dev)
    else:
        cell_ = (cell if cell.ndim == 3 else cell.unsqueeze(0)).to(dev)
        pbc_ = (pbc if pbc.ndim == 2 else pbc.unsqueeze(0)).to(dev)
    return cell_, pbc_


def _padded_outputs(n, device, has_pbc, max_neighbors, fill_value, neighbor_matrix, neighbor_matrix_shifts,
                    num_neighbors):
    """The allocate-or-reset block of the reference (naive.py:578-612): used where no kernel runs."""
    if neighbor_matrix is None:
        neighbor_matrix = torch.full((n, max_neighbors), fill_value, dtype=torch.int32, device=device)
    else:
        neighbor_matrix.fill_(fill_value)
    if num_neighbors is None:
        num_neighbors = torch.zeros(n, dtype=torch.int32, device=device)
    else:
        num_neighbors.zero_()
    if has_pbc:
        if neighbor_matrix_shifts is None:
            neighbor_matrix_shifts = torch.zeros((n, neighbor_matrix.shape[1], 3), dtype=torch.int32, device=device)
        else:
            neighbor_matrix_shifts.zero_()
    return neighbor_matrix, num_neighbors, neighbor_matrix_shifts


def _naive_route(positions, cutoff, cell, pbc, batch_idx, batch_ptr, max_neighbors, half_fill, fill_value,
                 return_neighbor_list, neighbor_matrix, neighbor_matrix_shifts, num_neighbors):
    """Shared body of naive_neighbor_list / batch_naive_neighbor_list after their argument handling."""
    n, dev = positions.shape[0], positions.device
    has_pbc = pbc is not None
    if max_neighbors is None and (neighbor_matrix is None or (neighbor_matrix_shifts is None and has_pbc)
                                  or num_neighbors is None):
        max_neighbors = estimate_max_neighbors(cutoff)
    if fill_value is None:
        fill_value = n
    ns = int(batch_ptr.shape[0] - 1) if batch_ptr is not None else 1
    cell_, pbc_ = _open_or_given_cell(positions, cell, pbc, ns)
    if (cutoff <= 0 or n == 0) and not return_neighbor_list:
        # no kernel: the reference still returns the (N, max_neighbors) buffers it allocated or reset
        _engine._dtype_code(positions.dtype)
        _engine._require_cuda(positions, "positions")
        nm, num, sh = _padded_outputs(n, dev, has_pbc, max_neighbors, fill_value, neighbor_matrix,
                                      neighbor_matrix_shifts, num_neighbors)
        return (nm, num, sh) if has_pbc else (nm, num)
    csq = _engine.cutoff_sq_in_dtype(cutoff, positions.dtype, python_double=True) if cutoff > 0 else None
    out = _run(positions, cutoff, cell_, pbc_, batch_idx, batch_ptr, max_neighbors, half_fill, fill_value,
               return_neighbor_list, neighbor_matrix, neighbor_matrix_shifts, num_neighbors, None, empty_fill=fill_value,
               cutoff_sq=csq)
    return out if has_pbc else (out[0], out[1])


def naive_neighbor_list(
    positions: torch.Tensor,
    cutoff: float,
    cell: torch.Tensor | None = None,
    pbc: torch.Tensor | None = None,
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
):
    """Neighbors of one system; same contract as the reference (naive.py:400-706).

    Returns ``(neighbor_matrix, num_neighbors)`` / ``(neighbor_list, neighbor_ptr)`` without PBC and
    ``(neighbor_matrix, num_neighbors, neighbor_matrix_shifts)`` / ``(neighbor_list, neighbor_ptr, shifts)`` with PBC.
    Pre-allocated tensors are filled in place and returned.  With ``cutoff <= 0`` and ``return_neighbor_list`` the
    reference returns an extra all-zero ``[N]`` tensor in second place (naive.py:622-650); that tuple is reproduced.
    """
    _check_cell_pbc(cell, pbc)
    if cutoff <= 0 and return_neighbor_list:
        n, dev = positions.shape[0], positions.device
        head = (torch.zeros((2, 0), dtype=torch.int32, device=dev), torch.zeros((n,), dtype=torch.int32, device=dev),
                torch.zeros((n + 1,), dtype=torch.int32, device=dev))
        return head + (torch.zeros((0, 3), dtype=torch.int32, device=dev),) if pbc is not None else head
    return _naive_route(positions, cutoff, cell, pbc, None, None, max_neighbors, half_fill, fill_value,
                        return_neighbor_list, neighbor_matrix, neighbor_matrix_shifts, num_neighbors)
