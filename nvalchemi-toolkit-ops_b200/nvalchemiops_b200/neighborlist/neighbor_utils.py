"""Host-side helpers of the neighbor-list path, mirroring nvalchemiops/neighborlist/neighbor_utils.py.

Same names, argument meaning and error behaviour as the reference:
``estimate_max_neighbors`` (:296-340), ``NeighborOverflowError`` (:343-349), ``assert_max_neighbors``
(:352-359), ``get_neighbor_list_from_neighbor_matrix`` (:362-441), ``_prepare_batch_idx_ptr`` (:444-491),
``allocate_cell_list`` (:494-539).  These are tensor plumbing (torch ops), not the hot path.
"""
from __future__ import annotations

import math

import torch


def estimate_max_neighbors(cutoff: float, atomic_density: float = 0.35, safety_factor: float = 5.0) -> int:
    """Upper bound on neighbors per atom: ceil(max(1, sf * rho * 4/3 pi rc^3) / 16) * 16."""
    if cutoff <= 0:
        return 0
    cutoff_sphere_volume = atomic_density * (4.0 / 3.0) * math.pi * (cutoff**3)
    expected_neighbors = max(1, safety_factor * cutoff_sphere_volume)
    return int(math.ceil(expected_neighbors / 16)) * 16


class NeighborOverflowError(Exception):
    """Raised when an atom has more neighbors than the padded matrix can hold."""

    def __init__(self, max_neighbors: int, num_neighbors: int):
        super().__init__(
            f"The number of neighbors is larger than the maximum allowed: {num_neighbors} > {max_neighbors}."
        )


def assert_max_neighbors(neighbor_matrix: torch.Tensor, num_neighbors: torch.Tensor):
    max_neighbors = 0 if num_neighbors.numel() == 0 else num_neighbors.max()
    if max_neighbors > neighbor_matrix.shape[1]:
        raise NeighborOverflowError(
            neighbor_matrix.shape[1],
            max_neighbors if isinstance(max_neighbors, int) else max_neighbors.item(),
        )


def get_neighbor_list_from_neighbor_matrix(
    neighbor_matrix: torch.Tensor,
    num_neighbors: torch.Tensor,
    neighbor_shift_matrix: torch.Tensor | None = None,
    fill_value: int = -1,
):
    """Padded matrix -> COO ``(neighbor_list [2,P], neighbor_ptr [N+1][, shifts [P,3]])``.

    Utility for matrices the caller already holds; ``cell_list(..., return_neighbor_list=True)`` does
    NOT go through it (the CUDA path writes COO directly).
    """
    if num_neighbors.shape[0] == 0:
        neighbor_list = torch.zeros(2, 0, dtype=neighbor_matrix.dtype, device=neighbor_matrix.device)
        neighbor_ptr = torch.zeros(1, dtype=torch.int32, device=neighbor_matrix.device)
        if neighbor_shift_matrix is None:
            return neighbor_list, neighbor_ptr
        shifts = torch.empty(0, 2, 3, dtype=neighbor_shift_matrix.dtype, device=neighbor_shift_matrix.device)
        return neighbor_list, neighbor_ptr, shifts
    assert_max_neighbors(neighbor_matrix, num_neighbors)
    mask = neighbor_matrix != fill_value
    dtype = neighbor_matrix.dtype
    i_idx = torch.where(mask)[0].to(dtype)
    j_idx = neighbor_matrix[mask].to(dtype)
    neighbor_list = torch.stack([i_idx, j_idx], dim=0)
    neighbor_ptr = torch.zeros(num_neighbors.shape[0] + 1, dtype=torch.int32, device=neighbor_matrix.device)
    torch.cumsum(num_neighbors, dim=0, out=neighbor_ptr[1:])
    if neighbor_shift_matrix is not None:
        return neighbor_list, neighbor_ptr, neighbor_shift_matrix[mask]
    return neighbor_list, neighbor_ptr


def _prepare_batch_idx_ptr(batch_idx, batch_ptr, num_atoms: int, device):
    """Derive whichever of ``batch_idx`` / ``batch_ptr`` is missing (reference :444-491)."""
    if batch_idx is None and batch_ptr is None:
        raise ValueError("Either batch_idx or batch_ptr must be provided.")
    if batch_idx is None:
        num_systems = batch_ptr.shape[0] - 1
        num_atoms_per_system = batch_ptr[1:] - batch_ptr[:-1]
        batch_idx = torch.repeat_interleave(
            torch.arange(num_systems, dtype=torch.int32, device=device), num_atoms_per_system
        )
    elif batch_ptr is None:
        num_systems = int(batch_idx.max()) + 1
        num_atoms_per_system = torch.bincount(batch_idx, minlength=num_systems)
        batch_ptr = torch.zeros(num_systems + 1, dtype=torch.int32, device=device)
        torch.cumsum(num_atoms_per_system, dim=0, out=batch_ptr[1:])
    return batch_idx, batch_ptr


def allocate_cell_list(total_atoms: int, max_total_cells: int, neighbor_search_radius: torch.Tensor, device):
    """Reference-shaped 7-tensor cell-list cache (reference :494-539).

    Kept for signature compatibility.  The CUDA path keeps its own opaque workspace
    (``nvnl_workspace_bytes``); of these tensors only ``cells_per_dimension`` and
    ``neighbor_search_radius`` are filled in (with the grid actually used).
    """
    cells_per_dimension = torch.zeros(
        (3,) if neighbor_search_radius.ndim == 1 else (neighbor_search_radius.shape[0], 3),
        dtype=torch.int32, device=device,
    )
    return (
        cells_per_dimension,
        neighbor_search_radius,
        torch.zeros((total_atoms, 3), dtype=torch.int32, device=device),
        torch.zeros((total_atoms, 3), dtype=torch.int32, device=device),
        torch.zeros((max_total_cells,), dtype=torch.int32, device=device),
        torch.zeros((max_total_cells,), dtype=torch.int32, device=device),
        torch.zeros((total_atoms,), dtype=torch.int32, device=device),
    )
