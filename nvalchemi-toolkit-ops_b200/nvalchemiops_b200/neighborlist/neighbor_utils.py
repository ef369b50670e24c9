"""Host-side helpers of the neighbor-list path — the contract of nvalchemiops/neighborlist/neighbor_utils.py
(``estimate_max_neighbors`` :296-340, ``NeighborOverflowError`` :343-349, ``assert_max_neighbors`` :352-359,
``get_neighbor_list_from_neighbor_matrix`` :362-441, ``_prepare_batch_idx_ptr`` :444-491, ``allocate_cell_list``
:494-539), written independently: same names, argument meaning, return shapes/dtypes and error behaviour.
Tensor plumbing (torch ops), not the hot path.
"""
from __future__ import annotations

import math

import torch

_SPHERE = 4.0 * math.pi / 3.0


def estimate_max_neighbors(cutoff: float, atomic_density: float = 0.35, safety_factor: float = 5.0) -> int:
    """Row width of the padded neighbor matrix: ``safety_factor`` times the atoms expected inside the cutoff sphere at
    number density ``atomic_density`` (at least one), rounded up to a multiple of 16; 0 for a non-positive cutoff."""
    if cutoff <= 0:
        return 0
    inside_sphere = _SPHERE * cutoff**3 * atomic_density
    want = safety_factor * inside_sphere
    if want < 1:
        want = 1
    return 16 * int(math.ceil(want / 16))


class NeighborOverflowError(Exception):
    """An atom has more neighbors than the padded matrix can hold (raised by the COO conversions only)."""

    def __init__(self, max_neighbors: int, num_neighbors: int):
        # message text kept identical to the reference's: callers match on it
        super().__init__(
            f"The number of neighbors is larger than the maximum allowed: {num_neighbors} > {max_neighbors}."
        )


def assert_max_neighbors(neighbor_matrix: torch.Tensor, num_neighbors: torch.Tensor):
    """Raise ``NeighborOverflowError`` if some atom's count exceeds the matrix width (one host sync)."""
    if num_neighbors.numel() == 0:
        return
    width = neighbor_matrix.shape[1]
    largest = int(num_neighbors.max().item())
    if largest > width:
        raise NeighborOverflowError(width, largest)


def get_neighbor_list_from_neighbor_matrix(
    neighbor_matrix: torch.Tensor,
    num_neighbors: torch.Tensor,
    neighbor_shift_matrix: torch.Tensor | None = None,
    fill_value: int = -1,
):
    """Padded matrix -> COO ``(neighbor_list [2,P], neighbor_ptr [N+1][, shifts [P,3]])``: every slot that does not
    hold ``fill_value`` becomes a pair, rows in order.  Utility for matrices the caller already holds;
    ``cell_list(..., return_neighbor_list=True)`` does NOT go through it (the CUDA path writes COO directly)."""
    device, dtype = neighbor_matrix.device, neighbor_matrix.dtype
    if num_neighbors.shape[0] == 0:
        empty = (torch.zeros((2, 0), dtype=dtype, device=device), torch.zeros((1,), dtype=torch.int32, device=device))
        if neighbor_shift_matrix is None:
            return empty
        return (*empty, torch.empty((0, 2, 3), dtype=neighbor_shift_matrix.dtype, device=neighbor_shift_matrix.device))
    assert_max_neighbors(neighbor_matrix, num_neighbors)
    src, slot = torch.nonzero(neighbor_matrix != fill_value, as_tuple=True)      # row-major: sources come out sorted
    neighbor_list = torch.stack((src.to(dtype), neighbor_matrix[src, slot]))
    neighbor_ptr = torch.cat((torch.zeros((1,), dtype=torch.int32, device=device),
                              torch.cumsum(num_neighbors, dim=0).to(torch.int32)))
    if neighbor_shift_matrix is None:
        return neighbor_list, neighbor_ptr
    return neighbor_list, neighbor_ptr, neighbor_shift_matrix[src, slot]


def _prepare_batch_idx_ptr(batch_idx, batch_ptr, num_atoms: int, device):
    """Return ``(batch_idx [N], batch_ptr [S+1])``, deriving the missing one from the other; ``ValueError`` when both
    are missing (same message as the reference)."""
    if batch_idx is None and batch_ptr is None:
        raise ValueError("Either batch_idx or batch_ptr must be provided.")
    if batch_idx is None:
        counts = torch.diff(batch_ptr)
        systems = torch.arange(counts.shape[0], dtype=torch.int32, device=device)
        batch_idx = systems.repeat_interleave(counts)
    elif batch_ptr is None:
        num_systems = int(batch_idx.max().item()) + 1
        counts = torch.bincount(batch_idx, minlength=num_systems)
        batch_ptr = torch.cat((torch.zeros((1,), dtype=torch.int32, device=device),
                               torch.cumsum(counts, dim=0).to(torch.int32)))
    return batch_idx, batch_ptr


def compute_naive_num_shifts(cell: torch.Tensor, cutoff: float, pbc: torch.Tensor):
    """``(shift_range [S,3] int32, shift_offset [S+1] int32, total_shifts int)`` — how many periodic images the
    reference's naive kernels enumerate per system (neighbor_utils.py:151-211, 233-293): per periodic dimension
    ``s_d = ceil(cutoff * |row d of inverse(cell)^T|)`` (= cutoff / face distance), 0 in open dimensions, and the
    half-space count ``s_0 (2 s_1 + 1)(2 s_2 + 1) + s_1 (2 s_2 + 1) + s_2 + 1``.  The B200 engine derives its images
    from the cell grid and does not need these; the function exists for callers that pre-compute them
    (``naive_neighbor_list(..., shift_range_per_dimension=, shift_offset=, total_shifts=)``).  Tensor arithmetic on
    ``cell.device`` in ``cell.dtype``; the one ``.item()`` is the reference's."""
    c = cell.reshape(-1, 3, 3)
    a0, a1, a2 = c[:, 0], c[:, 1], c[:, 2]
    # rows of inverse(cell)^T are the reciprocal vectors (a_j x a_k) / det
    r0, r1, r2 = torch.linalg.cross(a1, a2), torch.linalg.cross(a2, a0), torch.linalg.cross(a0, a1)
    det = (a0 * r0).sum(dim=1)
    inv_len = torch.stack([r0.norm(dim=1), r1.norm(dim=1), r2.norm(dim=1)], dim=1) / det.abs()[:, None]
    periodic = pbc.reshape(-1, 3).to(device=c.device, dtype=torch.bool)
    inv_len = torch.where(periodic, inv_len, torch.zeros_like(inv_len))
    s = torch.ceil(inv_len * torch.tensor(cutoff, dtype=c.dtype, device=c.device)).to(torch.int32)
    k1, k2 = 2 * s[:, 1] + 1, 2 * s[:, 2] + 1
    num_shifts = s[:, 0] * k1 * k2 + s[:, 1] * k2 + s[:, 2] + 1
    shift_offset = torch.zeros((c.shape[0] + 1,), dtype=torch.int32, device=c.device)
    shift_offset[1:] = torch.cumsum(num_shifts, dim=0)
    return s, shift_offset, int(shift_offset[-1].item())


def allocate_cell_list(total_atoms: int, max_total_cells: int, neighbor_search_radius: torch.Tensor, device):
    """The seven-tensor cell-list cache, zero-filled int32, in the reference's order: ``cells_per_dimension`` ([3], or
    [S,3] when ``neighbor_search_radius`` is [S,3]), ``neighbor_search_radius`` (passed through),
    ``atom_periodic_shifts`` [N,3], ``atom_to_cell_mapping`` [N,3], ``atoms_per_cell_count`` [C],
    ``cell_atom_start_indices`` [C], ``cell_atom_list`` [N].  ``build_cell_list`` / ``batch_build_cell_list`` fill all
    of them and keep the grid within the ``C = max_total_cells`` cells allocated here."""
    def zeros(*shape):
        return torch.zeros(shape, dtype=torch.int32, device=device)

    per_system = neighbor_search_radius.ndim > 1
    cpd = zeros(neighbor_search_radius.shape[0], 3) if per_system else zeros(3)
    return (cpd, neighbor_search_radius, zeros(total_atoms, 3), zeros(total_atoms, 3), zeros(max_total_cells),
            zeros(max_total_cells), zeros(total_atoms))
