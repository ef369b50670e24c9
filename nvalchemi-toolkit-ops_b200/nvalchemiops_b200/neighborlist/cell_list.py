"""``cell_list`` — single-system O(N) neighbor list on the B200 CUDA path.

Mirrors ``nvalchemiops/neighborlist/cell_list.py:1195-1443`` (signature, defaults, return tuples, empty
cases, in-place reuse of caller buffers); the work is done by ``libnvalchemi_nl_b200.so``.
"""
from __future__ import annotations

import torch

from . import _engine
from .neighbor_utils import estimate_max_neighbors, get_neighbor_list_from_neighbor_matrix

_CACHE_KEYS = ("cells_per_dimension", "neighbor_search_radius", "atom_periodic_shifts", "atom_to_cell_mapping",
               "atoms_per_cell_count", "cell_atom_start_indices", "cell_atom_list")


def _run(positions, cutoff, cell, pbc, batch_idx, batch_ptr, max_neighbors, half_fill, fill_value,
         return_neighbor_list, neighbor_matrix, neighbor_matrix_shifts, num_neighbors, cache, empty_fill,
         cutoff_sq=None):
    """Shared driver of cell_list / batch_cell_list / the dispatcher's naive routes."""
    total_atoms = positions.shape[0]
    device = positions.device
    _engine._dtype_code(positions.dtype)
    _engine._require_cuda(positions, "positions")

    # ---- empty cases: same shapes/dtypes as cell_list.py:1335-1349 / batch_cell_list.py:1358-1371 ----
    if total_atoms <= 0 or cutoff <= 0:
        if return_neighbor_list:
            return (
                torch.zeros((2, 0), dtype=torch.int32, device=device),
                torch.zeros((total_atoms + 1,), dtype=torch.int32, device=device),
                torch.zeros((0, 3), dtype=torch.int32, device=device),
            )
        return (
            torch.full((total_atoms, 0), empty_fill, dtype=torch.int32, device=device),
            torch.zeros((total_atoms,), dtype=torch.int32, device=device),
            torch.zeros((total_atoms, 0, 3), dtype=torch.int32, device=device),
        )

    if cutoff_sq is None:
        cutoff_sq = _engine.cutoff_sq_in_dtype(cutoff, positions.dtype)
    # The reference uses caller-provided cache tensors only when ALL seven are given (cell_list.py:1376-1410): it
    # zeroes them, builds into them and returns the same objects' contents.  Same here: the grid is then capped at
    # their capacity and all seven are filled.  Otherwise the library sizes its own workspace.
    full_cache = cache is not None and all(cache.get(k) is not None for k in _CACHE_KEYS)
    matrix_only = not (return_neighbor_list and neighbor_matrix is None)
    if matrix_only and not full_cache:
        # padded-matrix outputs: one mutation-only custom op (torch.compile keeps it in the graph, no host sync)
        user_buffers = neighbor_matrix is not None and neighbor_matrix_shifts is not None and num_neighbors is not None
        if max_neighbors is None and not user_buffers:
            max_neighbors = estimate_max_neighbors(cutoff)
        if neighbor_matrix is None:
            neighbor_matrix = torch.empty((total_atoms, max_neighbors), dtype=torch.int32, device=device)
        M = neighbor_matrix.shape[1]
        if neighbor_matrix_shifts is None:
            neighbor_matrix_shifts = torch.empty((total_atoms, M, 3), dtype=torch.int32, device=device)
        if num_neighbors is None:
            num_neighbors = torch.empty((total_atoms,), dtype=torch.int32, device=device)
        from .ops import neighbor_matrix_op

        neighbor_matrix_op(positions, float(cutoff), cell, pbc, batch_idx, batch_ptr, neighbor_matrix, neighbor_matrix_shifts,
                           num_neighbors, int(fill_value), bool(half_fill), float(cutoff_sq))
        if return_neighbor_list:
            return get_neighbor_list_from_neighbor_matrix(
                neighbor_matrix, num_neighbors=num_neighbors, neighbor_shift_matrix=neighbor_matrix_shifts,
                fill_value=fill_value,
            )
        return neighbor_matrix, num_neighbors, neighbor_matrix_shifts
    if full_cache:
        tensors = [cache[k] for k in _CACHE_KEYS]
        capacity = min(cache["atoms_per_cell_count"].numel(), cache["cell_atom_start_indices"].numel())
        h = _engine.build(positions, cutoff, cell, pbc, batch_idx=batch_idx, batch_ptr=batch_ptr, max_cells=capacity)
        _engine.export_cache(h, *tensors)
    else:
        h = _engine.build(positions, cutoff, cell, pbc, batch_idx=batch_idx, batch_ptr=batch_ptr)
    return _query(h, cutoff, cutoff_sq, max_neighbors, half_fill, fill_value, return_neighbor_list, neighbor_matrix,
                  neighbor_matrix_shifts, num_neighbors)


def _query(h, cutoff, cutoff_sq, max_neighbors, half_fill, fill_value, return_neighbor_list, neighbor_matrix,
           neighbor_matrix_shifts, num_neighbors):
    """One query of a built cell list: padded matrix (in place when buffers are given) or direct COO."""
    total_atoms, device = h.n, h.device
    user_buffers = neighbor_matrix is not None and neighbor_matrix_shifts is not None and num_neighbors is not None
    if max_neighbors is None and not user_buffers:
        max_neighbors = estimate_max_neighbors(cutoff)
    if return_neighbor_list and neighbor_matrix is None:
        # direct COO path: the padded matrix is never materialised
        neighbor_list, neighbor_ptr, shifts, _num = _engine.query_coo(h, cutoff_sq, half_fill, max_neighbors)
        return neighbor_list, neighbor_ptr, shifts

    if neighbor_matrix is None:
        neighbor_matrix = torch.empty((total_atoms, max_neighbors), dtype=torch.int32, device=device)
    M = neighbor_matrix.shape[1]
    if neighbor_matrix_shifts is None:
        neighbor_matrix_shifts = torch.empty((total_atoms, M, 3), dtype=torch.int32, device=device)
    if num_neighbors is None:
        num_neighbors = torch.empty((total_atoms,), dtype=torch.int32, device=device)
    # every slot (hits, padding with fill_value, zero shifts) is written by the kernel: the reference's
    # fill_()/zero_() of the outputs (cell_list.py:1358-1373) is fused into the sweep.
    _engine.query_matrix(h, cutoff_sq, neighbor_matrix, neighbor_matrix_shifts, num_neighbors, fill_value, half_fill)
    if return_neighbor_list:
        return get_neighbor_list_from_neighbor_matrix(
            neighbor_matrix, num_neighbors=num_neighbors, neighbor_shift_matrix=neighbor_matrix_shifts,
            fill_value=fill_value,
        )
    return neighbor_matrix, num_neighbors, neighbor_matrix_shifts


def _estimate_grid(cell: torch.Tensor, pbc: torch.Tensor, cutoff: float, max_nbins: int):
    """Per-system allocated cell count [S] (int64) and search radius [S,3] (int32) with the reference's arithmetic
    (_estimate_cell_list_sizes, cell_list.py:35-99 / batch_cell_list.py:35-99): face distances from the adjugate
    inverse in the input precision, cells per dimension = max(trunc(face / cutoff), 1), radius from the UN-halved grid,
    then all dimensions halved until the product fits ``max_nbins``.  Tensor ops on the input device, no host sync."""
    c = cell.reshape(-1, 3, 3)
    a, b, cc = c[:, 0, 0], c[:, 0, 1], c[:, 0, 2]
    d, e, f = c[:, 1, 0], c[:, 1, 1], c[:, 1, 2]
    g, h, i = c[:, 2, 0], c[:, 2, 1], c[:, 2, 2]
    det = a * (e * i - f * h) - b * (d * i - f * g) + cc * (d * h - e * g)
    # columns of the inverse = rows of its transpose; face distance d = 1 / |row d of inverse^T|
    col0 = torch.stack([e * i - f * h, f * g - d * i, d * h - e * g], dim=1)
    col1 = torch.stack([cc * h - b * i, a * i - cc * g, b * g - a * h], dim=1)
    col2 = torch.stack([b * f - cc * e, cc * d - a * f, a * e - b * d], dim=1)
    inv_t = torch.stack([col0, col1, col2], dim=1) * (1.0 / det)[:, None, None]
    face = 1.0 / torch.sqrt(inv_t[:, :, 0] * inv_t[:, :, 0] + inv_t[:, :, 1] * inv_t[:, :, 1] + inv_t[:, :, 2] * inv_t[:, :, 2])
    rc = torch.tensor(cutoff, dtype=cell.dtype, device=cell.device)
    cpd = torch.clamp((face / rc).to(torch.int64), min=1)
    open_single = (cpd == 1) & ~pbc.reshape(-1, 3).to(device=cell.device, dtype=torch.bool)
    radius = torch.where(open_single, torch.zeros_like(cpd), torch.ceil(rc * cpd.to(cell.dtype) / face).to(torch.int64))
    # halvings k = 0, 1, ...: the first k whose product fits (closed form of the reference's while loop)
    shifts = torch.arange(32, device=cell.device, dtype=torch.int64)
    cpd_k = torch.clamp(cpd[:, None, :] >> shifts[None, :, None], min=1)
    fits = cpd_k.prod(dim=2) <= max(int(max_nbins), 1)
    k = torch.argmax(fits.to(torch.int8), dim=1)
    cells = torch.gather(cpd_k.prod(dim=2), 1, k[:, None])[:, 0]
    return cells, radius.to(torch.int32)


def estimate_cell_list_sizes(cell: torch.Tensor, pbc: torch.Tensor, cutoff: float, max_nbins: int = 1000):
    """``(max_total_cells, neighbor_search_radius [3] int32 on cell.device)`` — cell_list.py:639-722, same formula and
    the same single ``.item()`` sync.  The result sizes ``allocate_cell_list``; ``build_cell_list`` keeps its grid within
    that capacity (so the default ``max_nbins = 1000`` gives large systems the reference's coarse, slow grid — pass a
    larger value, e.g. the number of atoms, for the split workflow; ``cell_list`` / ``neighbor_list`` size their own
    workspace and are not affected)."""
    if (cell.ndim == 3 and cell.shape[0] == 0) or cutoff <= 0:
        return 1, torch.zeros((3,), dtype=torch.int32, device=cell.device)
    _engine._dtype_code(cell.dtype)
    cells, radius = _estimate_grid(cell.reshape(-1, 3, 3)[:1], pbc.reshape(-1, 3)[:1], cutoff, max_nbins)
    return int(cells[0].item()), radius[0]


def cell_list(
    positions: torch.Tensor,
    cutoff: float,
    cell: torch.Tensor,
    pbc: torch.Tensor,
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
):
    """Neighbor list of one system with the cell-list algorithm (reference cell_list.py:1195-1443).

    Returns ``(neighbor_matrix [N,M] i32, num_neighbors [N] i32, neighbor_matrix_shifts [N,M,3] i32)`` or, with
    ``return_neighbor_list=True``, ``(neighbor_list [2,P] i32, neighbor_ptr [N+1] i32, shifts [P,3] i32)``.
    Pre-allocated output tensors are filled in place and returned as the same objects.  As in the reference, the seven
    cell-list cache tensors are used only when ALL of them are given: the grid is then capped at their capacity and all
    seven are filled in place (cell_list.py:1376-1410); otherwise the library sizes its own workspace.
    """
    total_atoms = positions.shape[0]
    cell = cell if cell.ndim == 3 else cell.unsqueeze(0)
    pbc = pbc.squeeze(0) if pbc.ndim == 2 else pbc
    if fill_value is None:
        fill_value = total_atoms
    cache = dict(zip(_CACHE_KEYS, (cells_per_dimension, neighbor_search_radius, atom_periodic_shifts, atom_to_cell_mapping,
                                   atoms_per_cell_count, cell_atom_start_indices, cell_atom_list)))
    return _run(positions, cutoff, cell[:1], pbc.reshape(1, 3), None, None, max_neighbors, half_fill, fill_value,
                return_neighbor_list, neighbor_matrix, neighbor_matrix_shifts, num_neighbors, cache,
                empty_fill=fill_value)


# ----------------------------------------------------------------------------------------------------------------
# split build / query (the MD workflow: build with cutoff + skin, re-query while atoms move) — reference
# cell_list.py:1037-1192 and docs/userguide/components/neighborlist.md:421-500.  Thin wrappers over the custom ops
# (ops.py); there is no hidden state: the cell list IS the seven tensors.
# ----------------------------------------------------------------------------------------------------------------
def build_cell_list(
    positions: torch.Tensor,
    cutoff: float,
    cell: torch.Tensor,
    pbc: torch.Tensor,
    cells_per_dimension: torch.Tensor,
    neighbor_search_radius: torch.Tensor,
    atom_periodic_shifts: torch.Tensor,
    atom_to_cell_mapping: torch.Tensor,
    atoms_per_cell_count: torch.Tensor,
    cell_atom_start_indices: torch.Tensor,
    cell_atom_list: torch.Tensor,
) -> None:
    """Build the cell list into the seven cache tensors (reference cell_list.py:1037-1105).  The grid has at most
    ``atoms_per_cell_count.numel()`` cells (``estimate_cell_list_sizes`` / ``allocate_cell_list``); the values are this
    implementation's binning and are all ``query_cell_list`` / ``cell_list_needs_rebuild`` need."""
    from .ops import build_cell_list_op

    _engine._require_cuda(positions, "positions")
    cell = cell if cell.ndim == 3 else cell.unsqueeze(0)
    build_cell_list_op(positions, float(cutoff), cell, pbc, cells_per_dimension, neighbor_search_radius,
                       atom_periodic_shifts, atom_to_cell_mapping, atoms_per_cell_count, cell_atom_start_indices,
                       cell_atom_list)


def query_cell_list(
    positions: torch.Tensor,
    cutoff: float,
    cell: torch.Tensor,
    pbc: torch.Tensor,
    cells_per_dimension: torch.Tensor,
    neighbor_search_radius: torch.Tensor,
    atom_periodic_shifts: torch.Tensor,
    atom_to_cell_mapping: torch.Tensor,
    atoms_per_cell_count: torch.Tensor,
    cell_atom_start_indices: torch.Tensor,
    cell_atom_list: torch.Tensor,
    neighbor_matrix: torch.Tensor,
    neighbor_matrix_shifts: torch.Tensor,
    num_neighbors: torch.Tensor,
    half_fill: bool = False,
) -> None:
    """Query a (possibly stale) cell list with the current ``positions`` and a cutoff <= the build cutoff
    (reference cell_list.py:1108-1192).  Writes hits and ``num_neighbors`` in place; like the reference op it does
    not reset the unused slots of ``neighbor_matrix`` / ``neighbor_matrix_shifts`` (the caller pre-fills them)."""
    from .ops import query_cell_list_op

    _engine._require_cuda(positions, "positions")
    cell = cell if cell.ndim == 3 else cell.unsqueeze(0)
    query_cell_list_op(positions, float(cutoff), cell, pbc, cells_per_dimension, neighbor_search_radius,
                       atom_periodic_shifts, atom_to_cell_mapping, atoms_per_cell_count, cell_atom_start_indices,
                       cell_atom_list, neighbor_matrix, neighbor_matrix_shifts, num_neighbors, bool(half_fill))
