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
    want_cache = cache is not None and any(v is not None for v in cache.values())
    matrix_only = not (return_neighbor_list and neighbor_matrix is None)
    if matrix_only and not want_cache:
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
    h = _engine.build(positions, cutoff, cell, pbc, batch_idx=batch_idx, batch_ptr=batch_ptr)
    if want_cache:
        cpd, rad = _engine.get_grid(h)
        if cache.get("cells_per_dimension") is not None:
            cache["cells_per_dimension"].copy_(cpd.reshape(cache["cells_per_dimension"].shape))
        if cache.get("neighbor_search_radius") is not None:
            cache["neighbor_search_radius"].copy_(rad.reshape(cache["neighbor_search_radius"].shape))
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


def estimate_cell_list_sizes(cell: torch.Tensor, pbc: torch.Tensor, cutoff: float, max_nbins: int = 1000):
    """Signature of cell_list.py:639-722.  The CUDA path sizes its own workspace (number of cells <= number of
    atoms, no device->host sync), so this only reports the stencil radius the reference would use per dimension
    computed on the host from ``cell``; ``max_nbins`` is accepted and ignored."""
    cell = cell.reshape(-1, 3, 3)[0].detach().double().cpu()
    pbc = pbc.reshape(-1)[:3].detach().cpu()
    if cutoff <= 0:
        return 1, torch.zeros((3,), dtype=torch.int32, device=pbc.device)
    inv = torch.linalg.inv(cell)
    face = 1.0 / torch.linalg.norm(inv, dim=0)
    cpd = torch.clamp((face / cutoff).floor().to(torch.int64), min=1)
    radius = torch.where((cpd == 1) & (~pbc.bool()), torch.zeros_like(cpd),
                         torch.ceil(cutoff * cpd / face).to(torch.int64))
    return int(cpd.prod().item()), radius.to(torch.int32).to(cell.device)


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
    Pre-allocated output tensors are filled in place and returned as the same objects.  The seven cell-list
    cache tensors of the reference are accepted for signature compatibility; only ``cells_per_dimension`` and
    ``neighbor_search_radius`` are written (the grid actually used) — the cache itself is an opaque workspace.
    """
    total_atoms = positions.shape[0]
    cell = cell if cell.ndim == 3 else cell.unsqueeze(0)
    pbc = pbc.squeeze(0) if pbc.ndim == 2 else pbc
    if fill_value is None:
        fill_value = total_atoms
    cache = {"cells_per_dimension": cells_per_dimension, "neighbor_search_radius": neighbor_search_radius}
    return _run(positions, cutoff, cell[:1], pbc.reshape(1, 3), None, None, max_neighbors, half_fill, fill_value,
                return_neighbor_list, neighbor_matrix, neighbor_matrix_shifts, num_neighbors, cache,
                empty_fill=fill_value)


# ----------------------------------------------------------------------------------------------------------------
# split build / query (the MD workflow: build with cutoff + skin, re-query while atoms move) — reference
# cell_list.py:1037-1192 and docs/userguide/components/neighborlist.md:421-500
# ----------------------------------------------------------------------------------------------------------------
_HANDLE_ATTR = "_nvnl_handle"


def _attach(handle, *tensors):
    for t in tensors:
        if t is not None:
            setattr(t, _HANDLE_ATTR, handle)


def _find_handle(*tensors):
    for t in tensors:
        h = getattr(t, _HANDLE_ATTR, None) if t is not None else None
        if h is not None:
            return h
    raise RuntimeError(
        "nvalchemiops_b200: these cache tensors were not produced by build_cell_list / batch_build_cell_list of this "
        "package (the cell list itself lives in an opaque device workspace attached to them)."
    )


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
    """Build the cell list (reference cell_list.py:1037-1105).  The seven cache tensors are filled with this
    implementation's grid / binning (for inspection and for ``cell_list_needs_rebuild``); the structure the queries
    actually use is an opaque device workspace attached to them — pass the SAME tensor objects to ``query_cell_list``."""
    if positions.shape[0] == 0 or cutoff <= 0:
        return
    cell = cell if cell.ndim == 3 else cell.unsqueeze(0)
    h = _engine.build(positions, cutoff, cell[:1], pbc.reshape(1, 3))
    _engine.export_cache(h, cells_per_dimension, neighbor_search_radius, atom_periodic_shifts, atom_to_cell_mapping,
                         atoms_per_cell_count, cell_atom_start_indices, cell_atom_list)
    _attach(h, cells_per_dimension, atom_periodic_shifts, atom_to_cell_mapping, atoms_per_cell_count,
            cell_atom_start_indices, cell_atom_list)


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
    if positions.shape[0] == 0 or cutoff <= 0:
        return
    h = _find_handle(cell_atom_list, atom_to_cell_mapping, atoms_per_cell_count, cell_atom_start_indices,
                     atom_periodic_shifts, cells_per_dimension)
    if cutoff > h.cutoff * (1.0 + 1e-12):
        raise ValueError(f"query cutoff {cutoff} exceeds the cutoff {h.cutoff} the cell list was built for")
    _engine.refresh_positions(h, positions)
    _engine.query_matrix(h, _engine.cutoff_sq_in_dtype(cutoff, positions.dtype), neighbor_matrix, neighbor_matrix_shifts,
                         num_neighbors, 0, half_fill, pad_rows=False)
