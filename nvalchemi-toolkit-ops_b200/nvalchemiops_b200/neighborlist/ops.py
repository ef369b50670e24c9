"""torch custom ops of the cell-list path, with the reference's operator schemas.

The reference's operator ABI for this path is four mutation-only ``torch.library.custom_op``s
(``nvalchemiops::build_cell_list`` cell_list.py:725-749, ``nvalchemiops::query_cell_list`` :892-912,
``nvalchemiops::batch_build_cell_list`` batch_cell_list.py:739-763, ``nvalchemiops::batch_query_cell_list`` :915-936).
The same four are registered here under this package's own namespace (the reference registers ``nvalchemiops::*`` at
import; double registration raises) with the SAME argument lists — note the batch query's ``(positions, cell, pbc,
cutoff, batch_idx, ...)`` order — so ``torch.compile`` keeps them in the graph with pre-allocated tensors
(reference test_cell_list.py:598-844):

    nvalchemiops_b200::build_cell_list(Tensor positions, float cutoff, Tensor cell, Tensor pbc,
        Tensor(a!) cells_per_dimension, Tensor(b!) neighbor_search_radius, Tensor(c!) atom_periodic_shifts,
        Tensor(d!) atom_to_cell_mapping, Tensor(e!) atoms_per_cell_count, Tensor(f!) cell_atom_start_indices,
        Tensor(g!) cell_atom_list) -> ()
    nvalchemiops_b200::query_cell_list(positions, cutoff, cell, pbc, <7 cache>, Tensor(a!) neighbor_matrix,
        Tensor(b!) neighbor_matrix_shifts, Tensor(c!) num_neighbors, bool half_fill=False) -> ()
    nvalchemiops_b200::batch_build_cell_list(positions, cutoff, cell, pbc, batch_idx, <7 cache>) -> ()
    nvalchemiops_b200::batch_query_cell_list(positions, cell, pbc, cutoff, batch_idx, <7 cache>, nm, shifts, num,
        half_fill=False) -> ()

No hidden state: the build op fills the seven cache tensors (grid capped at their capacity, like the reference's
``max_total_cells``), the query op rebuilds its device workspace from their VALUES and the current positions
(``nvnl_import_cache``).  The tensors may be cloned, moved or traced in between.

``nvalchemiops_b200::neighbor_matrix`` is an addition: build + query fused, workspace sized by the library (no cap).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _engine

_CACHE_MUTATED = ("cells_per_dimension", "neighbor_search_radius", "atom_periodic_shifts", "atom_to_cell_mapping",
                  "atoms_per_cell_count", "cell_atom_start_indices", "cell_atom_list")
_OUT_MUTATED = ("neighbor_matrix", "neighbor_matrix_shifts", "num_neighbors")


def _build(positions, cutoff, cell, pbc, batch_idx, cache):
    if positions.shape[0] == 0 or cutoff <= 0:
        return
    cpd, rad, ashift, amap, ccount, cstart, clist = cache
    capacity = min(ccount.numel(), cstart.numel())
    h = _engine.build(positions, cutoff, cell.reshape(-1, 3, 3), pbc.reshape(-1, 3), batch_idx=batch_idx,
                      max_cells=capacity)
    _engine.export_cache(h, cpd, rad, ashift, amap, ccount, cstart, clist)


def _query(positions, cutoff, cell, pbc, batch_idx, cache, neighbor_matrix, neighbor_matrix_shifts, num_neighbors,
           half_fill):
    if positions.shape[0] == 0 or cutoff <= 0:
        return
    h = _engine.import_cache(positions, cutoff, cell.reshape(-1, 3, 3), pbc.reshape(-1, 3), batch_idx, *cache)
    # like the reference op, only hits and num_neighbors are written: the caller pre-fills the padding
    _engine.query_matrix(h, _engine.cutoff_sq_in_dtype(cutoff, positions.dtype), neighbor_matrix, neighbor_matrix_shifts,
                         num_neighbors, 0, half_fill, pad_rows=False)


@torch.library.custom_op("nvalchemiops_b200::build_cell_list", mutates_args=_CACHE_MUTATED)
def build_cell_list_op(
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
    _build(positions, cutoff, cell, pbc, None,
           (cells_per_dimension, neighbor_search_radius, atom_periodic_shifts, atom_to_cell_mapping,
            atoms_per_cell_count, cell_atom_start_indices, cell_atom_list))


@torch.library.custom_op("nvalchemiops_b200::query_cell_list", mutates_args=_OUT_MUTATED)
def query_cell_list_op(
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
    _query(positions, cutoff, cell, pbc, None,
           (cells_per_dimension, neighbor_search_radius, atom_periodic_shifts, atom_to_cell_mapping,
            atoms_per_cell_count, cell_atom_start_indices, cell_atom_list),
           neighbor_matrix, neighbor_matrix_shifts, num_neighbors, half_fill)


@torch.library.custom_op("nvalchemiops_b200::batch_build_cell_list", mutates_args=_CACHE_MUTATED)
def batch_build_cell_list_op(
    positions: torch.Tensor,
    cutoff: float,
    cell: torch.Tensor,
    pbc: torch.Tensor,
    batch_idx: torch.Tensor,
    cells_per_dimension: torch.Tensor,
    neighbor_search_radius: torch.Tensor,
    atom_periodic_shifts: torch.Tensor,
    atom_to_cell_mapping: torch.Tensor,
    atoms_per_cell_count: torch.Tensor,
    cell_atom_start_indices: torch.Tensor,
    cell_atom_list: torch.Tensor,
) -> None:
    # (the reference lists neighbor_search_radius as read-only here, batch_cell_list.py:741-749; this op also WRITES
    # the radius of the grid it chose, so it is declared mutated — a superset of the reference's aliasing)
    _build(positions, cutoff, cell, pbc, batch_idx,
           (cells_per_dimension, neighbor_search_radius, atom_periodic_shifts, atom_to_cell_mapping,
            atoms_per_cell_count, cell_atom_start_indices, cell_atom_list))


@torch.library.custom_op("nvalchemiops_b200::batch_query_cell_list", mutates_args=_OUT_MUTATED)
def batch_query_cell_list_op(
    positions: torch.Tensor,
    cell: torch.Tensor,
    pbc: torch.Tensor,
    cutoff: float,
    batch_idx: torch.Tensor,
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
    _query(positions, cutoff, cell, pbc, batch_idx,
           (cells_per_dimension, neighbor_search_radius, atom_periodic_shifts, atom_to_cell_mapping,
            atoms_per_cell_count, cell_atom_start_indices, cell_atom_list),
           neighbor_matrix, neighbor_matrix_shifts, num_neighbors, half_fill)


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
    """Build + query fused (no cap on the grid, no cache export, no host sync, every output slot written once)."""
    h = _engine.build(positions, cutoff, cell, pbc, batch_idx=batch_idx, batch_ptr=batch_ptr)
    _engine.query_matrix(h, cutoff_sq, neighbor_matrix, neighbor_matrix_shifts, num_neighbors, fill_value, half_fill)
