"""``neighbor_list`` — the unified entry point (reference nvalchemiops/neighborlist/neighborlist.py:41-310).

Same signature, method auto-selection, kwargs forwarding and return arity as the reference.  Every method name
that the reference routes to an O(N^2) "naive" Warp kernel is served here by the same B200 cell-list engine
(the neighbor set does not depend on the search algorithm); only the return arity differs (2-tuples without
PBC).  The dual-cutoff methods build one cell list with the larger cutoff and query it twice.
"""
from __future__ import annotations

import torch

from . import _engine
from .batch_cell_list import batch_cell_list
from .cell_list import _query, _run, cell_list
from .neighbor_utils import _prepare_batch_idx_ptr

_NAIVE_KWARGS = {"max_neighbors", "neighbor_matrix", "neighbor_matrix_shifts", "num_neighbors",
                 "shift_range_per_dimension", "shift_offset", "total_shifts", "max_atoms_per_system"}


def _naive(positions, cutoff, cell, pbc, batch_idx, batch_ptr, half_fill, fill_value, return_neighbor_list, **kwargs):
    """naive / batch_naive routes (naive.py:400-706, batch_naive.py): same engine, reference return arity."""
    unknown = set(kwargs) - _NAIVE_KWARGS
    if unknown:
        raise TypeError(f"naive_neighbor_list() got an unexpected keyword argument '{sorted(unknown)[0]}'")
    if pbc is None and cell is not None:
        raise ValueError("If cell is provided, pbc must also be provided")
    if pbc is not None and cell is None:
        raise ValueError("If pbc is provided, cell must also be provided")
    n = positions.shape[0]
    dev = positions.device
    has_pbc = pbc is not None
    if fill_value is None:
        fill_value = n
    if batch_idx is not None:
        ns = int(batch_ptr.shape[0] - 1) if batch_ptr is not None else int(batch_idx.max().item()) + 1
    else:
        ns = 1
    if not has_pbc:
        cell_ = torch.eye(3, dtype=positions.dtype, device=dev).reshape(1, 3, 3).repeat(ns, 1, 1)
        pbc_ = torch.zeros((ns, 3), dtype=torch.bool, device=dev)
    else:
        cell_ = (cell if cell.ndim == 3 else cell.unsqueeze(0)).to(dev)
        pbc_ = (pbc if pbc.ndim == 2 else pbc.unsqueeze(0)).to(dev)
    if cutoff <= 0:
        # naive.py:622-662 keeps the allocated (N, max_neighbors) matrix for cutoff <= 0 in matrix mode; the
        # accelerated path returns the cell-list style empty shapes instead.
        pass
    # naive squares the cutoff in Python double and casts (naive.py:290)
    csq = _engine.cutoff_sq_in_dtype(cutoff, positions.dtype, python_double=True) if cutoff > 0 else None
    out = _run(positions, cutoff, cell_, pbc_, batch_idx, batch_ptr, kwargs.get("max_neighbors"), half_fill,
               fill_value, return_neighbor_list, kwargs.get("neighbor_matrix"), kwargs.get("neighbor_matrix_shifts"),
               kwargs.get("num_neighbors"), None, empty_fill=fill_value, cutoff_sq=csq)
    if has_pbc:
        return out
    return out[0], out[1]  # no PBC: 2-tuples (neighborlist.py:150-153)


_DUAL_KWARGS = {"max_neighbors1", "max_neighbors2", "neighbor_matrix1", "neighbor_matrix2", "neighbor_matrix_shifts1",
                "neighbor_matrix_shifts2", "num_neighbors1", "num_neighbors2", "shift_range_per_dimension", "shift_offset",
                "total_shifts", "max_atoms_per_system"}


def _dual_cutoff(positions, cutoff1, cutoff2, cell, pbc, batch_idx, batch_ptr, half_fill, fill_value,
                 return_neighbor_list, **kwargs):
    """naive_dual_cutoff / batch_naive_dual_cutoff routes (naive_dual_cutoff.py:544-919): ONE cell-list build with
    the larger cutoff, two queries.  Return arity of the reference: (data1, num1[, shifts1], data2, num2[, shifts2])."""
    unknown = set(kwargs) - _DUAL_KWARGS
    if unknown:
        raise TypeError(f"naive_neighbor_list_dual_cutoff() got an unexpected keyword argument '{sorted(unknown)[0]}'")
    if cutoff2 is None:
        raise ValueError("cutoff2 is required for the dual-cutoff methods")
    if pbc is None and cell is not None:
        raise ValueError("If cell is provided, pbc must also be provided")
    if pbc is not None and cell is None:
        raise ValueError("If pbc is provided, cell must also be provided")
    n, dev = positions.shape[0], positions.device
    has_pbc = pbc is not None
    if fill_value is None:
        fill_value = n
    ns = int(batch_ptr.shape[0] - 1) if batch_ptr is not None else 1
    if not has_pbc:
        cell_ = torch.eye(3, dtype=positions.dtype, device=dev).reshape(1, 3, 3).repeat(ns, 1, 1)
        pbc_ = torch.zeros((ns, 3), dtype=torch.bool, device=dev)
    else:
        cell_ = (cell if cell.ndim == 3 else cell.unsqueeze(0)).to(dev)
        pbc_ = (pbc if pbc.ndim == 2 else pbc.unsqueeze(0)).to(dev)
    m1 = kwargs.get("max_neighbors1")
    m2 = kwargs.get("max_neighbors2", m1)
    outs = []
    h = None
    if n > 0 and max(cutoff1, cutoff2) > 0:
        h = _engine.build(positions, max(cutoff1, cutoff2), cell_, pbc_, batch_idx=batch_idx, batch_ptr=batch_ptr)
    for rc, mx, nm, sh, num in ((cutoff1, m1, kwargs.get("neighbor_matrix1"), kwargs.get("neighbor_matrix_shifts1"),
                                 kwargs.get("num_neighbors1")),
                                (cutoff2, m2, kwargs.get("neighbor_matrix2"), kwargs.get("neighbor_matrix_shifts2"),
                                 kwargs.get("num_neighbors2"))):
        if h is None or rc <= 0:
            out = _run(positions, 0.0, cell_, pbc_, batch_idx, batch_ptr, mx, half_fill, fill_value, return_neighbor_list,
                       None, None, None, None, empty_fill=fill_value)
        else:
            csq = _engine.cutoff_sq_in_dtype(rc, positions.dtype, python_double=True)  # naive rule (naive.py:290)
            out = _query(h, rc, csq, mx, half_fill, fill_value, return_neighbor_list, nm, sh, num)
        outs.extend(out if has_pbc else out[:2])
    return tuple(outs)


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
            return _naive(positions, cutoff, cell, pbc, None, None, half_fill, fill_value, return_neighbor_list,
                          **kwargs)
        case "cell_list":
            return cell_list(positions, cutoff, cell, pbc, half_fill=half_fill, fill_value=fill_value,
                             return_neighbor_list=return_neighbor_list, **kwargs)
        case "batch_naive":
            if batch_idx is None or batch_ptr is None:
                batch_idx, batch_ptr = _prepare_batch_idx_ptr(batch_idx, batch_ptr, positions.shape[0],
                                                              positions.device)
            return _naive(positions, cutoff, cell, pbc, batch_idx, batch_ptr, half_fill, fill_value,
                          return_neighbor_list, **kwargs)
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
        case "naive_dual_cutoff" | "batch_naive_dual_cutoff":
            if method.startswith("batch_") and (batch_idx is None or batch_ptr is None):
                batch_idx, batch_ptr = _prepare_batch_idx_ptr(batch_idx, batch_ptr, positions.shape[0],
                                                              positions.device)
            return _dual_cutoff(positions, cutoff, cutoff2, cell, pbc, batch_idx, batch_ptr, half_fill, fill_value,
                                return_neighbor_list, **kwargs)
        case _:
            raise ValueError(f"Invalid method: {method}")
