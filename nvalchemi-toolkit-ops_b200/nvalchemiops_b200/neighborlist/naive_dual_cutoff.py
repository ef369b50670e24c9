"""``naive_neighbor_list_dual_cutoff`` — two neighbor lists (cutoff1, cutoff2) from one pass over the atoms
(reference nvalchemiops/neighborlist/naive_dual_cutoff.py:544-919), served by the B200 cell-list engine: ONE cell-list
build with the larger cutoff, then one sweep per cutoff over the same sorted records."""
from __future__ import annotations

import torch

from . import _engine
from .cell_list import _query
from .naive import _check_cell_pbc, _open_or_given_cell, _padded_outputs
from .neighbor_utils import estimate_max_neighbors


def _dual_route(positions, cutoff1, cutoff2, cell, pbc, batch_idx, batch_ptr, max_neighbors1, max_neighbors2, half_fill,
                fill_value, return_neighbor_list, neighbor_matrix1, neighbor_matrix2, neighbor_matrix_shifts1,
                neighbor_matrix_shifts2, num_neighbors1, num_neighbors2):
    """Shared body of the single and the batched dual-cutoff entry points.  Return pattern of the reference
    (naive_dual_cutoff.py:856-919): (data1, count-or-ptr1[, shifts1], data2, count-or-ptr2[, shifts2])."""
    if cutoff2 is None:
        raise ValueError("cutoff2 is required for the dual-cutoff methods")
    n, dev = positions.shape[0], positions.device
    has_pbc = pbc is not None
    if fill_value is None:
        fill_value = n
    # the reference sizes BOTH matrices from cutoff2 when no size is given (naive_dual_cutoff.py:761-772)
    if max_neighbors1 is None and (neighbor_matrix1 is None or neighbor_matrix2 is None
                                   or (neighbor_matrix_shifts1 is None and has_pbc)
                                   or (neighbor_matrix_shifts2 is None and has_pbc)
                                   or num_neighbors1 is None or num_neighbors2 is None):
        max_neighbors2 = estimate_max_neighbors(cutoff2)
        max_neighbors1 = max_neighbors2
    if max_neighbors2 is None:
        max_neighbors2 = max_neighbors1
    ns = int(batch_ptr.shape[0] - 1) if batch_ptr is not None else 1
    cell_, pbc_ = _open_or_given_cell(positions, cell, pbc, ns)
    _engine._dtype_code(positions.dtype)
    _engine._require_cuda(positions, "positions")
    h = None
    if n > 0 and max(cutoff1, cutoff2) > 0:
        h = _engine.build(positions, max(cutoff1, cutoff2), cell_, pbc_, batch_idx=batch_idx, batch_ptr=batch_ptr)
    outs = []
    for rc, mx, nm, sh, num in ((cutoff1, max_neighbors1, neighbor_matrix1, neighbor_matrix_shifts1, num_neighbors1),
                                (cutoff2, max_neighbors2, neighbor_matrix2, neighbor_matrix_shifts2, num_neighbors2)):
        if h is None or rc <= 0:
            # nothing within this cutoff: the reference's kernels find no pair and its buffers come back reset
            nm, num, sh = _padded_outputs(n, dev, has_pbc, mx, fill_value, nm, sh, num)
            if return_neighbor_list:
                out = (torch.zeros((2, 0), dtype=torch.int32, device=dev),
                       torch.zeros((n + 1,), dtype=torch.int32, device=dev),
                       torch.zeros((0, 3), dtype=torch.int32, device=dev))
            else:
                out = (nm, num, sh)
        else:
            csq = _engine.cutoff_sq_in_dtype(rc, positions.dtype, python_double=True)   # naive rule (naive.py:290)
            out = _query(h, rc, csq, mx, half_fill, fill_value, return_neighbor_list, nm, sh, num)
        outs.extend(out if has_pbc else out[:2])
    return tuple(outs)


def naive_neighbor_list_dual_cutoff(
    positions: torch.Tensor,
    cutoff1: float,
    cutoff2: float,
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
):
    """Neighbors of one system within ``cutoff1`` and within ``cutoff2``; same contract as the reference
    (naive_dual_cutoff.py:544-919): 4-tuple without PBC, 6-tuple with PBC, matrices or COO lists."""
    _check_cell_pbc(cell, pbc)
    return _dual_route(positions, cutoff1, cutoff2, cell, pbc, None, None, max_neighbors1, max_neighbors2, half_fill,
                       fill_value, return_neighbor_list, neighbor_matrix1, neighbor_matrix2, neighbor_matrix_shifts1,
                       neighbor_matrix_shifts2, num_neighbors1, num_neighbors2)
