"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of ``nvalchemiops.neighborlist`` (reference @ v0.2.0) for the cell-list
path, used as the checker by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs.  The product package never imports this module.

The arithmetic lives in ``nl_oracle.c`` / ``nl_oracle_impl.h`` (each function cites the reference
file:line it follows); this file restates the Python-side wrappers:

* ``neighbor_list``        <- nvalchemiops/neighborlist/neighborlist.py:41-310
* ``cell_list``            <- nvalchemiops/neighborlist/cell_list.py:1195-1443
* ``batch_cell_list``      <- nvalchemiops/neighborlist/batch_cell_list.py:1229-1468
* ``naive_neighbor_list``  <- nvalchemiops/neighborlist/naive.py:400-706 (no-PBC branch only)
* ``estimate_max_neighbors``, ``get_neighbor_list_from_neighbor_matrix``,
  ``NeighborOverflowError``, ``allocate_cell_list``, ``prepare_batch_idx_ptr``
                           <- nvalchemiops/neighborlist/neighbor_utils.py:296-539

All functions take / return numpy arrays (torch CPU tensors are converted).

Parity status: the reference (Warp) is not installable in the build sandbox (no network, no
``warp-lang``), so this restatement is pinned by the reference's own known-answer tests
(tests/golden/kat_structures.json) and an independent brute force — see DESIGN.md "Oracle".
"parity unpinned" applies to the FMA-contraction choice (``fma_mode``) that only matters for pairs
within ~1 ulp of the cutoff.
"""

from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnl_oracle.so")
_lib = None

# Default contraction mode of the distance predicate: 1 = mul + fma chain (what nvcc/NVRTC with
# fmad=true — Warp's default ``fuse_fp`` — emit), 0 = separately rounded mul/add.
DEFAULT_FMA_MODE = 1


def build(force: bool = False) -> str:
    """Compile the C oracle with the committed Makefile (gcc only)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH)
        < max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("nl_oracle.c", "nl_oracle_impl.h"))
    ):
        subprocess.check_call(["make", "-s", "-B", "-C", _HERE])
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.nlo_brute_force_f32.restype = ctypes.c_long
        _lib.nlo_brute_force_f64.restype = ctypes.c_long
    return _lib


def _np(x, dtype=None):
    if x is None:
        return None
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    x = np.ascontiguousarray(x)
    if dtype is not None and x.dtype != dtype:
        x = x.astype(dtype)
    return x


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _suffix(dtype):
    if dtype == np.float32:
        return "f32", ctypes.c_float
    if dtype == np.float64:
        return "f64", ctypes.c_double
    raise ValueError(f"Unsupported dtype: {dtype}")  # nvalchemiops/types.py:29


# --------------------------------------------------------------------------------------
# neighbor_utils.py
# --------------------------------------------------------------------------------------
def estimate_max_neighbors(cutoff: float, atomic_density: float = 0.35, safety_factor: float = 5.0) -> int:
    """neighbor_utils.py:296-340."""
    if cutoff <= 0:
        return 0
    vol = atomic_density * (4.0 / 3.0) * math.pi * (cutoff**3)
    expected = max(1, safety_factor * vol)
    return int(math.ceil(expected / 16)) * 16


class NeighborOverflowError(Exception):
    """neighbor_utils.py:343-349."""

    def __init__(self, max_neighbors: int, num_neighbors: int):
        super().__init__(
            f"The number of neighbors is larger than the maximum allowed: {num_neighbors} > {max_neighbors}."
        )


def get_neighbor_list_from_neighbor_matrix(neighbor_matrix, num_neighbors, neighbor_shift_matrix=None, fill_value=-1):
    """neighbor_utils.py:362-441: padded matrix -> COO via the ``!= fill_value`` mask."""
    nm = _np(neighbor_matrix)
    num = _np(num_neighbors)
    sh = _np(neighbor_shift_matrix)
    if num.shape[0] == 0:
        nl = np.zeros((2, 0), dtype=nm.dtype)
        ptr = np.zeros(1, dtype=np.int32)
        if sh is None:
            return nl, ptr
        return nl, ptr, np.empty((0, 2, 3), dtype=sh.dtype)
    mx = 0 if num.size == 0 else int(num.max())
    if mx > nm.shape[1]:
        raise NeighborOverflowError(nm.shape[1], mx)
    mask = nm != fill_value
    i_idx = np.nonzero(mask)[0].astype(nm.dtype)
    j_idx = nm[mask].astype(nm.dtype)
    nl = np.stack([i_idx, j_idx], axis=0)
    ptr = np.zeros(num.shape[0] + 1, dtype=np.int32)
    np.cumsum(num, out=ptr[1:])
    if sh is not None:
        return nl, ptr, sh[mask]
    return nl, ptr


def prepare_batch_idx_ptr(batch_idx, batch_ptr, num_atoms):
    """neighbor_utils.py:444-491."""
    if batch_idx is None and batch_ptr is None:
        raise ValueError("Either batch_idx or batch_ptr must be provided.")
    batch_idx = _np(batch_idx)
    batch_ptr = _np(batch_ptr)
    if batch_idx is None:
        counts = batch_ptr[1:] - batch_ptr[:-1]
        batch_idx = np.repeat(np.arange(len(counts), dtype=np.int32), counts)
    elif batch_ptr is None:
        ns = int(batch_idx.max()) + 1 if batch_idx.size else 0
        counts = np.bincount(batch_idx, minlength=ns)
        batch_ptr = np.zeros(ns + 1, dtype=np.int32)
        np.cumsum(counts, out=batch_ptr[1:])
    return batch_idx.astype(np.int32), batch_ptr.astype(np.int32)


def allocate_cell_list(total_atoms, max_total_cells, neighbor_search_radius):
    """neighbor_utils.py:494-539: the 7-tensor cache."""
    r = _np(neighbor_search_radius)
    cpd = np.zeros((3,) if r.ndim == 1 else (r.shape[0], 3), dtype=np.int32)
    return (
        cpd,
        r,
        np.zeros((total_atoms, 3), dtype=np.int32),
        np.zeros((total_atoms, 3), dtype=np.int32),
        np.zeros((max_total_cells,), dtype=np.int32),
        np.zeros((max_total_cells,), dtype=np.int32),
        np.zeros((total_atoms,), dtype=np.int32),
    )


# --------------------------------------------------------------------------------------
# cell_list.py / batch_cell_list.py
# --------------------------------------------------------------------------------------
def estimate_cell_list_sizes(cell, pbc, cutoff, max_nbins=1000):
    """cell_list.py:639-722.  Returns (max_total_cells, neighbor_search_radius[3])."""
    cell = _np(cell)
    pbc = _np(pbc, np.uint8).reshape(-1)
    if (cell.ndim == 3 and cell.shape[0] == 0) or cutoff <= 0:
        return 1, np.zeros(3, dtype=np.int32)
    suf, cf = _suffix(cell.dtype)
    cell = cell.reshape(-1, 9)[0].copy()
    radius = np.zeros(3, dtype=np.int32)
    fn = getattr(_load(), f"nlo_estimate_cell_list_sizes_{suf}")
    n = fn(_ptr(cell), _ptr(pbc), cf(cutoff), ctypes.c_int(max_nbins), _ptr(radius))
    return int(n), radius


def estimate_batch_cell_list_sizes(cell, pbc, cutoff, max_nbins=1000):
    """batch_cell_list.py:659-736: per-system cap, summed."""
    cell = _np(cell)
    ns = cell.shape[0]
    if ns == 0 or cutoff <= 0:
        return 1, np.zeros((ns, 3), dtype=np.int32)
    pbc = _np(pbc, np.uint8).reshape(ns, 3)
    suf, cf = _suffix(cell.dtype)
    cell = cell.reshape(ns, 9)
    radius = np.zeros((ns, 3), dtype=np.int32)
    fn = getattr(_load(), f"nlo_estimate_cell_list_sizes_{suf}")
    total = 0
    for s in range(ns):
        c = np.ascontiguousarray(cell[s])
        p = np.ascontiguousarray(pbc[s])
        r = np.zeros(3, dtype=np.int32)
        total += fn(_ptr(c), _ptr(p), cf(cutoff), ctypes.c_int(max_nbins), _ptr(r))
        radius[s] = r
    return int(total), radius


def build_cell_list(positions, cutoff, cell, pbc, cpd, radius, atom_shifts, atom_cell, count, start, clist,
                    batch_idx=None):
    """cell_list.py:725-889 / batch_cell_list.py:739-912 (mutates the cache arrays in place)."""
    pos = _np(positions)
    n = pos.shape[0]
    if n == 0 or cutoff <= 0:
        return
    suf, cf = _suffix(pos.dtype)
    cellm = _np(cell, pos.dtype).reshape(-1, 9)
    ns = cellm.shape[0]
    pbcm = _np(pbc, np.uint8).reshape(ns, 3)
    bidx = _np(batch_idx, np.int32)
    fn = getattr(_load(), f"nlo_build_cell_list_{suf}")
    fn(_ptr(pos), ctypes.c_int(n), _ptr(cellm), _ptr(pbcm), _ptr(bidx), ctypes.c_int(ns), cf(cutoff),
       ctypes.c_int(count.shape[0]), _ptr(cpd), _ptr(atom_shifts), _ptr(atom_cell), _ptr(count), _ptr(start),
       _ptr(clist))


def query_cell_list(positions, cutoff, cell, pbc, cpd, radius, atom_shifts, atom_cell, count, start, clist,
                    neighbor_matrix, neighbor_matrix_shifts, num_neighbors, half_fill=False, batch_idx=None,
                    fma_mode=None, nthreads=1, n_limit=0):
    """cell_list.py:892-1034 / batch_cell_list.py:915-1067 (mutates the three outputs in place)."""
    pos = _np(positions)
    n = pos.shape[0]
    if n == 0 or cutoff <= 0:
        return
    suf, cf = _suffix(pos.dtype)
    cellm = _np(cell, pos.dtype).reshape(-1, 9)
    ns = cellm.shape[0]
    pbcm = _np(pbc, np.uint8).reshape(ns, 3)
    bidx = _np(batch_idx, np.int32)
    fma = DEFAULT_FMA_MODE if fma_mode is None else int(fma_mode)
    fn = getattr(_load(), f"nlo_query_cell_list_{suf}")
    fn(_ptr(pos), ctypes.c_int(n), _ptr(cellm), _ptr(pbcm), _ptr(bidx), ctypes.c_int(ns), cf(cutoff),
       _ptr(np.ascontiguousarray(cpd)), _ptr(np.ascontiguousarray(radius)), _ptr(atom_shifts), _ptr(atom_cell),
       _ptr(count), _ptr(start), _ptr(clist), _ptr(neighbor_matrix), _ptr(neighbor_matrix_shifts),
       _ptr(num_neighbors), ctypes.c_int(neighbor_matrix.shape[1]), ctypes.c_int(bool(half_fill)),
       ctypes.c_int(fma), ctypes.c_int(nthreads), ctypes.c_int(int(n_limit)))


def _cell_list_impl(positions, cutoff, cell, pbc, batch_idx, max_neighbors, half_fill, fill_value,
                    return_neighbor_list, empty_fill, fma_mode, nthreads, max_nbins=1000):
    pos = _np(positions)
    _suffix(pos.dtype)
    n = pos.shape[0]
    batched = batch_idx is not None
    cellm = _np(cell, pos.dtype)
    cellm = cellm if cellm.ndim == 3 else cellm[None]
    pbcm = _np(pbc, np.uint8)
    pbcm = pbcm.reshape(cellm.shape[0], 3) if batched else pbcm.reshape(-1)[:3]
    if fill_value is None and not batched:
        fill_value = n  # cell_list.py:1331-1332
    if n <= 0 or cutoff <= 0:  # cell_list.py:1335-1349, batch_cell_list.py:1358-1371
        if return_neighbor_list:
            return (np.zeros((2, 0), np.int32), np.zeros((n + 1,), np.int32), np.zeros((0, 3), np.int32))
        ef = empty_fill if batched else fill_value
        return (np.full((n, 0), ef, np.int32), np.zeros((n,), np.int32), np.zeros((n, 0, 3), np.int32))
    if max_neighbors is None:
        max_neighbors = estimate_max_neighbors(cutoff)
    if fill_value is None:
        fill_value = n  # batch_cell_list.py:1376-1377
    nm = np.full((n, max_neighbors), fill_value, dtype=np.int32)
    sh = np.zeros((n, max_neighbors, 3), dtype=np.int32)
    num = np.zeros((n,), dtype=np.int32)
    if batched:
        max_cells, radius = estimate_batch_cell_list_sizes(cellm, pbcm, cutoff, max_nbins)
    else:
        max_cells, radius = estimate_cell_list_sizes(cellm, pbcm, cutoff, max_nbins)
    cache = allocate_cell_list(n, max_cells, radius)
    build_cell_list(pos, cutoff, cellm, pbcm, *cache, batch_idx=batch_idx)
    query_cell_list(pos, cutoff, cellm, pbcm, *cache, nm, sh, num, half_fill, batch_idx=batch_idx,
                    fma_mode=fma_mode, nthreads=nthreads)
    if return_neighbor_list:
        return get_neighbor_list_from_neighbor_matrix(nm, num, sh, fill_value)
    return nm, num, sh


def cell_list(positions, cutoff, cell, pbc, max_neighbors=None, half_fill=False, fill_value=None,
              return_neighbor_list=False, fma_mode=None, nthreads=1, max_nbins=1000):
    """cell_list.py:1195-1443."""
    return _cell_list_impl(positions, cutoff, cell, pbc, None, max_neighbors, half_fill, fill_value,
                           return_neighbor_list, None, fma_mode, nthreads, max_nbins)


def batch_cell_list(positions, cutoff, cell, pbc, batch_idx, max_neighbors=None, half_fill=False,
                    fill_value=None, return_neighbor_list=False, fma_mode=None, nthreads=1, max_nbins=1000):
    """batch_cell_list.py:1229-1468."""
    return _cell_list_impl(positions, cutoff, cell, pbc, _np(batch_idx, np.int32), max_neighbors, half_fill,
                           fill_value, return_neighbor_list, -1, fma_mode, nthreads, max_nbins)


def naive_neighbor_list(positions, cutoff, max_neighbors=None, half_fill=False, fill_value=None,
                        return_neighbor_list=False, fma_mode=None):
    """naive.py:400-706, no-PBC branch (2-tuple returns; cutoff² in double then cast, naive.py:290)."""
    pos = _np(positions)
    suf, cf = _suffix(pos.dtype)
    n = pos.shape[0]
    if max_neighbors is None:
        max_neighbors = estimate_max_neighbors(cutoff)
    if fill_value is None:
        fill_value = n
    nm = np.full((n, max_neighbors), fill_value, dtype=np.int32)
    num = np.zeros((n,), dtype=np.int32)
    if cutoff <= 0:
        if return_neighbor_list:
            return np.zeros((2, 0), np.int32), np.zeros((n,), np.int32), np.zeros((n + 1,), np.int32)
        return nm, num
    fma = DEFAULT_FMA_MODE if fma_mode is None else int(fma_mode)
    fn = getattr(_load(), f"nlo_naive_no_pbc_{suf}")
    fn(_ptr(pos), ctypes.c_int(n), cf(cutoff * cutoff), _ptr(nm), _ptr(num), ctypes.c_int(max_neighbors),
       ctypes.c_int(bool(half_fill)), ctypes.c_int(fma))
    if return_neighbor_list:
        return get_neighbor_list_from_neighbor_matrix(nm, num, None, fill_value)
    return nm, num


def neighbor_list(positions, cutoff, cell=None, pbc=None, batch_idx=None, batch_ptr=None, half_fill=False,
                  fill_value=None, return_neighbor_list=False, method=None, **kwargs):
    """neighborlist.py:213-310 restricted to naive (no PBC) / cell_list / batch_cell_list."""
    pos = _np(positions)
    if method is None:
        n = pos.shape[0]
        if n >= 5000:
            method = "cell_list"
            if cell is None or pbc is None:
                cell = np.eye(3, dtype=pos.dtype).reshape(1, 3, 3)
                pbc = np.array([False, False, False])
        else:
            method = "naive"
        if batch_idx is not None or batch_ptr is not None:
            method = "batch_" + method
            batch_idx, batch_ptr = prepare_batch_idx_ptr(batch_idx, batch_ptr, n)
    if method == "naive":
        if cell is not None or pbc is not None:
            raise NotImplementedError("oracle restates only the no-PBC naive branch")
        return naive_neighbor_list(pos, cutoff, half_fill=half_fill, fill_value=fill_value,
                                   return_neighbor_list=return_neighbor_list, **kwargs)
    if method == "cell_list":
        return cell_list(pos, cutoff, cell, pbc, half_fill=half_fill, fill_value=fill_value,
                         return_neighbor_list=return_neighbor_list, **kwargs)
    if method == "batch_cell_list":
        return batch_cell_list(pos, cutoff, cell, pbc, batch_idx, half_fill=half_fill, fill_value=fill_value,
                               return_neighbor_list=return_neighbor_list, **kwargs)
    raise ValueError(f"Invalid method: {method}")


# --------------------------------------------------------------------------------------
# Independent brute force and comparison helpers
# --------------------------------------------------------------------------------------
def brute_force(positions, cutoff, cell, pbc, fma_mode=None, extra_images=2):
    """All directed (i, j, sx, sy, sz) with |r_j - r_i + s.cell|^2 < rc^2, s in a box of images large
    enough for unwrapped inputs (ceil(rc/face)+extra per periodic dim).  O(N^2 * images): small inputs only.
    Returns an int32 array [P,5], lexicographically sorted."""
    pos = _np(positions)
    suf, cf = _suffix(pos.dtype)
    cellm = _np(cell, pos.dtype).reshape(3, 3)
    pbcv = _np(pbc, np.uint8).reshape(-1)[:3]
    inv = np.linalg.inv(cellm.astype(np.float64))
    face = 1.0 / np.linalg.norm(inv, axis=0)
    frac = pos.astype(np.float64) @ inv
    spread = np.ceil(frac.max(axis=0) - frac.min(axis=0)) if pos.shape[0] else np.zeros(3)
    K = np.array([int(math.ceil(cutoff / face[d]) + extra_images + spread[d]) if pbcv[d] else 0 for d in range(3)],
                 dtype=np.int32)
    fma = DEFAULT_FMA_MODE if fma_mode is None else int(fma_mode)
    fn = getattr(_load(), f"nlo_brute_force_{suf}")
    cap = 1 << 16
    while True:
        out = np.zeros((cap, 5), dtype=np.int32)
        cnt = fn(_ptr(pos), ctypes.c_int(pos.shape[0]), _ptr(np.ascontiguousarray(cellm)), _ptr(pbcv), cf(cutoff),
                 _ptr(K), ctypes.c_int(fma), _ptr(out), ctypes.c_long(cap))
        if cnt <= cap:
            return sort_records(out[:cnt])
        cap = int(cnt)


def sort_records(rec):
    """Lexicographic sort of [P,5] (i, j, sx, sy, sz) records."""
    rec = np.asarray(rec, dtype=np.int64).reshape(-1, 5)
    if rec.shape[0] == 0:
        return rec.astype(np.int32)
    order = np.lexsort((rec[:, 4], rec[:, 3], rec[:, 2], rec[:, 1], rec[:, 0]))
    return rec[order].astype(np.int32)


def records_from_matrix(neighbor_matrix, num_neighbors, neighbor_matrix_shifts=None):
    """Padded matrix outputs -> sorted [P,5] records (uses num_neighbors, not the fill value)."""
    nm = _np(neighbor_matrix)
    num = _np(num_neighbors)
    n, m = nm.shape
    cols = np.arange(m)[None, :]
    mask = cols < np.minimum(num, m)[:, None]
    i = np.nonzero(mask)[0]
    j = nm[mask]
    if neighbor_matrix_shifts is None:
        s = np.zeros((i.shape[0], 3), dtype=np.int32)
    else:
        s = _np(neighbor_matrix_shifts)[mask]
    return sort_records(np.concatenate([i[:, None], j[:, None], s], axis=1))


def records_from_coo(neighbor_list, shifts=None):
    """COO outputs -> sorted [P,5] records."""
    nl = _np(neighbor_list)
    if shifts is None:
        s = np.zeros((nl.shape[1], 3), dtype=np.int32)
    else:
        s = _np(shifts).reshape(-1, 3)
    return sort_records(np.concatenate([nl.T.astype(np.int64), s.astype(np.int64)], axis=1))


def canonical_undirected(rec):
    """Map directed records to canonical undirected ones (for half_fill comparisons):
    (i,j,s) and (j,i,-s) collapse onto the representative with i<j, or i==j and s lexicographically > 0."""
    rec = np.asarray(rec, dtype=np.int64).reshape(-1, 5).copy()
    i, j, s = rec[:, 0], rec[:, 1], rec[:, 2:]
    s_neg = (s[:, 0] < 0) | ((s[:, 0] == 0) & ((s[:, 1] < 0) | ((s[:, 1] == 0) & (s[:, 2] < 0))))
    flip = (i > j) | ((i == j) & s_neg)
    rec[flip] = np.concatenate([j[flip, None], i[flip, None], -s[flip]], axis=1)
    return sort_records(rec)
