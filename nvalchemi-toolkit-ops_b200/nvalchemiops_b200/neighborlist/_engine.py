"""Thin driver of the C ABI: tensor checks, workspace, stream, and the two output paths.

This is the only module that touches ``libnvalchemi_nl_b200.so``.  There is no other
implementation behind it: CPU tensors are rejected, a missing library is an ImportError.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .. import _lib, config
from .neighbor_utils import NeighborOverflowError

_DTYPES = {torch.float32: 0, torch.float64: 1}

ERR_MESSAGES = {
    1: "an atom lies more than 1e6 periodic images from the cell, or the search radius exceeds 63 cells",
    2: "batch_idx contains a system index outside [0, num_systems)",
    4: "a cell matrix is singular",
    8: "the cell-list cache tensors are inconsistent (they were not produced by build_cell_list of this package)",
}


def _dtype_code(dtype: torch.dtype) -> int:
    if dtype not in _DTYPES:
        # same error class/message family as nvalchemiops/types.py:29
        raise ValueError(f"Unsupported dtype: {dtype}")
    return _DTYPES[dtype]


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(
            f"nvalchemiops_b200: `{name}` is on {t.device}; this package runs on CUDA (sm_100a) only "
            "and has no CPU fallback."
        )


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def cutoff_sq_in_dtype(cutoff: float, dtype: torch.dtype, python_double: bool = False) -> float:
    """rc^2 as the reference kernels see it: squared in the kernel precision (cell_list.py:444), or squared
    in Python double and then cast (naive.py:290) when ``python_double``."""
    if dtype == torch.float32:
        if python_double:
            return float(np.float32(cutoff * cutoff))
        c = np.float32(cutoff)
        return float(np.float32(c * c))
    return float(cutoff) * float(cutoff)


class CellListHandle:
    """Result of ``build``: the opaque device workspace plus what the queries need."""

    __slots__ = ("ws", "dtype_code", "n", "ns", "batch_idx", "dtype", "device", "cutoff", "rows_overflow", "wide_stencil")

    def __init__(self, ws, dtype_code, n, ns, batch_idx, dtype, device, cutoff):
        self.ws, self.dtype_code, self.n, self.ns = ws, dtype_code, n, ns
        self.batch_idx, self.dtype, self.device, self.cutoff = batch_idx, dtype, device, cutoff
        self.rows_overflow = False
        self.wide_stencil = False   # (known after status(): some system searches more than one cell per side)


def _stream(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def build(positions, cutoff, cell, pbc, batch_idx=None, batch_ptr=None, workspace=None, max_cells=0) -> CellListHandle:
    """Grid + hash + counting sort (nvnl_build).  ``cell`` [S,3,3], ``pbc`` [S,3] bool.  ``max_cells`` > 0 caps the
    grid at the capacity of a caller-provided reference-shaped cache (see include/nvalchemi_nl_b200.h)."""
    _require_cuda(positions, "positions")
    code = _dtype_code(positions.dtype)
    if positions.ndim != 2 or positions.shape[1] != 3:
        raise ValueError("positions must have shape (total_atoms, 3)")
    n = positions.shape[0]
    dev = positions.device
    positions = positions.contiguous()
    cell = cell.to(device=dev, dtype=positions.dtype).reshape(-1, 3, 3).contiguous()
    ns = cell.shape[0]
    pbc_u8 = pbc.to(device=dev).reshape(-1, 3).to(torch.uint8).contiguous()
    if pbc_u8.shape[0] != ns:
        raise ValueError(f"pbc has {pbc_u8.shape[0]} systems but cell has {ns}")
    if batch_idx is not None:
        batch_idx = batch_idx.to(device=dev, dtype=torch.int32).contiguous()
        if batch_idx.shape[0] != n:
            raise ValueError("batch_idx must have one entry per atom")
    elif ns > 1:
        raise ValueError("batch_idx is required when more than one cell is given")
    if batch_ptr is not None:
        batch_ptr = batch_ptr.to(device=dev, dtype=torch.int32).contiguous()
        if batch_ptr.shape[0] != ns + 1:
            raise ValueError("batch_ptr must have num_systems + 1 entries")
    L = _lib.lib()
    nbytes = int(L.nvnl_workspace_bytes(n, ns, code))
    if workspace is None or workspace.numel() < nbytes or workspace.device != dev:
        workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(
            L.nvnl_build(_ptr(positions), code, n, _ptr(cell), _ptr(pbc_u8), _ptr(batch_idx), _ptr(batch_ptr), ns,
                         float(cutoff), int(max_cells), _ptr(workspace), workspace.numel(), _stream(dev)),
            "nvnl_build",
        )
    # keep the inputs alive until the stream has consumed them
    workspace._nvnl_keepalive = (positions, cell, pbc_u8, batch_idx, batch_ptr)
    return CellListHandle(workspace, code, n, ns, batch_idx, positions.dtype, dev, float(cutoff))


def import_cache(positions, cutoff, cell, pbc, batch_idx, cells_per_dimension, neighbor_search_radius,
                 atom_periodic_shifts, atom_to_cell_mapping, atoms_per_cell_count, cell_atom_start_indices,
                 cell_atom_list) -> CellListHandle:
    """Rebuild a workspace from the VALUES of the seven reference-shaped cache tensors and the current positions
    (nvnl_import_cache) — the query side of the split build/query workflow, no hidden state."""
    _require_cuda(positions, "positions")
    code = _dtype_code(positions.dtype)
    n, dev = positions.shape[0], positions.device
    positions = positions.contiguous()
    cell = cell.to(device=dev, dtype=positions.dtype).reshape(-1, 3, 3).contiguous()
    ns = cell.shape[0]
    pbc_u8 = pbc.to(device=dev).reshape(-1, 3).to(torch.uint8).contiguous()
    if pbc_u8.shape[0] != ns:
        raise ValueError(f"pbc has {pbc_u8.shape[0]} systems but cell has {ns}")
    if batch_idx is not None:
        batch_idx = batch_idx.to(device=dev, dtype=torch.int32).contiguous()
    elif ns > 1:
        raise ValueError("batch_idx is required when more than one cell is given")
    cache = [cells_per_dimension, neighbor_search_radius, atom_periodic_shifts, atom_to_cell_mapping,
             atoms_per_cell_count, cell_atom_start_indices, cell_atom_list]
    for k, t in enumerate(cache):
        if t.dtype != torch.int32 or t.device != dev:
            raise ValueError("cell-list cache tensors must be int32 tensors on the positions' device")
        cache[k] = t.contiguous()
    if cache[0].numel() != 3 * ns or cache[2].numel() != 3 * n or cache[3].numel() != 3 * n or cache[6].numel() != n:
        raise ValueError("cell-list cache tensors do not match (total_atoms, num_systems)")
    rad = cache[1] if cache[1].numel() == 3 * ns else None
    ncache = min(cache[4].numel(), cache[5].numel())
    L = _lib.lib()
    nbytes = int(L.nvnl_workspace_bytes(n, ns, code))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(
            L.nvnl_import_cache(_ptr(positions), code, n, _ptr(cell), _ptr(pbc_u8), _ptr(batch_idx), ns, float(cutoff),
                                _ptr(cache[0]), _ptr(rad), _ptr(cache[2]), _ptr(cache[3]), _ptr(cache[4]), _ptr(cache[5]),
                                ncache, _ptr(cache[6]), _ptr(ws), ws.numel(), _stream(dev)),
            "nvnl_import_cache",
        )
    ws._nvnl_keepalive = (positions, cell, pbc_u8, batch_idx, cache)
    return CellListHandle(ws, code, n, ns, batch_idx, positions.dtype, dev, float(cutoff))


def cells_changed_cache(positions, cell, pbc, batch_idx, cells_per_dimension, atom_to_cell_mapping) -> torch.Tensor:
    """int32 device flag [1]: 1 if any atom's cell on the cached grid differs from ``atom_to_cell_mapping``
    (nvnl_cells_changed_cache; stateless)."""
    _require_cuda(positions, "current_positions")
    code = _dtype_code(positions.dtype)
    n, dev = positions.shape[0], positions.device
    positions = positions.contiguous()
    cell = cell.to(device=dev, dtype=positions.dtype).reshape(-1, 3, 3).contiguous()
    ns = cell.shape[0]
    pbc_u8 = pbc.to(device=dev).reshape(-1, 3).to(torch.uint8).contiguous()
    cpd = cells_per_dimension.to(device=dev, dtype=torch.int32).reshape(-1, 3).contiguous()
    amap = atom_to_cell_mapping.to(device=dev, dtype=torch.int32).contiguous()
    if cpd.shape[0] != ns or pbc_u8.shape[0] != ns or amap.numel() != 3 * n:
        raise ValueError("cells_per_dimension / pbc / atom_to_cell_mapping do not match (total_atoms, num_systems)")
    if batch_idx is not None:
        batch_idx = batch_idx.to(device=dev, dtype=torch.int32).contiguous()
    elif ns > 1:
        raise ValueError("batch_idx is required when more than one cell is given")
    flag = torch.empty(1, dtype=torch.int32, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev):
        _lib.check(L.nvnl_cells_changed_cache(_ptr(positions), code, n, _ptr(cell), _ptr(pbc_u8), _ptr(batch_idx), ns,
                                              _ptr(cpd), _ptr(amap), _ptr(flag), _stream(dev)),
                   "nvnl_cells_changed_cache")
    return flag


def _raise_on_error_bits(bits: int):
    if bits:
        msgs = [m for b, m in ERR_MESSAGES.items() if bits & b]
        raise ValueError("nvalchemiops_b200: invalid input: " + "; ".join(msgs))


def status(h: CellListHandle):
    """(total_pairs, max_count, total_cells, error_bits, launch_hint) — synchronizes the stream.
    launch_hint (bit 0: atoms outside the primary image, bit 1: cells left to the general kernel, bit 5: single-cell
    systems that take the second single-sweep launch) lets the next stage launch only the kernels that have work."""
    L = _lib.lib()
    tp, mc, tc, eb, uw, hd, ro = (ctypes.c_int64(0), ctypes.c_int32(0), ctypes.c_int32(0), ctypes.c_int32(0),
                                  ctypes.c_int32(0), ctypes.c_int32(0), ctypes.c_int32(0))
    with torch.cuda.device(h.device):
        _lib.check(
            L.nvnl_status(_ptr(h.ws), h.dtype_code, h.n, h.ns, ctypes.byref(tp), ctypes.byref(mc), ctypes.byref(tc),
                          ctypes.byref(eb), ctypes.byref(uw), ctypes.byref(hd), ctypes.byref(ro), _stream(h.device)),
            "nvnl_status",
        )
    h.rows_overflow = bool(ro.value)
    h.wide_stencil = bool(uw.value & 2)
    return (tp.value, mc.value, tc.value, eb.value,
            (1 if (uw.value & 1) else 0) | (2 if (hd.value & 1) else 0) | (32 if (hd.value & 2) else 0))


def query_matrix(h: CellListHandle, cutoff_sq, neighbor_matrix, neighbor_matrix_shifts, num_neighbors, fill_value,
                 half_fill=False, pad_rows=True):
    """Fill the three padded outputs in place (nvnl_fill_matrix): no host sync, graph-capturable.
    ``pad_rows=False`` leaves the unused slots of every row untouched (the bare query op)."""
    for t, nm in ((neighbor_matrix, "neighbor_matrix"), (neighbor_matrix_shifts, "neighbor_matrix_shifts"),
                  (num_neighbors, "num_neighbors")):
        _require_cuda(t, nm)
        if t.dtype != torch.int32 or not t.is_contiguous():
            raise ValueError(f"{nm} must be a contiguous int32 tensor")
    M = neighbor_matrix.shape[1]
    if neighbor_matrix.shape[0] != h.n or neighbor_matrix_shifts.shape[:2] != (h.n, M) or num_neighbors.shape[0] != h.n:
        raise ValueError("output tensors do not match (total_atoms, max_neighbors)")
    L = _lib.lib()
    if config.check_inputs:
        # debug switch: the padded-matrix path is sync-free and does not read the device error word by default
        # (an out-of-range batch_idx is clamped, an over-long search radius truncated); this costs one host sync
        _raise_on_error_bits(status(h)[3])
    with torch.cuda.device(h.device):
        _lib.check(
            L.nvnl_fill_matrix(_ptr(h.ws), h.dtype_code, h.n, h.ns, _ptr(h.batch_idx), float(cutoff_sq),
                               int(bool(half_fill)), int(bool(config.fma)), _ptr(neighbor_matrix),
                               _ptr(neighbor_matrix_shifts), _ptr(num_neighbors), M, int(fill_value),
                               int(bool(pad_rows)), _stream(h.device)),
            "nvnl_fill_matrix",
        )


def use_rows(h: CellListHandle) -> bool:
    """fp32 inputs take the single-sweep COO path unless config.coo_path says otherwise."""
    if config.coo_path not in ("rows", "masks"):
        raise ValueError(f"config.coo_path must be 'rows' or 'masks', not {config.coo_path!r}")
    return config.coo_path == "rows" and h.dtype == torch.float32 and h.n < (1 << 27)


class _QueryHistory:
    """What the last COO query with a given signature (device, stream, atoms, systems, cutoff^2, half_fill) reported: its
    pair count and launch hint.  The next query with the same signature (MD steps, repeated evaluations) uses them to
    hand the sweep a shifts buffer to zero-fill WHILE it sweeps, to launch the output kernel before the size sync, and to
    skip kernel variants that had no work.  Only sizes and flags are remembered — never an output — and every guess is
    verified against the device's own report after the sync; a wrong guess costs a repeated stage, nothing else.
    Thread-safe; keyed per stream so that concurrent streams do not share entries."""

    def __init__(self, capacity=256):
        import threading

        self._lock = threading.Lock()
        self._entries: dict = {}
        self._capacity = capacity

    def get(self, key):
        with self._lock:
            return self._entries.get(key)

    def put(self, key, total, hint):
        with self._lock:
            if len(self._entries) >= self._capacity and key not in self._entries:
                self._entries.clear()
            self._entries[key] = (int(total), int(hint))

    def clear(self):
        with self._lock:
            self._entries.clear()


_pair_history = _QueryHistory()


def _history_key(h: CellListHandle, cutoff_sq, half_fill):
    dev = torch.device(h.device)
    stream = torch.cuda.current_stream(dev).cuda_stream if dev.type == "cuda" else 0
    return (dev.index or 0, stream, h.n, h.ns, float(cutoff_sq), bool(half_fill))


def _guess_shifts_buffer(h: CellListHandle, last):
    """Speculatively allocate a shifts buffer sized from the last query with this signature (or None)."""
    guess, last_hint = last if last is not None else (0, 0)
    # (inputs that were outside the primary periodic image last time take the two-pass kernels: nothing to overlap)
    if not config.prezero_shifts or not guess or guess < config.prezero_min_pairs or (last_hint & 1):
        return None
    return torch.empty(3 * (int(guess * 1.02) + 1024), dtype=torch.int32, device=h.device)


def count(h: CellListHandle, cutoff_sq, half_fill=False, want_ptr=True, rows=False, prezero=None, launch_hint=-1):
    """num_neighbors [N] and (optionally) neighbor_ptr [N+1] (nvnl_count / nvnl_count_rows); asynchronous.
    ``rows=True`` also leaves every atom's neighbors in the workspace's temporary row buffer for ``fill_coo(rows=True)``."""
    num = torch.empty(h.n, dtype=torch.int32, device=h.device)
    ptr = torch.empty(h.n + 1, dtype=torch.int32, device=h.device) if want_ptr else None
    L = _lib.lib()
    with torch.cuda.device(h.device):
        if rows:
            _lib.check(
                L.nvnl_count_rows(_ptr(h.ws), h.dtype_code, h.n, h.ns, _ptr(h.batch_idx), float(cutoff_sq),
                                  int(bool(half_fill)), int(bool(config.fma)), _ptr(num), _ptr(ptr), _ptr(prezero),
                                  prezero.numel() if prezero is not None else 0, int(launch_hint), _stream(h.device)),
                "nvnl_count_rows",
            )
        else:
            _lib.check(
                L.nvnl_count(_ptr(h.ws), h.dtype_code, h.n, h.ns, _ptr(h.batch_idx), float(cutoff_sq),
                             int(bool(half_fill)), int(bool(config.fma)), _ptr(num), _ptr(ptr), _stream(h.device)),
                "nvnl_count",
            )
    return num, ptr


def fill_coo(h: CellListHandle, cutoff_sq, neighbor_ptr, edge_index, shifts, num_pairs, half_fill=False,
             index_offset=0, launch_hint=-1, rows=False, row_stride=0):
    """Write COO rows at neighbor_ptr (nvnl_fill_coo, or nvnl_fill_rows after ``count(rows=True)``).  ``edge_index``
    is [2, num_pairs], ``shifts`` [num_pairs, 3]; with ``row_stride`` > 0 the two rows of ``edge_index`` are
    ``row_stride`` entries apart (a rank writing its own range of a larger global array)."""
    L = _lib.lib()
    fn, name = (L.nvnl_fill_rows, "nvnl_fill_rows") if rows else (L.nvnl_fill_coo, "nvnl_fill_coo")
    with torch.cuda.device(h.device):
        _lib.check(
            fn(_ptr(h.ws), h.dtype_code, h.n, h.ns, _ptr(h.batch_idx), float(cutoff_sq),
               int(bool(half_fill)), int(bool(config.fma)), _ptr(neighbor_ptr), _ptr(edge_index),
               int(num_pairs), int(row_stride), _ptr(shifts), int(index_offset), int(launch_hint), _stream(h.device)),
            name,
        )


def query_coo(h: CellListHandle, cutoff_sq, half_fill=False, max_neighbors=None):
    """COO outputs ``(neighbor_list [2,P], neighbor_ptr [N+1], shifts [P,3])``: count -> scan -> one sync for the
    size -> fill.  Raises NeighborOverflowError like the reference's COO conversion when an atom exceeds
    ``max_neighbors`` (neighbor_utils.py:352-359)."""
    key = _history_key(h, cutoff_sq, half_fill)
    last = _pair_history.get(key)
    zbuf = _guess_shifts_buffer(h, last) if use_rows(h) else None
    ebuf = None
    if zbuf is not None and config.speculative_fill:
        ebuf = torch.empty(2 * (zbuf.numel() // 3), dtype=torch.int32, device=h.device)
    num, ptr, total, max_count, err, hint, rows = count_and_size(h, cutoff_sq, half_fill, prezero=zbuf, spec_edge=ebuf,
                                                                 launch_hint=last[1] if last is not None else -1)
    _raise_on_error_bits(err)
    if max_neighbors is not None and max_count > max_neighbors:
        raise NeighborOverflowError(max_neighbors, max_count)
    if total > 2**31 - 1:
        raise OverflowError(f"{total} pairs do not fit int32 neighbor_ptr/neighbor_list indices")
    _pair_history.put(key, total, hint & 35)
    fits = (rows and zbuf is not None and 3 * total <= zbuf.numel() and 2 * 3 * total >= zbuf.numel()
            and not (hint & 1) and not (hint & 16))
    hint &= 15
    if fits and ebuf is not None:
        # the output kernel already ran (before the sync) into the speculative buffers: the outputs are their prefixes
        edge_index = ebuf[: 2 * total].view(2, total)
        shifts = zbuf[: 3 * total].view(total, 3)
        if total > 0 and (hint & 2):   # only the rows of the general kernel are still missing
            fill_coo(h, cutoff_sq, ptr, edge_index, shifts, total, half_fill, launch_hint=hint | 4 | 8, rows=True)
        return edge_index, ptr, shifts, num
    if fits:
        # the speculative buffer fits: shifts is its (contiguous) prefix, already zero when the fill starts
        edge_index = torch.empty((2, total), dtype=torch.int32, device=h.device)
        shifts = zbuf[: 3 * total].view(total, 3)
        hint |= 4
    else:
        # one allocation for both outputs (the GPU idles between the size sync and the first fill launch)
        buf = torch.empty(5 * total, dtype=torch.int32, device=h.device)
        edge_index = buf[: 2 * total].view(2, total)
        shifts = buf[2 * total:].view(total, 3)
    if total > 0:
        fill_coo(h, cutoff_sq, ptr, edge_index, shifts, total, half_fill, launch_hint=hint, rows=rows)
    return edge_index, ptr, shifts, num


def fill_rows_speculative(h: CellListHandle, neighbor_ptr, edge_buffer, shifts_zeroed, index_offset=0):
    """EXPERIMENTAL: launch the single-sweep path's output kernel before the pair count is known on the host
    (nvnl_fill_rows_speculative); ``edge_buffer`` holds 2 * cap and ``shifts_zeroed`` 3 * cap int32."""
    L = _lib.lib()
    with torch.cuda.device(h.device):
        _lib.check(
            L.nvnl_fill_rows_speculative(_ptr(h.ws), h.dtype_code, h.n, h.ns, _ptr(neighbor_ptr), _ptr(edge_buffer),
                                         edge_buffer.numel() // 2, _ptr(shifts_zeroed), int(index_offset), _stream(h.device)),
            "nvnl_fill_rows_speculative",
        )


def count_and_size(h: CellListHandle, cutoff_sq, half_fill=False, prezero=None, spec_edge=None, launch_hint=-1):
    """Count stage + the one host sync: ``(num, ptr, total, max_count, error_bits, launch_hint, rows)``.  ``rows``
    tells ``fill_coo`` which path the count ran on (single sweep unless fp64 / configured off / its temporary row
    buffer overflowed, in which case the count is repeated on the two-pass path).  ``launch_hint`` >= 0 (from an earlier
    query with this signature) skips the kernel variants that had no work then; if the device reports work for a
    skipped variant the count is repeated with everything launched."""
    pending = count_launch(h, cutoff_sq, half_fill, prezero, spec_edge, launch_hint)
    return count_finish(h, cutoff_sq, half_fill, pending)


def count_launch(h: CellListHandle, cutoff_sq, half_fill=False, prezero=None, spec_edge=None, launch_hint=-1):
    """First half of ``count_and_size``: the launches, no host sync (several independent cell lists can be queued before
    the first ``count_finish`` waits)."""
    rows = use_rows(h)
    num, ptr = count(h, cutoff_sq, half_fill, rows=rows, prezero=prezero if rows else None,
                     launch_hint=launch_hint if rows else -1)
    if rows and prezero is not None and spec_edge is not None:
        fill_rows_speculative(h, ptr, spec_edge, prezero)      # runs while the host waits for the size
    return num, ptr, rows, launch_hint


def count_finish(h: CellListHandle, cutoff_sq, half_fill, pending):
    """Second half of ``count_and_size``: the host sync for the sizes and the (rare) repeated counts."""
    num, ptr, rows, launch_hint = pending
    total, max_count, _cells, err, hint = status(h)
    if rows and launch_hint >= 0 and (hint & ~launch_hint & 35):
        # a variant that was not launched had work: repeat the count with every variant (the speculative outputs, if any,
        # are discarded by the caller because the totals cannot have been right)
        num, ptr = count(h, cutoff_sq, half_fill, rows=True, prezero=None, launch_hint=-1)
        total, max_count, _cells, err, hint = status(h)
        hint |= 16   # (tells the caller that speculative buffers were not filled consistently)
    if rows and h.rows_overflow:
        rows = False
        num, ptr = count(h, cutoff_sq, half_fill)
        total, max_count, _cells, err, hint = status(h)
        h.rows_overflow = True  # sticky: this query did not fit the temporary row buffer
    return num, ptr, total, max_count, err, hint, rows


def get_grid(h: CellListHandle):
    cpd = torch.empty((h.ns, 3), dtype=torch.int32, device=h.device)
    rad = torch.empty((h.ns, 3), dtype=torch.int32, device=h.device)
    L = _lib.lib()
    with torch.cuda.device(h.device):
        _lib.check(L.nvnl_get_grid(_ptr(h.ws), h.dtype_code, h.n, h.ns, _ptr(cpd), _ptr(rad), _stream(h.device)),
                   "nvnl_get_grid")
    return cpd, rad


def export_cache(h: CellListHandle, cells_per_dimension=None, neighbor_search_radius=None, atom_periodic_shifts=None,
                 atom_to_cell_mapping=None, atoms_per_cell_count=None, cell_atom_start_indices=None, cell_atom_list=None):
    """Write the reference-shaped cache tensors (int32, contiguous, on the handle's device) that are not None."""
    tensors = (cells_per_dimension, neighbor_search_radius, atom_periodic_shifts, atom_to_cell_mapping,
               atoms_per_cell_count, cell_atom_start_indices, cell_atom_list)
    for t in tensors:
        if t is not None and (t.dtype != torch.int32 or not t.is_contiguous() or t.device != h.device):
            raise ValueError("cell-list cache tensors must be contiguous int32 tensors on the positions' device")
    ncache = 0
    for t in (atoms_per_cell_count, cell_atom_start_indices):
        if t is not None:
            ncache = t.numel() if ncache == 0 else min(ncache, t.numel())
    L = _lib.lib()
    with torch.cuda.device(h.device):
        _lib.check(
            L.nvnl_export_cache(_ptr(h.ws), h.dtype_code, h.n, h.ns, _ptr(h.batch_idx), _ptr(cells_per_dimension),
                                _ptr(neighbor_search_radius), _ptr(atom_periodic_shifts), _ptr(atom_to_cell_mapping),
                                _ptr(atoms_per_cell_count), _ptr(cell_atom_start_indices), ncache, _ptr(cell_atom_list),
                                _stream(h.device)),
            "nvnl_export_cache",
        )


def refresh_positions(h: CellListHandle, positions):
    """Re-gather ``positions`` into the sorted records without re-binning (stale-cache query)."""
    _require_cuda(positions, "positions")
    if positions.shape[0] != h.n or positions.dtype != h.dtype:
        raise ValueError("positions do not match the cell list they are queried against")
    positions = positions.contiguous()
    L = _lib.lib()
    with torch.cuda.device(h.device):
        _lib.check(L.nvnl_refresh_positions(_ptr(h.ws), h.dtype_code, h.n, h.ns, _ptr(positions), _stream(h.device)),
                   "nvnl_refresh_positions")
    h.ws._nvnl_keepalive_pos = positions


def cells_changed(h: CellListHandle, positions) -> torch.Tensor:
    """int32 device flag [1]: 1 if any atom left the cell it was binned into by the last build."""
    _require_cuda(positions, "current_positions")
    positions = positions.contiguous()
    flag = torch.empty(1, dtype=torch.int32, device=h.device)
    L = _lib.lib()
    with torch.cuda.device(h.device):
        _lib.check(L.nvnl_cells_changed(_ptr(h.ws), h.dtype_code, h.n, h.ns, _ptr(positions), _ptr(h.batch_idx), _ptr(flag),
                                        _stream(h.device)), "nvnl_cells_changed")
    return flag


def moved_beyond(reference_positions, current_positions, threshold) -> torch.Tensor:
    _require_cuda(current_positions, "current_positions")
    _require_cuda(reference_positions, "reference_positions")
    code = _dtype_code(current_positions.dtype)
    if reference_positions.shape != current_positions.shape or reference_positions.dtype != current_positions.dtype:
        raise ValueError("reference_positions and current_positions must have the same shape and dtype")
    ref, cur = reference_positions.contiguous(), current_positions.contiguous()
    flag = torch.empty(1, dtype=torch.int32, device=cur.device)
    L = _lib.lib()
    with torch.cuda.device(cur.device):
        _lib.check(L.nvnl_moved_beyond(_ptr(ref), _ptr(cur), code, cur.shape[0], float(threshold), _ptr(flag),
                                       _stream(cur.device)), "nvnl_moved_beyond")
    return flag
