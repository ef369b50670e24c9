"""Real-space Coulomb / Ewald energies and forces on the B200 path.

Two entry points:

* ``coulomb_energy_forces`` / ``coulomb_energy`` / ``coulomb_forces`` — the reference's signatures
  (nvalchemiops/interactions/electrostatics/coulomb.py:1336, :1492, :1540): the consumer over an EXISTING neighbor list
  (COO ``neighbor_list`` + ``neighbor_ptr`` + ``neighbor_shifts``, or ``neighbor_matrix`` + ``neighbor_matrix_shifts``).
* ``fused_coulomb_energy_forces`` — SURVEY.md §8f rank 2: the same numbers straight from positions, with the stencil sweep and
  the pair consumer in ONE kernel, so the neighbor list (20 B per pair of HBM traffic) is never written.  Equivalent to
  ``neighbor_list(positions, cutoff, cell, pbc, ..., return_neighbor_list=True)`` followed by ``coulomb_energy_forces``.

Forward values only: the reference's autograd bridge (``warp_custom_op``) is tensor plumbing outside the hot path
(BASELINE.json north_star: "nvalchemiops.types tensor plumbing and autograd hooks stay").  float64 results like the reference
(coulomb.py:1626-1628 converts its inputs).  CUDA only; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes

import torch

from ... import _lib, config
from ...neighborlist import _engine


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _prep(positions, charges, cell, batch_idx, need_batch=True):
    _engine._require_cuda(positions, "positions")
    code = _engine._dtype_code(positions.dtype)
    dev = positions.device
    positions = positions.contiguous()
    n = positions.shape[0]
    q = charges.to(device=dev, dtype=torch.float64).reshape(-1).contiguous()
    if q.shape[0] != n:
        raise ValueError("charges must have one entry per atom")
    cell = cell.to(device=dev, dtype=positions.dtype).reshape(-1, 3, 3).contiguous()
    if batch_idx is not None:
        batch_idx = batch_idx.to(device=dev, dtype=torch.int32).contiguous()
    elif cell.shape[0] > 1 and need_batch:
        raise ValueError("batch_idx is required when more than one cell is given")
    return positions, q, cell, batch_idx, code, n, dev


def _list_consumer(positions, q, cell, batch_idx, code, cutoff, alpha, neighbor_ptr, neighbors, shifts, max_neighbors, fill_value):
    n, dev = positions.shape[0], positions.device
    energies = torch.empty(n, dtype=torch.float64, device=dev)
    forces = torch.empty((n, 3), dtype=torch.float64, device=dev)
    if n == 0:
        return energies, forces
    with torch.cuda.device(dev):
        _lib.check(
            _lib.lib().nvnl_coulomb_list(_ptr(positions), code, n, _ptr(cell), cell.shape[0], _ptr(batch_idx), _ptr(q),
                                         float(cutoff), float(alpha), _ptr(neighbor_ptr), _ptr(neighbors), _ptr(shifts),
                                         int(max_neighbors), int(fill_value), _ptr(energies), _ptr(forces),
                                         _engine._stream(dev)),
            "nvnl_coulomb_list")
    return energies, forces


def coulomb_energy_forces(positions, charges, cell, cutoff, alpha=0.0, neighbor_list=None, neighbor_ptr=None,
                          neighbor_shifts=None, neighbor_matrix=None, neighbor_matrix_shifts=None, fill_value=None,
                          batch_idx=None):
    """Per-atom energies [N] and forces [N,3] (float64) from an existing neighbor list — reference signature and checks
    (coulomb.py:1540-1700): exactly one of the COO and the matrix format; ``neighbor_ptr`` is required with the COO one."""
    use_list = neighbor_list is not None and neighbor_shifts is not None
    use_matrix = neighbor_matrix is not None and neighbor_matrix_shifts is not None
    if not use_list and not use_matrix:
        raise ValueError("Must provide either neighbor_list/neighbor_shifts or neighbor_matrix/neighbor_matrix_shifts")
    if use_list and use_matrix:
        raise ValueError("Cannot provide both neighbor list and neighbor matrix formats")
    positions, q, cell, batch_idx, code, n, dev = _prep(positions, charges, cell, batch_idx)
    if use_list:
        if neighbor_ptr is None:
            raise ValueError("neighbor_ptr is required when using neighbor_list format")
        idx_j = neighbor_list[1].to(device=dev, dtype=torch.int32).contiguous()
        ptr = neighbor_ptr.to(device=dev, dtype=torch.int32).contiguous()
        sh = neighbor_shifts.to(device=dev, dtype=torch.int32).contiguous()
        return _list_consumer(positions, q, cell, batch_idx, code, cutoff, alpha, ptr, idx_j, sh, 0, n)
    nm = neighbor_matrix.to(device=dev, dtype=torch.int32).contiguous()
    sh = neighbor_matrix_shifts.to(device=dev, dtype=torch.int32).contiguous()
    fv = n if fill_value is None else int(fill_value)
    return _list_consumer(positions, q, cell, batch_idx, code, cutoff, alpha, None, nm, sh, nm.shape[1], fv)


def coulomb_energy(*args, **kwargs):
    """Energies only (coulomb.py:1336)."""
    return coulomb_energy_forces(*args, **kwargs)[0]


def coulomb_forces(*args, **kwargs):
    """Forces only (coulomb.py:1492)."""
    return coulomb_energy_forces(*args, **kwargs)[1]


def fused_coulomb_energy_forces(positions, charges, cell, pbc, cutoff, alpha=0.0, batch_idx=None, batch_ptr=None,
                                return_path=False):
    """Energies [N] and forces [N,3] (float64) straight from positions: cell-list build, then ONE kernel that sweeps the
    stencil and evaluates the pair terms of the hits it finds — the neighbor list is never materialised.

    Same result (up to fp64 summation order) as the list path: ``neighbor_list(...)`` with this package's semantics (fp32
    predicate ``d^2 < cutoff^2``, full list) followed by ``coulomb_energy_forces``.  Inputs the fused sweep does not
    cover (float64 positions, atoms outside the primary periodic image, stencils wider than one cell, over-full cells)
    are detected on the device and served by the list path inside this call.  ``return_path=True`` appends "fused" or
    "list" (which path produced the numbers)."""
    from ...neighborlist.neighbor_utils import _prepare_batch_idx_ptr

    pos, q, cell3, _, code, n, dev = _prep(positions, charges, cell, batch_idx, need_batch=False)
    ns = cell3.shape[0]
    if batch_idx is not None or batch_ptr is not None:
        batch_idx, batch_ptr = _prepare_batch_idx_ptr(batch_idx, batch_ptr, n, dev)
    elif ns > 1:
        raise ValueError("Either batch_idx or batch_ptr must be given for more than one system")
    pbc = pbc.to(device=dev).reshape(-1, 3)
    if n == 0:
        out = (torch.zeros(0, dtype=torch.float64, device=dev), torch.zeros((0, 3), dtype=torch.float64, device=dev))
        return out + ("fused",) if return_path else out
    h = _engine.build(pos, cutoff, cell3, pbc, batch_idx=batch_idx, batch_ptr=batch_ptr)
    csq = _engine.cutoff_sq_in_dtype(cutoff, pos.dtype)
    if pos.dtype == torch.float32 and n < (1 << 27):
        energies = torch.empty(n, dtype=torch.float64, device=dev)
        forces = torch.empty((n, 3), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(
                _lib.lib().nvnl_coulomb_fused(_ptr(h.ws), h.dtype_code, n, ns, _ptr(h.batch_idx), float(csq),
                                              int(bool(config.fma)), _ptr(q), float(cutoff), float(alpha), _ptr(energies),
                                              _ptr(forces), _engine._stream(dev)),
                "nvnl_coulomb_fused")
        _total, _mc, _cells, err, hint = _engine.status(h)
        _engine._raise_on_error_bits(err)
        if not (hint & 3):
            return (energies, forces, "fused") if return_path else (energies, forces)
    # list path: the cell list is already built
    edge, ptr, shifts, _num = _engine.query_coo(h, csq)
    e, f = _list_consumer(pos, q, cell3, h.batch_idx, code, cutoff, alpha, ptr, edge[1].contiguous(), shifts, 0, n)
    return (e, f, "list") if return_path else (e, f)
