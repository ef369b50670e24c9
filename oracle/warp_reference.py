"""Adapter to the REAL reference (TEST INFRASTRUCTURE — only tests/, __graft_entry__.smoke() and bench.py may import it).

When ``warp`` (warp-lang >= 1.10) and the reference package ``nvalchemiops`` are importable in the running environment
(they are not in the build container: no network), the reference's own ``neighbor_list`` becomes the
primary oracle: ``reference_records`` runs it (Warp CPU device, or Warp CUDA when asked) and returns sorted
``(i, j, sx, sy, sz)`` records comparable with ``reference_oracle.records_from_*``.  ``decide_fma_mode`` evaluates the
knife-edge pairs of an input under both arithmetic variants of this repo (``config.fma`` True / False) and reports which
one reproduces the reference exactly — the one thing the restated oracle cannot pin.
Nothing here reads /root/reference.
"""
from __future__ import annotations

import importlib

import numpy as np


def available() -> bool:
    """True when both warp and the reference package can be imported here."""
    try:
        importlib.import_module("warp")
        importlib.import_module("nvalchemiops.neighborlist")
        return True
    except Exception:  # noqa: BLE001  (ImportError, or warp failing to initialise)
        return False


def why_unavailable() -> str:
    for name in ("warp", "nvalchemiops.neighborlist"):
        try:
            importlib.import_module(name)
        except Exception as e:  # noqa: BLE001
            return f"{name}: {type(e).__name__}: {e}"
    return ""


def reference_records(positions, cutoff, cell, pbc, batch_idx=None, batch_ptr=None, device="cpu", max_neighbors=None,
                      method=None):
    """Sorted (i, j, sx, sy, sz) int32 records of the real reference's neighbor_list on ``device``."""
    import torch
    from nvalchemiops.neighborlist import neighbor_list

    kw = {}
    if max_neighbors is not None:
        kw["max_neighbors"] = max_neighbors
    if method is not None:
        kw["method"] = method
    t = lambda x: None if x is None else x.to(device)  # noqa: E731
    out = neighbor_list(t(positions), cutoff, cell=t(cell), pbc=t(pbc), batch_idx=t(batch_idx), batch_ptr=t(batch_ptr),
                        return_neighbor_list=True, **kw)
    e, s = out[0].cpu().numpy(), out[2].cpu().numpy()
    rec = np.concatenate([e.T, s], axis=1).astype(np.int32)
    return rec[np.lexsort(rec.T[::-1])]


def decide_fma_mode(positions, cutoff, cell, pbc, ours, device="cpu", **kw):
    """``ours(fma: bool) -> sorted records`` of this repo's path.  Returns (mode or None, n_diff_fma, n_diff_nofma):
    mode is True / False when exactly that arithmetic reproduces the reference, None when both or neither do."""
    ref = reference_records(positions, cutoff, cell, pbc, device=device, **kw)

    def ndiff(a, b):
        if a.shape == b.shape and np.array_equal(a, b):
            return 0
        sa = {tuple(r) for r in a.tolist()}
        sb = {tuple(r) for r in b.tolist()}
        return len(sa ^ sb)

    d_fma, d_sep = ndiff(ours(True), ref), ndiff(ours(False), ref)
    mode = True if (d_fma == 0 and d_sep != 0) else (False if (d_sep == 0 and d_fma != 0) else None)
    return mode, d_fma, d_sep
