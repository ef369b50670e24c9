"""CPU oracle of the pair consumer fused with the sweep — TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's checks, never by the product package.

Restates, in numpy float64, the reference's real-space Coulomb / Ewald kernel for the neighbor-LIST format
(nvalchemiops/interactions/electrostatics/coulomb.py:206-292, _coulomb_energy_forces_kernel; batch variant :493-572) and
its erfc (nvalchemiops/math/math.py:52-93, wp_erfc: Abramowitz & Stegun 7.1.26, not the exact function).

Pinning: the reference needs warp-lang, which is not installed here, so it cannot be run to generate vectors; this
restatement is pinned by closed-form cases the reference's formulas imply (two charges: E_i = q1 q2 / (2 r),
F = q1 q2 / r^2 along r_ij; alpha > 0 with the A&S erfc evaluated independently) in tests/test_oracle_cpu.py —
"parity unpinned" against real Warp output for this consumer.
"""
from __future__ import annotations

import numpy as np

TWO_OVER_SQRT_PI = 1.1283791670955126   # coulomb.py:266


def wp_erfc(x):
    """math/math.py:52-93."""
    x = np.asarray(x, dtype=np.float64)
    ax = np.abs(x)
    t = 1.0 / (1.0 + 0.3275911 * ax)
    t2 = t * t
    t3 = t2 * t
    t4 = t3 * t
    t5 = t4 * t
    poly = 0.254829592 * t + -0.284496736 * t2 + 1.421413741 * t3 + -1.453152027 * t4 + 1.061405429 * t5
    v = poly * np.exp(-ax * ax)
    return np.where(x >= 0.0, v, 2.0 - v)


def coulomb_energy_forces_list(positions, charges, cell, cutoff, alpha, idx_i, idx_j, shifts, batch_idx=None):
    """coulomb.py:206-292 over COO arrays (idx_i = source atom of every entry, i.e. neighbor_list[0]).
    positions [N,3], charges [N], cell [S,3,3] (rows = lattice vectors), shifts [P,3] int.  Returns (E [N], F [N,3])."""
    pos = np.asarray(positions, dtype=np.float64)
    q = np.asarray(charges, dtype=np.float64).reshape(-1)
    cell = np.asarray(cell, dtype=np.float64).reshape(-1, 3, 3)
    i = np.asarray(idx_i, dtype=np.int64)
    j = np.asarray(idx_j, dtype=np.int64)
    s = np.asarray(shifts, dtype=np.float64).reshape(-1, 3)
    n = pos.shape[0]
    sys_i = np.zeros(i.shape[0], dtype=np.int64) if batch_idx is None else np.asarray(batch_idx, dtype=np.int64)[i]
    # shift_vec = cell^T · s  (coulomb.py:243 / :531)
    shift_vec = np.einsum("pk,pkd->pd", s, cell[sys_i])
    r_ij = pos[i] - pos[j] - shift_vec
    r = np.sqrt((r_ij * r_ij).sum(1))
    keep = ~((r >= cutoff) | (r < 1e-10))
    i, j, r_ij, r = i[keep], j[keep], r_ij[keep], r[keep]
    pre = 0.5 * q[i] * q[j]
    if alpha > 0.0:
        ar = alpha * r
        erfc_t = wp_erfc(ar)
        exp_t = np.exp(-(ar * ar))
        e = pre * erfc_t / r
        fm = pre * (erfc_t / (r * r * r) + TWO_OVER_SQRT_PI * alpha * exp_t / (r * r))
    else:
        e = pre / r
        fm = pre / (r * r * r)
    f = fm[:, None] * r_ij
    energies = np.zeros(n, dtype=np.float64)
    forces = np.zeros((n, 3), dtype=np.float64)
    np.add.at(energies, i, e)
    np.add.at(forces, i, f)
    np.add.at(forces, j, -f)
    return energies, forces
