"""Synthetic systems for the tests and the benchmark.

Generators follow the reference's test helpers (test/neighborlist/test_utils.py:32-249: simple cubic, random
``torch.rand(n,3)*L`` with ``torch.manual_seed(seed)``, triclinic from lattice parameters) and SURVEY.md §8(d)
for the BASELINE.json configurations (density 0.1 atoms/A^3, cutoff 6 A).
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_kat():
    with open(os.path.join(GOLDEN, "kat_structures.json")) as f:
        return json.load(f)


def kat_structure(name, dtype=torch.float32, device="cpu"):
    g = load_kat()[name]
    return (torch.tensor(g["positions"], dtype=dtype, device=device), torch.tensor(g["cell"], dtype=dtype, device=device),
            torch.tensor(g["pbc"], dtype=torch.bool, device=device), g["num_neighbors"])


def random_system(num_atoms=50, cell_size=5.0, dtype=torch.float32, seed=42, pbc_flag=True):
    g = torch.Generator().manual_seed(seed)
    positions = torch.rand(num_atoms, 3, dtype=dtype, generator=g) * cell_size
    cell = (torch.eye(3, dtype=dtype) * cell_size).reshape(1, 3, 3)
    pbc = torch.tensor(pbc_flag if isinstance(pbc_flag, list) else [pbc_flag] * 3).reshape(1, 3)
    return positions, cell, pbc


def triclinic_cell(a, b, c, alpha, beta, gamma, dtype=torch.float32):
    al, be, ga = np.deg2rad([alpha, beta, gamma])
    ca, cb, cg, sg = np.cos(al), np.cos(be), np.cos(ga), np.sin(ga)
    cz = c * np.sqrt(1.0 - ca**2 - cb**2 - cg**2 + 2.0 * ca * cb * cg) / sg
    m = np.array([[a, 0, 0], [b * cg, b * sg, 0], [c * cb, c * (ca - cb * cg) / sg, cz]], dtype=np.float64)
    return torch.tensor(m, dtype=dtype).reshape(1, 3, 3)


def triclinic_system(num_atoms=50, a=5.0, b=5.0, c=5.0, alpha=60.0, beta=60.0, gamma=60.0, dtype=torch.float32, seed=42,
                     pbc_flag=True, spread=(0.0, 1.0)):
    g = torch.Generator().manual_seed(seed)
    cell = triclinic_cell(a, b, c, alpha, beta, gamma, dtype)
    frac = torch.rand(num_atoms, 3, dtype=dtype, generator=g) * (spread[1] - spread[0]) + spread[0]
    positions = frac @ cell[0]
    pbc = torch.tensor(pbc_flag if isinstance(pbc_flag, list) else [pbc_flag] * 3).reshape(1, 3)
    return positions, cell, pbc


def bench_box(num_atoms, seed, density=0.1, dtype=torch.float32):
    """Uniform random periodic cubic box at the benchmark density: L rounded to fp32 first."""
    L = float(np.float32((num_atoms / density) ** (1.0 / 3.0)))
    g = torch.Generator().manual_seed(seed)
    positions = torch.rand(num_atoms, 3, dtype=dtype, generator=g) * L
    cell = (torch.eye(3, dtype=dtype) * L).reshape(1, 3, 3)
    pbc = torch.tensor([True, True, True]).reshape(1, 3)
    return positions, cell, pbc


def bench_batch(num_systems, atoms_lo, atoms_hi, seed, density=0.1, mixed_pbc=True, dtype=torch.float32):
    """Batch of cubic systems (configs 3 and 5): n_s ~ randint(lo, hi+1), L_s = (n_s/rho)^(1/3), pbc pattern
    s mod 8 as bits when ``mixed_pbc`` else fully periodic."""
    g = torch.Generator().manual_seed(seed)
    if atoms_lo == atoms_hi:
        counts = torch.full((num_systems,), atoms_lo, dtype=torch.int64)
    else:
        counts = torch.randint(atoms_lo, atoms_hi + 1, (num_systems,), generator=g)
    Ls = torch.tensor([float(np.float32((int(c) / density) ** (1.0 / 3.0))) for c in counts], dtype=dtype)
    total = int(counts.sum())
    batch_ptr = torch.zeros(num_systems + 1, dtype=torch.int32)
    batch_ptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
    batch_idx = torch.repeat_interleave(torch.arange(num_systems, dtype=torch.int32), counts)
    positions = torch.rand(total, 3, dtype=dtype, generator=g) * Ls[batch_idx.long()].unsqueeze(1)
    cell = torch.eye(3, dtype=dtype).unsqueeze(0) * Ls.reshape(-1, 1, 1)
    if mixed_pbc:
        s = torch.arange(num_systems)
        pbc = torch.stack([(s & 1) > 0, (s & 2) > 0, (s & 4) > 0], dim=1)
    else:
        pbc = torch.ones((num_systems, 3), dtype=torch.bool)
    return positions, cell, pbc, batch_idx, batch_ptr


def load_published_fcc():
    """Pair counts the reference publishes for its own FCC benchmark workload (tests/golden/make_fcc_golden.py)."""
    with open(os.path.join(GOLDEN, "reference_published_fcc.json")) as f:
        return json.load(f)


def fcc_benchmark_system(num_atoms, lattice_constant=4.0, dtype=torch.float32):
    """The reference benchmark's crystal (benchmarks/systems.py:874-971, restated vectorised): the first ``num_atoms``
    sites of the n^3 FCC supercell, n = ceil((num_atoms / 4)^(1/3)), enumerated in (i, j, k, basis) order; cubic
    periodic cell of edge n * a."""
    n = int(np.ceil((num_atoms / 4) ** (1 / 3)))
    basis = torch.tensor([[0.0, 0.0, 0.0], [0.5, 0.5, 0.0], [0.5, 0.0, 0.5], [0.0, 0.5, 0.5]], dtype=torch.float64)
    r = torch.arange(n, dtype=torch.float64)
    origin = torch.stack(torch.meshgrid(r, r, r, indexing="ij"), dim=-1).reshape(-1, 1, 3)     # i slowest, k fastest
    positions = ((origin + basis) * lattice_constant).reshape(-1, 3)[:num_atoms].to(dtype)
    cell = (torch.eye(3, dtype=dtype) * (n * lattice_constant)).reshape(1, 3, 3)
    pbc = torch.tensor([True, True, True]).reshape(1, 3)
    return positions, cell, pbc
