"""GPU parity of the pair consumer: the stencil sweep fused with the reference's real-space
Coulomb / Ewald kernel — the neighbor list is never written — and the same consumer over an existing list.

Checker: oracle/coulomb_oracle.py (numpy fp64 restatement of coulomb.py:206-292 + math.py:52-93) evaluated over the
ORACLE's neighbor list (oracle/reference_oracle.py), i.e. the reference pipeline neighbor_list -> coulomb_energy_forces.
Tolerance: 1e-11 relative to the largest |value| (fp64 everywhere; only the summation order differs)."""
import numpy as np
import pytest
import torch

import coulomb_oracle as co
import reference_oracle as ro
from systems import bench_batch, random_system, triclinic_system

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 1e-11


def _es():
    from nvalchemiops_b200.interactions import electrostatics
    return electrostatics


def _charges(n, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(n, generator=g, dtype=torch.float64) - 0.5) * 2.0


def _oracle(pos, q, cell, pbc, cutoff, alpha, batch_idx=None, half_fill=False):
    if batch_idx is None:
        o = ro.cell_list(pos, cutoff, cell, pbc, max_neighbors=2048, half_fill=half_fill, nthreads=8)
    else:
        o = ro.batch_cell_list(pos, cutoff, cell, pbc, batch_idx, max_neighbors=2048, half_fill=half_fill, nthreads=8)
    rec = ro.records_from_matrix(*o)
    return co.coulomb_energy_forces_list(pos.numpy(), q.numpy(), cell.numpy(), cutoff, alpha, rec[:, 0], rec[:, 1], rec[:, 2:],
                                         None if batch_idx is None else batch_idx.numpy()), rec


def _close(got, want, what):
    got = got.cpu().numpy()
    scale = max(1e-300, float(np.abs(want).max()))
    err = float(np.abs(got - want).max()) / scale
    assert err < RTOL, (what, err)


@pytest.mark.parametrize("alpha", [0.0, 0.3])
def test_fused_sweep_matches_list_pipeline_periodic_box(alpha):
    """Interior and boundary cells of a periodic box, plain and erfc-damped: fused == oracle pipeline, fused path taken."""
    pos, cell, pbc = random_system(6000, 40.0, torch.float32, seed=31)
    q = _charges(6000, 1)
    (e_want, f_want), rec = _oracle(pos, q, cell, pbc, 6.0, alpha)
    e, f, path = _es().fused_coulomb_energy_forces(pos.to(DEV), q.to(DEV), cell.to(DEV), pbc.to(DEV), 6.0, alpha, return_path=True)
    assert path == "fused"
    assert e.dtype == torch.float64 and f.dtype == torch.float64 and f.shape == (6000, 3)
    _close(e, e_want, "energies")
    _close(f, f_want, "forces")
    assert abs(float(f.sum(0).abs().max())) < 1e-9 * float(np.abs(f_want).max()) * 6000   # Newton's third law
    # the consumer over an existing list (this package's own list), COO and matrix formats
    from nvalchemiops_b200.neighborlist import neighbor_list
    d = [t.to(DEV) for t in (pos, cell, pbc)]
    nl, ptr, sh = neighbor_list(d[0], 6.0, cell=d[1], pbc=d[2], return_neighbor_list=True)
    e2, f2 = _es().coulomb_energy_forces(d[0], q.to(DEV), d[1], 6.0, alpha, neighbor_list=nl, neighbor_ptr=ptr, neighbor_shifts=sh)
    _close(e2, e_want, "energies (COO consumer)")
    _close(f2, f_want, "forces (COO consumer)")
    nm, num, nms = neighbor_list(d[0], 6.0, cell=d[1], pbc=d[2], max_neighbors=160)
    e3, f3 = _es().coulomb_energy_forces(d[0], q.to(DEV), d[1], 6.0, alpha, neighbor_matrix=nm, neighbor_matrix_shifts=nms,
                                         fill_value=6000)
    _close(e3, e_want, "energies (matrix consumer)")
    _close(f3, f_want, "forces (matrix consumer)")
    assert torch.equal(_es().coulomb_energy(d[0], q.to(DEV), d[1], 6.0, alpha, neighbor_list=nl, neighbor_ptr=ptr,
                                            neighbor_shifts=sh), e2)


def test_fused_sweep_batch_mixed_pbc_and_triclinic():
    """Batched systems with all PBC patterns (cells at open faces, 2-cell periodic boxes with image shifts) and a
    triclinic cell."""
    pos, cell, pbc, bidx, bptr = bench_batch(16, 300, 700, seed=8, mixed_pbc=True)
    q = _charges(pos.shape[0], 2)
    (e_want, f_want), _ = _oracle(pos, q, cell, pbc, 6.0, 0.25, batch_idx=bidx)
    for kw in ({"batch_idx": bidx.to(DEV)}, {"batch_ptr": bptr.to(DEV)}):
        e, f, path = _es().fused_coulomb_energy_forces(pos.to(DEV), q.to(DEV), cell.to(DEV), pbc.to(DEV), 6.0, 0.25,
                                                      return_path=True, **kw)
        assert path == "fused"
        _close(e, e_want, "batch energies")
        _close(f, f_want, "batch forces")
    pos, cell, pbc = triclinic_system(3000, 33.0, 37.0, 30.0, 70.0, 80.0, 100.0, torch.float32, seed=4)
    q = _charges(3000, 3)
    (e_want, f_want), _ = _oracle(pos, q, cell, pbc, 5.0, 0.0)
    e, f = _es().fused_coulomb_energy_forces(pos.to(DEV), q.to(DEV), cell.to(DEV), pbc.to(DEV), 5.0)
    _close(e, e_want, "triclinic energies")
    _close(f, f_want, "triclinic forces")


def test_fused_sweep_falls_back_to_the_list_path_when_it_must():
    """Atoms outside the primary image, boxes smaller than the cutoff (stencil wider than one cell) and float64 positions
    are served by list + consumer inside the same call — same numbers."""
    base, cell, pbc = random_system(3000, 32.0, torch.float32, seed=12)
    q = _charges(3000, 5)
    unw = base + 32.0 * (torch.arange(3000) % 3 - 1).float()[:, None]
    (e_want, f_want), _ = _oracle(unw, q, cell, pbc, 6.0, 0.2)
    e, f, path = _es().fused_coulomb_energy_forces(unw.to(DEV), q.to(DEV), cell.to(DEV), pbc.to(DEV), 6.0, 0.2, return_path=True)
    assert path == "list"
    _close(e, e_want, "unwrapped energies")
    _close(f, f_want, "unwrapped forces")
    pos, cell, pbc = random_system(40, 5.0, torch.float32, seed=3)           # 5 A box, 6 A cutoff: several images per pair
    q = _charges(40, 6)
    (e_want, f_want), _ = _oracle(pos, q, cell, pbc, 6.0, 0.0)
    e, f, path = _es().fused_coulomb_energy_forces(pos.to(DEV), q.to(DEV), cell.to(DEV), pbc.to(DEV), 6.0, return_path=True)
    assert path == "list"
    _close(e, e_want, "small-box energies")
    _close(f, f_want, "small-box forces")
    pos, cell, pbc = random_system(2000, 30.0, torch.float64, seed=14)
    q = _charges(2000, 7)
    o = ro.cell_list(pos, 6.0, cell, pbc, max_neighbors=512, nthreads=8)
    rec = ro.records_from_matrix(*o)
    e_want, f_want = co.coulomb_energy_forces_list(pos.numpy(), q.numpy(), cell.numpy(), 6.0, 0.3, rec[:, 0], rec[:, 1], rec[:, 2:])
    e, f, path = _es().fused_coulomb_energy_forces(pos.to(DEV), q.to(DEV), cell.to(DEV), pbc.to(DEV), 6.0, 0.3, return_path=True)
    assert path == "list"
    _close(e, e_want, "f64 energies")
    _close(f, f_want, "f64 forces")


def test_list_consumer_half_list_two_charges_and_errors():
    """Half lists (every pair stored once: the reaction on j comes from the atomics), the closed-form two-charge case of the
    reference's docstring formulas, and the reference's argument checks (coulomb.py:1607-1621)."""
    from nvalchemiops_b200.neighborlist import neighbor_list
    pos, cell, pbc = random_system(1500, 25.0, torch.float32, seed=41)
    q = _charges(1500, 9)
    (e_full, f_full), _ = _oracle(pos, q, cell, pbc, 6.0, 0.3)
    d = [t.to(DEV) for t in (pos, cell, pbc)]
    nl, ptr, sh = neighbor_list(d[0], 6.0, cell=d[1], pbc=d[2], half_fill=True, return_neighbor_list=True)
    e, f = _es().coulomb_energy_forces(d[0], q.to(DEV), d[1], 6.0, 0.3, neighbor_list=nl, neighbor_ptr=ptr, neighbor_shifts=sh)
    want = co.coulomb_energy_forces_list(pos.numpy(), q.numpy(), cell.numpy(), 6.0, 0.3, nl[0].cpu().numpy(), nl[1].cpu().numpy(),
                                         sh.cpu().numpy())
    _close(e, want[0], "half-list energies")
    _close(f, want[1], "half-list forces")
    # the reference's 1/2 prefactor assumes a FULL list (every pair stored twice): over a half list forces and the total
    # energy come out halved — reproduced, not "fixed"
    _close(2.0 * f, f_full, "2 x half-list forces == full-list forces")
    assert abs(2.0 * float(e.sum()) - float(e_full.sum())) < 1e-10 * abs(float(e_full.sum()))
    # two charges 2 A apart, no PBC: E_i = q1 q2 / (2 r), F on atom 0 = q1 q2 / r^2 along r_0 - r_1
    p2 = torch.tensor([[0.0, 0.0, 0.0], [2.0, 0.0, 0.0]], dtype=torch.float32, device=DEV)
    q2 = torch.tensor([1.5, -2.0], dtype=torch.float64, device=DEV)
    c2 = torch.eye(3, device=DEV).reshape(1, 3, 3) * 20.0
    e, f = _es().fused_coulomb_energy_forces(p2, q2, c2, torch.zeros(3, dtype=torch.bool, device=DEV), 5.0)
    assert np.allclose(e.cpu().numpy(), [1.5 * -2.0 / 4.0] * 2, rtol=1e-14)
    assert np.allclose(f.cpu().numpy(), [[1.5 * -2.0 / 4.0 * -1.0, 0, 0], [1.5 * -2.0 / 4.0, 0, 0]], rtol=1e-14, atol=1e-300)
    with pytest.raises(ValueError):
        _es().coulomb_energy_forces(d[0], q.to(DEV), d[1], 6.0)
    with pytest.raises(ValueError):
        _es().coulomb_energy_forces(d[0], q.to(DEV), d[1], 6.0, neighbor_list=nl, neighbor_shifts=sh)
    with pytest.raises(ValueError):
        _es().coulomb_energy_forces(d[0], q.to(DEV), d[1], 6.0, neighbor_list=nl, neighbor_ptr=ptr, neighbor_shifts=sh,
                                    neighbor_matrix=nl, neighbor_matrix_shifts=sh)
    with pytest.raises(RuntimeError):
        _es().fused_coulomb_energy_forces(pos, q, cell, pbc, 6.0)     # CPU tensors: no fallback
