"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same inputs.

Bar: neighbor sets bit-exact after sorting (i, j, sx, sy, sz) records — integer outputs, so no tolerance.
Modelled on the reference's own tests (test/neighborlist/test_cell_list.py, test_batch_cell_list.py,
test_neighborlist.py); the oracle plays the role the optional `vesin` comparison plays there.
"""
import numpy as np
import pytest
import torch

import reference_oracle as ro
from systems import (bench_batch, bench_box, kat_structure, load_kat, random_system, triclinic_system)

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _nl():
    from nvalchemiops_b200 import neighborlist

    return neighborlist


def _records_gpu_matrix(nm, num, sh):
    return ro.records_from_matrix(nm.cpu(), num.cpu(), sh.cpu())


def _check_matrix_padding(nm, num, sh, fill_value):
    nm, num, sh = nm.cpu(), num.cpu(), sh.cpu()
    M = nm.shape[1]
    cols = torch.arange(M)[None, :]
    pad = cols >= num.clamp(max=M)[:, None]
    assert (nm[pad] == fill_value).all(), "padding must be fill_value"
    assert (sh[pad] == 0).all(), "padding shifts must be zero"


# ----------------------------------------------------------------------------------------------
# known-answer tests of the reference (test_cell_list.py:391-419, test_batch_cell_list.py:516-542)
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["HoTlPd", "SiCu"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_known_answer_counts(name, dtype):
    pos, cell, pbc, expected = kat_structure(name, dtype, DEV)
    for k, rc in enumerate(load_kat()["cutoffs"]):
        nm, num, sh = _nl().cell_list(positions=pos, cutoff=rc, pbc=pbc, cell=cell)
        assert num.cpu().tolist() == expected[k]
        assert nm.dtype == torch.int32 and num.dtype == torch.int32 and sh.dtype == torch.int32
        assert nm.device == pos.device
        o = ro.cell_list(pos, rc, cell, pbc)
        assert np.array_equal(_records_gpu_matrix(nm, num, sh), ro.records_from_matrix(*o))
        nl, ptr, s = _nl().cell_list(pos, rc, cell, pbc, return_neighbor_list=True)
        assert np.array_equal(ro.records_from_coo(nl.cpu(), s.cpu()), ro.records_from_matrix(*o))
        assert ptr.cpu().tolist() == [0] + np.cumsum(expected[k]).tolist()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_known_answer_counts_batched(dtype):
    p1, c1, b1, e1 = kat_structure("HoTlPd", dtype, DEV)
    p2, c2, b2, e2 = kat_structure("SiCu", dtype, DEV)
    pos = torch.cat([p1, p2])
    cell = torch.stack([c1, c2])
    pbc = torch.stack([b1, b2])
    bidx = torch.tensor([0] * len(p1) + [1] * len(p2), dtype=torch.int32, device=DEV)
    for k, rc in enumerate(load_kat()["cutoffs"]):
        nm, num, sh = _nl().batch_cell_list(pos, rc, cell, pbc, bidx)
        assert num.cpu().tolist() == e1[k] + e2[k]
        o = ro.batch_cell_list(pos, rc, cell, pbc, bidx)
        assert np.array_equal(_records_gpu_matrix(nm, num, sh), ro.records_from_matrix(*o))


# ----------------------------------------------------------------------------------------------
# full (i, j, shift) set parity on random systems — the sweep of test_cell_list.py:316-389
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("pbc_flag", [[True, True, True], [True, False, True], [False, False, True],
                                      [False, False, False]])
@pytest.mark.parametrize("num_atoms", [10, 20, 50, 100])
@pytest.mark.parametrize("cutoff", [1.0, 3.0, 5.0])
@pytest.mark.parametrize("kind", ["cubic", "triclinic"])
def test_set_parity_small_systems(pbc_flag, num_atoms, cutoff, kind):
    for dtype in (torch.float32, torch.float64):
        if kind == "cubic":
            pos, cell, pbc = random_system(num_atoms, 3.0, dtype, seed=42, pbc_flag=pbc_flag)
        else:
            scale = (1 / 720.88) ** (1 / 3) * 3.0
            pos, cell, pbc = triclinic_system(num_atoms, 8.57 * scale, 12.9645 * scale, 7.2203 * scale, 90.74, 115.944,
                                              87.663, dtype, seed=42, pbc_flag=pbc_flag)
        o = ro.cell_list(pos, cutoff, cell, pbc, max_neighbors=8192)
        assert o[1].max() <= 8192
        want = ro.records_from_matrix(*o)
        nl, ptr, s = _nl().cell_list(pos.to(DEV), cutoff, cell.to(DEV), pbc.to(DEV), max_neighbors=8192,
                                     return_neighbor_list=True)
        got = ro.records_from_coo(nl.cpu(), s.cpu())
        assert np.array_equal(got, want), f"{kind} {dtype} n={num_atoms} rc={cutoff} pbc={pbc_flag}"
        src = nl[0].cpu().numpy()
        assert (np.diff(src) >= 0).all(), "COO source atoms must be sorted"
        assert np.array_equal(np.bincount(src, minlength=num_atoms), np.diff(ptr.cpu().numpy()))


@pytest.mark.parametrize("seed", range(12))
def test_set_parity_random_geometry(seed):
    """Random box sizes / cutoffs / mixed PBC / unwrapped atoms, both fma modes."""
    from nvalchemiops_b200 import config

    rng = np.random.default_rng(seed)
    n = int(rng.integers(30, 400))
    L = float(rng.uniform(3.0, 14.0))
    rc = float(rng.uniform(1.0, 5.5))
    pbc_flag = [bool(x) for x in rng.integers(0, 2, 3)]
    dtype = torch.float32 if seed % 2 == 0 else torch.float64
    if seed % 3 == 0:
        pos, cell, pbc = triclinic_system(n, L, 1.3 * L, 0.8 * L, 70.0, 105.0, 60.0, dtype, seed=seed, pbc_flag=pbc_flag,
                                          spread=(-1.4, 2.3))
    else:
        pos, cell, pbc = random_system(n, L, dtype, seed=seed, pbc_flag=pbc_flag)
        if seed % 3 == 1:
            pos = pos * 3.0 - L  # atoms up to one box outside on either side
    try:
        for fma in (True, False):
            config.fma = fma
            o = ro.cell_list(pos, rc, cell, pbc, max_neighbors=16384, fma_mode=int(fma))
            assert o[1].max() <= 16384
            want = ro.records_from_matrix(*o)
            nl, ptr, s = _nl().cell_list(pos.to(DEV), rc, cell.to(DEV), pbc.to(DEV), max_neighbors=16384,
                                         return_neighbor_list=True)
            assert np.array_equal(ro.records_from_coo(nl.cpu(), s.cpu()), want), f"seed {seed} fma {fma}"
    finally:
        config.fma = True


def test_matrix_output_padding_overflow_and_buffers():
    pos, cell, pbc = random_system(300, 9.0, torch.float32, seed=7)
    pos, cell, pbc = pos.to(DEV), cell.to(DEV), pbc.to(DEV)
    nl = _nl()
    nm, num, sh = nl.cell_list(pos, 3.0, cell, pbc, max_neighbors=96)
    assert nm.shape == (300, 96) and sh.shape == (300, 96, 3) and num.shape == (300,)
    want = ro.records_from_matrix(*ro.cell_list(pos, 3.0, cell, pbc, max_neighbors=96))
    assert np.array_equal(_records_gpu_matrix(nm, num, sh), want)
    _check_matrix_padding(nm, num, sh, 300)
    # explicit fill value
    nm2, num2, sh2 = nl.cell_list(pos, 3.0, cell, pbc, max_neighbors=96, fill_value=-1)
    _check_matrix_padding(nm2, num2, sh2, -1)
    # pre-allocated outputs are reused in place (test_neighborlist.py:817-855)
    bm = torch.full((300, 96), 12345, dtype=torch.int32, device=DEV)
    bs = torch.full((300, 96, 3), 7, dtype=torch.int32, device=DEV)
    bn = torch.full((300,), 9, dtype=torch.int32, device=DEV)
    r = nl.cell_list(pos, 3.0, cell, pbc, neighbor_matrix=bm, neighbor_matrix_shifts=bs, num_neighbors=bn)
    assert r[0] is bm and r[1] is bn and r[2] is bs
    assert torch.equal(bm, nm) or np.array_equal(_records_gpu_matrix(bm, bn, bs), want)
    _check_matrix_padding(bm, bn, bs, 300)
    # overflow: counts keep counting, stored entries are a subset of the true row (neighbor_utils.py:139-147)
    nm3, num3, sh3 = nl.cell_list(pos, 3.0, cell, pbc, max_neighbors=8)
    assert torch.equal(num3, num)
    assert num3.max().item() > 8
    got = ro.records_from_matrix(nm3.cpu(), num3.cpu(), sh3.cpu())
    want_set = {tuple(r) for r in want.tolist()}
    assert all(tuple(r) in want_set for r in got.tolist())
    assert got.shape[0] == int(num3.clamp(max=8).sum())
    with pytest.raises(nl.NeighborOverflowError):
        nl.cell_list(pos, 3.0, cell, pbc, max_neighbors=8, return_neighbor_list=True)


def test_half_fill_canonical_set():
    pos, cell, pbc = random_system(200, 5.0, torch.float32, seed=11)  # box < 2 rc: multi-image, self images
    full = ro.records_from_matrix(*ro.cell_list(pos, 3.5, cell, pbc, max_neighbors=4096))
    nl, ptr, s = _nl().cell_list(pos.to(DEV), 3.5, cell.to(DEV), pbc.to(DEV), half_fill=True, max_neighbors=4096,
                                 return_neighbor_list=True)
    half = ro.records_from_coo(nl.cpu(), s.cpu())
    assert 2 * half.shape[0] == full.shape[0]
    assert np.array_equal(ro.canonical_undirected(half), np.unique(ro.canonical_undirected(full), axis=0))
    nm, num, sh = _nl().cell_list(pos.to(DEV), 3.5, cell.to(DEV), pbc.to(DEV), half_fill=True, max_neighbors=2048)
    assert np.array_equal(ro.canonical_undirected(_records_gpu_matrix(nm, num, sh)), ro.canonical_undirected(half))


def test_no_pbc_shifts_zero_and_mixed_pbc_components():
    pos, cell, pbc = random_system(150, 8.0, torch.float32, seed=3, pbc_flag=[False, False, False])
    nm, num, sh = _nl().cell_list(pos.to(DEV), 4.0, cell.to(DEV), pbc.to(DEV))
    assert (sh == 0).all()
    pos, cell, pbc = random_system(150, 4.0, torch.float32, seed=3, pbc_flag=[True, True, False])
    nm, num, sh = _nl().cell_list(pos.to(DEV), 4.5, cell.to(DEV), pbc.to(DEV))
    assert (sh[..., 2] == 0).all() and (sh[..., 0] != 0).any()


def test_empty_and_zero_cutoff_shapes():
    nl = _nl()
    cell = torch.eye(3, device=DEV).reshape(1, 3, 3)
    pbc = torch.tensor([True, True, True], device=DEV)
    pos = torch.rand(10, 3, device=DEV)
    nm, num, sh = nl.cell_list(pos, 0.0, cell, pbc)
    assert nm.shape == (10, 0) and num.shape == (10,) and sh.shape == (10, 0, 3)
    e, p, s = nl.cell_list(pos, 0.0, cell, pbc, return_neighbor_list=True)
    assert e.shape == (2, 0) and p.shape == (11,) and s.shape == (0, 3)
    empty = torch.zeros((0, 3), device=DEV)
    nm, num, sh = nl.cell_list(empty, 1.0, cell, pbc)
    assert nm.shape == (0, 0) and num.shape == (0,) and sh.shape == (0, 0, 3)
    e, p, s = nl.cell_list(empty, 1.0, cell, pbc, return_neighbor_list=True)
    assert e.shape == (2, 0) and p.shape == (1,) and s.shape == (0, 3)
    one = torch.zeros((1, 3), device=DEV)
    nm, num, sh = nl.cell_list(one, 0.5, cell, pbc)
    assert num.tolist() == [0]
    nm, num, sh = nl.cell_list(one, 1.5, cell, pbc)  # self images: 6 at distance 1 + 12 at sqrt(2)
    assert num.tolist() == [18]
    nm, num, sh = nl.cell_list(one, 1.2, cell, pbc)
    assert num.tolist() == [6]


# ----------------------------------------------------------------------------------------------
# batches (BASELINE config 3 shape) and the dispatcher
# ----------------------------------------------------------------------------------------------
def test_batch_mixed_pbc_parity_and_no_cross_system_pairs():
    pos, cell, pbc, bidx, bptr = bench_batch(24, 150, 250, seed=3, mixed_pbc=True)
    want = ro.records_from_matrix(*ro.batch_cell_list(pos, 6.0, cell, pbc, bidx, max_neighbors=1024))
    nl = _nl()
    e, p, s = nl.batch_cell_list(pos.to(DEV), 6.0, cell.to(DEV), pbc.to(DEV), bidx.to(DEV), return_neighbor_list=True)  # default max_neighbors 1584
    got = ro.records_from_coo(e.cpu(), s.cpu())
    assert np.array_equal(got, want)
    b = bidx.numpy()
    assert (b[got[:, 0]] == b[got[:, 1]]).all()
    # through the dispatcher with batch_ptr only, and with shuffled (non-contiguous) batch_idx
    e2, p2, s2 = nl.neighbor_list(pos.to(DEV), 6.0, cell=cell.to(DEV), pbc=pbc.to(DEV), batch_ptr=bptr.to(DEV),
                                  return_neighbor_list=True, method=None if pos.shape[0] >= 5000 else "batch_cell_list",
                                  batch_idx=bidx.to(DEV))
    assert np.array_equal(ro.records_from_coo(e2.cpu(), s2.cpu()), want)
    perm = torch.randperm(pos.shape[0], generator=torch.Generator().manual_seed(0))
    e3, p3, s3 = nl.batch_cell_list(pos[perm].to(DEV), 6.0, cell.to(DEV), pbc.to(DEV), bidx[perm].to(DEV),
                                    return_neighbor_list=True)
    inv = perm.numpy()
    got3 = ro.records_from_coo(e3.cpu(), s3.cpu())
    got3[:, 0] = inv[got3[:, 0]]
    got3[:, 1] = inv[got3[:, 1]]
    assert np.array_equal(ro.sort_records(got3), want)


def test_dispatcher_naive_route_and_auto_selection():
    """BASELINE config 1 (256 atoms, no cell/pbc -> 2-tuple) and auto cell_list without a cell (>= 5000 atoms,
    test_neighborlist.py:91-114)."""
    nl = _nl()
    pos = random_system(256, 13.68, torch.float32, seed=1)[0]
    out = nl.neighbor_list(pos.to(DEV), 6.0)
    assert len(out) == 2
    nm, num = out
    o = ro.neighbor_list(pos, 6.0)
    assert np.array_equal(ro.records_from_matrix(nm.cpu(), num.cpu()), ro.records_from_matrix(*o))
    e, p = nl.neighbor_list(pos.to(DEV), 6.0, return_neighbor_list=True)
    assert np.array_equal(ro.records_from_coo(e.cpu()), ro.records_from_matrix(*o))
    big = random_system(6000, 40.0, torch.float32, seed=2)[0]
    out = nl.neighbor_list(big.to(DEV), 3.0, max_neighbors=64)
    assert len(out) == 3
    o = ro.neighbor_list(big, 3.0, max_neighbors=64)
    assert np.array_equal(_records_gpu_matrix(*out[:1], out[1], out[2]), ro.records_from_matrix(*o))
    with pytest.raises(TypeError):
        nl.neighbor_list(big.to(DEV), 3.0, not_a_kwarg=1)
    with pytest.raises(ValueError):
        nl.neighbor_list(big.to(DEV), 3.0, method="nope")


# ----------------------------------------------------------------------------------------------
# BASELINE configs at (near) full size
# ----------------------------------------------------------------------------------------------
def test_config2_50k_matrix_parity():
    pos, cell, pbc = bench_box(50_000, seed=2)
    want = ro.records_from_matrix(*ro.cell_list(pos, 6.0, cell, pbc, max_neighbors=160, nthreads=8))
    nm, num, sh = _nl().cell_list(pos.to(DEV), 6.0, cell.to(DEV), pbc.to(DEV), max_neighbors=160)
    assert np.array_equal(_records_gpu_matrix(nm, num, sh), want)
    _check_matrix_padding(nm, num, sh, 50_000)
    # default max_neighbors = 1584: same set, 1.27 GB of mostly padding written once
    nm, num, sh = _nl().cell_list(pos.to(DEV), 6.0, cell.to(DEV), pbc.to(DEV))
    assert nm.shape[1] == 1584
    assert np.array_equal(_records_gpu_matrix(nm, num, sh), want)


def _keys(i, j, s, n):
    """Total-order key of a directed pair record for |s_d| <= 3."""
    code = ((s[:, 0] + 3) * 7 + (s[:, 1] + 3)) * 7 + (s[:, 2] + 3)
    return (i.to(torch.int64) * n + j.to(torch.int64)) * 343 + code.to(torch.int64)


def test_config4_1m_atoms_properties_and_full_parity():
    """1 M atoms, rc = 6, periodic: size-independent properties on the full output, then full-set parity with
    the oracle run on the reference algorithm with max_nbins raised (same neighbor set, tractable on CPU)."""
    n = 1_000_000
    pos, cell, pbc = bench_box(n, seed=4)
    e, ptr, s = _nl().neighbor_list(pos.to(DEV), 6.0, cell=cell.to(DEV), pbc=pbc.to(DEV), return_neighbor_list=True)
    P = e.shape[1]
    assert ptr[-1].item() == P and ptr[0].item() == 0
    assert (ptr[1:] >= ptr[:-1]).all()
    assert (e[0, 1:] >= e[0, :-1]).all(), "source atoms sorted"
    counts = torch.bincount(e[0].long(), minlength=n)
    assert torch.equal(counts.to(torch.int32), ptr[1:] - ptr[:-1])
    assert s.abs().max().item() <= 1
    # symmetry: {(i,j,s)} == {(j,i,-s)}
    k_fwd = torch.sort(_keys(e[0], e[1], s, n)).values
    k_rev = torch.sort(_keys(e[1], e[0], -s, n)).values
    assert torch.equal(k_fwd, k_rev)
    assert (k_fwd[1:] != k_fwd[:-1]).all(), "no duplicate records"
    # geometry: every stored pair is inside the cutoff (fp64 re-evaluation, 1e-5 slack like the reference tests)
    p64 = pos.to(DEV).double()
    L = cell[0, 0, 0].double().item()
    d = p64[e[1].long()] - p64[e[0].long()] + s.double() * L
    assert (d.pow(2).sum(1).sqrt() < 6.0 + 1e-5).all()
    del d, p64, k_rev
    # expected density: 90.5 neighbors per atom
    assert abs(P / n - 90.48) < 0.2
    # full-set parity
    nm, num, sh = ro.cell_list(pos, 6.0, cell, pbc, max_neighbors=160, nthreads=8, max_nbins=1 << 20)
    assert num.max() <= 160
    assert np.array_equal(num, (ptr[1:] - ptr[:-1]).cpu().numpy())
    mask = np.arange(160)[None, :] < num[:, None]
    oi = torch.from_numpy(np.nonzero(mask)[0]).to(DEV)
    oj = torch.from_numpy(nm[mask]).to(DEV)
    os_ = torch.from_numpy(sh[mask]).to(DEV)
    k_or = torch.sort(_keys(oi, oj, os_, n)).values
    assert torch.equal(k_fwd, k_or)


def test_large_cells_multi_tile_path():
    """Few huge cells (cutoff ~ box/2): candidates exceed one shared-memory tile, rows exceed the staging row."""
    pos, cell, pbc = random_system(3000, 12.0, torch.float32, seed=9)
    for fl, mode in (([True, True, True], "pbc"), ([False, False, False], "open")):
        pbc = torch.tensor(fl).reshape(1, 3)
        o = ro.cell_list(pos, 5.5, cell, pbc, max_neighbors=4096, nthreads=8)
        assert o[1].max() <= 4096
        want = ro.records_from_matrix(*o)
        e, p, s = _nl().cell_list(pos.to(DEV), 5.5, cell.to(DEV), pbc.to(DEV), max_neighbors=4096,
                                  return_neighbor_list=True)
        assert np.array_equal(ro.records_from_coo(e.cpu(), s.cpu()), want), mode
        nm, num, sh = _nl().cell_list(pos.to(DEV), 5.5, cell.to(DEV), pbc.to(DEV), max_neighbors=2048)
        assert np.array_equal(_records_gpu_matrix(nm, num, sh), want), mode
        _check_matrix_padding(nm, num, sh, 3000)


def test_sharded_ranges_fill_pack_and_expand_on_one_gpu():
    """The multi-GPU data path without the collective: two 'ranks' are emulated on one device — each writes its own range
    of the FINAL global arrays in place (index_offset + row_stride) and packs its shifts into one byte per pair; what a
    peer would receive (target atoms, packed shifts, counts) is then expanded by nvnl_expand_gathered."""
    import ctypes

    from nvalchemiops_b200 import _lib
    from nvalchemiops_b200.neighborlist import _engine
    from nvalchemiops_b200.neighborlist.distributed import partition_systems

    pos, cell, pbc, bidx, bptr = bench_batch(10, 300, 700, seed=8, mixed_pbc=True)
    N = pos.shape[0]
    want = ro.records_from_matrix(*ro.batch_cell_list(pos, 6.0, cell, pbc, bidx, max_neighbors=1024))
    pos, cell, pbc, bptr_d = pos.to(DEV), cell.to(DEV), pbc.to(DEV), bptr.to(DEV)
    world = 2
    parts = partition_systems(bptr.tolist(), world)
    locals_ = []
    for (s0, s1) in parts:
        a0, a1 = int(bptr[s0]), int(bptr[s1])
        lptr = (bptr_d[s0:s1 + 1] - a0).to(torch.int32)
        lidx = torch.repeat_interleave(torch.arange(s1 - s0, dtype=torch.int32, device=DEV), (lptr[1:] - lptr[:-1]).long())
        h = _engine.build(pos[a0:a1], 6.0, cell[s0:s1], pbc[s0:s1], batch_idx=lidx, batch_ptr=lptr)
        num, ptr, total, max_count, err, hint, rows = _engine.count_and_size(h, 36.0)
        assert err == 0 and not h.wide_stencil and not (hint & 1)
        locals_.append((h, num, ptr, total, a0, a1, hint, rows))
    counts = [t[3] for t in locals_]
    P = sum(counts)
    edge = torch.full((2, P), -7, dtype=torch.int32, device=DEV)
    shifts = torch.full((P, 3), -7, dtype=torch.int32, device=DEV)
    packed = torch.zeros((P,), dtype=torch.uint8, device=DEV)
    num_all = torch.empty((N,), dtype=torch.int32, device=DEV)
    bad = torch.zeros((1,), dtype=torch.int32, device=DEV)
    off = 0
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for (h, num, ptr, total, a0, a1, hint, rows) in locals_:
        _engine.fill_coo(h, 36.0, ptr, edge[0, off:], shifts[off:off + total], total, False, a0, launch_hint=hint, rows=rows,
                         row_stride=P)
        num_all[a0:a1] = num
        _lib.check(_lib.lib().nvnl_pack_shifts(ctypes.c_void_p(shifts[off:].data_ptr()), total,
                                               ctypes.c_void_p(packed[off:].data_ptr()), ctypes.c_void_p(bad.data_ptr()), st), "pack")
        off += total
    torch.cuda.synchronize()
    assert int(bad.item()) == 0
    assert np.array_equal(ro.records_from_coo(edge.cpu(), shifts.cpu()), want)
    assert (edge[0, 1:] >= edge[0, :-1]).all()
    nptr = torch.zeros(N + 1, dtype=torch.int32, device=DEV)
    torch.cumsum(num_all, 0, out=nptr[1:])
    # what "rank 0" holds after the gather: its own range complete, of the peer only row 1, the packed shifts, the counts
    a0, a1 = locals_[0][4], locals_[0][5]
    e2, s2 = edge.clone(), shifts.clone()
    e2[0, counts[0]:] = -9
    s2[counts[0]:] = -9
    _lib.check(_lib.lib().nvnl_expand_gathered(ctypes.c_void_p(nptr.data_ptr()), N, a0, a1, ctypes.c_void_p(packed.data_ptr()),
                                               ctypes.c_void_p(e2.data_ptr()), ctypes.c_void_p(s2.data_ptr()), st), "expand")
    torch.cuda.synchronize()
    assert torch.equal(e2, edge) and torch.equal(s2, shifts)
    # the padded exchange (one in-place ncclAllGather per array): every rank's targets / packed shifts sit in its slot of
    # staging buffers of world x Pmax entries; nvnl_expand_padded writes row 1 everywhere and row 0 / shifts of the peers
    pmax = max(counts)
    offs = [0, counts[0], P]
    t_dst = torch.full((world * pmax,), -5, dtype=torch.int32, device=DEV)
    t_pk = torch.zeros((world * pmax,), dtype=torch.uint8, device=DEV)
    for g in range(world):
        t_dst[g * pmax: g * pmax + counts[g]] = edge[1, offs[g]:offs[g + 1]]
        t_pk[g * pmax: g * pmax + counts[g]] = packed[offs[g]:offs[g + 1]]
    ab = (ctypes.c_int64 * 3)(0, locals_[0][5], N)
    pb = (ctypes.c_int64 * 3)(*offs)
    for rank in range(world):
        e3, s3 = edge.clone(), shifts.clone()
        e3[1] = -9                                   # row 1 comes from the staging buffer on every rank
        lo, hi = offs[rank], offs[rank + 1]
        keep = torch.zeros(P, dtype=torch.bool, device=DEV); keep[lo:hi] = True
        e3[0, ~keep] = -9
        s3[~keep] = -9
        _lib.check(_lib.lib().nvnl_expand_padded(ctypes.c_void_p(nptr.data_ptr()), N, world, rank, ab, pb, pmax,
                                                 ctypes.c_void_p(t_dst.data_ptr()), ctypes.c_void_p(t_pk.data_ptr()),
                                                 ctypes.c_void_p(e3[0].data_ptr()), ctypes.c_void_p(e3[1].data_ptr()),
                                                 ctypes.c_void_p(s3.data_ptr()), st), "expand_padded")
        torch.cuda.synchronize()
        assert torch.equal(e3, edge) and torch.equal(s3, shifts), rank
    # the chunked exchange: every rank's atoms split into chunks (here at arbitrary atoms, one chunk empty), one staging
    # buffer set and one nvnl_expand_padded_ranges launch per chunk; atoms of other chunks must be left alone
    mid = locals_[0][5]
    chunk_atoms = [[(0, 37), (mid, mid + 1001)], [(37, mid - 5), (mid + 1001, mid + 1001)], [(mid - 5, mid), (mid + 1001, N)]]
    nptr_h = nptr.cpu().tolist()
    arr = ctypes.c_int64 * world
    for rank in range(world):
        e3, s3 = edge.clone(), shifts.clone()
        e3[1] = -9
        lo, hi = offs[rank], offs[rank + 1]
        keep = torch.zeros(P, dtype=torch.bool, device=DEV); keep[lo:hi] = True
        e3[0, ~keep] = -9
        s3[~keep] = -9
        for ranges in chunk_atoms:
            cnts = [nptr_h[b] - nptr_h[a] for a, b in ranges]
            pm = max(cnts)
            if pm == 0:
                continue
            c_dst = torch.full((world * pm,), -5, dtype=torch.int32, device=DEV)
            c_pk = torch.zeros((world * pm,), dtype=torch.uint8, device=DEV)
            for g, (a, b) in enumerate(ranges):
                c_dst[g * pm: g * pm + cnts[g]] = edge[1, nptr_h[a]:nptr_h[b]]
                c_pk[g * pm: g * pm + cnts[g]] = packed[nptr_h[a]:nptr_h[b]]
            _lib.check(_lib.lib().nvnl_expand_padded_ranges(
                ctypes.c_void_p(nptr.data_ptr()), N, world, rank, arr(*[a for a, _ in ranges]), arr(*[b for _, b in ranges]),
                arr(*[nptr_h[a] for a, _ in ranges]), pm, ctypes.c_void_p(c_dst.data_ptr()), ctypes.c_void_p(c_pk.data_ptr()),
                ctypes.c_void_p(e3[0].data_ptr()), ctypes.c_void_p(e3[1].data_ptr()), ctypes.c_void_p(s3.data_ptr()), st),
                "expand_padded_ranges")
            torch.cuda.synchronize()
        assert torch.equal(e3, edge) and torch.equal(s3, shifts), rank
    # one-word exchange: nvnl_pack_shifts_word ORs the packed shift into bits 26..31 of the targets in place; the
    # expansion takes the words as gathered_dst with gathered_packed = NULL
    words = edge[1].clone()
    bad.zero_()
    _lib.check(_lib.lib().nvnl_pack_shifts_word(ctypes.c_void_p(shifts.data_ptr()), P, ctypes.c_void_p(words.data_ptr()),
                                                ctypes.c_void_p(bad.data_ptr()), st), "pack_word")
    torch.cuda.synchronize()
    assert int(bad.item()) == 0
    assert torch.equal(words & ((1 << 26) - 1), edge[1])
    assert torch.equal(((words.long() & 0xFFFFFFFF) >> 26).to(torch.uint8), packed)
    for rank in range(world):
        e3, s3 = edge.clone(), shifts.clone()
        e3[1] = -9
        keep = torch.zeros(P, dtype=torch.bool, device=DEV); keep[offs[rank]:offs[rank + 1]] = True
        e3[0, ~keep] = -9
        s3[~keep] = -9
        for ranges in chunk_atoms:
            cnts = [nptr_h[b] - nptr_h[a] for a, b in ranges]
            pm = max(cnts)
            if pm == 0:
                continue
            c_dst = torch.full((world * pm,), -5, dtype=torch.int32, device=DEV)
            for g, (a, b) in enumerate(ranges):
                c_dst[g * pm: g * pm + cnts[g]] = words[nptr_h[a]:nptr_h[b]]
            _lib.check(_lib.lib().nvnl_expand_padded_ranges(
                ctypes.c_void_p(nptr.data_ptr()), N, world, rank, arr(*[a for a, _ in ranges]), arr(*[b for _, b in ranges]),
                arr(*[nptr_h[a] for a, _ in ranges]), pm, ctypes.c_void_p(c_dst.data_ptr()), None,
                ctypes.c_void_p(e3[0].data_ptr()), ctypes.c_void_p(e3[1].data_ptr()), ctypes.c_void_p(s3.data_ptr()), st),
                "expand_padded_ranges (word)")
            torch.cuda.synchronize()
        assert torch.equal(e3, edge) and torch.equal(s3, shifts), rank
    big = torch.tensor([5, 1 << 26], dtype=torch.int32, device=DEV)       # a target that does not fit 26 bits is reported
    bad.zero_()
    _lib.check(_lib.lib().nvnl_pack_shifts_word(ctypes.c_void_p(shifts.data_ptr()), 2, ctypes.c_void_p(big.data_ptr()),
                                                ctypes.c_void_p(bad.data_ptr()), st), "pack_word")
    assert int(bad.item()) == 1
    bad_ranges = arr(10, 5)
    assert _lib.lib().nvnl_expand_padded_ranges(ctypes.c_void_p(nptr.data_ptr()), N, world, 0, bad_ranges, arr(20, 30), arr(0, 0), 1,
                                                ctypes.c_void_p(t_dst.data_ptr()), ctypes.c_void_p(t_pk.data_ptr()),
                                                ctypes.c_void_p(e3[0].data_ptr()), ctypes.c_void_p(e3[1].data_ptr()),
                                                ctypes.c_void_p(s3.data_ptr()), st) != 0
    # a shift outside {-1, 0, 1} is reported by the pack kernel
    s3 = shifts[:100].clone(); s3[17, 1] = 2
    bad.zero_()
    _lib.check(_lib.lib().nvnl_pack_shifts(ctypes.c_void_p(s3.data_ptr()), 100, ctypes.c_void_p(packed.data_ptr()),
                                           ctypes.c_void_p(bad.data_ptr()), st), "pack")
    assert int(bad.item()) == 1


def test_chunked_exchange_pipeline_on_one_gpu(tmp_path):
    """The sharded path's exchange machinery on ONE GPU (a one-rank NCCL group): every chunk of the rank's systems is
    written into the final arrays with its targets in the chunk's staging slot, exchanged on the communication stream and
    re-assembled by nvnl_expand_padded_ranges — the result must equal the plain batched neighbor list for any number of
    chunks.  (Foreign ranges need peers: tests/test_distributed_gpu.py on >= 2 GPUs, and the emulation test above.)"""
    import torch.distributed as dist

    from nvalchemiops_b200.neighborlist.distributed import sharded_batch_neighbor_list

    if dist.is_initialized():
        pytest.skip("a process group is already initialised in this process")
    pos, cell, pbc, bidx, bptr = bench_batch(13, 200, 600, seed=21, mixed_pbc=True)
    pos, cell, pbc, bidx, bptr = pos.to(DEV), cell.to(DEV), pbc.to(DEV), bidx.to(DEV), bptr.to(DEV)
    e1, p1, s1 = _nl().neighbor_list(pos, 6.0, cell=cell, pbc=pbc, batch_idx=bidx, batch_ptr=bptr, return_neighbor_list=True,
                                     method="batch_cell_list")
    want = ro.records_from_coo(e1.cpu(), s1.cpu())
    dist.init_process_group("nccl", init_method=f"file://{tmp_path}/pg", rank=0, world_size=1, device_id=torch.device(DEV))
    try:
        from nvalchemiops_b200 import config
        for chunks, word in ((1, True), (2, True), (3, True), (16, True), (2, False), (3, False)):   # 16 > systems: empty chunks
            config.exchange_word = word       # one int32 per pair (target | shift << 26) or targets + a byte array
            e, p, s, stats = sharded_batch_neighbor_list(pos, 6.0, cell, pbc, bptr, return_stats=True, chunks=chunks,
                                                         _exchange_when_alone=True)
            config.exchange_word = True
            assert stats["packed"] and stats["chunks"] == chunks and stats["peer_bytes"] == 0
            assert stats["bytes_per_pair"] == (4 if word else 5)
            assert torch.equal(p, p1), chunks
            assert bool((e[0, 1:] >= e[0, :-1]).all()), chunks
            assert np.array_equal(ro.records_from_coo(e.cpu(), s.cpu()), want), chunks
        # the world-of-one short cut (what bench.py's N = 1 leg calls) gives the same list
        e, p, s = sharded_batch_neighbor_list(pos, 6.0, cell, pbc, bptr)
        assert torch.equal(p, p1) and np.array_equal(ro.records_from_coo(e.cpu(), s.cpu()), want)
        with pytest.raises(_nl().NeighborOverflowError):
            sharded_batch_neighbor_list(pos, 6.0, cell, pbc, bptr, max_neighbors=1, _exchange_when_alone=True)
    finally:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
# split build / query workflow and rebuild detection (reference cell_list.py:1037-1192,
# rebuild_detection.py, docs/userguide/components/neighborlist.md:421-500)
# ----------------------------------------------------------------------------------------------
def _alloc_cache(n, ncell_cap, ns=None):
    shape3 = (3,) if ns is None else (ns, 3)
    z = lambda *s: torch.zeros(s, dtype=torch.int32, device=DEV)  # noqa: E731
    return [z(*shape3), z(*shape3), z(n, 3), z(n, 3), z(ncell_cap), z(ncell_cap), z(n)]


def test_build_query_split_with_moving_atoms():
    """Build once with cutoff + skin, then re-query with displaced atoms and the bare cutoff: every query must equal
    a from-scratch oracle run on the displaced positions (the stale cell assignment stays valid below skin/2)."""
    nl = _nl()
    rc, skin = 4.0, 1.0
    for pbc_flag in ([True, True, True], [True, False, True]):
        pos, cell, pbc = random_system(800, 22.0, torch.float32, seed=21, pbc_flag=pbc_flag)
        cache = _alloc_cache(800, 4096)
        nl.build_cell_list(pos.to(DEV), rc + skin, cell.to(DEV), pbc.to(DEV), *cache)
        assert cache[0].tolist() == [4, 4, 4] or not all(pbc_flag)      # 22 / (5 * 1.001) -> 4 cells per periodic dim
        assert int(cache[4].sum()) == 800 and sorted(cache[6].cpu().tolist()) == list(range(800))
        g = torch.Generator().manual_seed(3)
        cur = pos.clone()
        for step in range(4):
            if step:
                cur = cur + (torch.rand(800, 3, generator=g) - 0.5) * 0.2   # |d| <= 0.17 per step, 0.52 total < skin/2... 
            nm = torch.full((800, 64), 800, dtype=torch.int32, device=DEV)
            sh = torch.zeros((800, 64, 3), dtype=torch.int32, device=DEV)
            num = torch.zeros((800,), dtype=torch.int32, device=DEV)
            nl.query_cell_list(cur.to(DEV), rc, cell.to(DEV), pbc.to(DEV), *cache, nm, sh, num)
            want = ro.records_from_matrix(*ro.cell_list(cur, rc, cell, pbc, max_neighbors=64))
            assert np.array_equal(_records_gpu_matrix(nm, num, sh), want), f"step {step} pbc {pbc_flag}"
            _check_matrix_padding(nm, num, sh, 800)   # untouched slots keep the caller's pre-fill
        # a query cutoff ABOVE the build cutoff is still exact for atoms that have not moved since the build (the stencil
        # radius is recomputed for the query cutoff on the cached grid; the reference has no check either)
        nm = torch.full((800, 160), 800, dtype=torch.int32, device=DEV)
        sh = torch.zeros((800, 160, 3), dtype=torch.int32, device=DEV)
        num = torch.zeros((800,), dtype=torch.int32, device=DEV)
        nl.query_cell_list(pos.to(DEV), rc + 2 * skin, cell.to(DEV), pbc.to(DEV), *cache, nm, sh, num)
        want = ro.records_from_matrix(*ro.cell_list(pos, rc + 2 * skin, cell, pbc, max_neighbors=160))
        assert np.array_equal(_records_gpu_matrix(nm, num, sh), want)
    # no hidden state: the cache survives clone() (new storage, new tensor objects)
    cloned = [t.clone() for t in cache]
    nm2 = torch.full((800, 64), 800, dtype=torch.int32, device=DEV)
    sh2 = torch.zeros((800, 64, 3), dtype=torch.int32, device=DEV)
    num2 = torch.zeros((800,), dtype=torch.int32, device=DEV)
    nl.query_cell_list(cur.to(DEV), rc, cell.to(DEV), pbc.to(DEV), *cloned, nm2, sh2, num2)
    want = ro.records_from_matrix(*ro.cell_list(cur, rc, cell, pbc, max_neighbors=64))
    assert np.array_equal(_records_gpu_matrix(nm2, num2, sh2), want)
    # a cache that no build filled is reported when input checking is on (the matrix path is sync-free otherwise)
    from nvalchemiops_b200 import config
    config.check_inputs = True
    try:
        with pytest.raises(ValueError, match="cache"):
            nl.query_cell_list(cur.to(DEV), rc, cell.to(DEV), pbc.to(DEV), *_alloc_cache(800, 4096), nm2, sh2, num2)
    finally:
        config.check_inputs = False


def test_estimate_allocate_build_query_end_to_end():
    """The reference's documented split workflow (docs/userguide/components/neighborlist.md:421-500):
    estimate_cell_list_sizes -> allocate_cell_list -> build_cell_list -> query_cell_list, single and batched, on CUDA
    tensors; the estimates equal the oracle's restatement of cell_list.py:35-99 / batch_cell_list.py:35-99."""
    nl = _nl()
    for L, rc, pbc_flag, nbins in ((22.0, 4.0, [True, True, True], 1000), (60.4, 3.0, [True, False, True], 1000),
                                   (60.4, 3.0, [True, True, True], 100000), (7.0, 4.0, [True, True, False], 1000)):
        pos, cell, pbc = random_system(700, L, torch.float32, seed=5, pbc_flag=pbc_flag)
        if L < 10:
            pos = pos[:100].clone().repeat(7, 1) + torch.arange(7).repeat_interleave(100)[:, None] * torch.tensor([0.0, 0.0, 40.0])   # sparse stack along the open z
        pos_d, cell_d, pbc_d = pos.to(DEV), cell.to(DEV), pbc.to(DEV)
        cells, radius = nl.estimate_cell_list_sizes(cell_d, pbc_d, rc, max_nbins=nbins)
        want_cells, want_radius = ro.estimate_cell_list_sizes(cell, pbc, rc, max_nbins=nbins)
        assert cells == want_cells and radius.device == pos_d.device and radius.dtype == torch.int32
        assert radius.cpu().tolist() == list(np.asarray(want_radius).reshape(-1))
        cache = nl.allocate_cell_list(700, cells, radius, pos_d.device)
        assert [tuple(t.shape) for t in cache] == [(3,), (3,), (700, 3), (700, 3), (cells,), (cells,), (700,)]
        assert all(t.dtype == torch.int32 and t.device == pos_d.device for t in cache)
        nl.build_cell_list(pos_d, rc, cell_d, pbc_d, *cache)
        assert int(cache[0].prod()) <= cells                       # the grid respects the allocated capacity
        assert int(cache[4].sum()) == 700 and sorted(cache[6].cpu().tolist()) == list(range(700))
        assert torch.equal(cache[5][: int(cache[0].prod())].cpu(),
                           (torch.cumsum(cache[4][: int(cache[0].prod())], 0) - cache[4][: int(cache[0].prod())]).cpu().to(torch.int32))
        M = 128
        nm = torch.full((700, M), 700, dtype=torch.int32, device=DEV)
        sh = torch.zeros((700, M, 3), dtype=torch.int32, device=DEV)
        num = torch.zeros((700,), dtype=torch.int32, device=DEV)
        nl.query_cell_list(pos_d, rc, cell_d, pbc_d, *cache, nm, sh, num)
        want = ro.records_from_matrix(*ro.cell_list(pos, rc, cell, pbc, max_neighbors=M))
        assert np.array_equal(_records_gpu_matrix(nm, num, sh), want), (L, rc, pbc_flag, nbins)
    # batched
    pos, cell, pbc, bidx, bptr = bench_batch(9, 100, 300, seed=13, mixed_pbc=True)
    n = pos.shape[0]
    pos_d, cell_d, pbc_d, bidx_d = pos.to(DEV), cell.to(DEV), pbc.to(DEV), bidx.to(DEV)
    cells, radius = nl.estimate_batch_cell_list_sizes(cell_d, pbc_d, 5.0)
    want_cells, want_radius = ro.estimate_batch_cell_list_sizes(cell, pbc, 5.0)
    assert cells == want_cells and radius.shape == (9, 3) and radius.device == pos_d.device
    assert np.array_equal(radius.cpu().numpy(), np.asarray(want_radius).reshape(9, 3))
    cache = nl.allocate_cell_list(n, cells, radius, pos_d.device)
    assert tuple(cache[0].shape) == (9, 3)
    nl.batch_build_cell_list(pos_d, 5.0, cell_d, pbc_d, bidx_d, *cache)
    assert int(cache[0].prod(dim=1).sum()) <= cells and int(cache[0].prod(dim=1).max()) <= cells // 9
    nm = torch.full((n, 200), -1, dtype=torch.int32, device=DEV)
    sh = torch.zeros((n, 200, 3), dtype=torch.int32, device=DEV)
    num = torch.zeros((n,), dtype=torch.int32, device=DEV)
    nl.batch_query_cell_list(pos_d, cell_d, pbc_d, 5.0, bidx_d, *cache, nm, sh, num)
    want = ro.records_from_matrix(*ro.batch_cell_list(pos, 5.0, cell, pbc, bidx, max_neighbors=200))
    assert np.array_equal(_records_gpu_matrix(nm, num, sh), want)
    # empty / zero-cutoff estimates (cell_list.py:683-684, batch_cell_list.py:697-698)
    assert nl.estimate_cell_list_sizes(cell_d[:1], pbc_d[:1], 0.0)[0] == 1
    c0, r0 = nl.estimate_batch_cell_list_sizes(cell_d[:0], pbc_d[:0], 5.0)
    assert c0 == 1 and tuple(r0.shape) == (0, 3)


def test_cell_list_fills_all_seven_preallocated_cache_tensors():
    """cell_list() / batch_cell_list() with a full set of pre-allocated tensors: outputs AND caches are the same objects,
    reset and filled in place (reference cell_list.py:1395-1410; test_neighborlist.py:817-855)."""
    nl = _nl()
    pos, cell, pbc = random_system(600, 20.0, torch.float32, seed=8)
    pos_d, cell_d, pbc_d = pos.to(DEV), cell.to(DEV), pbc.to(DEV)
    cells, radius = nl.estimate_cell_list_sizes(cell_d, pbc_d, 4.0)
    cache = nl.allocate_cell_list(600, cells, radius, pos_d.device)
    for t in cache:
        t.fill_(-7)                                               # stale garbage the call must overwrite
    names = ("cells_per_dimension", "neighbor_search_radius", "atom_periodic_shifts", "atom_to_cell_mapping",
             "atoms_per_cell_count", "cell_atom_start_indices", "cell_atom_list")
    nm = torch.full((600, 96), -3, dtype=torch.int32, device=DEV)
    sh = torch.full((600, 96, 3), -3, dtype=torch.int32, device=DEV)
    num = torch.full((600,), -3, dtype=torch.int32, device=DEV)
    out = nl.cell_list(pos_d, 4.0, cell_d, pbc_d, neighbor_matrix=nm, neighbor_matrix_shifts=sh, num_neighbors=num,
                       **dict(zip(names, cache)))
    assert out[0] is nm and out[1] is num and out[2] is sh
    want = ro.records_from_matrix(*ro.cell_list(pos, 4.0, cell, pbc, max_neighbors=96))
    assert np.array_equal(_records_gpu_matrix(nm, num, sh), want)
    _check_matrix_padding(nm, num, sh, 600)
    ncell = int(cache[0].prod())
    assert cache[0].tolist() == [4, 4, 4] and ncell <= cells and cache[1].tolist() == [1, 1, 1]
    assert int(cache[4].sum()) == 600 and int(cache[4][ncell:].abs().sum()) == 0
    assert sorted(cache[6].cpu().tolist()) == list(range(600))
    assert int(cache[2].abs().sum()) == 0                         # wrapped positions: no periodic images
    w = 20.0 / 4
    assert torch.equal(cache[3].cpu().long(), torch.floor(pos / w).long().clamp(0, 3))
    # the filled cache is a valid input of query_cell_list
    nm2 = torch.full((600, 96), 600, dtype=torch.int32, device=DEV)
    sh2 = torch.zeros((600, 96, 3), dtype=torch.int32, device=DEV)
    num2 = torch.zeros((600,), dtype=torch.int32, device=DEV)
    nl.query_cell_list(pos_d, 4.0, cell_d, pbc_d, *cache, nm2, sh2, num2)
    assert np.array_equal(_records_gpu_matrix(nm2, num2, sh2), want)
    # COO output with a full cache
    e, p, s = nl.cell_list(pos_d, 4.0, cell_d, pbc_d, return_neighbor_list=True, max_neighbors=96, **dict(zip(names, cache)))
    assert np.array_equal(ro.records_from_coo(e.cpu(), s.cpu()), want)


@pytest.mark.parametrize("batched", [False, True])
def test_torch_compile_build_and_query_ops(batched):
    """The four reference-schema custom ops under torch.compile(fullgraph=True) with pre-allocated tensors: compiled ==
    eager == oracle (reference test_cell_list.py:598-844, test_batch_cell_list.py compile tests)."""
    nl = _nl()
    rc, M = 4.0, 96
    if batched:
        pos, cell, pbc, bidx, bptr = bench_batch(5, 80, 160, seed=19, mixed_pbc=True)
        n = pos.shape[0]
        want = ro.records_from_matrix(*ro.batch_cell_list(pos, rc, cell, pbc, bidx, max_neighbors=M))
        bidx_d = bidx.to(DEV)
        cells, radius = nl.estimate_batch_cell_list_sizes(cell.to(DEV), pbc.to(DEV), rc)
    else:
        pos, cell, pbc = random_system(500, 14.0, torch.float32, seed=41)
        n = 500
        want = ro.records_from_matrix(*ro.cell_list(pos, rc, cell, pbc, max_neighbors=M))
        bidx_d = None
        cells, radius = nl.estimate_cell_list_sizes(cell.to(DEV), pbc.to(DEV), rc)
    pos_d, cell_d, pbc_d = pos.to(DEV), cell.to(DEV), pbc.to(DEV)

    def fn(positions, cache, nm, sh, num):
        if batched:
            nl.batch_build_cell_list(positions, rc, cell_d, pbc_d, bidx_d, *cache)
            nl.batch_query_cell_list(positions, cell_d, pbc_d, rc, bidx_d, *cache, nm, sh, num, False)
        else:
            nl.build_cell_list(positions, rc, cell_d, pbc_d, *cache)
            nl.query_cell_list(positions, rc, cell_d, pbc_d, *cache, nm, sh, num, False)
        return num.sum()

    def fresh():
        return (list(nl.allocate_cell_list(n, cells, radius.clone(), pos_d.device)),
                torch.full((n, M), n, dtype=torch.int32, device=DEV), torch.zeros((n, M, 3), dtype=torch.int32, device=DEV),
                torch.zeros((n,), dtype=torch.int32, device=DEV))

    cache, nm, sh, num = fresh()
    eager_total = fn(pos_d, cache, nm, sh, num).item()
    assert np.array_equal(_records_gpu_matrix(nm, num, sh), want)
    cache2, nm2, sh2, num2 = fresh()
    compiled = torch.compile(fn, fullgraph=True)
    total = compiled(pos_d, cache2, nm2, sh2, num2).item()
    assert total == eager_total == want.shape[0]
    assert np.array_equal(_records_gpu_matrix(nm2, num2, sh2), want)
    for a, b in zip(cache[:6], cache2[:6]):
        assert torch.equal(a, b)                                  # compiled build == eager build, tensor by tensor
    # (the order of atoms inside a cell is not deterministic: compare cell_atom_list cell by cell)
    starts, counts = cache[5].cpu().tolist(), cache[4].cpu().tolist()
    la, lb = cache[6].cpu().tolist(), cache2[6].cpu().tolist()
    for st, cn in zip(starts, counts):
        assert sorted(la[st:st + cn]) == sorted(lb[st:st + cn])


def test_batch_build_query_split():
    nl = _nl()
    pos, cell, pbc, bidx, bptr = bench_batch(6, 200, 400, seed=12, mixed_pbc=True)
    n = pos.shape[0]
    cache = _alloc_cache(n, 4096, ns=6)
    nl.batch_build_cell_list(pos.to(DEV), 6.0, cell.to(DEV), pbc.to(DEV), bidx.to(DEV), *cache)
    nm = torch.full((n, 256), -1, dtype=torch.int32, device=DEV)
    sh = torch.zeros((n, 256, 3), dtype=torch.int32, device=DEV)
    num = torch.zeros((n,), dtype=torch.int32, device=DEV)
    nl.batch_query_cell_list(pos.to(DEV), cell.to(DEV), pbc.to(DEV), 5.0, bidx.to(DEV), *cache, nm, sh, num)
    want = ro.records_from_matrix(*ro.batch_cell_list(pos, 5.0, cell, pbc, bidx, max_neighbors=256))
    assert np.array_equal(_records_gpu_matrix(nm, num, sh), want)


def test_rebuild_detection():
    nl = _nl()
    pos, cell, pbc = random_system(2000, 30.0, torch.float32, seed=31)
    pos_d, cell_d, pbc_d = pos.to(DEV), cell.to(DEV), pbc.to(DEV)
    # neighbor_list_needs_rebuild: displacement vs skin (rebuild_detection.py:168-217)
    assert nl.neighbor_list_needs_rebuild(pos_d, pos_d.clone(), 0.5).tolist() == [False]
    moved = pos_d.clone(); moved[1234, 1] += 0.6
    r = nl.neighbor_list_needs_rebuild(pos_d, moved, 0.5)
    assert r.dtype == torch.bool and r.shape == (1,) and r.tolist() == [True]
    assert nl.neighbor_list_needs_rebuild(pos_d, moved, 0.7).tolist() == [False]
    assert nl.check_neighbor_list_rebuild_needed(pos_d, moved, 0.5) is True
    assert nl.neighbor_list_needs_rebuild(pos_d[:10], moved, 0.5).tolist() == [True]          # shape mismatch
    assert nl.neighbor_list_needs_rebuild(pos_d[:0], pos_d[:0], 0.5).tolist() == [False]     # empty
    # oracle restatement of the same predicate on random displacements
    g = torch.Generator().manual_seed(1)
    for thr in (0.05, 0.2, 1.0):
        cur = pos + (torch.rand(2000, 3, generator=g) - 0.5) * 0.4
        want = bool((np.sqrt(((cur - pos).numpy().astype(np.float64) ** 2).sum(1)) > thr).any())
        assert nl.check_neighbor_list_rebuild_needed(pos_d, cur.to(DEV), thr) is want
    # cell_list_needs_rebuild: cell crossing under the grid of the last build (rebuild_detection.py:36-121)
    cache = _alloc_cache(2000, 4096)
    nl.build_cell_list(pos_d, 5.0, cell_d, pbc_d, *cache)
    cpd = cache[0].cpu().numpy()
    assert cpd.tolist() == [5, 5, 5]
    assert nl.cell_list_needs_rebuild(pos_d, cache[3], cache[0], cell_d, pbc_d).tolist() == [False]
    w = 30.0 / 5
    inside = pos.clone(); inside[7] = torch.tensor([2.5 * w, 2.5 * w, 2.5 * w])       # cell centre
    cache2 = _alloc_cache(2000, 4096)
    nl.build_cell_list(inside.to(DEV), 5.0, cell_d, pbc_d, *cache2)
    small = inside.clone(); small[7, 0] += 0.3 * w                                      # stays in its cell
    assert nl.check_cell_list_rebuild_needed(*cache2, small.to(DEV), cell_d, pbc_d, 5.0) is False
    big = inside.clone(); big[7, 0] += 0.6 * w                                          # crosses into the next cell
    assert nl.check_cell_list_rebuild_needed(*cache2, big.to(DEV), cell_d, pbc_d, 5.0) is True
    wrap = inside.clone(); wrap[7, 0] += 30.0                                           # a full period: same cell
    assert nl.check_cell_list_rebuild_needed(*cache2, wrap.to(DEV), cell_d, pbc_d, 5.0) is False
    # the exported mapping agrees with the geometry
    cells = torch.floor(pos / w).long().clamp(0, 4)
    assert torch.equal(cache[3].cpu().long(), cells)
    # stateless: cloned tensors (no hidden handle), a non-periodic dimension, and a batch
    assert nl.check_cell_list_rebuild_needed(*[t.clone() for t in cache2], big.to(DEV), cell_d, pbc_d, 5.0) is True
    pbc_open = torch.tensor([True, False, True])
    cache3 = _alloc_cache(2000, 4096)
    nl.build_cell_list(inside.to(DEV), 5.0, cell_d, pbc_open.to(DEV), *cache3)
    up = inside.clone(); up[7, 1] += 0.6 * w
    assert nl.check_cell_list_rebuild_needed(*cache3, inside.to(DEV), cell_d, pbc_open.to(DEV), 5.0) is False
    assert nl.check_cell_list_rebuild_needed(*cache3, up.to(DEV), cell_d, pbc_open.to(DEV), 5.0) is True
    bpos, bcell, bpbc, bidx, bptr = bench_batch(4, 150, 250, seed=3, mixed_pbc=True)
    bcache = _alloc_cache(bpos.shape[0], 4096, ns=4)
    nl.batch_build_cell_list(bpos.to(DEV), 3.0, bcell.to(DEV), bpbc.to(DEV), bidx.to(DEV), *bcache)
    assert nl.cell_list_needs_rebuild(bpos.to(DEV), bcache[3], bcache[0], bcell.to(DEV), bpbc.to(DEV), bidx.to(DEV)).tolist() == [False]


def test_dual_cutoff_routes():
    """cutoff2 -> naive_dual_cutoff: one build, two queries; reference return arity (neighborlist.py:150-176)."""
    nl = _nl()
    pos, cell, pbc = random_system(400, 12.0, torch.float32, seed=17)
    out = nl.neighbor_list(pos.to(DEV), 2.5, cell=cell.to(DEV), pbc=pbc.to(DEV), cutoff2=5.0, max_neighbors1=64,
                           max_neighbors2=256)
    assert len(out) == 6
    for k, rc, M in ((0, 2.5, 64), (3, 5.0, 256)):
        want = ro.records_from_matrix(*ro.cell_list(pos, rc, cell, pbc, max_neighbors=M))
        assert out[k].shape == (400, M)
        assert np.array_equal(_records_gpu_matrix(out[k], out[k + 1], out[k + 2]), want)
    out = nl.neighbor_list(pos.to(DEV), 2.5, cutoff2=4.0, max_neighbors1=64, max_neighbors2=128)   # no PBC: 4-tuple
    assert len(out) == 4
    free = torch.tensor([False] * 3)
    for k, rc, M in ((0, 2.5, 64), (2, 4.0, 128)):
        want = ro.records_from_matrix(*ro.cell_list(pos, rc, torch.eye(3).reshape(1, 3, 3), free, max_neighbors=M))
        assert np.array_equal(ro.records_from_matrix(out[k].cpu(), out[k + 1].cpu()), want)
    out = nl.neighbor_list(pos.to(DEV), 2.5, cell=cell.to(DEV), pbc=pbc.to(DEV), cutoff2=5.0, return_neighbor_list=True,
                           max_neighbors1=64, max_neighbors2=256)
    assert len(out) == 6 and out[0].shape[0] == 2 and out[3].shape[0] == 2
    assert np.array_equal(ro.records_from_coo(out[3].cpu(), out[5].cpu()),
                          ro.records_from_matrix(*ro.cell_list(pos, 5.0, cell, pbc, max_neighbors=256)))


def test_public_naive_entry_points():
    """The reference's naive functions as public names (naive.py:400, batch_naive.py:480, naive_dual_cutoff.py:544,
    batch_naive_dual_cutoff.py:592): reference signatures, return arity, in-place buffers, cutoff <= 0, default sizes —
    all answered by the cell-list engine with the same neighbor sets."""
    nl = _nl()
    pos, cell, pbc = random_system(300, 11.0, torch.float32, seed=31)
    pd, cd, bd = pos.to(DEV), cell.to(DEV), pbc.to(DEV)
    want = ro.records_from_matrix(*ro.cell_list(pos, 2.5, cell, pbc, max_neighbors=64))
    # keyword order of the reference's docstring example: pbc before cell
    nm, num, sh = nl.naive_neighbor_list(pd, 2.5, pbc=bd, cell=cd, max_neighbors=64)
    assert nm.shape == (300, 64) and np.array_equal(_records_gpu_matrix(nm, num, sh), want)
    _check_matrix_padding(nm, num, sh, 300)
    # pre-allocated buffers are filled in place and returned as the same objects; pre-computed shift ranges are accepted
    rng, off, tot = nl.compute_naive_num_shifts(cd, 2.5, bd)
    assert rng.device.type == "cuda" and rng.cpu().tolist() == [[1, 1, 1]] and tot == 14
    b_nm = torch.full((300, 48), -7, dtype=torch.int32, device=DEV)
    b_sh = torch.full((300, 48, 3), 9, dtype=torch.int32, device=DEV)
    b_num = torch.full((300,), 5, dtype=torch.int32, device=DEV)
    out = nl.naive_neighbor_list(pd, 2.5, cd, bd, fill_value=-1, neighbor_matrix=b_nm, neighbor_matrix_shifts=b_sh,
                                 num_neighbors=b_num, shift_range_per_dimension=rng, shift_offset=off, total_shifts=tot)
    assert out[0] is b_nm and out[1] is b_num and out[2] is b_sh
    assert np.array_equal(ro.records_from_matrix(b_nm.cpu(), b_num.cpu(), b_sh.cpu()), want)
    _check_matrix_padding(b_nm, b_num, b_sh, -1)
    e, ptr, s = nl.naive_neighbor_list(pd, 2.5, cd, bd, max_neighbors=64, return_neighbor_list=True)
    assert np.array_equal(ro.records_from_coo(e.cpu(), s.cpu()), want) and ptr[-1].item() == want.shape[0]
    with pytest.raises(nl.NeighborOverflowError):
        nl.naive_neighbor_list(pd, 2.5, cd, bd, max_neighbors=2, return_neighbor_list=True)
    # cutoff <= 0, matrix mode: the (N, max_neighbors) buffers the reference allocates, all padding
    nm0, num0, sh0 = nl.naive_neighbor_list(pd, 0.0, cd, bd, max_neighbors=8)
    assert nm0.shape == (300, 8) and (nm0 == 300).all() and num0.sum().item() == 0 and sh0.shape == (300, 8, 3)
    nm0, num0 = nl.naive_neighbor_list(pd, -1.0, max_neighbors=8)
    assert nm0.shape == (300, 8) and num0.shape == (300,)
    # batch: one of batch_idx / batch_ptr is enough; pairs never cross systems; 2-tuple without PBC
    bpos, bcell, bpbc, bidx, bptr = bench_batch(5, 60, 90, seed=9, mixed_pbc=True)
    wantb = ro.records_from_matrix(*ro.batch_cell_list(bpos, 5.0, bcell, bpbc, bidx, max_neighbors=256))
    for kw in ({"batch_idx": bidx.to(DEV)}, {"batch_ptr": bptr.to(DEV)}, {"batch_idx": bidx.to(DEV), "batch_ptr": bptr.to(DEV)}):
        nmb, numb, shb = nl.batch_naive_neighbor_list(bpos.to(DEV), 5.0, pbc=bpbc.to(DEV), cell=bcell.to(DEV),
                                                      max_neighbors=256, max_atoms_per_system=90, **kw)
        assert np.array_equal(_records_gpu_matrix(nmb, numb, shb), wantb)
    open_cell = torch.eye(3).reshape(1, 3, 3).repeat(5, 1, 1)
    want_open = ro.records_from_matrix(*ro.batch_cell_list(bpos, 5.0, open_cell, torch.zeros(5, 3, dtype=torch.bool), bidx,
                                                            max_neighbors=256))
    outb = nl.batch_naive_neighbor_list(bpos.to(DEV), 5.0, bidx.to(DEV), max_neighbors=256, return_neighbor_list=True)
    assert len(outb) == 2 and np.array_equal(ro.records_from_coo(outb[0].cpu()), want_open)
    assert nl.neighbor_list(bpos.to(DEV), 5.0, batch_ptr=bptr.to(DEV), method="batch_naive", max_neighbors=256)[0].shape == (bpos.shape[0], 256)
    # dual cutoff: both matrices are sized from cutoff2 when no size is given (naive_dual_cutoff.py:761-772)
    out = nl.naive_neighbor_list_dual_cutoff(pd, 2.5, 4.0, pbc=bd, cell=cd)
    M2 = nl.estimate_max_neighbors(4.0)
    assert len(out) == 6 and out[0].shape == (300, M2) and out[3].shape == (300, M2)
    assert np.array_equal(_records_gpu_matrix(out[0], out[1], out[2]), want)
    assert np.array_equal(_records_gpu_matrix(out[3], out[4], out[5]),
                          ro.records_from_matrix(*ro.cell_list(pos, 4.0, cell, pbc, max_neighbors=M2)))
    out = nl.naive_neighbor_list_dual_cutoff(pd, 0.0, 2.5, pbc=bd, cell=cd, max_neighbors1=16, max_neighbors2=64)
    assert out[0].shape == (300, 16) and out[1].sum().item() == 0 and (out[0] == 300).all()
    assert np.array_equal(_records_gpu_matrix(out[3], out[4], out[5]), want)
    outb = nl.batch_naive_neighbor_list_dual_cutoff(bpos.to(DEV), 2.5, 5.0, batch_ptr=bptr.to(DEV), pbc=bpbc.to(DEV),
                                                    cell=bcell.to(DEV), max_neighbors1=128, max_neighbors2=256,
                                                    return_neighbor_list=True)
    assert len(outb) == 6 and np.array_equal(ro.records_from_coo(outb[3].cpu(), outb[5].cpu()), wantb)
    assert np.array_equal(ro.records_from_coo(outb[0].cpu(), outb[2].cpu()),
                          ro.records_from_matrix(*ro.batch_cell_list(bpos, 2.5, bcell, bpbc, bidx, max_neighbors=128)))


def test_dispatcher_scenarios_of_the_reference_suite():
    """The dispatcher-level scenarios of test/neighborlist/test_neighborlist.py, restated: auto-selection with and without
    a batch (:43-292), kwargs forwarding and pre-allocated tensors (:735-899), empty / single-atom inputs through the naive
    route (:905-943), wrapper == direct call (:328-357, 418-463)."""
    nl = _nl()
    pos, cell, pbc = random_system(50, 10.0, torch.float32, seed=42)
    pd, cd, bd = pos.to(DEV), cell.to(DEV), pbc.to(DEV)
    for kw, width in (({"method": "naive", "max_neighbors": 20}, 20), ({"method": "cell_list", "max_neighbors": 30}, 30),
                      ({"max_neighbors": 25}, 25)):
        out = nl.neighbor_list(pd, 5.0, cell=cd, pbc=bd, **kw)
        assert len(out) == 3 and out[0].shape == (50, width) and out[0].dtype == torch.int32 and out[0].device == pd.device
    nm1, _, _, nm2, _, _ = nl.neighbor_list(pd, 2.5, cell=cd, pbc=bd, cutoff2=3.5, method="naive_dual_cutoff",
                                            max_neighbors1=15, max_neighbors2=25)
    assert nm1.shape[1] == 15 and nm2.shape[1] == 25
    bufs = (torch.full((50, 20), 50, dtype=torch.int32, device=DEV), torch.zeros(50, dtype=torch.int32, device=DEV),
            torch.zeros((50, 20, 3), dtype=torch.int32, device=DEV))
    out = nl.neighbor_list(pd, 5.0, cell=cd, pbc=bd, method="naive", neighbor_matrix=bufs[0], num_neighbors=bufs[1],
                           neighbor_matrix_shifts=bufs[2])
    assert out[0] is bufs[0] and out[1] is bufs[1] and out[2] is bufs[2]
    with pytest.raises(TypeError):
        nl.neighbor_list(pd, 2.0, method="naive", invalid_parameter_name=123)
    # wrapper == direct call, both output formats
    a = nl.neighbor_list(pd, 3.0, cell=cd, pbc=bd, method="cell_list", max_neighbors=64)
    b = nl.cell_list(pd, 3.0, cd, bd, max_neighbors=64)
    assert np.array_equal(_records_gpu_matrix(*a), _records_gpu_matrix(*b))
    a = nl.neighbor_list(pd, 3.0, cell=cd, pbc=bd, method="naive", max_neighbors=64, return_neighbor_list=True)
    assert np.array_equal(ro.records_from_coo(a[0].cpu(), a[2].cpu()), _records_gpu_matrix(*b))
    # empty and single-atom systems through the naive route: 2-tuples, no pairs
    e, ptr = nl.neighbor_list(torch.empty(0, 3, device=DEV), 2.0, method="naive", return_neighbor_list=True)
    assert e.shape == (2, 0) and ptr.tolist() == [0]
    e, ptr = nl.neighbor_list(torch.randn(1, 3, device=DEV), 2.0, method="naive", return_neighbor_list=True)
    assert e.shape == (2, 0) and ptr.tolist() == [0, 0]
    # auto-selection with a batch: batch_naive below 5000 atoms (2-tuple without PBC), batch_naive_dual_cutoff with cutoff2
    bpos, bcell, bpbc, bidx, bptr = bench_batch(3, 40, 60, seed=11, mixed_pbc=False)
    out = nl.neighbor_list(bpos.to(DEV), 3.0, batch_idx=bidx.to(DEV), max_neighbors=64)
    assert len(out) == 2 and out[0].shape == (bpos.shape[0], 64)
    i = torch.arange(bpos.shape[0]).repeat_interleave(out[1].cpu().long())
    j = out[0].cpu()[out[0].cpu() < bpos.shape[0]].long()
    assert (bidx[i] == bidx[j]).all(), "pairs must not cross systems"
    out = nl.neighbor_list(bpos.to(DEV), 2.0, cell=bcell.to(DEV), pbc=bpbc.to(DEV), batch_ptr=bptr.to(DEV), cutoff2=3.0,
                           max_neighbors1=64, max_neighbors2=64)
    assert len(out) == 6
    want = ro.records_from_matrix(*ro.batch_cell_list(bpos, 3.0, bcell, bpbc, bidx, max_neighbors=64))
    assert np.array_equal(_records_gpu_matrix(out[3], out[4], out[5]), want)
    # >= 5000 atoms with a batch: batch_cell_list is selected (3-tuple even without a cell)
    big, _, _, gidx, gptr = bench_batch(4, 1300, 1300, seed=12, mixed_pbc=False)
    out = nl.neighbor_list(big.to(DEV), 3.0, batch_ptr=gptr.to(DEV), max_neighbors=64)
    assert len(out) == 3 and out[0].shape == (5200, 64)


@pytest.mark.parametrize("pbc_flag", [[True, True, True], [True, True, False]])
def test_unwrapped_coordinates_medium_box(pbc_flag):
    """MD-style unwrapped coordinates (atoms up to two lattice vectors outside the box) on a box large enough for the
    unwrapped variant of the fast kernel: full set parity, COO and matrix, plus a half-fill run."""
    pos, cell, pbc = random_system(6000, 40.0, torch.float32, seed=23, pbc_flag=pbc_flag)
    g = torch.Generator().manual_seed(5)
    img = torch.randint(-2, 3, (6000, 3), generator=g).float()
    img[:, 2] = img[:, 2] if pbc_flag[2] else 0.0
    pos = pos + img * 40.0
    o = ro.cell_list(pos, 6.0, cell, pbc, max_neighbors=256, nthreads=8)
    assert o[1].max() <= 256
    want = ro.records_from_matrix(*o)
    e, p, s = _nl().cell_list(pos.to(DEV), 6.0, cell.to(DEV), pbc.to(DEV), max_neighbors=256, return_neighbor_list=True)
    assert np.array_equal(ro.records_from_coo(e.cpu(), s.cpu()), want)
    assert s.abs().max().item() >= 2
    nm, num, sh = _nl().cell_list(pos.to(DEV), 6.0, cell.to(DEV), pbc.to(DEV), max_neighbors=256)
    assert np.array_equal(_records_gpu_matrix(nm, num, sh), want)
    _check_matrix_padding(nm, num, sh, 6000)
    eh, ph, sh2 = _nl().cell_list(pos.to(DEV), 6.0, cell.to(DEV), pbc.to(DEV), max_neighbors=256, half_fill=True,
                                  return_neighbor_list=True)
    assert 2 * eh.shape[1] == want.shape[0]
    assert np.array_equal(ro.canonical_undirected(ro.records_from_coo(eh.cpu(), sh2.cpu())),
                          np.unique(ro.canonical_undirected(want), axis=0))


def test_batch_with_empty_systems_unaligned_positions_and_f64():
    """Edge cases of the batch path: systems without atoms, positions that are a non-16-byte-aligned view (scalar load
    path of the hash / scatter kernels), float64 inputs, half-fill in a batch."""
    nl = _nl()
    pos, cell, pbc, bidx, bptr = bench_batch(7, 120, 260, seed=14, mixed_pbc=True)
    # insert two empty systems (ids 2 and 5 get no atoms)
    remap = torch.tensor([0, 1, 3, 4, 6, 7, 8])
    bidx2 = remap[bidx.long()].to(torch.int32)
    cell2 = torch.eye(3).repeat(9, 1, 1) * 15.0
    pbc2 = torch.ones(9, 3, dtype=torch.bool)
    cell2[remap] = cell
    pbc2[remap] = pbc
    for dtype in (torch.float32, torch.float64):
        p = pos.to(dtype)
        c = cell2.to(dtype)
        want = ro.records_from_matrix(*ro.batch_cell_list(p, 6.0, c, pbc2, bidx2, max_neighbors=512))
        big = torch.zeros((p.shape[0] + 1, 3), dtype=dtype, device=DEV)
        big[1:] = p.to(DEV)
        view = big[1:]                                  # data_ptr offset by 12 / 24 bytes: not 16-byte aligned
        assert view.data_ptr() % 16 != 0
        e, ptr, s = nl.batch_cell_list(view, 6.0, c.to(DEV), pbc2.to(DEV), bidx2.to(DEV), max_neighbors=512,
                                       return_neighbor_list=True)
        assert np.array_equal(ro.records_from_coo(e.cpu(), s.cpu()), want), dtype
        eh, ph, sh = nl.batch_cell_list(view, 6.0, c.to(DEV), pbc2.to(DEV), bidx2.to(DEV), max_neighbors=512,
                                        half_fill=True, return_neighbor_list=True)
        assert np.array_equal(ro.canonical_undirected(ro.records_from_coo(eh.cpu(), sh.cpu())),
                              np.unique(ro.canonical_undirected(want), axis=0)), dtype


def test_invalid_inputs_are_reported():
    """Device-side validation surfaces as Python exceptions on the COO path (the one place with a host sync)."""
    nl = _nl()
    pos, cell, pbc = random_system(100, 10.0, torch.float32, seed=2)
    bad = pos.clone(); bad[3, 0] = 3.0e9                      # > 1e6 periodic images away
    with pytest.raises(ValueError, match="periodic images"):
        nl.cell_list(bad.to(DEV), 3.0, cell.to(DEV), pbc.to(DEV), return_neighbor_list=True)
    sing = cell.clone(); sing[0, 2] = sing[0, 1]                # singular cell
    with pytest.raises(ValueError, match="singular"):
        nl.cell_list(pos.to(DEV), 3.0, sing.to(DEV), pbc.to(DEV), return_neighbor_list=True)
    bidx = torch.zeros(100, dtype=torch.int32); bidx[7] = 5     # only 2 systems given
    cells = cell.repeat(2, 1, 1); pbcs = pbc.repeat(2, 1)
    with pytest.raises(ValueError, match="batch_idx"):
        nl.batch_cell_list(pos.to(DEV), 3.0, cells.to(DEV), pbcs.to(DEV), bidx.to(DEV), return_neighbor_list=True)
    # the workspace is reusable after an error: a valid call right after still matches the oracle
    e, p, s = nl.cell_list(pos.to(DEV), 3.0, cell.to(DEV), pbc.to(DEV), return_neighbor_list=True, max_neighbors=256)
    assert np.array_equal(ro.records_from_coo(e.cpu(), s.cpu()),
                          ro.records_from_matrix(*ro.cell_list(pos, 3.0, cell, pbc, max_neighbors=256)))


def test_torch_compile_keeps_the_matrix_op_in_the_graph():
    """Mutation-only custom op with pre-allocated outputs under torch.compile(fullgraph=True) — the property the
    reference tests for its build/query ops (test_cell_list.py:598-844)."""
    import nvalchemiops_b200.neighborlist.ops  # noqa: F401  (registers nvalchemiops_b200::neighbor_matrix)

    pos, cell, pbc = random_system(500, 14.0, torch.float32, seed=41)
    pos, cell, pbc = pos.to(DEV), cell.to(DEV), pbc.to(DEV)
    want = ro.records_from_matrix(*ro.cell_list(pos, 4.0, cell, pbc, max_neighbors=96))

    def fn(positions, nm, sh, num):
        torch.ops.nvalchemiops_b200.neighbor_matrix(positions * 1.0, 4.0, cell, pbc.reshape(1, 3), None, None, nm, sh, num, 500,
                                                    False, 16.0)
        return num.sum()

    nm = torch.empty((500, 96), dtype=torch.int32, device=DEV)
    sh = torch.empty((500, 96, 3), dtype=torch.int32, device=DEV)
    num = torch.empty((500,), dtype=torch.int32, device=DEV)
    eager_total = fn(pos, nm, sh, num).item()
    assert np.array_equal(_records_gpu_matrix(nm, num, sh), want)
    nm.fill_(-5); sh.fill_(-5); num.fill_(-5)
    compiled = torch.compile(fn, fullgraph=True)
    total = compiled(pos, nm, sh, num).item()
    assert total == eager_total == want.shape[0]
    assert np.array_equal(_records_gpu_matrix(nm, num, sh), want)
    _check_matrix_padding(nm, num, sh, 500)


# ----------------------------------------------------------------------------------------------
# the two COO paths (config.coo_path): "rows" = single sweep + streamed output (default for fp32),
# "masks" = count -> hit masks -> fill.  Everything above runs on the default; these pin both.
# ----------------------------------------------------------------------------------------------
@pytest.fixture(params=["rows", "masks"])
def coo_path(request):
    from nvalchemiops_b200 import config

    old = config.coo_path
    config.coo_path = request.param
    yield request.param
    config.coo_path = old


@pytest.mark.parametrize("pbc_flag", [[True, True, True], [True, False, True], [False, False, False]])
def test_coo_paths_medium_box(coo_path, pbc_flag):
    """6000 atoms in a 40 A box (7 cells per periodic dim: interior cells, boundary cells with image shifts, open
    faces): COO full and half fill, both fma modes, sorted sources, CSR consistency."""
    from nvalchemiops_b200 import config

    pos, cell, pbc = random_system(6000, 40.0, torch.float32, seed=31, pbc_flag=pbc_flag)
    try:
        for fma in (True, False):
            config.fma = fma
            o = ro.cell_list(pos, 6.0, cell, pbc, max_neighbors=256, nthreads=8, fma_mode=int(fma))
            assert o[1].max() <= 256
            want = ro.records_from_matrix(*o)
            e, p, s = _nl().cell_list(pos.to(DEV), 6.0, cell.to(DEV), pbc.to(DEV), max_neighbors=256,
                                      return_neighbor_list=True)
            assert np.array_equal(ro.records_from_coo(e.cpu(), s.cpu()), want), (coo_path, fma)
            assert (e[0, 1:] >= e[0, :-1]).all()
            assert torch.equal(torch.bincount(e[0].long(), minlength=6000).to(torch.int32), p[1:] - p[:-1])
            assert np.array_equal(o[1], (p[1:] - p[:-1]).cpu().numpy())
    finally:
        config.fma = True
    eh, ph, sh = _nl().cell_list(pos.to(DEV), 6.0, cell.to(DEV), pbc.to(DEV), max_neighbors=256, half_fill=True,
                                 return_neighbor_list=True)
    assert 2 * eh.shape[1] == want.shape[0]
    assert np.array_equal(ro.canonical_undirected(ro.records_from_coo(eh.cpu(), sh.cpu())),
                          np.unique(ro.canonical_undirected(want), axis=0))


def test_coo_paths_coincident_atoms_and_small_periodic_box(coo_path):
    """Distinct atoms at identical coordinates are neighbors (d = 0 < rc) while (i, i, 0) is not; a periodic box
    with two cells per dimension makes every stencil hold the same cells under several image shifts."""
    pos, cell, pbc = random_system(1500, 25.0, torch.float32, seed=5)
    pos[100] = pos[7]
    pos[101] = pos[7]
    pos[900] = pos[899]
    for rc in (6.0, 4.0):
        o = ro.cell_list(pos, rc, cell, pbc, max_neighbors=256, nthreads=8)
        want = ro.records_from_matrix(*o)
        e, p, s = _nl().cell_list(pos.to(DEV), rc, cell.to(DEV), pbc.to(DEV), max_neighbors=256, return_neighbor_list=True)
        got = ro.records_from_coo(e.cpu(), s.cpu())
        assert np.array_equal(got, want), (coo_path, rc)
        assert not ((got[:, 0] == got[:, 1]) & (got[:, 2:] == 0).all(1)).any()
        assert ((got[:, 0] == 100) & (got[:, 1] == 7) & (got[:, 2:] == 0).all(1)).any()
    pos, cell, pbc = random_system(400, 12.5, torch.float32, seed=6)     # cpd = 2: images of the same cells
    want = ro.records_from_matrix(*ro.cell_list(pos, 6.0, cell, pbc, max_neighbors=512, nthreads=8))
    e, p, s = _nl().cell_list(pos.to(DEV), 6.0, cell.to(DEV), pbc.to(DEV), max_neighbors=512, return_neighbor_list=True)
    assert np.array_equal(ro.records_from_coo(e.cpu(), s.cpu()), want), coo_path


def test_coo_paths_single_cell_systems_and_many_target_cells(coo_path):
    """Systems that are ONE cell (box edge < 2 rc): every image of the stencil is the same run of records — staged once,
    aliased by up to 27 shift segments — and cells with more than 64 targets, which several CTAs sweep in parts.
    All 8 PBC patterns, both fill modes, batched next to ordinary systems; plus one dense box whose 8 cells hold
    ~250 atoms each."""
    from nvalchemiops_b200 import config

    g = torch.Generator().manual_seed(77)
    counts = [40, 97, 150, 172, 60, 130, 166, 33, 171, 90, 165, 120, 20, 173, 101, 149, 400, 650]
    Ls = [9.5, 10.0, 11.45, 11.98, 7.0, 11.0, 11.9, 6.5, 11.95, 8.0, 11.8, 10.5, 6.2, 12.0, 9.9, 11.4, 16.0, 19.0]
    pos, cells, pbcs, bidx = [], [], [], []
    for s, (n, L) in enumerate(zip(counts, Ls)):
        pos.append(torch.rand(n, 3, generator=g) * L)
        cells.append(torch.eye(3) * L)
        pbcs.append(torch.tensor([(s & 1) > 0, (s & 2) > 0, (s & 4) > 0]) if s % 9 != 8 else torch.tensor([True, True, True]))
        bidx.append(torch.full((n,), s, dtype=torch.int32))
    # every fully periodic single-cell case once more, so that 27-image tiles of several sizes occur
    for n, L in ((170, 11.9), (64, 9.0), (33, 6.3)):
        s = len(pos)
        pos.append(torch.rand(n, 3, generator=g) * L); cells.append(torch.eye(3) * L)
        pbcs.append(torch.tensor([True, True, True])); bidx.append(torch.full((n,), s, dtype=torch.int32))
    pos, cell, pbc, bidx = torch.cat(pos), torch.stack(cells), torch.stack(pbcs), torch.cat(bidx)
    d = [t.to(DEV) for t in (pos, cell, pbc, bidx)]
    for half in (False, True):
        for fma in (True, False):
            config.fma = fma
            try:
                want = ro.records_from_matrix(*ro.batch_cell_list(pos, 6.0, cell, pbc, bidx, max_neighbors=1536, half_fill=half,
                                                                  fma_mode=int(fma), nthreads=8))
                e, p, sft = _nl().batch_cell_list(d[0], 6.0, d[1], d[2], d[3], half_fill=half, return_neighbor_list=True)
            finally:
                config.fma = True
            got = ro.records_from_coo(e.cpu(), sft.cpu())
            if half:
                assert np.array_equal(np.unique(ro.canonical_undirected(got), axis=0),
                                      np.unique(ro.canonical_undirected(want), axis=0)), (coo_path, fma)
                assert got.shape[0] == want.shape[0]
            else:
                assert np.array_equal(got, want), (coo_path, fma)
            num = (p[1:] - p[:-1]).cpu().numpy()
            assert np.array_equal(num, np.bincount(got[:, 0], minlength=pos.shape[0]))
    # a dense periodic box: 2 x 2 x 2 cells of ~250 atoms (> 64 targets per cell -> swept in parts), repeated queries
    pos, cell, pbc = random_system(2000, 12.6, torch.float32, seed=9)
    want = ro.records_from_matrix(*ro.cell_list(pos, 6.0, cell, pbc, max_neighbors=2048, nthreads=8))
    for _ in range(2):
        e, p, sft = _nl().cell_list(pos.to(DEV), 6.0, cell.to(DEV), pbc.to(DEV), return_neighbor_list=True)
        assert np.array_equal(ro.records_from_coo(e.cpu(), sft.cpu()), want), coo_path


def test_matrix_path_is_cuda_graph_capturable():
    """The padded-matrix path has no host sync: build + query with pre-allocated outputs can be captured in a CUDA graph
    (the kernels are launched with the programmatic-dependent-launch attribute, also under capture) and replayed on new
    positions in the same buffers."""
    n, M = 4000, 256       # (134 neighbors per atom on average: no row overflows, so the stored sets are comparable)
    pos_a, cell, pbc = random_system(n, 30.0, torch.float32, seed=3)
    pos_b, _, _ = random_system(n, 30.0, torch.float32, seed=4)
    d_pos, d_cell, d_pbc = pos_a.to(DEV), cell.to(DEV), pbc.to(DEV)
    nm = torch.empty((n, M), dtype=torch.int32, device=DEV)
    sh = torch.empty((n, M, 3), dtype=torch.int32, device=DEV)
    num = torch.empty((n,), dtype=torch.int32, device=DEV)

    def call():
        return _nl().cell_list(d_pos, 6.0, d_cell, d_pbc, max_neighbors=M, neighbor_matrix=nm, neighbor_matrix_shifts=sh,
                               num_neighbors=num)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        call()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = call()
    assert out[0] is nm and out[1] is num and out[2] is sh
    for pos in (pos_b, pos_a):
        d_pos.copy_(pos.to(DEV))
        nm.fill_(-3); num.fill_(-3); sh.fill_(-3)
        g.replay()
        torch.cuda.synchronize()
        want = ro.records_from_matrix(*ro.cell_list(pos, 6.0, cell, pbc, max_neighbors=M, nthreads=8))
        assert np.array_equal(_records_gpu_matrix(nm, num, sh), want)
        _check_matrix_padding(nm, num, sh, n)


def test_coo_paths_batch_and_sharded_blocks(coo_path):
    """Batched mixed-PBC systems through the public API, and the rank-sharded fill (index_offset, block layout)."""
    from nvalchemiops_b200.neighborlist import _engine

    pos, cell, pbc, bidx, bptr = bench_batch(12, 300, 900, seed=21, mixed_pbc=True)
    want = ro.records_from_matrix(*ro.batch_cell_list(pos, 6.0, cell, pbc, bidx, max_neighbors=1024))
    d = [t.to(DEV) for t in (pos, cell, pbc, bidx, bptr)]
    e, p, s = _nl().batch_cell_list(d[0], 6.0, d[1], d[2], d[3], return_neighbor_list=True)
    assert np.array_equal(ro.records_from_coo(e.cpu(), s.cpu()), want), coo_path
    # sharded block: second half of the systems as a "rank" with global indices
    s0, s1 = 6, 12
    a0, a1 = int(bptr[s0]), int(bptr[s1])
    lptr = (d[4][s0:s1 + 1] - a0).to(torch.int32)
    lidx = (d[3][a0:a1] - s0).to(torch.int32)
    h = _engine.build(d[0][a0:a1], 6.0, d[1][s0:s1], d[2][s0:s1], batch_idx=lidx, batch_ptr=lptr)
    num, ptr, total, max_count, err, hint, rows = _engine.count_and_size(h, 36.0)
    assert rows == (coo_path == "rows")
    pmax = total + 5
    block = torch.full((5 * pmax,), -7, dtype=torch.int32, device=DEV)
    _engine.fill_coo(h, 36.0, ptr, block[:2 * pmax], block[2 * pmax:], pmax, False, a0, launch_hint=hint, rows=rows)
    torch.cuda.synchronize()
    edge = torch.stack([block[:total], block[pmax:pmax + total]])
    shifts = block[2 * pmax:2 * pmax + 3 * total].reshape(total, 3)
    sel = (want[:, 0] >= a0) & (want[:, 0] < a1)
    assert np.array_equal(ro.records_from_coo(edge.cpu(), shifts.cpu()), want[sel]), coo_path
    assert (block[total:pmax] == -7).all() and (block[2 * pmax + 3 * total:] == -7).all()


def test_rows_path_overflow_falls_back_to_masks():
    """A temporary row buffer that is too small must not change the result: the engine repeats the query on the
    two-pass path (nvnl_status.rows_overflow)."""
    from nvalchemiops_b200 import _lib, config
    from nvalchemiops_b200.neighborlist import _engine

    assert config.coo_path == "rows"
    pos, cell, pbc = random_system(6000, 40.0, torch.float32, seed=33)
    want = ro.records_from_matrix(*ro.cell_list(pos, 6.0, cell, pbc, max_neighbors=256, nthreads=8))
    L = _lib.lib()
    L.nvnl_set_rows_budget(8, 0)
    try:
        h = _engine.build(pos.to(DEV), 6.0, cell.to(DEV), pbc.to(DEV))
        num, ptr, total, max_count, err, hint, rows = _engine.count_and_size(h, 36.0)
        assert h.rows_overflow and not rows
        e, p, s = _nl().cell_list(pos.to(DEV), 6.0, cell.to(DEV), pbc.to(DEV), max_neighbors=256, return_neighbor_list=True)
        assert np.array_equal(ro.records_from_coo(e.cpu(), s.cpu()), want)
    finally:
        L.nvnl_set_rows_budget(-1, -1)
    h = _engine.build(pos.to(DEV), 6.0, cell.to(DEV), pbc.to(DEV))
    num, ptr, total, max_count, err, hint, rows = _engine.count_and_size(h, 36.0)
    assert rows and not h.rows_overflow and total == want.shape[0]


def test_config4_1m_atoms_masks_path_matches_rows_path():
    """Config 4 at full size on both COO paths: identical (i, j, s) sets."""
    from nvalchemiops_b200 import config

    n = 1_000_000
    pos, cell, pbc = bench_box(n, seed=4)
    pos, cell, pbc = pos.to(DEV), cell.to(DEV), pbc.to(DEV)
    keys = {}
    old = config.coo_path
    try:
        for path in ("rows", "masks"):
            config.coo_path = path
            e, ptr, s = _nl().neighbor_list(pos, 6.0, cell=cell, pbc=pbc, return_neighbor_list=True)
            assert (e[0, 1:] >= e[0, :-1]).all()
            keys[path] = (torch.sort(_keys(e[0], e[1], s, n)).values, ptr.clone())
            del e, s
    finally:
        config.coo_path = old
    assert torch.equal(keys["rows"][0], keys["masks"][0])
    assert torch.equal(keys["rows"][1], keys["masks"][1])


def test_rows_path_prezeroed_shifts_on_repeated_queries():
    """Second and later COO queries with the same (atoms, cutoff) signature zero the shifts buffer on a side stream
    while the sweep runs (sized from the previous pair count).  Results must not depend on it: same positions, more
    pairs than guessed (falls back), far fewer pairs than guessed (falls back), boundary rows, deferred cells."""
    from nvalchemiops_b200 import config
    from nvalchemiops_b200.neighborlist import _engine

    assert config.coo_path == "rows" and config.prezero_shifts
    _engine._pair_history.clear()
    old_min, config.prezero_min_pairs = config.prezero_min_pairs, 1
    n = 6000
    base, cell, pbc = random_system(n, 40.0, torch.float32, seed=41)
    dense = base.clone()
    dense[: n // 2] = dense[: n // 2] * 0.25 + 15.0           # half of the atoms squeezed into a (10 A)^3 corner
    sparse = base.clone()
    sparse[:, 0] = torch.linspace(0.0, 39.9, n)               # same count, very different pair total
    sparse[:, 1:] *= 0.02
    seq = [base, base, base * 1.0, dense, dense, sparse, sparse, base]
    for k, pos in enumerate(seq):
        o = ro.cell_list(pos, 6.0, cell, pbc, max_neighbors=4096, nthreads=8)
        assert o[1].max() <= 4096
        want = ro.records_from_matrix(*o)
        e, p, s = _nl().cell_list(pos.to(DEV), 6.0, cell.to(DEV), pbc.to(DEV), max_neighbors=4096, return_neighbor_list=True)
        assert np.array_equal(ro.records_from_coo(e.cpu(), s.cpu()), want), k
        assert e.is_contiguous() and s.is_contiguous() and s.shape == (want.shape[0], 3)
    # batch of small systems: everything deferred to the general kernel, buffer still pre-zeroed on the second call
    bp, bc, bb, bi, bptr = bench_batch(16, 150, 250, seed=3)
    want = ro.records_from_matrix(*ro.batch_cell_list(bp, 6.0, bc, bb, bi, max_neighbors=1024))
    for _ in range(3):
        e, p, s = _nl().batch_cell_list(bp.to(DEV), 6.0, bc.to(DEV), bb.to(DEV), bi.to(DEV), return_neighbor_list=True)
        assert np.array_equal(ro.records_from_coo(e.cpu(), s.cpu()), want)
    config.prezero_shifts = False
    try:
        e, p, s = _nl().cell_list(base.to(DEV), 6.0, cell.to(DEV), pbc.to(DEV), max_neighbors=4096, return_neighbor_list=True)
        want = ro.records_from_matrix(*ro.cell_list(base, 6.0, cell, pbc, max_neighbors=4096, nthreads=8))
        assert np.array_equal(ro.records_from_coo(e.cpu(), s.cpu()), want)
    finally:
        config.prezero_shifts = True
        config.prezero_min_pairs = old_min


def test_speculative_fill_matches_the_regular_path():
    """The output kernel launched before the size sync (nvnl_fill_rows_speculative) must give the same COO outputs:
    interior/boundary cells, a batch with cells left to the general kernel, a guess that is too small, unwrapped input."""
    from nvalchemiops_b200 import config
    from nvalchemiops_b200.neighborlist import _engine

    old_min, old_spec = config.prezero_min_pairs, config.speculative_fill
    config.prezero_min_pairs, config.speculative_fill = 1, True
    _engine._pair_history.clear()
    try:
        n = 6000
        base, cell, pbc = random_system(n, 40.0, torch.float32, seed=51)
        dense = base.clone()
        dense[: n // 2] = dense[: n // 2] * 0.25 + 15.0
        unw = base + 40.0 * (torch.arange(n) % 3 - 1).float()[:, None]
        for k, pos in enumerate([base, base, base, dense, dense, base, unw, unw, base, base]):
            o = ro.cell_list(pos, 6.0, cell, pbc, max_neighbors=4096, nthreads=8)
            want = ro.records_from_matrix(*o)
            e, p, s = _nl().cell_list(pos.to(DEV), 6.0, cell.to(DEV), pbc.to(DEV), max_neighbors=4096, return_neighbor_list=True)
            assert np.array_equal(ro.records_from_coo(e.cpu(), s.cpu()), want), k
            assert e.is_contiguous() and s.is_contiguous() and e.shape == (2, want.shape[0])
            assert np.array_equal(o[1], (p[1:] - p[:-1]).cpu().numpy())
        bp, bc, bb, bi, bptr = bench_batch(24, 150, 900, seed=13)
        want = ro.records_from_matrix(*ro.batch_cell_list(bp, 6.0, bc, bb, bi, max_neighbors=1024))
        for _ in range(3):
            e, p, s = _nl().batch_cell_list(bp.to(DEV), 6.0, bc.to(DEV), bb.to(DEV), bi.to(DEV), return_neighbor_list=True)
            assert np.array_equal(ro.records_from_coo(e.cpu(), s.cpu()), want)
    finally:
        config.prezero_min_pairs, config.speculative_fill = old_min, old_spec
        _engine._pair_history.clear()


# ----------------------------------------------------------------------------------------------
# BASELINE configs 3 and 5 at FULL size, and the reference's own published workload
# ----------------------------------------------------------------------------------------------
def test_config3_full_size_512_systems_mixed_pbc(coo_path):
    """Config 3 as BASELINE.json states it: 512 systems x 150-250 atoms, the 8 PBC patterns 64 times each, batch_idx +
    batch_ptr, COO output — exact set equality with the oracle, on both COO paths."""
    pos, cell, pbc, bidx, bptr = bench_batch(512, 150, 250, seed=3, mixed_pbc=True)
    assert sorted({tuple(p) for p in pbc.tolist()}) == sorted({(bool(a), bool(b), bool(c)) for a in (0, 1) for b in (0, 1) for c in (0, 1)})
    o = ro.batch_cell_list(pos, 6.0, cell, pbc, bidx, max_neighbors=320, nthreads=8)
    assert o[1].max() <= 320
    want = ro.records_from_matrix(*o)
    e, p, s = _nl().neighbor_list(pos.to(DEV), 6.0, cell=cell.to(DEV), pbc=pbc.to(DEV), batch_idx=bidx.to(DEV),
                                  batch_ptr=bptr.to(DEV), return_neighbor_list=True, method="batch_cell_list")
    assert np.array_equal(ro.records_from_coo(e.cpu(), s.cpu()), want)
    assert np.array_equal((p[1:] - p[:-1]).cpu().numpy(), o[1])
    assert (e[0, 1:] >= e[0, :-1]).all()
    b = bidx.to(DEV).long()
    assert torch.equal(b[e[0].long()], b[e[1].long()]), "no pair crosses a system boundary"


def test_config5_full_size_4096_systems_of_1000_atoms():
    """Config 5 at full size (4096 x 1000 atoms, 3 cells per dimension: every cell touches the periodic boundary):
    invariants on the complete 3.7e8-pair output, exact per-atom counts and exact set equality with the oracle on every
    16th system (256 systems, 2.3e7 pairs)."""
    S, n1 = 4096, 1000
    pos, cell, pbc, bidx, bptr = bench_batch(S, n1, n1, seed=5, mixed_pbc=False)
    n = S * n1
    e, p, s = _nl().neighbor_list(pos.to(DEV), 6.0, cell=cell.to(DEV), pbc=pbc.to(DEV), batch_idx=bidx.to(DEV),
                                  batch_ptr=bptr.to(DEV), return_neighbor_list=True, method="batch_cell_list")
    P = e.shape[1]
    assert p[0].item() == 0 and p[-1].item() == P and (p[1:] >= p[:-1]).all()
    assert (e[0, 1:] >= e[0, :-1]).all()
    assert torch.equal(torch.bincount(e[0].long(), minlength=n).to(torch.int32), p[1:] - p[:-1])
    assert torch.equal(e[0] // n1, e[1] // n1), "no pair crosses a system boundary"
    assert s.abs().max().item() <= 1 and abs(P / n - 90.4) < 0.5
    # symmetry {(i, j, s)} == {(j, i, -s)} through an order-independent 64-bit checksum of a mixing hash of the records
    def mix(i, j, sh):
        k = (i.long() * 4194301 + j.long()) * 27 + (sh[:, 0].long() + 1) * 9 + (sh[:, 1].long() + 1) * 3 + (sh[:, 2].long() + 1)
        k = (k ^ (k >> 29)) * -4658895280553007687
        return int((k ^ (k >> 32)).sum().item())
    assert mix(e[0], e[1], s) == mix(e[1], e[0], -s)
    # exact comparison on a sample of the systems
    sel = torch.arange(0, S, 16)
    atoms = (sel[:, None] * n1 + torch.arange(n1)[None, :]).reshape(-1)
    o = ro.batch_cell_list(pos[atoms], 6.0, cell[sel], pbc[sel], torch.arange(sel.numel(), dtype=torch.int32).repeat_interleave(n1),
                           max_neighbors=176, nthreads=8)
    assert o[1].max() <= 176
    counts = (p[1:] - p[:-1]).cpu()
    assert np.array_equal(counts[atoms].numpy(), o[1])
    want = ro.records_from_matrix(*o)                              # indices local to the sample
    in_sample = ((e[0] // n1) % 16 == 0)
    ee, ss = e[:, in_sample].cpu(), s[in_sample].cpu()
    remap = lambda g: (g // n1) // 16 * n1 + g % n1                # noqa: E731  global -> sample-local atom index
    got = ro.records_from_coo(torch.stack([remap(ee[0]), remap(ee[1])]), ss)
    assert np.array_equal(got, want)


def test_reference_published_fcc_benchmark_pair_counts():
    """The reference's own benchmark workload (FCC a = 4 A, r_cut = 5 A, fp32, max_neighbors = 192,
    benchmarks/neighborlist/benchmark_config.yaml) at every size it publishes results for: the total neighbor count must
    equal the number the real Warp kernels produced on an H100 (tests/golden/reference_published_fcc.json), through the
    matrix API with pre-allocated outputs exactly as the benchmark calls it, and through the COO path."""
    from systems import fcc_benchmark_system, load_published_fcc

    g = load_published_fcc()
    for n_str, want in g["cell_list_total_neighbors"].items():
        n = int(n_str)
        pos, cell, pbc = fcc_benchmark_system(n)
        pos_d, cell_d, pbc_d = pos.to(DEV), cell.to(DEV), pbc.to(DEV)
        nm = torch.empty((n, 192), dtype=torch.int32, device=DEV)
        sh = torch.empty((n, 192, 3), dtype=torch.int32, device=DEV)
        num = torch.empty((n,), dtype=torch.int32, device=DEV)
        _nl().neighbor_list(pos_d, 5.0, cell=cell_d, pbc=pbc_d, method="cell_list", neighbor_matrix=nm,
                            neighbor_matrix_shifts=sh, num_neighbors=num)
        assert int(num.sum().item()) == want, (n, int(num.sum().item()), want)
        e, p, s = _nl().neighbor_list(pos_d, 5.0, cell=cell_d, pbc=pbc_d, method="cell_list", return_neighbor_list=True,
                                      max_neighbors=192)
        assert e.shape[1] == want and torch.equal(p[1:] - p[:-1], num)


def test_real_warp_reference_when_installed():
    """Primary oracle when available: the REAL reference (Warp CPU and Warp CUDA) on a knife-edge-rich input — decides
    the FMA mode by evidence.  Skipped where warp-lang / nvalchemiops are not installed."""
    import warp_reference as wr

    if not wr.available():
        pytest.skip("real reference not importable here: " + wr.why_unavailable())
    from nvalchemiops_b200 import config

    pos, cell, pbc = bench_box(200_000, seed=4)

    def ours(fma):
        old = config.fma
        config.fma = fma
        try:
            e, p, s = _nl().neighbor_list(pos.to(DEV), 6.0, cell=cell.to(DEV), pbc=pbc.to(DEV), return_neighbor_list=True)
        finally:
            config.fma = old
        return ro.records_from_coo(e.cpu(), s.cpu())

    for device in ("cpu", DEV):
        mode, d_fma, d_sep = wr.decide_fma_mode(pos, 6.0, cell, pbc, ours, device=device, max_neighbors=160)
        assert min(d_fma, d_sep) == 0, f"neither arithmetic variant reproduces the reference on {device}: {d_fma} / {d_sep}"
        if mode is not None:
            assert mode == config.fma, f"config.fma default disagrees with the real reference on {device}"
