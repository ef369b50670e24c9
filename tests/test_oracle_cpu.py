"""CPU tests (no GPU): the oracle against the reference's golden vectors and an independent brute force,
the reference-side host helpers, and the C-ABI library surface."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import reference_oracle as ro
from systems import kat_structure, load_kat, random_system, triclinic_system

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["HoTlPd", "SiCu"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("fma", [0, 1])
def test_oracle_reproduces_reference_known_answers(name, dtype, fma):
    """test/neighborlist/test_cell_list.py:391-419 — per-atom counts at rc = 1, 4, 6."""
    pos, cell, pbc, expected = kat_structure(name, dtype)
    for k, rc in enumerate(load_kat()["cutoffs"]):
        _, num, _ = ro.cell_list(pos, rc, cell, pbc, fma_mode=fma)
        assert num.tolist() == expected[k]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_oracle_batch_known_answers(dtype):
    """test/neighborlist/test_batch_cell_list.py:516-542 — both structures in one batch."""
    p1, c1, b1, e1 = kat_structure("HoTlPd", dtype)
    p2, c2, b2, e2 = kat_structure("SiCu", dtype)
    pos = torch.cat([p1, p2])
    cell = torch.stack([c1, c2])
    pbc = torch.stack([b1, b2])
    bidx = torch.tensor([0] * len(p1) + [1] * len(p2), dtype=torch.int32)
    for k, rc in enumerate(load_kat()["cutoffs"]):
        _, num, _ = ro.batch_cell_list(pos, rc, cell, pbc, bidx)
        assert num.tolist() == e1[k] + e2[k]


def test_oracle_matches_independent_brute_force():
    """Random small systems, mixed PBC, cubic and triclinic, partly unwrapped — restatement == brute force."""
    rng = np.random.default_rng(0)
    for trial in range(40):
        n = int(rng.integers(5, 60))
        L = float(rng.uniform(2.0, 9.0))
        rc = float(rng.uniform(0.8, 5.0))
        pbc = [bool(x) for x in rng.integers(0, 2, 3)]
        dtype = torch.float32 if trial % 2 == 0 else torch.float64
        if trial % 3 == 0:
            pos, cell, pb = triclinic_system(n, L, L * 1.2, L * 0.9, 75.0, 100.0, 65.0, dtype, seed=trial, pbc_flag=pbc,
                                             spread=(-0.3, 1.3))
        else:
            pos, cell, pb = random_system(n, L, dtype, seed=trial, pbc_flag=pbc)
            if trial % 3 == 1:
                pos = pos * 1.6 - 0.3 * L  # unwrapped
        for fma in (0, 1):
            nm, num, sh = ro.cell_list(pos, rc, cell, pb, max_neighbors=4096, fma_mode=fma)
            assert num.max() <= 4096
            rec = ro.records_from_matrix(nm, num, sh)
            bf = ro.brute_force(pos, rc, cell, pb, fma_mode=fma)
            assert np.array_equal(rec, bf), f"trial {trial} fma {fma}"


def test_oracle_half_fill_is_half_of_full():
    pos, cell, pbc = random_system(40, 6.0, torch.float32, seed=5)
    nm, num, sh = ro.cell_list(pos, 3.0, cell, pbc)
    nmh, numh, shh = ro.cell_list(pos, 3.0, cell, pbc, half_fill=True)
    full = ro.canonical_undirected(ro.records_from_matrix(nm, num, sh))
    half = ro.canonical_undirected(ro.records_from_matrix(nmh, numh, shh))
    assert num.sum() == 2 * numh.sum()
    assert np.array_equal(np.unique(full, axis=0), half)


def test_oracle_naive_no_pbc_equals_cell_list():
    """BASELINE config 1 plumbing: naive (2-tuple) == cell-list set without PBC."""
    pos = random_system(256, 13.68, torch.float32, seed=1)[0]
    nm, num = ro.neighbor_list(pos, 6.0)
    cell = torch.eye(3).reshape(1, 3, 3)
    nm2, num2, sh2 = ro.cell_list(pos, 6.0, cell, torch.tensor([False] * 3))
    assert np.array_equal(ro.records_from_matrix(nm, num), ro.records_from_matrix(nm2, num2, sh2))
    assert (sh2 == 0).all()


def test_prefix_sum_and_coo_conversion_known_answers():
    g = load_kat()
    assert np.concatenate([[0], np.cumsum(g["prefix_sum"]["counts"][:-1])]).tolist() == g["prefix_sum"]["starts"]
    c = g["coo_conversion"]
    nm = np.array(c["neighbor_matrix"], dtype=np.int32)
    num = np.array(c["num_neighbors"], dtype=np.int32)
    nl, ptr = ro.get_neighbor_list_from_neighbor_matrix(nm, num, fill_value=c["fill_value"])
    assert nl.shape == (2, c["num_pairs"]) and ptr.tolist() == [0, 2, 4, 6, 6]
    with pytest.raises(ro.NeighborOverflowError):
        ro.get_neighbor_list_from_neighbor_matrix(nm, np.array(c["overflow_num_neighbors"], dtype=np.int32),
                                                  fill_value=c["fill_value"])
    e = g["estimate_max_neighbors"]
    assert ro.estimate_max_neighbors(6.0) == e["cutoff_6_default"]
    assert ro.estimate_max_neighbors(5.0, 0.35, 1.0) == e["cutoff_5_density_035_sf_1"]


# ---- host-side mirror of the reference API (pure torch plumbing, runs on CPU) ----
def test_host_helpers_match_reference_contract():
    from nvalchemiops_b200.neighborlist import (NeighborOverflowError, estimate_max_neighbors,
                                                get_neighbor_list_from_neighbor_matrix)
    from nvalchemiops_b200.neighborlist.neighbor_utils import _prepare_batch_idx_ptr

    g = load_kat()
    assert estimate_max_neighbors(6.0) == 1584 and estimate_max_neighbors(0.0) == 0
    c = g["coo_conversion"]
    nm = torch.tensor(c["neighbor_matrix"], dtype=torch.int32)
    num = torch.tensor(c["num_neighbors"], dtype=torch.int32)
    nl, ptr = get_neighbor_list_from_neighbor_matrix(nm, num, fill_value=-1)
    assert nl.shape == (2, 6) and nl.dtype == torch.int32 and ptr.tolist() == [0, 2, 4, 6, 6]
    shm = torch.zeros((4, 4, 3), dtype=torch.int32)
    nl, ptr, sh = get_neighbor_list_from_neighbor_matrix(nm, num, shm, fill_value=-1)
    assert sh.shape == (6, 3)
    with pytest.raises(NeighborOverflowError):
        get_neighbor_list_from_neighbor_matrix(nm, torch.tensor([8, 8, 8, 8], dtype=torch.int32), shm, fill_value=-1)
    bi, bp = _prepare_batch_idx_ptr(None, torch.tensor([0, 3, 5], dtype=torch.int32), 5, "cpu")
    assert bi.tolist() == [0, 0, 0, 1, 1]
    bi, bp = _prepare_batch_idx_ptr(torch.tensor([0, 0, 1, 1, 1], dtype=torch.int32), None, 5, "cpu")
    assert bp.tolist() == [0, 2, 5]
    with pytest.raises(ValueError):
        _prepare_batch_idx_ptr(None, None, 5, "cpu")


def test_product_path_has_no_cpu_fallback():
    from nvalchemiops_b200.neighborlist import batch_cell_list, cell_list, neighbor_list

    pos, cell, pbc = random_system(20, 5.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        cell_list(pos, 2.0, cell, pbc)
    with pytest.raises(RuntimeError, match="CUDA"):
        neighbor_list(pos, 2.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        batch_cell_list(pos, 2.0, cell, pbc, torch.zeros(20, dtype=torch.int32))
    with pytest.raises(ValueError, match="Invalid method"):
        neighbor_list(pos, 2.0, method="bogus")
    with pytest.raises(ValueError, match="Unsupported dtype"):
        cell_list(pos.half(), 2.0, cell.half(), pbc)


def test_c_abi_library_exports_every_declared_symbol():
    """include/nvalchemi_nl_b200.h <-> libnvalchemi_nl_b200.so <-> the ctypes binding (no compute calls)."""
    from nvalchemiops_b200 import _lib

    header = open(os.path.join(ROOT, "include", "nvalchemi_nl_b200.h")).read()
    declared = set(re.findall(r"\b(nvnl_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    L = ctypes.CDLL(_lib.library_path())
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.declared_symbols())
    lib = _lib.lib()
    assert lib.nvnl_abi_version() == _lib.ABI_VERSION == 6
    assert lib.nvnl_workspace_bytes(1000, 1, 0) > 1000 * 16
    assert lib.nvnl_workspace_bytes(1000, 4, 1) > lib.nvnl_workspace_bytes(1000, 4, 0)
    # argument validation happens before any CUDA call
    assert lib.nvnl_build(None, 0, 0, None, None, None, None, 1, 1.0, 0, None, 0, None) != 0
    assert b"positive" in lib.nvnl_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "nvalchemi-toolkit-ops_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "reference_oracle" not in src and "nl_oracle" not in src, f


def test_rows_path_c_abi_argument_validation_and_budget():
    """nvnl_count_rows / nvnl_fill_rows reject bad arguments before touching CUDA; nvnl_set_rows_budget sizes the
    temporary row buffer inside the workspace."""
    from nvalchemiops_b200 import _lib

    lib = _lib.lib()
    dummy = ctypes.c_void_p(4096)          # never dereferenced: validation fails first
    assert lib.nvnl_count_rows(dummy, 1, 100, 1, None, 36.0, 0, 1, dummy, dummy, None, 0, -1, None) != 0
    assert b"fp32 only" in lib.nvnl_last_error()
    assert lib.nvnl_count_rows(dummy, 0, 1 << 27, 1, None, 36.0, 0, 1, dummy, dummy, None, 0, -1, None) != 0
    assert b"2^27" in lib.nvnl_last_error()
    assert lib.nvnl_count_rows(dummy, 0, 100, 1, None, 36.0, 0, 1, dummy, dummy, ctypes.c_void_p(4100), 64, -1, None) != 0
    assert b"16-byte aligned" in lib.nvnl_last_error()
    assert lib.nvnl_fill_rows(dummy, 0, 100, 1, None, 36.0, 0, 1, dummy, dummy, 10, 0, dummy, 0, -1, None) != 0
    assert b"launch_hint" in lib.nvnl_last_error()
    assert lib.nvnl_fill_rows(dummy, 1, 100, 1, None, 36.0, 0, 1, dummy, dummy, 10, 0, dummy, 0, 0, None) != 0
    base = lib.nvnl_workspace_bytes(100000, 1, 0)
    try:
        lib.nvnl_set_rows_budget(8, 0)
        small = lib.nvnl_workspace_bytes(100000, 1, 0)
        lib.nvnl_set_rows_budget(320, -1)
        big = lib.nvnl_workspace_bytes(100000, 1, 0)
    finally:
        lib.nvnl_set_rows_budget(-1, -1)
    assert small < base < big
    assert base - small >= (160 - 8) * 100000 * 4
    assert lib.nvnl_workspace_bytes(100000, 1, 0) == base


def test_engine_coo_path_selection_fallback_and_speculative_shifts(monkeypatch):
    """Host logic of the COO query (no GPU): which path runs, the repeat on rows_overflow, and when the speculatively
    sized shifts buffer is handed to the sweep / accepted as the output (launch_hint bit 2)."""
    from nvalchemiops_b200 import config
    from nvalchemiops_b200.neighborlist import _engine

    class H:
        ws = None; dtype_code = 0; ns = 1; batch_idx = None; device = torch.device("cpu"); cutoff = 6.0
        rows_overflow = False

        def __init__(self, n, dtype=torch.float32):
            self.n, self.dtype = n, dtype

    assert _engine.use_rows(H(1000)) and not _engine.use_rows(H(1000, torch.float64)) and not _engine.use_rows(H(1 << 27))
    monkeypatch.setattr(config, "coo_path", "masks")
    assert not _engine.use_rows(H(1000))
    monkeypatch.setattr(config, "coo_path", "bogus")
    with pytest.raises(ValueError):
        _engine.use_rows(H(1000))
    monkeypatch.setattr(config, "coo_path", "rows")

    calls = []
    state = {"total": 5_000_000, "hint": 0, "overflow_once": False}

    hints = []

    def fake_count(h, csq, half_fill=False, want_ptr=True, rows=False, prezero=None, launch_hint=-1):
        hints.append(launch_hint)
        calls.append(("count", rows, None if prezero is None else prezero.numel()))
        return torch.zeros(h.n, dtype=torch.int32), torch.zeros(h.n + 1, dtype=torch.int32)

    def fake_status(h):
        h.rows_overflow = state["overflow_once"] and calls[-1][1]
        return state["total"], 100, 10, 0, state["hint"]

    def fake_fill(h, csq, ptr, edge, shifts, total, half_fill=False, index_offset=0, launch_hint=-1, rows=False, row_stride=0):
        calls.append(("fill", rows, launch_hint, tuple(edge.shape), tuple(shifts.shape), shifts.is_contiguous()))

    monkeypatch.setattr(_engine, "count", fake_count)
    monkeypatch.setattr(_engine, "status", fake_status)
    monkeypatch.setattr(_engine, "fill_coo", fake_fill)
    monkeypatch.setattr(config, "speculative_fill", False)
    _engine._pair_history.clear()
    h = H(60_000)
    # first query: nothing known -> no speculative buffer, the output kernel zero-fills (bit 2 clear)
    _engine.query_coo(h, 36.0)
    assert calls == [("count", True, None), ("fill", True, 0, (2, 5_000_000), (5_000_000, 3), True)]
    # second query, same signature: buffer of the last size + 2 % goes to the sweep and becomes the output
    calls.clear()
    e, p, s, num = _engine.query_coo(h, 36.0)
    assert calls[0] == ("count", True, 3 * (int(5_000_000 * 1.02) + 1024))
    assert calls[1] == ("fill", True, 4, (2, 5_000_000), (5_000_000, 3), True) and s.shape == (5_000_000, 3)
    # more pairs than guessed: the guess is dropped
    calls.clear(); state["total"] = 5_500_000
    _engine.query_coo(h, 36.0)
    assert calls[0][2] is not None and calls[1][2] == 0
    # far fewer pairs than guessed: do not pin a mostly unused buffer
    calls.clear(); state["total"] = 1_200_000
    _engine.query_coo(h, 36.0)
    assert calls[0][2] is not None and calls[1][2] == 0
    # a different cutoff has its own history
    calls.clear()
    _engine.query_coo(h, 25.0)
    assert calls[0] == ("count", True, None)
    # unwrapped inputs: the hint of the last query (0: only the lean kernel was launched) misses the bit -> the count is
    # repeated with every variant; two-pass kernels served the query -> never "pre-zeroed", and no speculation next time
    calls.clear(); hints.clear(); state["hint"] = 1
    _engine.query_coo(h, 36.0)
    assert [c[0] for c in calls] == ["count", "count", "fill"] and hints == [0, -1] and calls[2][2] == 1
    calls.clear(); hints.clear()
    _engine.query_coo(h, 36.0)
    assert calls[0] == ("count", True, None) and calls[1][2] == 1 and hints == [1]
    # temporary rows overflow: the count is repeated on the two-pass path and the fill follows it
    calls.clear(); state.update(hint=0, overflow_once=True)
    _engine.query_coo(h, 36.0)
    assert [c[:2] for c in calls] == [("count", True), ("count", False), ("fill", False)] and h.rows_overflow
    # small results are not worth the extra buffer
    calls.clear(); state.update(total=10_000, overflow_once=False)
    _engine.query_coo(h, 4.0); _engine.query_coo(h, 4.0)
    assert calls[2] == ("count", True, None)
    # output kernel launched before the size sync into buffers of the guessed size (the default)
    monkeypatch.setattr(config, "speculative_fill", True)
    monkeypatch.setattr(_engine, "fill_rows_speculative",
                        lambda h, ptr, ebuf, zbuf, index_offset=0: calls.append(("spec", ebuf.numel(), zbuf.numel())))
    _engine._pair_history.clear()
    state.update(total=5_000_000, hint=0)
    _engine.query_coo(h, 36.0)                      # no history yet: regular path
    calls.clear()
    e, p, s, num = _engine.query_coo(h, 36.0)       # guess fits: no fill launch after the sync at all
    cap = int(5_000_000 * 1.02) + 1024
    assert calls == [("count", True, 3 * cap), ("spec", 2 * cap, 3 * cap)]
    assert e.shape == (2, 5_000_000) and e.is_contiguous() and s.shape == (5_000_000, 3) and s.is_contiguous()
    calls.clear(); hints.clear(); state["hint"] = 2   # cells left to the general kernel, but the hint (0) did not launch it:
    _engine.query_coo(h, 36.0)                        # count repeated with every variant, speculative outputs dropped
    assert [c[0] for c in calls] == ["count", "spec", "count", "fill"] and hints == [0, -1] and calls[3][:3] == ("fill", True, 2)
    calls.clear(); hints.clear()                      # next time the hint is right: only the general kernel's rows remain
    _engine.query_coo(h, 36.0)
    assert calls[1][0] == "spec" and calls[2][:3] == ("fill", True, 2 | 4 | 8) and hints == [2]
    calls.clear(); state.update(total=6_000_000, hint=0)   # guess too small: the regular fill, nothing pre-zeroed
    _engine.query_coo(h, 36.0)
    assert calls[1][0] == "spec" and calls[2][:3] == ("fill", True, 0) and calls[2][3] == (2, 6_000_000)
    _engine._pair_history.clear()


def test_estimate_cell_list_sizes_matches_the_oracle_restatement():
    """estimate_cell_list_sizes / estimate_batch_cell_list_sizes are torch ops (they run on CPU tensors too): same cell
    counts and radii as the oracle's restatement of cell_list.py:35-99 / batch_cell_list.py:35-99, incl. max_nbins."""
    import numpy as np
    from nvalchemiops_b200.neighborlist import (allocate_cell_list, estimate_batch_cell_list_sizes,
                                                estimate_cell_list_sizes)

    g = torch.Generator().manual_seed(11)
    cells, pbcs = [], []
    for k in range(40):
        L = 3.0 + 60.0 * float(torch.rand(1, generator=g))
        c = torch.eye(3) * L
        if k % 2:
            c = c + (torch.rand(3, 3, generator=g) - 0.5) * 0.3 * L       # triclinic
        p = torch.rand(3, generator=g) > 0.4
        cells.append(c); pbcs.append(p)
        for dtype in (torch.float32, torch.float64):
            for rc, nbins in ((2.7, 1000), (6.0, 1000), (6.0, 27), (1.1, 10**6)):
                got = estimate_cell_list_sizes(c.to(dtype).reshape(1, 3, 3), p.reshape(1, 3), rc, max_nbins=nbins)
                want = ro.estimate_cell_list_sizes(c.to(dtype).reshape(1, 3, 3), p, rc, max_nbins=nbins)
                assert got[0] == want[0] and got[1].tolist() == list(np.asarray(want[1]).reshape(-1)), (k, dtype, rc, nbins)
    cell, pbc = torch.stack(cells), torch.stack(pbcs)
    got = estimate_batch_cell_list_sizes(cell, pbc, 4.0, max_nbins=64)
    want = ro.estimate_batch_cell_list_sizes(cell, pbc, 4.0, max_nbins=64)
    assert got[0] == want[0] and np.array_equal(got[1].numpy(), np.asarray(want[1]).reshape(-1, 3))
    assert estimate_cell_list_sizes(cell[:1], pbc[:1], -1.0)[0] == 1
    cache = allocate_cell_list(10, 7, got[1][:3], torch.device("cpu"))
    assert [tuple(t.shape) for t in cache] == [(3, 3), (3, 3), (10, 3), (10, 3), (7,), (7,), (10,)]
    assert cache[1] is got[1][:3] or torch.equal(cache[1], got[1][:3])


def test_oracle_reproduces_the_pair_counts_the_reference_publishes():
    """Pins the oracle to OUTPUTS OF THE REAL REFERENCE: total_neighbors of its own FCC benchmark workload as measured
    with the Warp kernels and published in docs/benchmarks/benchmark_results/*.csv (tests/golden/make_fcc_golden.py).
    The cell-list restatement and the independent image-sum brute force must both give the published totals."""
    from systems import fcc_benchmark_system, load_published_fcc

    g = load_published_fcc()
    assert g["workload"]["cutoff"] == 5.0 and g["workload"]["lattice_constant"] == 4.0
    for n_str, want in g["cell_list_total_neighbors"].items():
        n = int(n_str)
        if n > 16384:
            continue                                       # (larger sizes: GPU suite, against the same published numbers)
        pos, cell, pbc = fcc_benchmark_system(n)
        assert pos.shape[0] == n
        nm, num, sh = ro.cell_list(pos, 5.0, cell, pbc, max_neighbors=192)
        assert int(num.sum()) == want, (n, int(num.sum()), want)
    for n_str, want in g["naive_total_neighbors"].items():
        n = int(n_str)
        if n > 1536:
            continue
        # (the published naive totals equal the cell-list ones; checked here against the independent brute force)
        pos, cell, pbc = fcc_benchmark_system(n)
        assert ro.brute_force(pos, 5.0, cell, pbc).shape[0] == want, ("brute force", n)


def test_coulomb_oracle_closed_forms_and_reference_erfc():
    """The pair-consumer oracle (coulomb.py:206-292, math.py:52-93 restated): two charges reproduce the closed forms the
    reference's docstrings state; the damped case equals the formulas evaluated by hand with the A&S erfc, which itself
    stays within its documented 1.5e-7 of the exact function; forces obey Newton's third law on a full periodic list."""
    import math

    import coulomb_oracle as co

    q1, q2, r = 1.5, -2.0, 2.0
    pos = np.array([[0.0, 0.0, 0.0], [r, 0.0, 0.0]])
    cell = np.eye(3)[None] * 20.0
    e, f = co.coulomb_energy_forces_list(pos, [q1, q2], cell, 5.0, 0.0, [0, 1], [1, 0], np.zeros((2, 3)))
    assert np.allclose(e, [q1 * q2 / (2 * r)] * 2, rtol=1e-15)
    assert np.allclose(f, [[-q1 * q2 / r**2, 0, 0], [q1 * q2 / r**2, 0, 0]], rtol=1e-15)
    alpha = 0.35
    e, f = co.coulomb_energy_forces_list(pos, [q1, q2], cell, 5.0, alpha, [0, 1], [1, 0], np.zeros((2, 3)))
    erfc_as = float(co.wp_erfc(alpha * r))
    assert abs(erfc_as - math.erfc(alpha * r)) < 1.5e-7
    assert np.allclose(e, [0.5 * q1 * q2 * erfc_as / r] * 2, rtol=1e-15)
    fm = q1 * q2 * (erfc_as / r**3 + 2 * alpha / math.sqrt(math.pi) * math.exp(-(alpha * r) ** 2) / r**2)
    assert np.allclose(f[0], [fm * -r, 0, 0], rtol=1e-13)
    xs = np.linspace(-3.0, 5.0, 161)
    assert np.abs(co.wp_erfc(xs) - np.array([math.erfc(x) for x in xs])).max() < 1.5e-7
    # beyond the cutoff / coincident atoms are skipped (coulomb.py:247)
    e, f = co.coulomb_energy_forces_list(pos, [q1, q2], cell, 1.9, 0.0, [0, 1], [1, 0], np.zeros((2, 3)))
    assert not e.any() and not f.any()
    # periodic system over the neighbor-list oracle: total force vanishes, a shifted image enters through cell^T · s
    p, c, b = random_system(60, 9.0, torch.float32, seed=2)
    rec = ro.records_from_matrix(*ro.cell_list(p, 4.0, c, b, max_neighbors=256))
    qs = np.linspace(-1, 1, 60)
    e, f = co.coulomb_energy_forces_list(p.numpy(), qs, c.numpy(), 4.0, 0.3, rec[:, 0], rec[:, 1], rec[:, 2:])
    assert np.abs(f.sum(0)).max() < 1e-12 and (rec[:, 2:] != 0).any()


def test_pair_consumer_host_checks_need_no_gpu():
    """Argument checks of the consumer API mirror the reference's (coulomb.py:1607-1621) and run before any CUDA call; CPU
    tensors are rejected (no fallback)."""
    from nvalchemiops_b200.interactions import electrostatics as es

    pos, cell, pbc = random_system(8, 5.0, torch.float32, seed=1)
    q = torch.ones(8, dtype=torch.float64)
    nl = torch.zeros((2, 3), dtype=torch.int32)
    sh = torch.zeros((3, 3), dtype=torch.int32)
    with pytest.raises(ValueError, match="Must provide either"):
        es.coulomb_energy_forces(pos, q, cell, 3.0)
    with pytest.raises(ValueError, match="Cannot provide both"):
        es.coulomb_energy_forces(pos, q, cell, 3.0, neighbor_list=nl, neighbor_shifts=sh, neighbor_matrix=nl, neighbor_matrix_shifts=sh)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        es.coulomb_energy_forces(pos, q, cell, 3.0, neighbor_list=nl, neighbor_ptr=torch.zeros(9, dtype=torch.int32), neighbor_shifts=sh)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        es.fused_coulomb_energy_forces(pos, q, cell, pbc, 3.0)
    assert set(es.__all__) == {"coulomb_energy", "coulomb_forces", "coulomb_energy_forces", "fused_coulomb_energy_forces"}


def test_public_api_matches_the_reference_signatures():
    """Every name of the reference's ``nvalchemiops.neighborlist.__all__`` exists here with the reference's parameter
    names, order, kinds and defaults (tests/golden/reference_signatures.json, generated from the reference sources by
    tests/golden/make_reference_signatures.py).  Allowed difference: extra trailing parameters that default to None
    (``batch_cell_list(..., batch_ptr=None)``, ``cell_list_needs_rebuild(..., batch_idx=None)``)."""
    import inspect
    import json

    import nvalchemiops_b200.neighborlist as nl

    with open(os.path.join(ROOT, "tests", "golden", "reference_signatures.json")) as f:
        want = json.load(f)["functions"]
    assert len(want) == 18
    for name, spec in want.items():
        assert name in nl.__all__ and hasattr(nl, name), f"{name} is not exported"
        params = list(inspect.signature(getattr(nl, name)).parameters.values())
        ref = spec["params"]
        assert len(params) >= len(ref), f"{name}: parameters missing"
        for p, (rname, rdefault, rkind) in zip(params, ref):
            assert p.name == rname, f"{name}: parameter {p.name} where the reference has {rname}"
            assert p.kind.name.lower() == rkind, f"{name}.{rname}: kind {p.kind.name}"
            if rdefault is None:
                assert p.default is inspect.Parameter.empty, f"{name}.{rname} has a default here"
            else:
                assert p.default is not inspect.Parameter.empty and repr(p.default) == rdefault, \
                    f"{name}.{rname}: default {p.default!r} vs reference {rdefault}"
        for p in params[len(ref):]:
            assert p.default is None, f"{name}: extra parameter {p.name} must be optional"


def test_compute_naive_num_shifts_restates_the_reference_kernel():
    """neighbor_utils.py:194-211: s_d = ceil(cutoff / face distance) in periodic dims, 0 in open ones;
    num_shifts = s0 (2 s1 + 1)(2 s2 + 1) + s1 (2 s2 + 1) + s2 + 1; offsets = exclusive scan (test_naive.py:415-423
    uses a 3 A cubic cell with cutoff 1.5: one image per direction -> 14 half-space shifts)."""
    from nvalchemiops_b200.neighborlist import compute_naive_num_shifts

    cell = (torch.eye(3) * 3.0).reshape(1, 3, 3)
    rng, off, total = compute_naive_num_shifts(cell, 1.5, torch.tensor([[True, True, True]]))
    assert rng.dtype == torch.int32 and rng.tolist() == [[1, 1, 1]] and off.tolist() == [0, 14] and total == 14
    assert isinstance(total, int) and off.dtype == torch.int32
    rng, off, total = compute_naive_num_shifts(cell, 1.5, torch.tensor([[True, False, True]]))
    assert rng.tolist() == [[1, 0, 1]] and total == 1 * 1 * 3 + 0 + 1 + 1
    # a batch: cubic 10 A with rc 12 (two images), an open system, a triclinic cell checked against numpy
    tri = torch.tensor([[8.0, 0.0, 0.0], [2.0, 7.0, 0.0], [1.0, 1.5, 9.0]], dtype=torch.float64)
    cells = torch.stack([torch.eye(3, dtype=torch.float64) * 10.0, torch.eye(3, dtype=torch.float64) * 10.0, tri])
    pbc = torch.tensor([[True, True, True], [False, False, False], [True, True, False]])
    rng, off, total = compute_naive_num_shifts(cells, 12.0, pbc)
    inv_t = np.linalg.inv(tri.numpy()).T
    s_tri = [int(np.ceil(np.linalg.norm(inv_t[d]) * 12.0)) if pbc[2, d] else 0 for d in range(3)]
    assert rng.tolist() == [[2, 2, 2], [0, 0, 0], s_tri]
    count = lambda s: s[0] * (2 * s[1] + 1) * (2 * s[2] + 1) + s[1] * (2 * s[2] + 1) + s[2] + 1   # noqa: E731
    assert off.tolist() == [0, 63, 64, 64 + count(s_tri)] and total == off[-1].item()


def test_naive_entry_points_host_logic():
    """What the naive entry points decide before any kernel runs (naive.py:560-662, batch_naive.py:653-707,
    naive_dual_cutoff.py:751-772): argument errors, no CPU fallback, and the reference's tuples for cutoff <= 0."""
    from nvalchemiops_b200.neighborlist import (batch_naive_neighbor_list, batch_naive_neighbor_list_dual_cutoff,
                                                naive_neighbor_list, naive_neighbor_list_dual_cutoff)

    pos, cell, pbc = random_system(12, 5.0)
    for fn, args in ((naive_neighbor_list, (pos, 2.0)), (naive_neighbor_list_dual_cutoff, (pos, 1.0, 2.0)),
                     (batch_naive_neighbor_list, (pos, 2.0, torch.zeros(12, dtype=torch.int32))),
                     (batch_naive_neighbor_list_dual_cutoff, (pos, 1.0, 2.0, torch.zeros(12, dtype=torch.int32)))):
        with pytest.raises(ValueError, match="pbc must also be provided"):
            fn(*args, cell=cell)
        with pytest.raises(ValueError, match="cell must also be provided"):
            fn(*args, pbc=pbc)
        with pytest.raises(RuntimeError, match="CUDA"):
            fn(*args)
    with pytest.raises(ValueError, match="Either batch_idx or batch_ptr"):
        batch_naive_neighbor_list(pos, 2.0)
    # cutoff <= 0 with return_neighbor_list: the reference's 3- / 4-tuples with the extra zero [N] tensor
    out = naive_neighbor_list(pos, 0.0, return_neighbor_list=True)
    assert [tuple(t.shape) for t in out] == [(2, 0), (12,), (13,)]
    out = naive_neighbor_list(pos, -1.0, cell=cell, pbc=pbc, return_neighbor_list=True)
    assert [tuple(t.shape) for t in out] == [(2, 0), (12,), (13,), (0, 3)]
    assert all(t.dtype == torch.int32 for t in out)
