#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 cell-list neighbor-list path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

Metric (BASELINE.json): directed neighbor pairs/s (and atoms/s) of the cell-list build for a single periodic box
of 1,000,000 atoms, density 0.1 /A^3, r_cut = 6 A, COO output (config 4).  A "step" is one complete
``neighbor_list(positions, 6.0, cell, pbc, return_neighbor_list=True)`` call through the public API: grid + hash +
counting sort + the stencil sweep (k_rows: every distance once, compact temporary rows) + scan + (the one size sync) +
output allocation + the output kernel (k_rows_out: edge_index / shifts in atom order).  ``--coo-path masks`` times the
older two-pass path (count sweep -> hit masks -> fill sweep).

* N = 1: config 4.  N > 1 (torchrun): the single box does not shard ("replicas only"), so
  every rank runs its own config-4 replica (weak scaling, no data-path collective); the line also carries a
  ``sharded_batch`` object: config 5 (4096 systems x 1000 atoms) sharded by batch_ptr with the NCCL all-gather.
* ``value``: whole-job pairs/s, inputs resident in HBM, CUDA events, max over ranks.
* ``e2e``: same call with HOST buffers: pinned positions -> H2D -> neighbor_list -> D2H of the full COO result.
* ``roofline``: the dominant kernel stage measured live with CUDA events; algorithmic bytes per SURVEY.md §8d.
* ``cpu_baseline`` / ``--impl reference``: the reference algorithm (oracle port of the Warp kernels, reference grid
  with its 1000-cell cap) on the host cores, on a bounded sample of the same workload.

L2 note: every step writes ~1.8 GB of output (> 126 MB L2), which evicts the 12 MB of inputs between steps; an
explicit 256 MB flush is additionally issued between timed steps, outside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "nvalchemi-toolkit-ops_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

CUTOFF = 6.0
METRIC = "neighbor_pairs_per_s (cell_list build, 1M atoms, r_cut=6A, COO)"


# ------------------------------------------------------------------------------------------------
def _ensure_library():
    """The CUDA library is built in-tree (git-ignored, shipped with the snapshot): build it if it is missing."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("nvnl_build_script", os.path.join(ROOT, "nvalchemi-toolkit-ops_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if not os.path.exists(mod.LIB):
        if int(os.environ.get("LOCAL_RANK", "0")) == 0:
            mod.build()
        else:
            for _ in range(600):
                if os.path.exists(mod.LIB):
                    break
                time.sleep(0.5)


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(n_atoms, n_pairs, batched_systems=0):
    """inputs read once + API-mandated outputs written once."""
    b = 12 * n_atoms + 20 * n_pairs + 4 * (n_atoms + 1)
    if batched_systems:
        b += 4 * n_atoms + 4 * (batched_systems + 1) + 39 * batched_systems
    return b


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_run(n_atoms, seed, budget_s, nthreads, steps=1, warmup=0):
    """Reference algorithm (cell_list.py:35-556 restated in oracle/, reference grid: max_nbins=1000) on config 4.
    The cell list is built for the full box; the query runs the first ``n_limit`` per-atom threads, n_limit chosen
    from a probe so that one step costs about ``budget_s`` seconds.  Returns atoms/s, pairs/s and a description."""
    import reference_oracle as ro
    from systems import bench_box

    pos, cell, pbc = bench_box(n_atoms, seed=seed)
    pos_n, cell_n = pos.numpy(), cell.numpy()
    pbc_n = pbc.numpy().reshape(-1).astype(np.uint8)
    M = 160
    max_cells, radius = ro.estimate_cell_list_sizes(cell_n, pbc_n, CUTOFF)
    cache = ro.allocate_cell_list(n_atoms, max_cells, radius)
    nm = np.empty((n_atoms, M), dtype=np.int32)
    sh = np.empty((n_atoms, M, 3), dtype=np.int32)
    num = np.empty((n_atoms,), dtype=np.int32)

    def one(n_limit):
        t0 = time.perf_counter()
        nm.fill(n_atoms); sh.fill(0); num.fill(0)               # cell_list.py:1358-1373
        for c in cache:
            if c is not radius:
                c.fill(0)
        ro.build_cell_list(pos_n, CUTOFF, cell_n, pbc_n, *cache)
        t1 = time.perf_counter()
        ro.query_cell_list(pos_n, CUTOFF, cell_n, pbc_n, *cache, nm, sh, num, False, nthreads=nthreads, n_limit=n_limit)
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    probe = min(n_atoms, 400 * max(1, nthreads))
    tb, tq = one(probe)
    per_atom = tq / probe
    n_limit = int(min(n_atoms, max(probe, budget_s / max(per_atom, 1e-12))))
    times = []
    for k in range(warmup + steps):
        tb, tq = one(n_limit)
        if k >= warmup:
            times.append((tb, tq))
    tb = float(np.mean([t[0] for t in times])); tq = float(np.mean([t[1] for t in times]))
    # one full step = setup (memsets + build, measured on the full box) + query extrapolated to all atoms
    t_full = tb + tq * (n_atoms / n_limit)
    pairs_per_atom = 90.485084  # config 4, seed 4 (tests/test_gpu_parity.py pins the exact count)
    atoms_s = n_atoms / t_full
    return {
        "atoms_per_s": atoms_s, "pairs_per_s": atoms_s * pairs_per_atom, "t_step_s": t_full, "cores": nthreads,
        "sample": f"config 4 box ({n_atoms} atoms): full build + query of the first {n_limit} atoms "
                  f"(reference grid {max_cells} cells, max_neighbors={M}), query time scaled by {n_atoms / n_limit:.1f}x",
        "measured_s": tb + tq,
    }


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--atoms", type=int, default=1_000_000)
    ap.add_argument("--cpu-budget-s", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sharded", action="store_true")
    ap.add_argument("--coo-path", default=None, choices=["rows", "masks"],
                    help="override nvalchemiops_b200.config.coo_path (default: the package default)")
    args = ap.parse_args()
    steps, warmup = max(1, args.steps), max(0, args.warmup)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # ------------------------------------------------ reference arm ------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        nthreads = host_threads()
        k = min(steps, 3)
        r = cpu_reference_run(args.atoms, 4, args.cpu_budget_s, nthreads, steps=k, warmup=min(warmup, 1))
        line = {
            "impl": "reference", "metric": METRIC, "value": r["pairs_per_s"], "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": k, "warmup": min(warmup, 1), "ms_per_step": r["t_step_s"] * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "atoms_per_s": r["atoms_per_s"],
            "config": {"workload": f"config 4: single periodic box, {args.atoms} atoms, rho=0.1/A^3, r_cut=6A, COO (seed 4)",
                       "implementation": "oracle port of the reference Warp kernels (cell_list.py:35-556), reference grid "
                                         "(max_nbins=1000); the reference itself needs warp-lang, which cannot be installed here"},
            "cpu_baseline": {"value": r["pairs_per_s"], "unit": "pairs/s", "cores": r["cores"], "kind": "port",
                             "sample": r["sample"]},
            "e2e": {"value": r["pairs_per_s"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------ B200 arm ------------------------------------------------
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist

    if world > 1:
        # keep stdout to the one JSON line: NCCL_DEBUG=VERSION (set on some boxes) prints a banner to stdout
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    _ensure_library()
    from nvalchemiops_b200 import config as nl_config, launch_count
    from nvalchemiops_b200.neighborlist import _engine, neighbor_list

    if args.coo_path:
        nl_config.coo_path = args.coo_path
    from systems import bench_batch, bench_box

    n = args.atoms
    pos_h, cell_h, pbc_h = bench_box(n, seed=4 + rank)
    pos_pin = pos_h.pin_memory()
    pos, cell, pbc = pos_h.to(dev), cell_h.to(dev), pbc_h.to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step():
        return neighbor_list(pos, CUTOFF, cell=cell, pbc=pbc, return_neighbor_list=True)

    for _ in range(max(warmup, 3)):
        out = step()
    torch.cuda.synchronize()
    P = int(out[0].shape[1])
    del out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed: K steps of the public API, inputs resident in HBM ----
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    l0 = launch_count()
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.__enter__()                         # sampled through the timed loop, the stage timings and the e2e loop
    for k in range(steps):
        flush.zero_()                          # L2 flush, outside the timed region
        ev[k][0].record()
        out = step()
        ev[k][1].record()
        del out
    torch.cuda.synchronize()
    barrier()
    launches = launch_count() - l0
    t_steps = [a.elapsed_time(b) for a, b in ev]
    t_total_ms = float(sum(t_steps))
    if world > 1:
        tt = torch.tensor([t_total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_total_ms = float(tt.item())
    ms_per_step = t_total_ms / steps
    value = world * P / (ms_per_step * 1e-3)
    atoms_s = world * n / (ms_per_step * 1e-3)

    # ---- stage timings (C-ABI calls bracketed by events): which kernel dominates, and its roofline ----
    csq = _engine.cutoff_sq_in_dtype(CUTOFF, pos.dtype)
    st = {"build": [], "count": [], "fill_coo": []}
    edge = torch.empty((2, P), dtype=torch.int32, device=dev)
    shf = torch.empty((P, 3), dtype=torch.int32, device=dev)
    rows_path = nl_config.coo_path == "rows"
    for k in range(max(5, min(steps, 20))):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        flush.zero_()
        e[0].record(); h = _engine.build(pos, CUTOFF, cell, pbc)
        e[1].record(); num, ptr = _engine.count(h, csq, rows=rows_path)
        e[2].record(); hint = _engine.status(h)[4]          # the size sync of the COO path (not part of either stage)
        assert not h.rows_overflow
        e[3].record(); _engine.fill_coo(h, csq, ptr, edge, shf, P, launch_hint=hint, rows=rows_path)
        e[4].record(); torch.cuda.synchronize()
        st["build"].append(e[0].elapsed_time(e[1])); st["count"].append(e[1].elapsed_time(e[2]))
        st["fill_coo"].append(e[3].elapsed_time(e[4]))
    stage_ms = {k: float(np.median(v)) for k, v in st.items()}
    del edge, shf
    peak, peak_src = measured_peak_gbs()
    B = algorithmic_bytes(n, P)
    dom = max(("count", "fill_coo"), key=lambda k: stage_ms[k])
    # the fill stage is the one that moves the API-mandated bytes (rows path: k_rows_out streams them in atom order
    # from the temporary rows the sweep left; masks path: k_fast<FILL_COO> expands hit masks into rows)
    fill_gbs = B / (stage_ms["fill_coo"] * 1e-3) / 1e9
    kernel_name = ("nvnl::k_rows_out (nvnl_fill_rows stage: temporary rows -> edge_index/shifts in atom order)" if rows_path
                   else "nvnl::k_fast<float, FILL_COO> (nvnl_fill_coo stage: mask expansion + COO row writes)")
    roofline = {
        "bound": "hbm", "kernel": kernel_name, "coo_path": nl_config.coo_path,
        "achieved": fill_gbs, "peak": peak, "unit": "GB/s", "frac": fill_gbs / peak, "traffic": None,
        "peak_source": peak_src, "algorithmic_bytes": B,
        "stages_ms": stage_ms, "longest_stage": dom,
        "pipeline": {"achieved": B / (ms_per_step * 1e-3) / 1e9, "frac": B / (ms_per_step * 1e-3) / 1e9 / peak,
                     "note": "all stages of one neighbor_list call (incl. the size sync and output allocation)"},
        "api_note": "stages_ms times nvnl_fill_rows writing every output byte itself (the kernel the roofline line is about); in "
                    "the public-API loop above, repeated queries let the sweep kernel zero-fill the shifts buffer while it "
                    "sweeps (fused into k_rows' producer warps), so ms_per_step is below the sum of the stages",
        "count_stage_note": "the sweep (k_rows / k_fast<COUNT>) is fp32-issue bound (583 distance tests/atom), not an HBM kernel: "
                            f"{n * 583 / (stage_ms['count'] * 1e-3) / 1e12:.2f} T tests/s",
    }
    ncu_json = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(ncu_json):
        try:
            roofline["traffic"] = json.load(open(ncu_json)).get("rows_out_dram_bytes" if rows_path else "fill_coo_dram_bytes")
        except Exception:
            pass

    # ---- e2e: host buffers; H2D of the inputs and D2H of the full COO result inside the timed region ----
    edge_pin = torch.empty((2, P), dtype=torch.int32).pin_memory()
    ptr_pin = torch.empty((n + 1,), dtype=torch.int32).pin_memory()
    sh_pin = torch.empty((P, 3), dtype=torch.int32).pin_memory()
    cell_pin, pbc_pin = cell_h.pin_memory(), pbc_h.pin_memory()

    def step_e2e():
        p = pos_pin.to(dev, non_blocking=True)
        c = cell_pin.to(dev, non_blocking=True)
        b = pbc_pin.to(dev, non_blocking=True)
        e_, p_, s_ = neighbor_list(p, CUTOFF, cell=c, pbc=b, return_neighbor_list=True)
        edge_pin.copy_(e_, non_blocking=True); ptr_pin.copy_(p_, non_blocking=True); sh_pin.copy_(s_, non_blocking=True)

    k_e2e = max(3, min(steps, 10))
    for _ in range(2):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k_e2e):
        step_e2e()
    e1.record()
    torch.cuda.synchronize()
    t_e2e = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e = float(tt.item())
    e2e_ms = t_e2e / k_e2e
    clocks.__exit__()
    e2e = {"value": world * P / (e2e_ms * 1e-3), "unit": "pairs/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": 12 * n + 36 + 3, "d2h_bytes_per_step": 20 * P + 4 * (n + 1),
           "note": "pinned host positions/cell/pbc -> H2D -> neighbor_list -> D2H of edge_index, neighbor_ptr, shifts"}
    del edge_pin, sh_pin

    # ---- N > 1: config 5 sharded by batch_ptr, with and without the NCCL all-gather ----
    sharded = None
    if world > 1 and not args.no_sharded:
        from nvalchemiops_b200.neighborlist.distributed import sharded_batch_neighbor_list

        bp, bc, bb, bi, bptr = bench_batch(4096, 1000, 1000, seed=5, mixed_pbc=False)
        bp, bc, bb, bptr = bp.to(dev), bc.to(dev), bb.to(dev), bptr.to(dev)
        res = {}
        for name, gather in (("kernels_only", False), ("with_allgather", True)):
            for _ in range(3):
                o = sharded_batch_neighbor_list(bp, CUTOFF, bc, bb, bptr, gather=gather)
            Pb = int(o[0].shape[1])
            del o
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                o = sharded_batch_neighbor_list(bp, CUTOFF, bc, bb, bptr, gather=gather)
                del o
            b.record(); torch.cuda.synchronize()
            tt = torch.tensor([a.elapsed_time(b) / 5], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            tot = torch.tensor([Pb], device=dev, dtype=torch.int64)
            if not gather:
                dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            res[name] = {"ms_per_step": float(tt.item()), "pairs_per_s": float(tot.item()) / (float(tt.item()) * 1e-3),
                         "atoms_per_s": 4096 * 1000 / (float(tt.item()) * 1e-3), "pairs": int(tot.item())}
        sharded = {"workload": "config 5: 4096 systems x 1000 atoms, periodic, COO, sharded by batch_ptr (strong scaling)",
                   **res}

    # ---- cpu_baseline (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        nthreads = host_threads()
        r = cpu_reference_run(n, 4, args.cpu_budget_s, nthreads)
        cpu = {"value": r["pairs_per_s"], "unit": "pairs/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "atoms_per_s": r["atoms_per_s"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": steps, "warmup": max(warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "atoms_per_s": atoms_s,
            "config": {"workload": f"config 4: single periodic box, {n} atoms, rho=0.1/A^3, r_cut=6A, COO output "
                                   f"(seed 4+rank); {P} directed pairs per box" + ("; one replica per GPU" if world > 1 else ""),
                       "l2": "1.8 GB written per step (> L2) + explicit 256 MB flush between timed steps",
                       "parallelism": "replicas" if world > 1 else "single GPU"},
            "e2e": e2e, "gpu_launches": int(launches), "gpu_launches_per_step": launches / steps,
            "clocks": clocks.summary(), "roofline": roofline,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if sharded is not None:
            line["sharded_batch"] = sharded
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
