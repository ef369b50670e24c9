#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 cell-list neighbor-list path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores
    python bench.py --config 2|3|4|5                         # headline on another BASELINE config (default 4)

Metric (BASELINE.json): directed neighbor pairs/s (and atoms/s) of the cell-list build for a single periodic box
of 1,000,000 atoms, density 0.1 /A^3, r_cut = 6 A, COO output (config 4).  A "step" is one complete
``neighbor_list(positions, 6.0, cell, pbc, return_neighbor_list=True)`` call through the public API: grid + hash +
counting sort + the stencil sweep (k_rows: every distance once, compact temporary rows, the shifts output zero-filled
while it sweeps) + scan + (the one size sync) + output allocation + the output kernel (k_rows_out: edge_index / boundary
shifts in atom order).

* N = 1: config 4.  N > 1 (torchrun): the single box does not shard ("replicas only"), so every
  rank runs its own config-4 replica (weak scaling, no data-path collective).
* ``value``: whole-job pairs/s, inputs resident in HBM, CUDA events, max over ranks.  Steady state: from the second call
  with the same signature on, the engine sizes the shifts buffer from the previous pair count so that the sweep can
  zero-fill it; ``first_call_ms`` is the cold first call.
* ``e2e``: same call with HOST buffers: pinned positions -> H2D -> neighbor_list -> D2H of the full COO result.
* ``roofline``: whole-call algorithmic bytes / ms_per_step against the measured HBM peak (``frac``), plus
  every stage timed live with CUDA events in the configuration the API loop runs, each with its own bound: the sweep is
  fp32-issue bound (distance tests/s against the FP32 peak), the output kernel HBM bound (GB/s).
* ``other_configs`` (N = 1): BASELINE configs 2, 3 and 5 on one GPU through the public API, each with its own
  whole-call roofline fraction.
* ``sharded_batch`` (every N): config 5 (4096 systems x 1000 atoms) sharded by batch_ptr; kernels only / with the NCCL
  gather; bytes received per rank, NVLink fraction, and a checksum of the gathered result that must agree on all ranks.
* ``cpu_baseline`` / ``--impl reference``: the reference algorithm (oracle port of the Warp kernels, reference grid with
  its 1000-cell cap) on the host cores.  A step is the full build + the query of a bounded sample of the atoms; the
  value is the pairs found for that sample over the measured time (nothing is extrapolated into ms_per_step).

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
# FP32 peak measured on this pool's B200 (profiles/r1_microbench_f32x2.txt): 35.1-36.9 T lane-ops/s; one distance test of
# the reference's predicate is 3 subtractions + 1 multiply + 2 fused multiply-adds + 1 compare = 7 lane operations
FP32_LANE_OPS_PEAK = 36.0e12
LANE_OPS_PER_TEST = 7.0
NVLINK_GBS_PER_DIR = 900.0


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


def algorithmic_bytes(n_atoms, n_pairs, batched_systems=0, matrix_width=0):
    """inputs read once + API-mandated outputs written once."""
    if matrix_width:
        return 12 * n_atoms + 16 * n_atoms * matrix_width + 4 * n_atoms
    b = 12 * n_atoms + 20 * n_pairs + 4 * (n_atoms + 1)
    if batched_systems:
        b += 4 * n_atoms + 4 * (batched_systems + 1) + 39 * batched_systems
    return b


def bind_to_gpu_numa_node(index):
    """Pin this process (and therefore the pinned host buffers it first-touches) to the CPUs NVML reports as local to
    the GPU.  Returns a description for the JSON line."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        allowed = os.sched_getaffinity(0)
        pick = cpus & allowed
        if pick and pick != allowed:
            os.sched_setaffinity(0, pick)
            return f"bound to the GPU's {len(pick)} local CPUs"
        return f"GPU-local CPU set = the whole allowed set ({len(allowed)} CPUs): nothing to bind"
    except Exception as e:  # noqa: BLE001
        return f"not bound ({type(e).__name__})"


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
    One step = memsets + full build + the query of the first ``n_limit`` atoms (all of them if that fits ``budget_s``),
    n_limit chosen from a probe.  Nothing is extrapolated: the step time is what was measured and the value is the
    number of pairs the step found over that time."""
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
        ro.query_cell_list(pos_n, CUTOFF, cell_n, pbc_n, *cache, nm, sh, num, False, nthreads=nthreads, n_limit=n_limit)
        return time.perf_counter() - t0

    probe = min(n_atoms, 400 * max(1, nthreads))
    t_probe = one(probe)
    per_atom = max(t_probe / probe, 1e-12)
    n_limit = int(min(n_atoms, max(probe, budget_s / per_atom)))
    times, pairs = [], 0
    for k in range(warmup + steps):
        t = one(n_limit)
        if k >= warmup:
            times.append(t)
            # every pair found from the half stencil of a sampled atom is stored in both directions (full fill):
            # the directed pairs this step produced = the sum of all counters
            pairs = int(num.sum())
    t_step = float(np.mean(times))
    return {
        "atoms_per_s": n_limit / t_step, "pairs_per_s": pairs / t_step, "t_step_s": t_step, "cores": nthreads,
        "sample": f"config 4 box ({n_atoms} atoms, reference grid {max_cells} cells, max_neighbors={M}): memsets + full build + "
                  f"query threads of the first {n_limit} atoms ({pairs} directed pairs stored) per step",
        "sample_atoms": n_limit, "sample_pairs": pairs,
    }


# ------------------------------------------------------------------------------------------------
def time_api(fn, reps, flush=None):
    """Median / all CUDA-event times (ms) of ``fn()`` with an optional L2 flush before each (outside the timing)."""
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
        del out
    return float(np.median(ts)), ts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=[2, 3, 4, 5], help="BASELINE config used as the headline")
    ap.add_argument("--atoms", type=int, default=1_000_000)
    ap.add_argument("--cpu-budget-s", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sharded", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
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
        w = min(warmup, 1)
        r = cpu_reference_run(args.atoms, 4, args.cpu_budget_s, nthreads, steps=k, warmup=w)
        line = {
            "impl": "reference", "metric": METRIC, "value": r["pairs_per_s"], "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": k, "warmup": w, "ms_per_step": r["t_step_s"] * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "atoms_per_s": r["atoms_per_s"],
            "config": {"workload": f"config 4: single periodic box, {args.atoms} atoms, rho=0.1/A^3, r_cut=6A, COO (seed 4); "
                                   f"each step processes a bounded sample of it: {r['sample']}",
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

    numa_note = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    _ensure_library()
    from nvalchemiops_b200 import config as nl_config, launch_count
    from nvalchemiops_b200.neighborlist import _engine, neighbor_list

    if args.coo_path:
        nl_config.coo_path = args.coo_path
    from systems import bench_batch, bench_box

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    peak, peak_src = measured_peak_gbs()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- workloads -------------------------------------------------------------------------------------------------
    def make_workload(cfg, seed_shift=0):
        """(description, call, host tensors, n_atoms, n_systems, matrix_width)"""
        if cfg == 4:
            n = args.atoms
            p, c, b = bench_box(n, seed=4 + seed_shift)
            t = [x.to(dev) for x in (p, c, b)]
            return (f"config 4: single periodic box, {n} atoms, rho=0.1/A^3, r_cut=6A, COO output (seed {4 + seed_shift})",
                    lambda: neighbor_list(t[0], CUTOFF, cell=t[1], pbc=t[2], return_neighbor_list=True), (p, c, b), n, 0, 0)
        if cfg == 2:
            p, c, b = bench_box(50_000, seed=2)
            t = [x.to(dev) for x in (p, c, b)]
            return ("config 2: 50,000 atoms, cubic PBC, r_cut=6A, padded neighbor_matrix (default max_neighbors=1584)",
                    lambda: neighbor_list(t[0], CUTOFF, cell=t[1], pbc=t[2]), (p, c, b), 50_000, 0, 1584)
        if cfg == 3:
            p, c, b, bi, bp = bench_batch(512, 150, 250, seed=3, mixed_pbc=True)
            t = [x.to(dev) for x in (p, c, b, bi, bp)]
            return ("config 3: 512 systems x 150-250 atoms, 8 mixed PBC patterns, batch_idx/batch_ptr, COO output",
                    lambda: neighbor_list(t[0], CUTOFF, cell=t[1], pbc=t[2], batch_idx=t[3], batch_ptr=t[4],
                                          return_neighbor_list=True, method="batch_cell_list"), (p, c, b, bi, bp),
                    int(p.shape[0]), 512, 0)
        p, c, b, bi, bp = bench_batch(4096, 1000, 1000, seed=5, mixed_pbc=False)
        t = [x.to(dev) for x in (p, c, b, bi, bp)]
        return ("config 5 on ONE GPU: 4096 systems x 1000 atoms, periodic, COO output",
                lambda: neighbor_list(t[0], CUTOFF, cell=t[1], pbc=t[2], batch_idx=t[3], batch_ptr=t[4],
                                      return_neighbor_list=True, method="batch_cell_list"), (p, c, b, bi, bp),
                int(p.shape[0]), 4096, 0)

    desc, step, host, n, nsys, mwidth = make_workload(args.config, seed_shift=rank if args.config == 4 else 0)
    is_coo = mwidth == 0

    # ---- cold first call, then warm-up ----
    a0, b0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step(); torch.cuda.synchronize()                 # (library load, cudaFuncSetAttribute, allocator growth)
    _engine._pair_history.clear()
    flush.zero_()
    a0.record(); out = step(); b0.record(); torch.cuda.synchronize()
    first_call_ms = a0.elapsed_time(b0)
    for _ in range(max(warmup, 3)):
        out = step()
    torch.cuda.synchronize()
    P = int(out[0].shape[1]) if is_coo else int(out[1].sum().item())
    del out

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
    B = algorithmic_bytes(n, P, nsys, mwidth)
    call_gbs = B / (ms_per_step * 1e-3) / 1e9

    # ---- stage timings (C-ABI calls bracketed by events) in the configuration the API loop runs ----
    roofline = {"bound": "hbm", "achieved": call_gbs, "peak": peak, "unit": "GB/s", "frac": call_gbs / peak,
                "kernel": "whole neighbor_list call (all kernels, the size sync and the output allocation)",
                "traffic": None, "peak_source": peak_src, "algorithmic_bytes": B, "coo_path": nl_config.coo_path}
    if args.config == 4 and is_coo:
        pos_d, cell_d, pbc_d = [x.to(dev) for x in host]
        csq = _engine.cutoff_sq_in_dtype(CUTOFF, pos_d.dtype)
        rows_path = nl_config.coo_path == "rows"
        st = {"build": [], "sweep": [], "size_sync": [], "output": []}
        edge = torch.empty((2, P), dtype=torch.int32, device=dev)
        shf = torch.empty((P, 3), dtype=torch.int32, device=dev)
        for k in range(max(5, min(steps, 20))):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            flush.zero_()
            e[0].record(); h = _engine.build(pos_d, CUTOFF, cell_d, pbc_d)
            e[1].record(); num, ptr = _engine.count(h, csq, rows=rows_path, prezero=shf.view(-1) if rows_path else None)
            e[2].record(); hint = _engine.status(h)[4]
            assert not h.rows_overflow
            e[3].record(); _engine.fill_coo(h, csq, ptr, edge, shf, P, launch_hint=hint | (4 if rows_path else 0), rows=rows_path)
            e[4].record(); torch.cuda.synchronize()
            for name, (i, j) in zip(st, ((0, 1), (1, 2), (2, 3), (3, 4))):
                st[name].append(e[i].elapsed_time(e[j]))
        stage_ms = {k: float(np.median(v)) for k, v in st.items()}
        del edge, shf
        tests = 27.0 * n * n / max(1, _engine.status(h)[2])         # nominal: 27 stencil cells x mean cell population per atom
        out_bytes = 8 * P + 4 * P + 8 * n                          # edge_index written + temporary rows read + ptr/row_ref read
        sweep_tests_s = tests / (stage_ms["sweep"] * 1e-3)
        fp32_peak_tests = FP32_LANE_OPS_PEAK / LANE_OPS_PER_TEST
        longest = max(("build", "sweep", "output"), key=lambda k: stage_ms[k])
        roofline.update({
            "stages_ms": stage_ms, "longest_stage": longest,
            "stages": {
                "build": {"kernels": "k_init, k_bbox, k_grid, k_hash, k_scan, k_scatter", "bound": "latency (6 small launches)",
                          "ms": stage_ms["build"]},
                "sweep": {"kernels": "k_rows (+ fused zero-fill of the shifts output) + k_scan", "bound": "fp32 issue",
                          "ms": stage_ms["sweep"], "achieved": sweep_tests_s, "peak": fp32_peak_tests, "unit": "distance tests/s",
                          "frac": sweep_tests_s / fp32_peak_tests,
                          "note": f"{tests / n:.0f} nominal tests per atom; peak = {FP32_LANE_OPS_PEAK:.3g} measured fp32 lane-ops/s / "
                                  f"{LANE_OPS_PER_TEST:.0f} lane-ops per test; also writes 4 B/pair temporary rows and 12 B/pair zeros"},
                "output": {"kernels": "k_rows_out", "bound": "hbm", "ms": stage_ms["output"],
                           "achieved": out_bytes / (stage_ms["output"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                           "frac": out_bytes / (stage_ms["output"] * 1e-3) / 1e9 / peak,
                           "note": "bytes this kernel moves: 8 B/pair edge_index written + 4 B/pair temporary rows read "
                                   "(the 12 B/pair shifts were zero-filled by the sweep)"},
            },
            "first_call_ms": first_call_ms,
            "traffic_static": {"note": "dram__bytes per launch from the committed ncu capture, not from this run",
                               "source": "profiles/ncu_traffic.json"},
        })
        ncu_json = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(ncu_json):
            try:
                tj = json.load(open(ncu_json))
                roofline["traffic"] = tj.get("call_dram_bytes")
                roofline["traffic_static"].update({k: v for k, v in tj.items() if k != "call_dram_bytes"})
            except Exception:
                pass
        del pos_d, cell_d, pbc_d

    # ---- e2e: host buffers; H2D of the inputs and D2H of the full result inside the timed region ----
    pins_in = [x.pin_memory() for x in host]
    if is_coo:
        pins_out = [torch.empty((2, P), dtype=torch.int32).pin_memory(), torch.empty((n + 1,), dtype=torch.int32).pin_memory(),
                    torch.empty((P, 3), dtype=torch.int32).pin_memory()]
    else:
        pins_out = [torch.empty((n, mwidth), dtype=torch.int32).pin_memory(), torch.empty((n,), dtype=torch.int32).pin_memory(),
                    torch.empty((n, mwidth, 3), dtype=torch.int32).pin_memory()]

    def step_e2e():
        d = [x.to(dev, non_blocking=True) for x in pins_in]
        if args.config in (2, 4):
            res = neighbor_list(d[0], CUTOFF, cell=d[1], pbc=d[2], return_neighbor_list=is_coo)
        else:
            res = neighbor_list(d[0], CUTOFF, cell=d[1], pbc=d[2], batch_idx=d[3], batch_ptr=d[4], return_neighbor_list=True,
                                method="batch_cell_list")
        for dst, src in zip(pins_out, res):
            dst.copy_(src, non_blocking=True)

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
    h2d = int(sum(x.numel() * x.element_size() for x in pins_in))
    d2h = int(sum(x.numel() * x.element_size() for x in pins_out))
    e2e = {"value": world * P / (e2e_ms * 1e-3), "unit": "pairs/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "note": "pinned host inputs -> H2D -> neighbor_list -> D2H of the complete result (COO: edge_index, neighbor_ptr, "
                   "shifts); bound by the PCIe read-back of the result, and at N > 1 by the host memory the N read-backs share",
           "host_placement": numa_note}
    del pins_in, pins_out

    # ---- BASELINE configs 2, 3, 5 on one GPU (N = 1 only; the headline stays config 4) ----
    others = None
    if world == 1 and not args.no_other_configs and args.config == 4:
        others = {}
        for cfg in (2, 3, 5):
            d_, fn, _h, n_, ns_, mw_ = make_workload(cfg)
            for _ in range(4):
                o = fn()
            torch.cuda.synchronize()
            P_ = int(o[0].shape[1]) if mw_ == 0 else int(o[1].sum().item())
            del o
            ms, _ = time_api(fn, 12 if cfg != 5 else 6, flush)
            B_ = algorithmic_bytes(n_, P_, ns_, mw_)
            others[f"config{cfg}"] = {"workload": d_, "ms_per_call": ms, "pairs": P_, "pairs_per_s": P_ / (ms * 1e-3),
                                      "atoms_per_s": n_ / (ms * 1e-3), "algorithmic_bytes": B_,
                                      "roofline_frac": B_ / (ms * 1e-3) / 1e9 / peak}
            if cfg == 2:
                t2 = [x.to(dev) for x in _h]
                ms160, _ = time_api(lambda: neighbor_list(t2[0], CUTOFF, cell=t2[1], pbc=t2[2], max_neighbors=160), 12, flush)
                B160 = algorithmic_bytes(n_, P_, 0, 160)
                others["config2"]["max_neighbors_160"] = {"ms_per_call": ms160, "algorithmic_bytes": B160,
                                                          "roofline_frac": B160 / (ms160 * 1e-3) / 1e9 / peak}
            torch.cuda.empty_cache()

    # ---- config 5 sharded by batch_ptr, with and without the NCCL gather (every N; N = 1 anchors the curve) ----
    sharded = None
    if not args.no_sharded and args.config == 4:
        from nvalchemiops_b200.neighborlist.distributed import sharded_batch_neighbor_list

        bp, bc, bb, bi, bptr = bench_batch(4096, 1000, 1000, seed=5, mixed_pbc=False)
        bp, bc, bb, bptr = bp.to(dev), bc.to(dev), bb.to(dev), bptr.to(dev)
        res = {}
        for name, gather in (("kernels_only", False), ("with_gather", True)):
            for _ in range(3):
                o = sharded_batch_neighbor_list(bp, CUTOFF, bc, bb, bptr, gather=gather, return_stats=gather or world == 1)
            Pb = int(o[0].shape[1])
            stats = o[3] if isinstance(o[3], dict) else None
            check = None
            if gather or world == 1:
                # order-independent checksum of the gathered list: must agree on every rank (and with the N = 1 run)
                e_, s_ = o[0].long(), o[2].long()
                key = (e_[0] * 1000003 + e_[1]) * 27 + (s_[:, 0] + 1) * 9 + (s_[:, 1] + 1) * 3 + (s_[:, 2] + 1)
                check = int((key % 2147483647).sum().item()) ^ int(o[1].long().sum().item())
            del o
            barrier()
            reps = 5
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                o = sharded_batch_neighbor_list(bp, CUTOFF, bc, bb, bptr, gather=gather)
                del o
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / reps
            tot = Pb
            if world > 1:
                tt = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                ms = float(tt.item())
                if not gather:
                    tp = torch.tensor([Pb], device=dev, dtype=torch.int64)
                    dist.all_reduce(tp, op=dist.ReduceOp.SUM)
                    tot = int(tp.item())
            res[name] = {"ms_per_step": ms, "pairs_per_s": tot / (ms * 1e-3), "atoms_per_s": 4096 * 1000 / (ms * 1e-3), "pairs": tot}
            if check is not None:
                agree = True
                if world > 1:
                    lo = torch.tensor([check], device=dev, dtype=torch.int64)
                    hi = lo.clone()
                    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
                    agree = bool(lo.item() == hi.item())
                res[name].update({"checksum": check, "checksum_agrees_on_all_ranks": agree})
            if stats is not None and gather and world > 1:
                if "phase_ms" in stats:
                    res[name]["phase_ms_rank0_one_call"] = stats["phase_ms"]
                res[name].update({"peer_bytes_per_rank": stats["peer_bytes"], "packed_exchange": stats["packed"],
                                  "exchange_chunks_per_rank": stats.get("chunks"), "wire_bytes_per_pair": stats.get("bytes_per_pair", 20),
                                  "full_payload_bytes_per_rank": 20 * (tot - tot // world) + 4 * (4096000 - 4096000 // world)})
        if world > 1 and "with_gather" in res and "kernels_only" in res:
            t_gather = max(res["with_gather"]["ms_per_step"] - res["kernels_only"]["ms_per_step"], 1e-6)
            res["with_gather"]["gather_ms"] = t_gather
            res["with_gather"]["nvlink_frac"] = res["with_gather"].get("peer_bytes_per_rank", 0) / (t_gather * 1e-3) / 1e9 / NVLINK_GBS_PER_DIR
            res["with_gather"]["nvlink_note"] = ("bytes a rank receives from its peers / (with_gather - kernels_only time) / 900 GB/s per "
                                                 "direction; the exchange is wire_bytes_per_pair (target atom + packed shift) instead of 20 B/pair, "
                                                 "so the same result needs 4-5x fewer NVLink bytes than full_payload_bytes_per_rank; it runs in "
                                                 "exchange_chunks_per_rank chunks on a communication stream under the kernels of the "
                                                 "neighbouring chunks (phase_ms_rank0_one_call.nccl_chunks_on_comm_stream)")
        sharded = {"workload": "config 5: 4096 systems x 1000 atoms, periodic, COO, sharded by batch_ptr (strong scaling)", **res}
        del bp, bc, bb, bptr

    # ---- the sweep fused with a pair consumer (N = 1): config-4 box, real-space Ewald, list never written ----
    fused = None
    if world == 1 and not args.no_other_configs and args.config == 4:
        from nvalchemiops_b200.interactions.electrostatics import coulomb_energy_forces, fused_coulomb_energy_forces

        p4, c4, b4 = [x.to(dev) for x in bench_box(args.atoms, seed=4)]
        q4 = ((torch.rand(args.atoms, generator=torch.Generator().manual_seed(1), dtype=torch.float64) - 0.5) * 2).to(dev)

        def list_pipeline(alpha):
            nl_, ptr_, sh_ = neighbor_list(p4, CUTOFF, cell=c4, pbc=b4, return_neighbor_list=True)
            return coulomb_energy_forces(p4, q4, c4, CUTOFF, alpha, neighbor_list=nl_, neighbor_ptr=ptr_, neighbor_shifts=sh_)

        fused = {"workload": f"config-4 box ({args.atoms} atoms) + real-space Coulomb/Ewald energies and forces in fp64 "
                             "(reference coulomb.py:1540): fused = build + ONE sweep+consumer kernel, list = neighbor_list + consumer"}
        for alpha in (0.0, 0.3):
            for _ in range(2):
                of = fused_coulomb_energy_forces(p4, q4, c4, b4, CUTOFF, alpha, return_path=True)
                ol = list_pipeline(alpha)
            ms_f, _ = time_api(lambda: fused_coulomb_energy_forces(p4, q4, c4, b4, CUTOFF, alpha), 8, flush)
            ms_l, _ = time_api(lambda: list_pipeline(alpha), 8, flush)
            fused[f"alpha_{alpha}"] = {
                "fused_ms": ms_f, "list_pipeline_ms": ms_l, "path": of[2], "pairs_per_s_fused": P / (ms_f * 1e-3),
                "max_rel_diff_energy": float((of[0] - ol[0]).abs().max() / ol[0].abs().max()),
                "max_rel_diff_forces": float((of[1] - ol[1]).abs().max() / ol[1].abs().max())}
            del of, ol
        del p4, c4, b4, q4
        torch.cuda.empty_cache()

    # ---- cpu_baseline (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        nthreads = host_threads()
        r = cpu_reference_run(args.atoms, 4, args.cpu_budget_s, nthreads)
        cpu = {"value": r["pairs_per_s"], "unit": "pairs/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "atoms_per_s": r["atoms_per_s"]}

    if rank == 0:
        line = {
            "metric": METRIC if args.config == 4 else f"neighbor_pairs_per_s ({desc})", "value": value, "unit": "pairs/s",
            "n_gpus": world, "steps": steps, "warmup": max(warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "atoms_per_s": atoms_s, "first_call_ms": first_call_ms,
            "config": {"workload": desc + f"; {P} directed pairs" + ("; one replica per GPU" if world > 1 else ""),
                       "l2": "outputs (> L2) evict the inputs between steps + explicit 256 MB flush between timed steps",
                       "parallelism": "replicas" if world > 1 else "single GPU",
                       "steady_state": "timed steps reuse the previous call's pair count to size the shifts buffer the sweep "
                                       "zero-fills; first_call_ms is the cold call"},
            "e2e": e2e, "gpu_launches": int(launches), "gpu_launches_per_step": launches / steps,
            "clocks": clocks.summary(), "roofline": roofline,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if others is not None:
            line["other_configs"] = others
        if sharded is not None:
            line["sharded_batch"] = sharded
        if fused is not None:
            line["fused_consumer"] = fused
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
