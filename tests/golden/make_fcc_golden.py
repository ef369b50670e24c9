"""Golden vectors the reference itself publishes for this path: total neighbor counts of its own benchmark workload
(FCC lattice a = 4.0 A, r_cut = 5.0 A, full PBC, fp32 — benchmarks/neighborlist/benchmark_config.yaml:6-10,
benchmarks/systems.py:874-971) measured with the real Warp kernels on an H100
(docs/benchmarks/benchmark_results/neighbor_list_benchmark_cell-list_h100-80gb-hbm3.csv, column total_neighbors;
the batch_cell_list file gives the same per-system counts).  Run in the build container (reads /root/reference):

    python tests/golden/make_fcc_golden.py
"""
import csv
import json
import os

REF = "/root/reference/docs/benchmarks/benchmark_results"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_published_fcc.json")

rows = {}
with open(os.path.join(REF, "neighbor_list_benchmark_cell-list_h100-80gb-hbm3.csv")) as f:
    for r in csv.DictReader(f):
        if r["success"] == "True" and int(r["batch_size"]) == 1:
            rows[int(r["total_atoms"])] = int(r["total_neighbors"])
naive = {}
p = os.path.join(REF, "neighbor_list_benchmark_naive_h100-80gb-hbm3.csv")
if os.path.exists(p):
    with open(p) as f:
        for r in csv.DictReader(f):
            if r["success"] == "True" and int(r["batch_size"]) == 1:
                naive[int(r["total_atoms"])] = int(r["total_neighbors"])
json.dump({
    "source": "docs/benchmarks/benchmark_results/neighbor_list_benchmark_{cell-list,naive}_h100-80gb-hbm3.csv (reference v0.2.0)",
    "workload": {"lattice": "fcc", "lattice_constant": 4.0, "cutoff": 5.0, "dtype": "float32", "pbc": [True, True, True],
                 "generator": "benchmarks/systems.py:874-971 (create_crystal_system: first num_atoms sites of the "
                              "ceil((n/4)^(1/3))^3 supercell in i, j, k, basis order)"},
    "cell_list_total_neighbors": {str(k): v for k, v in sorted(rows.items())},
    "naive_total_neighbors": {str(k): v for k, v in sorted(naive.items())},
}, open(OUT, "w"), indent=1)
print(OUT, len(rows), len(naive))
