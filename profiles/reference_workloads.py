"""The reference's own published benchmark workloads (benchmarks/neighborlist/benchmark_neighborlist.py,
benchmarks/systems.py:904-971): FCC lattice a = 4.0 A, rc = 5.0 A, full PBC, fp32, pre-allocated padded-matrix outputs
with max_neighbors = 192, median CUDA-event time of the whole neighbor_list(**inputs) call (10 warm-up / 100 timed).
Published numbers are H100; this script measures the same workloads on the local GPU."""
import json, os, sys, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + '/nvalchemi-toolkit-ops_b200', R + '/tests', R + '/oracle']
from nvalchemiops_b200.neighborlist import neighbor_list
dev = 'cuda:0'

def fcc(n_side, a=4.0):
    base = torch.tensor([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]], dtype=torch.float32)
    g = torch.stack(torch.meshgrid(*[torch.arange(n_side, dtype=torch.float32)] * 3, indexing='ij'), -1).reshape(-1, 1, 3)
    pos = ((g + base) * a).reshape(-1, 3)
    return pos, torch.eye(3).reshape(1, 3, 3) * (n_side * a)

def bench(fn, warm=10, n=100):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2]

H100 = {'cell_list 32768': 0.878, 'cell_list 131072': 6.713, 'cell_list 262144': 9.815, 'cell_list 524288': 18.440,
        'batch 512x256': 0.827, 'batch 512x1024(864)': 2.330, 'batch 512x6144(6912)': 13.061}
out = {}
M = 192
for n_side in (16, 32, 40, 51):          # 16384*? -> 4*n^3 atoms: 16384, 131072, 256000, 530604
    pos, cell = fcc(n_side)
    n = pos.shape[0]
    pos, cell = pos.to(dev), cell.to(dev); pbc = torch.ones(1, 3, dtype=torch.bool, device=dev)
    nm = torch.empty((n, M), dtype=torch.int32, device=dev); sh = torch.empty((n, M, 3), dtype=torch.int32, device=dev)
    num = torch.empty((n,), dtype=torch.int32, device=dev)
    f = lambda: neighbor_list(pos, 5.0, cell=cell, pbc=pbc, method='cell_list', neighbor_matrix=nm, neighbor_matrix_shifts=sh,
                              num_neighbors=num)
    t = bench(f)
    out[f'cell_list {n}'] = {'ms': t, 'atoms_per_s': n / t * 1e3, 'pairs': int(num.sum()), 'pairs_per_s': int(num.sum()) / t * 1e3,
                             'max_num': int(num.max())}
for n_side, S in ((4, 512), (6, 512), (12, 512)):     # 256, 864, 6912 atoms per system
    pos1, cell1 = fcc(n_side)
    n1 = pos1.shape[0]
    pos = pos1.repeat(S, 1).to(dev); cell = cell1.repeat(S, 1, 1).to(dev); pbc = torch.ones(S, 3, dtype=torch.bool, device=dev)
    bidx = torch.arange(S, dtype=torch.int32).repeat_interleave(n1).to(dev)
    bptr = (torch.arange(S + 1, dtype=torch.int32) * n1).to(dev)
    n = pos.shape[0]
    nm = torch.empty((n, M), dtype=torch.int32, device=dev); sh = torch.empty((n, M, 3), dtype=torch.int32, device=dev)
    num = torch.empty((n,), dtype=torch.int32, device=dev)
    f = lambda: neighbor_list(pos, 5.0, cell=cell, pbc=pbc, batch_idx=bidx, batch_ptr=bptr, method='batch_cell_list',
                              neighbor_matrix=nm, neighbor_matrix_shifts=sh, num_neighbors=num)
    t = bench(f, n=50)
    out[f'batch {S}x{n1}'] = {'ms': t, 'atoms': n, 'atoms_per_s': n / t * 1e3, 'pairs': int(num.sum()),
                              'pairs_per_s': int(num.sum()) / t * 1e3, 'max_num': int(num.max())}
out['_published_H100_ms'] = H100
print(json.dumps(out, indent=1))
