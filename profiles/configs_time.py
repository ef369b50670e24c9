"""Times BASELINE.json configs 2, 3, 5 (per-GPU shard) and a spatially coherent variant of config 4 through the public API."""
import os, sys, json, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + '/nvalchemi-toolkit-ops_b200', R + '/tests', R + '/oracle']
from systems import bench_box, bench_batch
from nvalchemiops_b200.neighborlist import neighbor_list, cell_list, batch_cell_list, _engine
dev = 'cuda:0'

def timeit(fn, n=20, warm=5):
    for _ in range(warm): out = fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); out = fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], out

res = {}
# config 2: 50k atoms, matrix output, default max_neighbors (1584) and 160, pre-allocated outputs
pos, cell, pbc = [t.to(dev) for t in bench_box(50_000, seed=2)]
for M in (1584, 160):
    nm = torch.empty((50_000, M), dtype=torch.int32, device=dev); sh = torch.empty((50_000, M, 3), dtype=torch.int32, device=dev)
    num = torch.empty((50_000,), dtype=torch.int32, device=dev)
    t, out = timeit(lambda: cell_list(pos, 6.0, cell, pbc, neighbor_matrix=nm, neighbor_matrix_shifts=sh, num_neighbors=num))
    B = 12 * 50_000 + 16 * 50_000 * M + 4 * 50_000
    res[f'config2_matrix_M{M}'] = {'ms': t, 'atoms_per_s': 50_000 / t * 1e3, 'pairs_per_s': int(num.sum()) / t * 1e3, 'alg_GBps': B / t / 1e6}
# config 3: 512 systems x ~200 atoms, mixed PBC, COO
bp, bc, bb, bi, bptr = [t.to(dev) for t in bench_batch(512, 150, 250, seed=3, mixed_pbc=True)]
t, out = timeit(lambda: neighbor_list(bp, 6.0, cell=bc, pbc=bb, batch_idx=bi, batch_ptr=bptr, return_neighbor_list=True, method='batch_cell_list'))
P = out[0].shape[1]
res['config3_batch512'] = {'ms': t, 'atoms': bp.shape[0], 'pairs': P, 'atoms_per_s': bp.shape[0] / t * 1e3, 'pairs_per_s': P / t * 1e3}
# config 5 on one GPU: 4096 x 1000 periodic
bp, bc, bb, bi, bptr = [t.to(dev) for t in bench_batch(4096, 1000, 1000, seed=5, mixed_pbc=False)]
t, out = timeit(lambda: neighbor_list(bp, 6.0, cell=bc, pbc=bb, batch_idx=bi, batch_ptr=bptr, return_neighbor_list=True, method='batch_cell_list'), n=5, warm=2)
P = out[0].shape[1]; del out
res['config5_1gpu'] = {'ms': t, 'atoms': bp.shape[0], 'pairs': P, 'atoms_per_s': bp.shape[0] / t * 1e3, 'pairs_per_s': P / t * 1e3,
                       'alg_GBps': (16 * bp.shape[0] + 20 * P) / t / 1e6}
del bp, bc, bb, bi, bptr
# config 4 with spatially coherent atom order (atoms pre-sorted by cell): rows land sequentially
pos, cell, pbc = bench_box(1_000_000, seed=4)
L = cell[0, 0, 0].item(); cpd = 35
c = torch.floor(pos / L * cpd).long().clamp(max=cpd - 1)
order = torch.argsort(c[:, 0] + cpd * (c[:, 1] + cpd * c[:, 2]), stable=True)
for name, p in (('random_order', pos), ('cell_sorted_order', pos[order])):
    p, cl, pb = p.to(dev), cell.to(dev), pbc.to(dev)
    csq = 36.0
    def stages():
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record(); h = _engine.build(p, 6.0, cl, pb); e[1].record(); n_, ptr = _engine.count(h, csq); e[2].record()
        tot = _engine.status(h)[0]
        ei = torch.empty((2, tot), dtype=torch.int32, device=dev); shf = torch.empty((tot, 3), dtype=torch.int32, device=dev)
        e[3] = torch.cuda.Event(enable_timing=True); e[3].record(); _engine.fill_coo(h, csq, ptr, ei, shf, tot)
        e4 = torch.cuda.Event(enable_timing=True); e4.record(); torch.cuda.synchronize()
        return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[3].elapsed_time(e4)
    for _ in range(3): s = stages()
    ss = [stages() for _ in range(7)]
    med = [sorted(x[k] for x in ss)[3] for k in range(3)]
    res[f'config4_{name}'] = {'build_ms': med[0], 'count_ms': med[1], 'fill_ms': med[2], 'sum_ms': sum(med)}
print(json.dumps(res, indent=1))
