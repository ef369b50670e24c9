"""Workloads that exercise the GENERAL sweep kernel: config 3 (many small boxes, 23 % of them smaller than 2 rc) and the
config-4 box with unwrapped coordinates (half of the atoms displaced by one lattice vector)."""
import os, sys, json, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + '/nvalchemi-toolkit-ops_b200', R + '/tests', R + '/oracle']
from systems import bench_box, bench_batch
from nvalchemiops_b200.neighborlist import neighbor_list
dev = 'cuda:0'
def timeit(fn, n=15, warm=4):
    for _ in range(warm): out = fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); out = fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], out
res = {}
bp, bc, bb, bi, bptr = [t.to(dev) for t in bench_batch(512, 150, 250, seed=3, mixed_pbc=True)]
t, out = timeit(lambda: neighbor_list(bp, 6.0, cell=bc, pbc=bb, batch_idx=bi, batch_ptr=bptr, return_neighbor_list=True, method='batch_cell_list'))
res['config3_batch512'] = {'ms': t, 'atoms': bp.shape[0], 'pairs': out[0].shape[1]}
# tiny periodic crystals: 2048 systems x 32 atoms in 8 A boxes (box < 2 rc), fully periodic
g = torch.Generator().manual_seed(9)
S, n1, L = 2048, 32, 8.0
pos = (torch.rand(S * n1, 3, generator=g) * L).to(dev)
cell = (torch.eye(3) * L).repeat(S, 1, 1).to(dev); pbc = torch.ones(S, 3, dtype=torch.bool, device=dev)
bidx = torch.arange(S, dtype=torch.int32).repeat_interleave(n1).to(dev); bp2 = (torch.arange(S + 1, dtype=torch.int32) * n1).to(dev)
t, out = timeit(lambda: neighbor_list(pos, 6.0, cell=cell, pbc=pbc, batch_idx=bidx, batch_ptr=bp2, return_neighbor_list=True,
                                      method='batch_cell_list', max_neighbors=4096))
res['tiny_boxes_2048x32_L8'] = {'ms': t, 'atoms': S * n1, 'pairs': out[0].shape[1]}
pos, cell, pbc = bench_box(1_000_000, seed=4)
Lb = cell[0, 0, 0].item()
pos = pos.clone(); pos[::2, 0] += Lb
pos, cell, pbc = pos.to(dev), cell.to(dev), pbc.to(dev)
t, out = timeit(lambda: neighbor_list(pos, 6.0, cell=cell, pbc=pbc, return_neighbor_list=True), n=7, warm=2)
res['config4_unwrapped'] = {'ms': t, 'pairs': out[0].shape[1]}
print(json.dumps(res))
