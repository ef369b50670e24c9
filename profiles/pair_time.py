"""Fused sweep + Coulomb consumer vs list pipeline (neighbor_list -> consumer), config-4 box: python profiles/pair_time.py [n] [alpha]"""
import os, sys, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + '/nvalchemi-toolkit-ops_b200', R + '/tests', R + '/oracle']
from systems import bench_box
from nvalchemiops_b200.neighborlist import neighbor_list
from nvalchemiops_b200.interactions.electrostatics import coulomb_energy_forces, fused_coulomb_energy_forces
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
alpha = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
dev = 'cuda:0'
pos, cell, pbc = [t.to(dev) for t in bench_box(n, seed=4)]
q = ((torch.rand(n, generator=torch.Generator().manual_seed(1), dtype=torch.float64) - 0.5) * 2).to(dev)
def fused():
    return fused_coulomb_energy_forces(pos, q, cell, pbc, 6.0, alpha, return_path=True)
def listed():
    nl, ptr, sh = neighbor_list(pos, 6.0, cell=cell, pbc=pbc, return_neighbor_list=True)
    return coulomb_energy_forces(pos, q, cell, 6.0, alpha, neighbor_list=nl, neighbor_ptr=ptr, neighbor_shifts=sh)
def timeit(fn, reps=8):
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2], out
tf, of = timeit(fused)
tl, ol = timeit(listed)
err_e = float((of[0] - ol[0]).abs().max() / ol[0].abs().max()); err_f = float((of[1] - ol[1]).abs().max() / ol[1].abs().max())
print('n=%d alpha=%.2f  fused %.3f ms (path %s)  list pipeline %.3f ms  speed-up %.2fx  max rel diff E %.2e F %.2e' % (n, alpha, tf, of[2], tl, tl / tf, err_e, err_f))
