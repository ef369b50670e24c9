"""One neighbor_list call of a BASELINE config (for ncu launch lists): python profiles/one_call.py {3|5} [repeats] [rows|masks]"""
import os, sys, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + '/nvalchemi-toolkit-ops_b200', R + '/tests', R + '/oracle']
from systems import bench_batch
from nvalchemiops_b200.neighborlist import neighbor_list
from nvalchemiops_b200 import config
cfg = int(sys.argv[1]); rep = int(sys.argv[2]) if len(sys.argv) > 2 else 2
if len(sys.argv) > 3:
    config.coo_path = sys.argv[3]
if cfg == 3:
    a = bench_batch(512, 150, 250, seed=3, mixed_pbc=True)
else:
    a = bench_batch(4096, 1000, 1000, seed=5, mixed_pbc=False)
bp, bc, bb, bi, bptr = [t.to('cuda:0') for t in a]
for _ in range(rep):
    out = neighbor_list(bp, 6.0, cell=bc, pbc=bb, batch_idx=bi, batch_ptr=bptr, return_neighbor_list=True, method='batch_cell_list')
    torch.cuda.synchronize()
print(out[0].shape)
