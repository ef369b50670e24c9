"""Config 4 (1 M atoms, COO) through the public API, for ncu: python profiles/cfg4_calls.py [reps] [rows|masks] [atoms]"""
import os, sys, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + '/nvalchemi-toolkit-ops_b200', R + '/tests', R + '/oracle']
from systems import bench_box
from nvalchemiops_b200 import config
from nvalchemiops_b200.neighborlist import neighbor_list
rep = int(sys.argv[1]) if len(sys.argv) > 1 else 2
if len(sys.argv) > 2:
    config.coo_path = sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
pos, cell, pbc = [t.to('cuda:0') for t in bench_box(n, seed=4)]
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for k in range(rep):
    ev[0].record()
    out = neighbor_list(pos, 6.0, cell=cell, pbc=pbc, return_neighbor_list=True)
    ev[1].record()
    torch.cuda.synchronize()
    print(config.coo_path, n, 'pairs', out[0].shape[1], 'ms %.3f' % ev[0].elapsed_time(ev[1]))
    del out
