"""One BASELINE config through the public API, for ncu launch lists: python profiles/cfg_calls.py <2|3|4|5> [reps]"""
import os, sys, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + '/nvalchemi-toolkit-ops_b200', R + '/tests', R + '/oracle']
from systems import bench_batch, bench_box
from nvalchemiops_b200.neighborlist import neighbor_list
cfg = int(sys.argv[1]); rep = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = 'cuda:0'
if cfg in (2, 4):
    t = [x.to(dev) for x in bench_box(50_000 if cfg == 2 else 1_000_000, seed=cfg)]
    fn = (lambda: neighbor_list(t[0], 6.0, cell=t[1], pbc=t[2])) if cfg == 2 else \
         (lambda: neighbor_list(t[0], 6.0, cell=t[1], pbc=t[2], return_neighbor_list=True))
else:
    t = [x.to(dev) for x in (bench_batch(512, 150, 250, seed=3, mixed_pbc=True) if cfg == 3 else
                             bench_batch(4096, 1000, 1000, seed=5, mixed_pbc=False))]
    fn = lambda: neighbor_list(t[0], 6.0, cell=t[1], pbc=t[2], batch_idx=t[3], batch_ptr=t[4], return_neighbor_list=True,
                               method='batch_cell_list')
for k in range(rep):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = fn(); b.record(); torch.cuda.synchronize()
    print('cfg', cfg, 'call', k, 'ms %.3f' % a.elapsed_time(b), 'pairs', int(out[0].shape[1]) if cfg != 2 else int(out[1].sum()))
    del out
