"""Config 3 split by system class (which systems make the tail?): python profiles/cfg3_split.py <class> [reps]
classes: all | big (n >= 174: 2 cells per dim) | small3 / small2 / small1 (n < 174 with 3 / 2 / <= 1 periodic dims)"""
import os, sys, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + '/nvalchemi-toolkit-ops_b200', R + '/tests', R + '/oracle']
from systems import bench_batch
from nvalchemiops_b200.neighborlist import neighbor_list
cls = sys.argv[1]; rep = int(sys.argv[2]) if len(sys.argv) > 2 else 3
p, c, b, bi, bp = bench_batch(512, 150, 250, seed=3, mixed_pbc=True)
n = (bp[1:] - bp[:-1]).long()
nper = b.long().sum(1)
sel = {'all': n > 0, 'big': n >= 174, 'small3': (n < 174) & (nper == 3), 'small2': (n < 174) & (nper == 2),
       'small1': (n < 174) & (nper <= 1)}[cls]
ids = torch.nonzero(sel).flatten()
pos = torch.cat([p[bp[s]:bp[s + 1]] for s in ids.tolist()])
cnt = n[ids]
ptr = torch.zeros(len(ids) + 1, dtype=torch.int32); ptr[1:] = torch.cumsum(cnt, 0).to(torch.int32)
idx = torch.repeat_interleave(torch.arange(len(ids), dtype=torch.int32), cnt)
t = [x.to('cuda:0') for x in (pos, c[ids], b[ids], idx, ptr)]
for k in range(rep):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = neighbor_list(t[0], 6.0, cell=t[1], pbc=t[2], batch_idx=t[3], batch_ptr=t[4], return_neighbor_list=True, method='batch_cell_list')
    e.record(); torch.cuda.synchronize()
    print('cfg3', cls, 'systems', len(ids), 'atoms', pos.shape[0], 'call', k, 'ms %.3f' % a.elapsed_time(e), 'pairs', int(out[0].shape[1]))
    del out
