"""Back-to-back config-4 COO calls without host syncs in between (the bench.py loop), prezero on/off, flush on/off."""
import os, sys, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + '/nvalchemi-toolkit-ops_b200', R + '/tests', R + '/oracle']
from systems import bench_box
from nvalchemiops_b200 import config
from nvalchemiops_b200.neighborlist import neighbor_list
dev = torch.device('cuda:0')
pos, cell, pbc = [t.to(dev) for t in bench_box(1_000_000, seed=4)]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for prezero in (False, True):
    for do_flush in (False, True):
        for keep in (False, True):
            config.prezero_shifts = prezero
            for _ in range(4):
                out = neighbor_list(pos, 6.0, cell=cell, pbc=pbc, return_neighbor_list=True)
            del out
            torch.cuda.synchronize()
            K = 20
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
            for k in range(K):
                if do_flush:
                    flush.zero_()
                ev[k][0].record()
                out = neighbor_list(pos, 6.0, cell=cell, pbc=pbc, return_neighbor_list=True)
                ev[k][1].record()
                if not keep:
                    del out
            torch.cuda.synchronize()
            ts = sorted(a.elapsed_time(b) for a, b in ev)
            print('prezero=%d flush=%d keep_out=%d  median %.3f  min %.3f  max %.3f ms' % (prezero, do_flush, keep, ts[K // 2], ts[0], ts[-1]))
