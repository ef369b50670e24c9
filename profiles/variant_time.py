"""Stage timings + a small parity check for a library variant: python profiles/variant_time.py <lib.so> [n_atoms]"""
import os, sys, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + '/nvalchemi-toolkit-ops_b200', R + '/tests', R + '/oracle']
from nvalchemiops_b200 import _lib
_lib._LIB_PATH = os.path.abspath(sys.argv[1])
from systems import bench_box, bench_batch
from nvalchemiops_b200.neighborlist import _engine, neighbor_list
from nvalchemiops_b200 import config
dev = 'cuda:0'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
# parity of the rows path against the two-pass path on a medium box and a small batch
def canon(out):
    ei, ptr, sh = out
    key = torch.stack([ei[0].long(), ei[1].long(), sh[:, 0].long() + 8, sh[:, 1].long() + 8, sh[:, 2].long() + 8])
    k = ((key[0] * 4000000 + key[1]) * 16 + key[2]) * 256 + key[3] * 16 + key[4]
    return torch.sort(k).values, ptr
for (p, c, b, kw) in [(*[t.to(dev) for t in bench_box(30_000, seed=7)], {}),]:
    res = {}
    for path in ('rows', 'masks'):
        config.coo_path = path
        res[path] = canon(neighbor_list(p, 6.0, cell=c, pbc=b, return_neighbor_list=True))
    ok = torch.equal(res['rows'][0], res['masks'][0]) and torch.equal(res['rows'][1], res['masks'][1])
    print('parity rows==masks (30k box):', ok)
b5 = [t.to(dev) for t in bench_batch(64, 1000, 1000, seed=5, mixed_pbc=False)]
res = {}
for path in ('rows', 'masks'):
    config.coo_path = path
    res[path] = canon(neighbor_list(b5[0], 6.0, cell=b5[1], pbc=b5[2], batch_idx=b5[3], batch_ptr=b5[4],
                                    return_neighbor_list=True, method='batch_cell_list'))
print('parity rows==masks (64x1000 batch):', torch.equal(res['rows'][0], res['masks'][0]) and torch.equal(res['rows'][1], res['masks'][1]))
config.coo_path = 'rows'
pos, cell, pbc = [t.to(dev) for t in bench_box(n, seed=4)]
csq = _engine.cutoff_sq_in_dtype(6.0, torch.float32)
h = _engine.build(pos, 6.0, cell, pbc)
num, ptr = _engine.count(h, csq, rows=True)
tot = _engine.status(h)
P = tot[0]
ei = torch.empty((2, P), dtype=torch.int32, device=dev)
sh = torch.empty((P, 3), dtype=torch.int32, device=dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for prezero in (False, True):
    res = []
    for it in range(10):
        flush.zero_()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(); h = _engine.build(pos, 6.0, cell, pbc)
        ev[1].record(); num, ptr = _engine.count(h, csq, rows=True, prezero=sh.view(-1) if prezero else None)
        ev[2].record(); _engine.fill_coo(h, csq, ptr, ei, sh, P, launch_hint=(4 if prezero else 0), rows=True)
        ev[3].record(); torch.cuda.synchronize()
        res.append([ev[k].elapsed_time(ev[k + 1]) for k in range(3)])
    res = res[3:]
    med = [sorted(r[k] for r in res)[len(res) // 2] for k in range(3)]
    print('%s n=%d prezero=%d  build %.3f  sweep+scan %.3f  output %.3f  total %.3f ms' % (os.path.basename(sys.argv[1]), n, prezero, med[0], med[1], med[2], sum(med)))
# config 5 on one GPU (4096 x 1000) through the API
b5 = [t.to(dev) for t in bench_batch(4096, 1000, 1000, seed=5, mixed_pbc=False)]
ts = []
for it in range(6):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = neighbor_list(b5[0], 6.0, cell=b5[1], pbc=b5[2], batch_idx=b5[3], batch_ptr=b5[4], return_neighbor_list=True, method='batch_cell_list'); b.record()
    torch.cuda.synchronize(); ts.append(a.elapsed_time(b)); del out
print('%s cfg5 api ms %.3f' % (os.path.basename(sys.argv[1]), sorted(ts)[len(ts) // 2]))
# config 5 without the fused zero-fill (every row of a 3-cell box carries shifts), and config 3
config.prezero_shifts = False
ts = []
for it in range(6):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = neighbor_list(b5[0], 6.0, cell=b5[1], pbc=b5[2], batch_idx=b5[3], batch_ptr=b5[4], return_neighbor_list=True, method='batch_cell_list'); b.record()
    torch.cuda.synchronize(); ts.append(a.elapsed_time(b)); del out
print('%s cfg5 api ms (prezero off) %.3f' % (os.path.basename(sys.argv[1]), sorted(ts)[len(ts) // 2]))
config.prezero_shifts = True
del b5
b3 = [t.to(dev) for t in bench_batch(512, 150, 250, seed=3, mixed_pbc=True)]
ts = []
for it in range(12):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = neighbor_list(b3[0], 6.0, cell=b3[1], pbc=b3[2], batch_idx=b3[3], batch_ptr=b3[4], return_neighbor_list=True, method='batch_cell_list'); b.record()
    torch.cuda.synchronize(); ts.append(a.elapsed_time(b)); del out
print('%s cfg3 api ms %.3f' % (os.path.basename(sys.argv[1]), sorted(ts)[len(ts) // 2]))
