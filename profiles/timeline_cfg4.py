"""Event timeline of one rows-path COO query at config 4 (where do the microseconds between the kernels go?)."""
import ctypes, os, sys, time, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + '/nvalchemi-toolkit-ops_b200', R + '/tests', R + '/oracle']
from systems import bench_box
from nvalchemiops_b200 import _lib, config
from nvalchemiops_b200.neighborlist import _engine, neighbor_list
dev = torch.device('cuda:0')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
pos, cell, pbc = [t.to(dev) for t in bench_box(n, seed=4)]
for _ in range(3):
    out = neighbor_list(pos, 6.0, cell=cell, pbc=pbc, return_neighbor_list=True)
P = out[0].shape[1]; del out
torch.cuda.synchronize()
csq = _engine.cutoff_sq_in_dtype(6.0, pos.dtype)
E = lambda: torch.cuda.Event(enable_timing=True)
for rep in range(4):
    for prezero in (False, True):
        ev = {k: E() for k in ('start', 'built', 'counted', 'synced', 'filled')}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev['start'].record()
        h = _engine.build(pos, 6.0, cell, pbc)
        ev['built'].record()
        zbuf = torch.empty(3 * (P + 1000), dtype=torch.int32, device=dev) if prezero else None
        num, ptr = _engine.count(h, csq, rows=True, prezero=zbuf)
        ev['counted'].record()
        t1 = time.perf_counter()
        total, mc, _c, err, hint = _engine.status(h)
        t2 = time.perf_counter()
        ev['synced'].record()
        if prezero:
            edge = torch.empty((2, total), dtype=torch.int32, device=dev); shf = zbuf[:3 * total].view(total, 3)
            hint |= 4
        else:
            buf = torch.empty(5 * total, dtype=torch.int32, device=dev); edge = buf[:2 * total].view(2, total); shf = buf[2 * total:].view(total, 3)
        _engine.fill_coo(h, csq, ptr, edge, shf, total, launch_hint=hint, rows=True)
        ev['filled'].record()
        t3 = time.perf_counter()
        torch.cuda.synchronize()
        s = ev['start']
        line = ' '.join('%s=%.3f' % (k, s.elapsed_time(ev[k])) for k in ('built', 'counted', 'synced', 'filled'))
        print('prezero=%d  %s | host: launch %.3f sync-wait %.3f post %.3f ms' % (prezero, line, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
