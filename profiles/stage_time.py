"""Stage timings of the single-sweep COO path through the C ABI (CUDA events on torch's current stream):
build / sweep (k_rows + scan) / output (k_rows_out), in the configuration the API loop runs (shifts pre-zeroed by the
sweep) and without.  usage: python profiles/stage_time.py [n_atoms ...]"""
import os, sys, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + '/nvalchemi-toolkit-ops_b200', R + '/tests', R + '/oracle']
from systems import bench_box
from nvalchemiops_b200.neighborlist import _engine
dev = 'cuda:0'
sizes = [int(a) for a in sys.argv[1:]] or [1_000_000]
for n in sizes:
    pos, cell, pbc = [t.to(dev) for t in bench_box(n, seed=4)]
    csq = _engine.cutoff_sq_in_dtype(6.0, torch.float32)
    h = _engine.build(pos, 6.0, cell, pbc)
    num, ptr = _engine.count(h, csq, rows=True)
    tot = _engine.status(h)
    P = tot[0]
    print(n, 'status', tot, 'rows_overflow', h.rows_overflow)
    ei = torch.empty((2, P), dtype=torch.int32, device=dev)
    sh = torch.empty((P, 3), dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for prezero in (False, True):
        res = []
        for it in range(8):
            flush.zero_()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record(); h = _engine.build(pos, 6.0, cell, pbc)
            ev[1].record(); num, ptr = _engine.count(h, csq, rows=True, prezero=sh.view(-1) if prezero else None)
            ev[2].record(); _engine.fill_coo(h, csq, ptr, ei, sh, P, launch_hint=(4 if prezero else 0), rows=True)
            ev[3].record(); torch.cuda.synchronize()
            res.append([ev[k].elapsed_time(ev[k + 1]) for k in range(3)])
        res = res[2:]
        med = [sorted(r[k] for r in res)[len(res) // 2] for k in range(3)]
        B = 12 * n + 20 * P + 4 * (n + 1)
        t = sum(med) * 1e-3
        print('n=%d prezero=%d  build %.3f  sweep+scan %.3f  output %.3f  total %.3f ms  %.0f GB/s alg  pairs/s %.3e'
              % (n, prezero, med[0], med[1], med[2], t * 1e3, B / t / 1e9, P / t))
