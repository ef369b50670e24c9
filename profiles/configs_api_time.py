"""BASELINE configs 2, 3, 5 (+ config 4 unwrapped / cell-sorted) through the public API, both COO paths.
usage: python profiles/configs_api_time.py [out.json]"""
import json, os, sys, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + '/nvalchemi-toolkit-ops_b200', R + '/tests', R + '/oracle']
from systems import bench_batch, bench_box
from nvalchemiops_b200 import config
from nvalchemiops_b200.neighborlist import neighbor_list
dev = 'cuda:0'
res = {}

def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], out

p2 = [t.to(dev) for t in bench_box(50_000, seed=2)]
b3 = [t.to(dev) for t in bench_batch(512, 150, 250, seed=3, mixed_pbc=True)]
b5 = [t.to(dev) for t in bench_batch(4096, 1000, 1000, seed=5, mixed_pbc=False)]
p4 = [t.to(dev) for t in bench_box(1_000_000, seed=4)]
p4u = p4[0] + (torch.arange(1_000_000, device=dev) % 2).float()[:, None] * p4[1][0, 0, 0]
for path in ('rows', 'masks'):
    config.coo_path = path
    t, o = timeit(lambda: neighbor_list(p2[0], 6.0, cell=p2[1], pbc=p2[2], return_neighbor_list=True))
    res[f'cfg2_50k_coo_{path}_ms'] = t
    t, o = timeit(lambda: neighbor_list(b3[0], 6.0, cell=b3[1], pbc=b3[2], batch_idx=b3[3], batch_ptr=b3[4],
                                        return_neighbor_list=True, method='batch_cell_list'))
    res[f'cfg3_512x200_coo_{path}_ms'] = t; res['cfg3_pairs'] = int(o[0].shape[1])
    t, o = timeit(lambda: neighbor_list(b5[0], 6.0, cell=b5[1], pbc=b5[2], batch_idx=b5[3], batch_ptr=b5[4],
                                        return_neighbor_list=True, method='batch_cell_list'), reps=8)
    res[f'cfg5_4096x1000_coo_{path}_ms'] = t; res['cfg5_pairs'] = int(o[0].shape[1])
    del o
    t, o = timeit(lambda: neighbor_list(p4[0], 6.0, cell=p4[1], pbc=p4[2], return_neighbor_list=True), reps=10)
    res[f'cfg4_1m_coo_{path}_ms'] = t
    del o
    t, o = timeit(lambda: neighbor_list(p4u, 6.0, cell=p4[1], pbc=p4[2], return_neighbor_list=True), reps=6)
    res[f'cfg4_1m_unwrapped_coo_{path}_ms'] = t
    del o
t, o = timeit(lambda: neighbor_list(p2[0], 6.0, cell=p2[1], pbc=p2[2], max_neighbors=160))
res['cfg2_50k_matrix_M160_ms'] = t
t, o = timeit(lambda: neighbor_list(p2[0], 6.0, cell=p2[1], pbc=p2[2]))
res['cfg2_50k_matrix_M1584_ms'] = t
print(json.dumps(res, indent=1))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], 'w'), indent=1)
