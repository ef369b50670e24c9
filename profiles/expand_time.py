"""Time of the re-assembly kernel of the padded exchange on ONE GPU (config 5, emulating rank 0 of `world` ranks):
python profiles/expand_time.py [world]"""
import ctypes, os, sys, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + '/nvalchemi-toolkit-ops_b200', R + '/tests', R + '/oracle']
from systems import bench_batch
from nvalchemiops_b200 import _lib
from nvalchemiops_b200.neighborlist import neighbor_list
from nvalchemiops_b200.neighborlist.distributed import partition_systems
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = 'cuda:0'
p, c, b, bi, bp = bench_batch(4096, 1000, 1000, seed=5, mixed_pbc=False)
t = [x.to(dev) for x in (p, c, b, bi, bp)]
e, ptr, s = neighbor_list(t[0], 6.0, cell=t[1], pbc=t[2], batch_idx=t[3], batch_ptr=t[4], return_neighbor_list=True, method='batch_cell_list')
N, P = p.shape[0], e.shape[1]
parts = partition_systems(bp.tolist(), world)
ab = [int(bp[a]) for a, _ in parts] + [N]
offs = [int(ptr[a].item()) for a in ab]
pmax = max(offs[g + 1] - offs[g] for g in range(world))
packed = ((s[:, 0] + 1) | ((s[:, 1] + 1) << 2) | ((s[:, 2] + 1) << 4)).to(torch.uint8)
t_dst = torch.zeros(world * pmax, dtype=torch.int32, device=dev); t_pk = torch.zeros(world * pmax, dtype=torch.uint8, device=dev)
for g in range(world):
    t_dst[g * pmax: g * pmax + offs[g + 1] - offs[g]] = e[1, offs[g]:offs[g + 1]]
    t_pk[g * pmax: g * pmax + offs[g + 1] - offs[g]] = packed[offs[g]:offs[g + 1]]
e2 = e.clone(); s2 = s.clone(); e2[1] = -1; e2[0, offs[1]:] = -1; s2[offs[1]:] = -7
A = (ctypes.c_int64 * (world + 1))(*ab); Bp = (ctypes.c_int64 * (world + 1))(*offs)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
ts = []
for it in range(6):
    a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a_.record()
    _lib.check(_lib.lib().nvnl_expand_padded(ctypes.c_void_p(ptr.data_ptr()), N, world, 0, A, Bp, pmax, ctypes.c_void_p(t_dst.data_ptr()),
                                             ctypes.c_void_p(t_pk.data_ptr()), ctypes.c_void_p(e2[0].data_ptr()), ctypes.c_void_p(e2[1].data_ptr()),
                                             ctypes.c_void_p(s2.data_ptr()), st), 'expand')
    b_.record(); torch.cuda.synchronize(); ts.append(a_.elapsed_time(b_))
ok = torch.equal(e2, e) and torch.equal(s2, s)
bytes_moved = 4 * P * 2 + (P - offs[1]) * (4 + 12 + 1) + 8 * N
print('world %d: nvnl_expand_padded %.3f ms (median of %d), correct %s, %.2f TB/s of useful traffic' % (world, sorted(ts)[len(ts) // 2], len(ts), ok, bytes_moved / sorted(ts)[len(ts) // 2] / 1e9))
