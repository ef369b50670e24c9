"""Config 5 (4096 x 1000 atoms) sharded over the ranks of a torchrun job: time of sharded_batch_neighbor_list per number of
exchange chunks, phase times of one call, and an order-independent checksum that must agree for every chunk count and rank.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 profiles/sharded_time.py [chunks ...]
"""
import json, os, sys, torch, torch.distributed as dist
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + '/nvalchemi-toolkit-ops_b200', R + '/tests']
from systems import bench_batch
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
dev = torch.device('cuda', torch.cuda.current_device())
dist.init_process_group('nccl', device_id=dev)
from nvalchemiops_b200.neighborlist.distributed import sharded_batch_neighbor_list
ns = int(os.environ.get('NS', 4096))
bp, bc, bb, bi, bptr = [t.to(dev) for t in bench_batch(ns, 1000, 1000, seed=5, mixed_pbc=False)]
out = {}
for chunks in [int(a) for a in sys.argv[1:]] or [1, 2, 4]:
    for _ in range(3):
        o = sharded_batch_neighbor_list(bp, 6.0, bc, bb, bptr, return_stats=True, chunks=chunks)
    e_, s_ = o[0].long(), o[2].long()
    key = (e_[0] * 1000003 + e_[1]) * 27 + (s_[:, 0] + 1) * 9 + (s_[:, 1] + 1) * 3 + (s_[:, 2] + 1)
    check = int((key % 2147483647).sum().item()) ^ int(o[1].long().sum().item())
    sorted_ok = bool((o[0][0, 1:] >= o[0][0, :-1]).all())
    stats = o[3]
    del o, e_, s_, key
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 6
    a.record()
    for _ in range(reps):
        o = sharded_batch_neighbor_list(bp, 6.0, bc, bb, bptr, chunks=chunks)
        del o
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    lo = torch.tensor([check], device=dev, dtype=torch.int64); hi = lo.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    out[chunks] = {'ms_per_call_max_over_ranks': float(t.item()), 'checksum': check, 'agree': bool(lo.item() == hi.item()),
                   'sorted': sorted_ok, 'phase_ms_rank0': stats.get('phase_ms'), 'peer_bytes': stats['peer_bytes']}
if rank == 0:
    print(json.dumps({'world': world, 'systems': ns, 'by_chunks': out}))
    cs = {v['checksum'] for v in out.values()}
    print('checksums equal across chunk counts:', len(cs) == 1, 'agree on all ranks:', all(v['agree'] for v in out.values()))
dist.barrier()
dist.destroy_process_group()
