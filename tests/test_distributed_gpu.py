"""NCCL test of the batch-sharded path on real GPUs (needs >= 2 visible devices; `gpurun --gpus 2`): every rank's
gathered result must equal the single-GPU neighbor list of the whole batch — sorted (i, j, s) records and neighbor_ptr —
for the packed exchange (5 B/pair) and for the int32 fallback (a batch with a box smaller than the cutoff)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _paths():
    for p in (ROOT, os.path.join(ROOT, "nvalchemi-toolkit-ops_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)


def _records(e, s):
    rec = torch.cat([e.t().long(), s.long()], dim=1).cpu().numpy()
    return rec[np.lexsort(rec.T[::-1])]


def _worker(rank, world, port, out_dir):
    _paths()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from nvalchemiops_b200.neighborlist import neighbor_list
    from nvalchemiops_b200.neighborlist.distributed import sharded_batch_neighbor_list
    from systems import bench_batch

    results = {}
    for name, (ns, lo, hi, mixed, small_box) in {"packed": (37, 300, 900, True, False), "fallback": (6, 40, 60, False, True)}.items():
        pos, cell, pbc, bidx, bptr = bench_batch(ns, lo, hi, seed=5, mixed_pbc=mixed)
        if small_box:                                  # boxes of ~5 A with a 6 A cutoff: shifts of +-2 occur
            pos, cell = pos * 0.45, cell * 0.45
        pos, cell, pbc, bidx, bptr = pos.to(dev), cell.to(dev), pbc.to(dev), bidx.to(dev), bptr.to(dev)
        for rep in range(2):                           # second call: cached partition
            e, ptr, s, stats = sharded_batch_neighbor_list(pos, 6.0, cell, pbc, bptr, return_stats=True)
        assert stats["packed"] == (name == "packed"), (name, stats)
        e1, ptr1, s1 = neighbor_list(pos, 6.0, cell=cell, pbc=pbc, batch_idx=bidx, batch_ptr=bptr, return_neighbor_list=True,
                                     method="batch_cell_list")
        assert torch.equal(ptr, ptr1), name
        assert bool((e[0, 1:] >= e[0, :-1]).all()), name
        assert np.array_equal(_records(e, s), _records(e1, s1)), name
        results[name] = int(e.shape[1])
    # errors are raised consistently on every rank (no rank is left waiting in a collective)
    from nvalchemiops_b200.neighborlist import NeighborOverflowError
    try:
        sharded_batch_neighbor_list(pos, 6.0, cell, pbc, bptr, max_neighbors=1)
        raised = False
    except NeighborOverflowError:
        raised = True
    assert raised
    torch.save(results, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_nccl_sharded_batch_equals_single_gpu(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(torch.cuda.device_count(), 4)
    port = 29700 + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    assert all(r == res[0] for r in res) and res[0]["packed"] > 0 and res[0]["fallback"] > 0
