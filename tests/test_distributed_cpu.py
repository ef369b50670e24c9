"""world_size-2/3 gloo tests (CPU) of the batch-sharding logic: partitioning, global index offsets, the size
exchange, in-place writes of a rank's own range, the variable-size broadcasts and the expansion of the foreign ranges.  The per-rank neighbor lists come from the oracle
through the module's test hook — the collective plumbing is what is under test here; the CUDA local path is
covered by the -m gpu tests."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _OracleShard:
    """Test hook for ``sharded_batch_neighbor_list(_local_shard=...)``: a rank's own systems computed by the oracle."""

    def __init__(self, positions, cutoff, cell, pbc, batch_idx, batch_ptr, half_fill, index_offset):
        import reference_oracle as ro

        e, p, s = ro.batch_cell_list(positions, cutoff, cell, pbc, batch_idx, max_neighbors=2048, half_fill=half_fill,
                                     return_neighbor_list=True)
        self.e, self.s, self.off = e, np.ascontiguousarray(s), index_offset
        self.num = torch.from_numpy(np.diff(p).astype(np.int32))
        self.total = int(e.shape[1])
        self.max_count = int(self.num.max()) if self.num.numel() else 0
        self.err = 0
        self.packable = bool(np.abs(self.s).max() <= 1) if self.total else True

    def fill(self, edge_rows, shifts, row_stride):
        t = self.total
        edge_rows[:t] = torch.from_numpy(self.e[0]) + self.off
        edge_rows[row_stride:row_stride + t] = torch.from_numpy(self.e[1]) + self.off
        shifts[:t] = torch.from_numpy(self.s)


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "nvalchemi-toolkit-ops_b200"), os.path.join(ROOT, "oracle"),
              os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nvalchemiops_b200.neighborlist.distributed import sharded_batch_neighbor_list
    from systems import bench_batch

    pos, cell, pbc, bidx, bptr = bench_batch(9, 60, 140, seed=5, mixed_pbc=True)
    e, ptr, s = sharded_batch_neighbor_list(pos, 6.0, cell, pbc, bptr, _local_shard=_OracleShard)
    torch.save({"e": e, "ptr": ptr, "s": s}, os.path.join(out_dir, f"r{rank}.pt"))
    # the chunked exchange (config.exchange_chunks = 2 above): any number of chunks per rank gives the same arrays,
    # also when some chunks of some ranks are empty (9 systems over 3 ranks x 4 chunks)
    import reference_oracle as ro
    from nvalchemiops_b200 import config
    for chunks, word in ((1, True), (3, True), (4, True), (2, False), (3, False)):
        # word: target atom and packed shift in ONE int32 per pair (4 B); otherwise targets + a byte array (5 B)
        config.exchange_word = word
        e2, ptr2, s2, st = sharded_batch_neighbor_list(pos, 6.0, cell, pbc, bptr, _local_shard=_OracleShard, chunks=chunks,
                                                       return_stats=True)
        assert st["packed"] and st["chunks"] == chunks and st["bytes_per_pair"] == (4 if word else 5)
        assert torch.equal(ptr2, ptr), chunks
        assert np.array_equal(ro.records_from_coo(e2, s2), ro.records_from_coo(e, s)), chunks
        assert bool((e2[0, 1:] >= e2[0, :-1]).all()), chunks
    config.exchange_word = True
    shard = sharded_batch_neighbor_list(pos, 6.0, cell, pbc, bptr, gather=False, _local_shard=_OracleShard)
    torch.save({"e": shard[0], "range": shard[3]}, os.path.join(out_dir, f"s{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_batch_matches_single_process(tmp_path, world):
    import reference_oracle as ro
    from systems import bench_batch

    port = 29600 + world + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    pos, cell, pbc, bidx, bptr = bench_batch(9, 60, 140, seed=5, mixed_pbc=True)
    e, p, s = ro.batch_cell_list(pos, 6.0, cell, pbc, bidx, max_neighbors=2048, return_neighbor_list=True)
    want = ro.records_from_coo(e, s)
    covered = []
    for r in range(world):
        got = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert np.array_equal(ro.records_from_coo(got["e"], got["s"]), want), f"rank {r}"
        assert np.array_equal(got["ptr"].numpy(), p)
        assert (np.diff(got["e"][0].numpy()) >= 0).all(), "global COO stays sorted by source atom"
        sh = torch.load(os.path.join(tmp_path, f"s{r}.pt"))
        lo, hi = sh["range"]
        src = sh["e"][0].numpy()
        assert ((src >= lo) & (src < hi)).all()
        covered.append((lo, hi))
    covered.sort()
    assert covered[0][0] == 0 and covered[-1][1] == pos.shape[0]
    assert all(covered[k][1] == covered[k + 1][0] for k in range(world - 1))


def test_chunk_partition_covers_every_rank_range():
    """_partition: every rank's systems split into contiguous chunks that cover exactly the rank's range."""
    from nvalchemiops_b200.neighborlist.distributed import _partition

    bptr = torch.tensor([0, 7, 19, 19, 40, 41, 90, 120, 121, 200], dtype=torch.int32)
    for world in (1, 2, 3, 5):
        for chunks in (1, 2, 4):
            parts, atoms, n, sub, sub_atoms = _partition(bptr, world, chunks)
            assert n == 200 and len(sub) == world and all(len(row) == chunks for row in sub)
            for g in range(world):
                assert sub[g][0][0] == parts[g][0] and sub[g][-1][1] == parts[g][1]
                assert all(sub[g][k][1] == sub[g][k + 1][0] for k in range(chunks - 1))
                assert sub_atoms[g][0][0] == atoms[g][0] and sub_atoms[g][-1][1] == atoms[g][1]
                assert all(a <= b for a, b in sub_atoms[g])


def test_partition_balances_atoms():
    from nvalchemiops_b200.neighborlist.distributed import partition_systems

    ptr = [0, 10, 20, 30, 40, 50, 60, 70, 80]
    assert partition_systems(ptr, 2) == [(0, 4), (4, 8)]
    assert partition_systems(ptr, 8) == [(k, k + 1) for k in range(8)]
    parts = partition_systems([0, 1000, 1010, 1020, 1030], 2)
    assert parts[0][0] == 0 and parts[-1][1] == 4 and all(a <= b for a, b in parts)
    parts = partition_systems([0, 5, 10], 4)  # more ranks than systems: some ranks get nothing
    assert parts[0][0] == 0 and parts[-1][1] == 2 and sum(b - a for a, b in parts) == 2
