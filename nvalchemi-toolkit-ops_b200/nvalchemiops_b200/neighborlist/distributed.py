"""Batch-sharded multi-GPU neighbor lists: one process per GPU, ONE payload all-gather.

The reference has no distributed code (no NCCL / torch.distributed call sites); this module is
the multi-GPU step BASELINE.json's north_star defines: independent systems are split by ``batch_ptr`` across the
ranks of one node, every rank builds the COO list of its own systems with GLOBAL atom indices, and a single
``all_gather_into_tensor`` over NCCL (NVLink 5 / NVSwitch) re-assembles the global ``edge_index`` / ``shifts`` /
``neighbor_ptr`` on every rank.  Semantics: identical to running the whole batch on one GPU, up to the order of
entries inside a source atom's row (which the reference leaves unspecified).

Wire format of a rank's block (int32, all ranks padded to the same length):
    [ src (Pmax) | dst (Pmax) | shifts (3*Pmax) | num_neighbors (Nmax) ]
The fill kernel writes src/dst/shifts straight into the block (no pack pass); ``nvnl_unpack_gathered`` writes the
global arrays from the gathered blocks (one kernel).
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from .neighbor_utils import NeighborOverflowError


def partition_systems(batch_ptr_host, world_size: int):
    """Contiguous split of the systems into ``world_size`` chunks balancing the atom counts.
    ``batch_ptr_host``: sequence of S+1 ints.  Returns a list of (s0, s1) system ranges."""
    ptr = [int(x) for x in batch_ptr_host]
    S = len(ptr) - 1
    n = ptr[-1]
    bounds = [0]
    s = 0
    for r in range(1, world_size):
        target = n * r / world_size
        while s < S and ptr[s + 1] <= target:
            s += 1
        # choose the closer boundary
        if s < S and (target - ptr[s]) > (ptr[s + 1] - target):
            s += 1
        s = max(s, bounds[-1])
        bounds.append(min(s, S))
    bounds.append(S)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def _default_local_coo(positions, cutoff, cell, pbc, batch_idx, batch_ptr, half_fill, index_offset, block_builder):
    """CUDA path: count -> (sizes exchanged by the caller) -> fill straight into the send block."""
    from . import _engine

    h = _engine.build(positions, cutoff, cell, pbc, batch_idx=batch_idx, batch_ptr=batch_ptr)
    csq = _engine.cutoff_sq_in_dtype(cutoff, positions.dtype)
    num, ptr, total, max_count, err, hint, rows = _engine.count_and_size(h, csq, half_fill)
    _engine._raise_on_error_bits(err)

    def fill(block, pmax):
        if total > 0:
            _engine.fill_coo(h, csq, ptr, block[: 2 * pmax], block[2 * pmax: 5 * pmax], pmax, half_fill, index_offset,
                             launch_hint=hint, rows=rows)

    return num, total, max_count, fill


def _torch_unpack(recv, world, pmax, nmax, counts, natoms, total, device):
    """Reference re-assembly with torch slicing (CPU/gloo tests; the CUDA path uses nvnl_unpack_gathered)."""
    blk = 5 * pmax + nmax
    edge = torch.empty((2, total), dtype=torch.int32, device=device)
    shifts = torch.empty((total, 3), dtype=torch.int32, device=device)
    o = 0
    for g in range(world):
        b = recv[g * blk:(g + 1) * blk]
        c = counts[g]
        edge[0, o:o + c] = b[:c]
        edge[1, o:o + c] = b[pmax:pmax + c]
        shifts[o:o + c] = b[2 * pmax:2 * pmax + 3 * c].reshape(c, 3)
        o += c
    return edge, shifts


def sharded_batch_neighbor_list(positions, cutoff, cell, pbc, batch_ptr, half_fill=False, max_neighbors=None,
                                group=None, gather=True, _local_coo=None):
    """COO neighbor list of a batch, sharded over the ranks of ``group`` (default: WORLD).

    Every rank passes the SAME global tensors (on its own device): ``positions`` [N,3], ``cell`` [S,3,3],
    ``pbc`` [S,3], ``batch_ptr`` [S+1] (atoms of a system contiguous).  Returns on every rank
    ``(neighbor_list [2,P] int32, neighbor_ptr [N+1] int32, shifts [P,3] int32)`` with global atom indices.
    With ``gather=False`` the collective is skipped and the rank's own shard is returned as
    ``(neighbor_list, neighbor_ptr_local, shifts, (atom_lo, atom_hi))`` — the "kernels only" figure of the bench.

    ``_local_coo`` is a test hook (CPU/gloo tests inject the oracle); the product path leaves it None.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = positions.device
    ptr_host = batch_ptr.detach().cpu().tolist()
    S = len(ptr_host) - 1
    N = positions.shape[0]
    s0, s1 = partition_systems(ptr_host, world)[rank]
    a0, a1 = ptr_host[s0], ptr_host[s1]
    n_loc = a1 - a0
    local_fn = _local_coo or _default_local_coo
    if n_loc > 0:
        lptr = (batch_ptr[s0:s1 + 1] - a0).to(torch.int32)
        lidx = torch.repeat_interleave(torch.arange(s1 - s0, dtype=torch.int32, device=dev),
                                       (lptr[1:] - lptr[:-1]).long())
        num, total, max_count, fill = local_fn(positions[a0:a1], cutoff, cell[s0:s1], pbc[s0:s1], lidx, lptr,
                                               half_fill, a0, None)
    else:
        num = torch.zeros(0, dtype=torch.int32, device=dev)
        total, max_count = 0, 0

        def fill(block, pmax):
            return None

    if max_neighbors is not None and max_count > max_neighbors:
        raise NeighborOverflowError(max_neighbors, max_count)

    if not gather or world == 1:
        block = torch.empty(5 * max(total, 1), dtype=torch.int32, device=dev)
        fill(block, total)
        edge = block[:2 * total].reshape(2, total)
        shifts = block[2 * total:5 * total].reshape(total, 3)
        lp = torch.zeros(n_loc + 1, dtype=torch.int32, device=dev)
        torch.cumsum(num, 0, out=lp[1:])
        if world == 1:
            return edge, lp, shifts
        return edge, lp, shifts, (a0, a1)

    # ---- sizes (tiny all-gather), then ONE payload all-gather ----
    sizes = torch.tensor([total, n_loc], dtype=torch.int64, device=dev)
    all_sizes = torch.empty(2 * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_sizes, sizes, group=group)
    all_sizes = all_sizes.cpu().tolist()
    counts = all_sizes[0::2]
    natoms = all_sizes[1::2]
    pmax, nmax = max(max(counts), 1), max(max(natoms), 1)
    P = sum(counts)
    if P > 2**31 - 1:
        raise OverflowError(f"{P} pairs do not fit int32 indices")
    blk = 5 * pmax + nmax
    block = torch.empty(blk, dtype=torch.int32, device=dev)
    fill(block, pmax)
    block[5 * pmax:5 * pmax + n_loc] = num
    recv = torch.empty(world * blk, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(recv, block, group=group)

    if dev.type == "cuda" and _local_coo is None:
        from .. import _lib

        edge = torch.empty((2, P), dtype=torch.int32, device=dev)
        shifts = torch.empty((P, 3), dtype=torch.int32, device=dev)
        cnt_arr = (ctypes.c_int64 * world)(*counts)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().nvnl_unpack_gathered(ctypes.c_void_p(recv.data_ptr()), world, pmax, blk, cnt_arr,
                                                       ctypes.c_void_p(edge.data_ptr()), P,
                                                       ctypes.c_void_p(shifts.data_ptr()),
                                                       ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                       "nvnl_unpack_gathered")
    else:
        edge, shifts = _torch_unpack(recv, world, pmax, nmax, counts, natoms, P, dev)
    num_all = torch.cat([recv[g * blk + 5 * pmax: g * blk + 5 * pmax + natoms[g]] for g in range(world)])
    neighbor_ptr = torch.zeros(N + 1, dtype=torch.int32, device=dev)
    torch.cumsum(num_all, 0, out=neighbor_ptr[1:])
    return edge, neighbor_ptr, shifts
