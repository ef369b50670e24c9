"""Batch-sharded multi-GPU neighbor lists: one process per GPU, ranks exchange only what cannot be recomputed.

The reference has no distributed code (no NCCL / torch.distributed call sites); this module is the
multi-GPU step BASELINE.json's north_star defines: independent systems are split by ``batch_ptr`` across the ranks of
one node, every rank builds the COO list of its own systems with GLOBAL atom indices, and an all-gather over NCCL
(NVLink 5 / NVSwitch) re-assembles the global ``edge_index`` / ``shifts`` / ``neighbor_ptr`` on every rank.
Semantics: identical to running the whole batch on one GPU, up to the order of entries inside a source atom's row
(which the reference leaves unspecified).

Data flow (no padded blocks, no re-assembly pass over the payload):
  1. local build + sweep (counts);  ONE small all-gather of (pairs, atoms, max count, error bits, flags) per rank — every
     rank then knows every offset and raises the same errors at the same point;
  2. every rank writes its OWN range straight into the final global arrays (``nvnl_fill_rows`` with the global row
     stride) and packs its shifts into one byte per pair (``nvnl_pack_shifts``);
  3. variable-size gather = one NCCL broadcast per rank and array, IN PLACE on slices of the final arrays, grouped into
     one NCCL group: per pair only the target atom (4 B) and the packed shift (1 B) travel, per atom ``num_neighbors``
     — 5 B/pair instead of 20;
  4. ``neighbor_ptr`` = scan of the gathered counts; ``nvnl_expand_gathered`` writes the source atoms and int32 shifts of
     the foreign ranges from it.
If some shift does not fit the packed byte (unwrapped coordinates, boxes smaller than the cutoff) the source atoms and
int32 shifts are broadcast instead (20 B/pair, still in place).
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from .neighbor_utils import NeighborOverflowError

_partition_cache: dict = {}


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


def _partition(batch_ptr, world):
    """Partition of ``batch_ptr`` (device tensor) over ``world`` ranks, cached per tensor identity + version so that
    repeated calls on the same batch pay no device->host copy."""
    key = (batch_ptr.data_ptr(), batch_ptr._version, batch_ptr.shape[0], world, str(batch_ptr.device))
    hit = _partition_cache.get(key)
    if hit is None:
        ptr_host = batch_ptr.detach().cpu().tolist()
        parts = partition_systems(ptr_host, world)
        hit = (parts, [(ptr_host[a], ptr_host[b]) for a, b in parts], ptr_host[-1])
        if len(_partition_cache) > 64:
            _partition_cache.clear()
        _partition_cache[key] = hit
    return hit


class _CudaShard:
    """A rank's own systems on the CUDA path: build + count now, fill into caller-provided (global) arrays later."""

    def __init__(self, positions, cutoff, cell, pbc, batch_idx, batch_ptr, half_fill, index_offset):
        from . import _engine

        self._e = _engine
        self.h = _engine.build(positions, cutoff, cell, pbc, batch_idx=batch_idx, batch_ptr=batch_ptr)
        self.csq = _engine.cutoff_sq_in_dtype(cutoff, positions.dtype)
        self.half_fill, self.index_offset = half_fill, index_offset
        self.num, self.ptr, self.total, self.max_count, self.err, self.hint, self.rows = _engine.count_and_size(
            self.h, self.csq, half_fill)
        # shifts fit one byte when no atom lies outside the primary image and no stencil is wider than one cell
        self.packable = not (self.hint & 1) and not self.h.wide_stencil

    def fill(self, edge_rows, shifts, row_stride):
        """edge_rows: flat view of the global edge_index starting at this rank's first pair of row 0 (row 1 is
        ``row_stride`` entries further); shifts: view of the rank's own rows of the global shifts."""
        if self.total > 0:
            self._e.fill_coo(self.h, self.csq, self.ptr, edge_rows, shifts, self.total, self.half_fill, self.index_offset,
                             launch_hint=self.hint, rows=self.rows, row_stride=row_stride)


def _pack_shifts(shifts_own, packed_own):
    if shifts_own.is_cuda:
        from .. import _lib

        with torch.cuda.device(shifts_own.device):
            _lib.check(_lib.lib().nvnl_pack_shifts(ctypes.c_void_p(shifts_own.data_ptr()), shifts_own.shape[0],
                                                   ctypes.c_void_p(packed_own.data_ptr()), None,
                                                   ctypes.c_void_p(torch.cuda.current_stream(shifts_own.device).cuda_stream)),
                       "nvnl_pack_shifts")
    else:  # gloo tests of the plumbing (CPU tensors only ever get here through the _local_shard test hook)
        s = (shifts_own + 1).to(torch.uint8)
        packed_own.copy_(s[:, 0] | (s[:, 1] << 2) | (s[:, 2] << 4))


def _expand(neighbor_ptr, n_atoms, a0, a1, packed, edge, shifts):
    if edge.is_cuda:
        from .. import _lib

        with torch.cuda.device(edge.device):
            _lib.check(_lib.lib().nvnl_expand_gathered(ctypes.c_void_p(neighbor_ptr.data_ptr()), n_atoms, a0, a1,
                                                       ctypes.c_void_p(packed.data_ptr()), ctypes.c_void_p(edge.data_ptr()),
                                                       ctypes.c_void_p(shifts.data_ptr()),
                                                       ctypes.c_void_p(torch.cuda.current_stream(edge.device).cuda_stream)),
                       "nvnl_expand_gathered")
    else:
        counts = torch.diff(neighbor_ptr).long()
        src = torch.repeat_interleave(torch.arange(n_atoms, dtype=torch.int32), counts)
        p = packed.to(torch.int32)
        full = torch.stack([(p & 3) - 1, ((p >> 2) & 3) - 1, ((p >> 4) & 3) - 1], dim=1).to(torch.int32)
        lo, hi = int(neighbor_ptr[a0]), int(neighbor_ptr[a1])
        keep_e, keep_s = edge[0, lo:hi].clone(), shifts[lo:hi].clone()
        edge[0].copy_(src)
        shifts.copy_(full)
        edge[0, lo:hi] = keep_e
        shifts[lo:hi] = keep_s


def _expand_padded(neighbor_ptr, n_atoms, world, rank, atom_ranges, offs, pmax, t_dst, t_packed, edge, shifts):
    """edge[1] for every pair, edge[0] / shifts for the other ranks' pairs, from the padded gathered staging buffers."""
    if edge.is_cuda:
        from .. import _lib

        ab = (ctypes.c_int64 * (world + 1))(*([a for a, _ in atom_ranges] + [atom_ranges[-1][1]]))
        pb = (ctypes.c_int64 * (world + 1))(*offs)
        with torch.cuda.device(edge.device):
            _lib.check(_lib.lib().nvnl_expand_padded(ctypes.c_void_p(neighbor_ptr.data_ptr()), n_atoms, world, rank, ab, pb, pmax,
                                                     ctypes.c_void_p(t_dst.data_ptr()), ctypes.c_void_p(t_packed.data_ptr()),
                                                     ctypes.c_void_p(edge[0].data_ptr()), ctypes.c_void_p(edge[1].data_ptr()),
                                                     ctypes.c_void_p(shifts.data_ptr()),
                                                     ctypes.c_void_p(torch.cuda.current_stream(edge.device).cuda_stream)),
                       "nvnl_expand_padded")
        return
    counts = torch.diff(neighbor_ptr).long()
    src = torch.repeat_interleave(torch.arange(n_atoms, dtype=torch.int32), counts)
    for g in range(world):
        lo, hi = offs[g], offs[g + 1]
        edge[1, lo:hi] = t_dst[g * pmax: g * pmax + (hi - lo)]
        if g == rank:
            continue
        pk = t_packed[g * pmax: g * pmax + (hi - lo)].to(torch.int32)
        edge[0, lo:hi] = src[lo:hi]
        shifts[lo:hi] = torch.stack([(pk & 3) - 1, ((pk >> 2) & 3) - 1, ((pk >> 4) & 3) - 1], dim=1).to(torch.int32)


def _gather_slices(arrays_by_rank, rank, group):
    """Variable-size all-gather IN PLACE: ``arrays_by_rank[k][g]`` is the contiguous view of array k that rank g owns
    (already filled on rank g); afterwards every view is filled on every rank.
    NCCL: one ``all_gather`` per array — for uneven sizes ProcessGroupNCCL issues it as ONE group of ncclBroadcasts
    straight into the output views (no staging copy).  Other backends (the gloo tests): one broadcast per view."""
    world = len(arrays_by_rank[0])
    dev = arrays_by_rank[0][0].device
    if dev.type == "cuda":
        for views in arrays_by_rank:
            if sum(v.numel() for v in views) > 0:
                dist.all_gather(list(views), views[rank], group=group)
        return
    root = (lambda g: dist.get_global_rank(group, g)) if group is not None else (lambda g: g)
    for views in arrays_by_rank:
        for g in range(world):
            if views[g].numel() > 0:
                dist.broadcast(views[g], src=root(g), group=group)


def sharded_batch_neighbor_list(positions, cutoff, cell, pbc, batch_ptr, half_fill=False, max_neighbors=None,
                                group=None, gather=True, _local_shard=None, return_stats=False):
    """COO neighbor list of a batch, sharded over the ranks of ``group`` (default: WORLD).

    Every rank passes the SAME global tensors (on its own device): ``positions`` [N,3], ``cell`` [S,3,3],
    ``pbc`` [S,3], ``batch_ptr`` [S+1] (atoms of a system contiguous).  Returns on every rank
    ``(neighbor_list [2,P] int32, neighbor_ptr [N+1] int32, shifts [P,3] int32)`` with global atom indices.
    With ``gather=False`` the collective is skipped and the rank's own shard is returned as
    ``(neighbor_list, neighbor_ptr_local, shifts, (atom_lo, atom_hi))`` — the "kernels only" figure of the bench.
    ``return_stats=True`` appends a dict (bytes received from peers, packed or not).

    ``_local_shard`` is a test hook (CPU/gloo tests inject an oracle-backed shard); the product path leaves it None.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = positions.device
    N = positions.shape[0]
    # phase marks for return_stats (CUDA events on the current stream; read after the call's last kernel)
    marks = []

    def mark(name):
        if return_stats and dev.type == "cuda":
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    mark("start")
    parts, atom_ranges, n_total = _partition(batch_ptr, world)
    if n_total != N:
        raise ValueError("batch_ptr[-1] must equal the number of atoms")
    s0, s1 = parts[rank]
    a0, a1 = atom_ranges[rank]
    n_loc = a1 - a0
    make = _local_shard or _CudaShard
    if n_loc > 0:
        lptr = (batch_ptr[s0:s1 + 1] - a0).to(torch.int32)
        lidx = torch.repeat_interleave(torch.arange(s1 - s0, dtype=torch.int32, device=dev),
                                       (lptr[1:] - lptr[:-1]).long())
        shard = make(positions[a0:a1], cutoff, cell[s0:s1], pbc[s0:s1], lidx, lptr, half_fill, a0)
        total, max_count, err, packable, num = shard.total, shard.max_count, shard.err, shard.packable, shard.num
    else:
        shard, total, max_count, err, packable = None, 0, 0, 0, True
        num = torch.zeros(0, dtype=torch.int32, device=dev)

    if not gather or world == 1:
        from ._engine import _raise_on_error_bits

        _raise_on_error_bits(err)
        if max_neighbors is not None and max_count > max_neighbors:
            raise NeighborOverflowError(max_neighbors, max_count)
        edge = torch.empty((2, total), dtype=torch.int32, device=dev)
        shifts = torch.empty((total, 3), dtype=torch.int32, device=dev)
        if shard is not None:
            shard.fill(edge.view(-1), shifts, total)
        lp = torch.zeros(n_loc + 1, dtype=torch.int32, device=dev)
        torch.cumsum(num, 0, out=lp[1:])
        if world == 1:
            return (edge, lp, shifts, {"peer_bytes": 0, "packed": False}) if return_stats else (edge, lp, shifts)
        return edge, lp, shifts, (a0, a1)

    mark("build_sweep_count")
    # ---- 1. one small all-gather: every rank learns every size and raises the same errors at the same point ----
    mine = torch.tensor([total, max_count, err, 1 if packable else 0], dtype=torch.int64, device=dev)
    everyone = torch.empty(4 * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(everyone, mine, group=group)
    everyone = everyone.cpu().tolist()
    counts = everyone[0::4]
    from ._engine import _raise_on_error_bits

    err_all = 0
    for e in everyone[2::4]:
        err_all |= int(e)
    _raise_on_error_bits(err_all)
    worst = max(everyone[1::4])
    if max_neighbors is not None and worst > max_neighbors:
        raise NeighborOverflowError(max_neighbors, worst)
    P = sum(counts)
    if P > 2**31 - 1:
        raise OverflowError(f"{P} pairs do not fit int32 indices")
    packed_ok = all(everyone[3::4])
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + c)

    mark("size_exchange")
    # ---- 2. every rank writes its own range of the final arrays ----
    # Packed exchange (the usual case): the targets (row 1 of edge_index) and the packed shifts travel through staging
    # buffers of world x Pmax entries, each rank writing ITS slot directly (the fill kernel's row 1 lands there), so that
    # the exchange is ONE in-place ncclAllGather per array (grouped broadcasts into uneven views run at a third of its
    # bandwidth); nvnl_expand_padded then writes edge_index row 1 everywhere and row 0 / shifts of the other ranks.
    shifts = torch.empty((P, 3), dtype=torch.int32, device=dev)
    num_all = torch.empty((N,), dtype=torch.int32, device=dev)
    o0, o1 = offs[rank], offs[rank + 1]
    pmax = max(counts)
    if packed_ok:
        ebuf = torch.empty(2 * P + world * pmax, dtype=torch.int32, device=dev)
        edge = ebuf[:2 * P].view(2, P)
        t_dst = ebuf[2 * P:]
        t_packed = torch.empty(world * pmax, dtype=torch.uint8, device=dev)
        if shard is not None:
            # row 0 at the rank's offset of the final array, row 1 in the rank's slot of the staging buffer
            shard.fill(ebuf[o0:], shifts[o0:o1], (2 * P + rank * pmax) - o0)
            num_all[a0:a1] = num
            if o1 > o0:
                _pack_shifts(shifts[o0:o1], t_packed[rank * pmax: rank * pmax + (o1 - o0)])
    else:
        edge = torch.empty((2, P), dtype=torch.int32, device=dev)
        if shard is not None:
            shard.fill(edge.view(-1)[o0:], shifts[o0:o1], P)
            num_all[a0:a1] = num

    mark("alloc_fill_own_pack")
    # ---- 3. gather ----
    counts_views = [num_all[atom_ranges[g][0]:atom_ranges[g][1]] for g in range(world)]
    if packed_ok:
        _gather_slices([counts_views], rank, group)
        if pmax > 0:
            if dev.type == "cuda":
                dist.all_gather_into_tensor(t_dst, t_dst[rank * pmax:(rank + 1) * pmax], group=group)
                dist.all_gather_into_tensor(t_packed, t_packed[rank * pmax:(rank + 1) * pmax], group=group)
            else:   # gloo tests of the plumbing
                dist.all_gather([t_dst[g * pmax:(g + 1) * pmax] for g in range(world)], t_dst[rank * pmax:(rank + 1) * pmax].clone(),
                                group=group)
                dist.all_gather([t_packed[g * pmax:(g + 1) * pmax] for g in range(world)],
                                t_packed[rank * pmax:(rank + 1) * pmax].clone(), group=group)
    else:
        _gather_slices([[edge[1, offs[g]:offs[g + 1]] for g in range(world)], counts_views,
                        [edge[0, offs[g]:offs[g + 1]] for g in range(world)],
                        [shifts[offs[g]:offs[g + 1]] for g in range(world)]], rank, group)

    mark("nccl_gather")
    # ---- 4. neighbor_ptr, then the targets everywhere and the source atoms / shifts of the foreign ranges ----
    neighbor_ptr = torch.zeros(N + 1, dtype=torch.int32, device=dev)
    torch.cumsum(num_all, 0, out=neighbor_ptr[1:])
    if packed_ok:
        _expand_padded(neighbor_ptr, N, world, rank, atom_ranges, offs, pmax, t_dst, t_packed, edge, shifts)
    mark("ptr_scan_expand")
    if return_stats:
        per_pair = 5 if packed_ok else 20
        stats = {"peer_bytes": per_pair * (P - counts[rank]) + 4 * (N - n_loc), "packed": packed_ok}
        if marks:
            torch.cuda.synchronize(dev)
            stats["phase_ms"] = {marks[k + 1][0]: marks[k][1].elapsed_time(marks[k + 1][1]) for k in range(len(marks) - 1)}
        return edge, neighbor_ptr, shifts, stats
    return edge, neighbor_ptr, shifts
