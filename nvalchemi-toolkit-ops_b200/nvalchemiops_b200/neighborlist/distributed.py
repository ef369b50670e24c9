"""Batch-sharded multi-GPU neighbor lists: one process per GPU, ranks exchange only what cannot be recomputed.

The reference has no distributed code (no NCCL / torch.distributed call sites); this module is the
multi-GPU step BASELINE.json's north_star defines: independent systems are split by ``batch_ptr`` across the ranks of
one node, every rank builds the COO list of its own systems with GLOBAL atom indices, and an all-gather over NCCL
(NVLink 5 / NVSwitch) re-assembles the global ``edge_index`` / ``shifts`` / ``neighbor_ptr`` on every rank.
Semantics: identical to running the whole batch on one GPU, up to the order of entries inside a source atom's row
(which the reference leaves unspecified).

Data flow (4 B per pair on the wire, no re-assembly pass over the payload, exchange overlapped with the kernels):
  1. every rank splits its own systems into ``config.exchange_chunks`` chunks; local build + sweep (counts) of every chunk;
     ONE small all-gather of (pairs, max count, error bits, flags) per rank and chunk — every rank then knows every offset
     and raises the same errors at the same point; ``num_neighbors`` is gathered and scanned into ``neighbor_ptr``;
  2. chunk by chunk every rank writes its OWN range straight into the final global arrays (``nvnl_fill_rows`` with the
     global row stride, the targets landing in the rank's slot of the chunk's staging buffer) and packs its shifts into one
     byte per pair (``nvnl_pack_shifts``);
  3. the exchange of a chunk is ONE in-place ``ncclAllGather`` of one word per pair (target atom in bits 0..25, packed shift
     in bits 26..31; with 2^26 atoms or more: targets 4 B/pair + a byte array of packed shifts, two all-gathers) on a
     communication stream: it runs while the next chunk is being written and the previous one re-assembled;
  4. ``nvnl_expand_padded_ranges`` (one launch per chunk) writes the targets everywhere and the source atoms / int32
     shifts of the foreign ranges.
If some shift does not fit the packed byte (unwrapped coordinates, boxes smaller than the cutoff) the source atoms and
int32 shifts are broadcast instead (20 B/pair, in place, not chunked).
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


def _partition(batch_ptr, world, chunks=1):
    """Partition of ``batch_ptr`` (device tensor) over ``world`` ranks and of every rank's systems into ``chunks`` chunks,
    cached per tensor identity + version so that repeated calls on the same batch pay no device->host copy.
    Returns (system ranges, atom ranges, atoms, per-rank chunk system ranges, per-rank chunk atom ranges)."""
    key = (batch_ptr.data_ptr(), batch_ptr._version, batch_ptr.shape[0], world, chunks, str(batch_ptr.device))
    hit = _partition_cache.get(key)
    if hit is None:
        ptr_host = batch_ptr.detach().cpu().tolist()
        parts = partition_systems(ptr_host, world)
        sub = []
        for a, b in parts:
            local = partition_systems([ptr_host[s] - ptr_host[a] for s in range(a, b + 1)], chunks)
            sub.append([(a + x, a + y) for x, y in local])
        hit = (parts, [(ptr_host[a], ptr_host[b]) for a, b in parts], ptr_host[-1], sub,
               [[(ptr_host[x], ptr_host[y]) for x, y in row] for row in sub])
        if len(_partition_cache) > 64:
            _partition_cache.clear()
        _partition_cache[key] = hit
    return hit


_comm_streams: dict = {}


def _comm_stream(dev):
    """One side stream per device for the chunk exchanges (created once: stream creation is not free)."""
    key = (dev.type, dev.index)
    st = _comm_streams.get(key)
    if st is None:
        st = _comm_streams[key] = torch.cuda.Stream(device=dev)
    return st


class _CudaShard:
    """Some of a rank's own systems on the CUDA path: build + count queued now, sizes read by ``finish()`` (the chunks of a
    rank are all queued before the first host sync), fill into caller-provided (global) arrays later."""

    def __init__(self, positions, cutoff, cell, pbc, batch_idx, batch_ptr, half_fill, index_offset):
        from . import _engine

        self._e = _engine
        self.h = _engine.build(positions, cutoff, cell, pbc, batch_idx=batch_idx, batch_ptr=batch_ptr)
        self.csq = _engine.cutoff_sq_in_dtype(cutoff, positions.dtype)
        self.half_fill, self.index_offset = half_fill, index_offset
        self._pending = _engine.count_launch(self.h, self.csq, half_fill)

    def finish(self):
        self.num, self.ptr, self.total, self.max_count, self.err, self.hint, self.rows = self._e.count_finish(
            self.h, self.csq, self.half_fill, self._pending)
        self._pending = None
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


_WORD_SHIFT = 26          # one-word exchange: target atom in bits 0..25, packed shift in bits 26..31
_WORD_MASK = (1 << _WORD_SHIFT) - 1


def _pack_shifts_word(shifts_own, targets_own):
    """targets_own[p] |= packed shift << 26, in place on the rank's slot of the staging buffer."""
    if shifts_own.is_cuda:
        from .. import _lib

        with torch.cuda.device(shifts_own.device):
            _lib.check(_lib.lib().nvnl_pack_shifts_word(ctypes.c_void_p(shifts_own.data_ptr()), shifts_own.shape[0],
                                                        ctypes.c_void_p(targets_own.data_ptr()), None,
                                                        ctypes.c_void_p(torch.cuda.current_stream(shifts_own.device).cuda_stream)),
                       "nvnl_pack_shifts_word")
    else:  # gloo tests of the plumbing
        s = (shifts_own + 1).to(torch.int64)
        w = targets_own.to(torch.int64) | ((s[:, 0] | (s[:, 1] << 2) | (s[:, 2] << 4)) << _WORD_SHIFT)
        targets_own.copy_(torch.where(w >= 2**31, w - 2**32, w).to(torch.int32))      # the same 32 bits as an int32


def _expand_chunk(neighbor_ptr, n_atoms, world, rank, begins, ends, pair_begins, counts, pmax, t_dst, t_packed, edge, shifts):
    """One chunk of the exchange: ``t_dst`` / ``t_packed`` are the chunk's world x pmax staging buffers; rank g's atoms of
    the chunk are [begins[g], ends[g]) and its ``counts[g]`` pairs start at ``pair_begins[g]``."""
    if edge.is_cuda:
        from .. import _lib

        arr = ctypes.c_int64 * world
        with torch.cuda.device(edge.device):
            _lib.check(_lib.lib().nvnl_expand_padded_ranges(
                ctypes.c_void_p(neighbor_ptr.data_ptr()), n_atoms, world, rank, arr(*begins), arr(*ends), arr(*pair_begins), pmax,
                ctypes.c_void_p(t_dst.data_ptr()), ctypes.c_void_p(t_packed.data_ptr()) if t_packed is not None else None,
                ctypes.c_void_p(edge[0].data_ptr()),
                ctypes.c_void_p(edge[1].data_ptr()), ctypes.c_void_p(shifts.data_ptr()),
                ctypes.c_void_p(torch.cuda.current_stream(edge.device).cuda_stream)), "nvnl_expand_padded_ranges")
        return
    src = torch.repeat_interleave(torch.arange(n_atoms, dtype=torch.int32), torch.diff(neighbor_ptr).long())
    for g in range(world):
        lo, hi = pair_begins[g], pair_begins[g] + counts[g]
        got = t_dst[g * pmax: g * pmax + (hi - lo)]
        if t_packed is None:                                   # one-word exchange
            w = got.to(torch.int64) & 0xFFFFFFFF
            edge[1, lo:hi] = (w & _WORD_MASK).to(torch.int32)
            pk = (w >> _WORD_SHIFT).to(torch.int32)
        else:
            edge[1, lo:hi] = got
            pk = t_packed[g * pmax: g * pmax + (hi - lo)].to(torch.int32)
        if g == rank:
            continue
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
                                group=None, gather=True, _local_shard=None, return_stats=False, chunks=None,
                                _exchange_when_alone=False):
    """COO neighbor list of a batch, sharded over the ranks of ``group`` (default: WORLD).

    Every rank passes the SAME global tensors (on its own device): ``positions`` [N,3], ``cell`` [S,3,3],
    ``pbc`` [S,3], ``batch_ptr`` [S+1] (atoms of a system contiguous).  Returns on every rank
    ``(neighbor_list [2,P] int32, neighbor_ptr [N+1] int32, shifts [P,3] int32)`` with global atom indices.
    With ``gather=False`` the collective is skipped and the rank's own shard is returned as
    ``(neighbor_list, neighbor_ptr_local, shifts, (atom_lo, atom_hi))`` — the "kernels only" figure of the bench.
    ``return_stats=True`` appends a dict (bytes received from peers, packed or not, phase times).
    ``chunks`` (default ``config.exchange_chunks``): chunks per rank of the overlapped exchange.

    ``_local_shard`` is a test hook (CPU/gloo tests inject an oracle-backed shard); the product path leaves it None.
    ``_exchange_when_alone`` is a test hook too: a world of ONE rank still goes through the chunked exchange (staging
    slots, communication stream, re-assembly), so that this plumbing is covered on a single GPU.
    """
    from .. import config
    from ._engine import _raise_on_error_bits

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = positions.device
    N = positions.shape[0]
    exchange = gather and (world > 1 or (_exchange_when_alone and dist.is_initialized()))
    K = max(1, int(config.exchange_chunks if chunks is None else chunks)) if exchange else 1
    # phase marks for return_stats (CUDA events on the current stream; read after the call's last kernel)
    marks = []

    def mark(name):
        if return_stats and dev.type == "cuda":
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    mark("start")
    parts, atom_ranges, n_total, sub_sys, sub_atoms = _partition(batch_ptr, world, K)
    if n_total != N:
        raise ValueError("batch_ptr[-1] must equal the number of atoms")
    a0, a1 = atom_ranges[rank]
    n_loc = a1 - a0
    make = _local_shard or _CudaShard
    shards = []
    for k in range(K):
        (s0, s1), (c0, c1) = sub_sys[rank][k], sub_atoms[rank][k]
        if c1 > c0:
            lptr = (batch_ptr[s0:s1 + 1] - c0).to(torch.int32)
            lidx = torch.repeat_interleave(torch.arange(s1 - s0, dtype=torch.int32, device=dev),
                                           (lptr[1:] - lptr[:-1]).long())
            shards.append(make(positions[c0:c1], cutoff, cell[s0:s1], pbc[s0:s1], lidx, lptr, half_fill, c0))
        else:
            shards.append(None)
    for sh in shards:                                  # the host syncs, after every chunk has been queued
        if sh is not None and hasattr(sh, "finish"):
            sh.finish()

    if not exchange:
        shard = shards[0]
        total, max_count, err = (shard.total, shard.max_count, shard.err) if shard is not None else (0, 0, 0)
        num = shard.num if shard is not None else torch.zeros(0, dtype=torch.int32, device=dev)
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
    mine = []
    for sh in shards:
        mine += [sh.total, sh.max_count, sh.err, 1 if sh.packable else 0] if sh is not None else [0, 0, 0, 1]
    mine = torch.tensor(mine, dtype=torch.int64, device=dev)
    everyone = torch.empty(4 * K * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(everyone, mine, group=group)
    everyone = everyone.cpu().tolist()
    counts = [[everyone[4 * (g * K + k)] for k in range(K)] for g in range(world)]     # pairs of (rank, chunk)
    err_all = 0
    for e in everyone[2::4]:
        err_all |= int(e)
    _raise_on_error_bits(err_all)
    worst = max(everyone[1::4])
    if max_neighbors is not None and worst > max_neighbors:
        raise NeighborOverflowError(max_neighbors, worst)
    packed_ok = all(everyone[3::4])
    offs, P = [], 0                                    # first pair of (rank, chunk): global atom order = (rank, chunk) order
    for g in range(world):
        row = []
        for k in range(K):
            row.append(P)
            P += counts[g][k]
        offs.append(row)
    if P > 2**31 - 1:
        raise OverflowError(f"{P} pairs do not fit int32 indices")
    rank_lo = [offs[g][0] for g in range(world)] + [P]         # pairs of rank g: [rank_lo[g], rank_lo[g + 1])
    pmax = max(max(row) for row in counts)

    mark("size_exchange")
    shifts = torch.empty((P, 3), dtype=torch.int32, device=dev)
    num_all = torch.empty((N,), dtype=torch.int32, device=dev)
    for k, sh in enumerate(shards):
        if sh is not None:
            num_all[sub_atoms[rank][k][0]:sub_atoms[rank][k][1]] = sh.num
    counts_views = [num_all[atom_ranges[g][0]:atom_ranges[g][1]] for g in range(world)]
    neighbor_ptr = torch.zeros(N + 1, dtype=torch.int32, device=dev)
    stats = {}

    if packed_ok:
        # neighbor_ptr first (the per-atom counts are all the re-assembly needs besides the payload)
        _gather_slices([counts_views], rank, group)
        torch.cumsum(num_all, 0, out=neighbor_ptr[1:])
        mark("counts_gather_scan")
        # ---- 2./3. chunk by chunk: own range into the final arrays (targets into the chunk's staging slot, shifts packed),
        #      then ONE in-place all-gather per array on the communication stream ----
        ebuf = torch.empty(2 * P + K * world * pmax, dtype=torch.int32, device=dev)
        edge = ebuf[:2 * P].view(2, P)
        t_dst = ebuf[2 * P:]
        word = bool(config.exchange_word) and N <= _WORD_MASK       # every atom index fits 26 bits: 4 B per pair, one array
        t_packed = None if word else torch.empty(K * world * pmax, dtype=torch.uint8, device=dev)
        on_gpu = dev.type == "cuda"
        cur = torch.cuda.current_stream(dev) if on_gpu else None
        comm = _comm_stream(dev) if on_gpu and K > 1 else cur
        done, comm_marks = [], []
        for k, sh in enumerate(shards):
            base = k * world * pmax                     # the chunk's staging buffers: [base, base + world * pmax)
            slot = base + rank * pmax
            o0, cnt = offs[rank][k], counts[rank][k]
            if sh is not None:
                # row 0 at the chunk's offset of the final array, row 1 in the rank's slot of the chunk's staging buffer
                sh.fill(ebuf[o0:], shifts[o0:o0 + cnt], (2 * P + slot) - o0)
                if cnt > 0 and word:
                    _pack_shifts_word(shifts[o0:o0 + cnt], t_dst[slot:slot + cnt])
                elif cnt > 0:
                    _pack_shifts(shifts[o0:o0 + cnt], t_packed[slot:slot + cnt])
            if pmax == 0:
                continue
            if on_gpu:
                if comm is not cur:
                    ready = torch.cuda.Event()
                    ready.record(cur)
                    comm.wait_event(ready)
                with torch.cuda.stream(comm):
                    if return_stats:
                        b = torch.cuda.Event(enable_timing=True)
                        b.record(comm)
                    dist.all_gather_into_tensor(t_dst[base:base + world * pmax], t_dst[slot:slot + pmax], group=group)
                    if not word:
                        dist.all_gather_into_tensor(t_packed[base:base + world * pmax], t_packed[slot:slot + pmax], group=group)
                    ev = torch.cuda.Event(enable_timing=return_stats)
                    ev.record(comm)
                    if return_stats:
                        comm_marks.append((b, ev))
                done.append(ev)
            else:   # gloo tests of the plumbing
                dist.all_gather([t_dst[base + g * pmax: base + (g + 1) * pmax] for g in range(world)],
                                t_dst[slot:slot + pmax].clone(), group=group)
                if not word:
                    dist.all_gather([t_packed[base + g * pmax: base + (g + 1) * pmax] for g in range(world)],
                                    t_packed[slot:slot + pmax].clone(), group=group)
        mark("alloc_fill_own_pack")
        # ---- 4. per chunk, as its exchange completes: the targets everywhere, source atoms / shifts of the foreign ranges ----
        for k in range(K):
            if pmax == 0:
                break
            if on_gpu and comm is not cur:
                cur.wait_event(done[k])
            base = k * world * pmax
            _expand_chunk(neighbor_ptr, N, world, rank, [sub_atoms[g][k][0] for g in range(world)],
                          [sub_atoms[g][k][1] for g in range(world)], [offs[g][k] for g in range(world)],
                          [counts[g][k] for g in range(world)], pmax, t_dst[base:base + world * pmax],
                          None if word else t_packed[base:base + world * pmax], edge, shifts)
        mark("exchange_wait_expand")
        stats["bytes_per_pair"] = 4 if word else 5
        if comm_marks:
            stats["_comm_marks"] = comm_marks
    else:
        # ---- fallback: int32 shifts and source atoms travel too (20 B/pair), in place, one exchange ----
        edge = torch.empty((2, P), dtype=torch.int32, device=dev)
        for k, sh in enumerate(shards):
            if sh is not None:
                o0, cnt = offs[rank][k], counts[rank][k]
                sh.fill(edge.view(-1)[o0:], shifts[o0:o0 + cnt], P)
        mark("alloc_fill_own")
        _gather_slices([[edge[1, rank_lo[g]:rank_lo[g + 1]] for g in range(world)], counts_views,
                        [edge[0, rank_lo[g]:rank_lo[g + 1]] for g in range(world)],
                        [shifts[rank_lo[g]:rank_lo[g + 1]] for g in range(world)]], rank, group)
        mark("nccl_gather")
        torch.cumsum(num_all, 0, out=neighbor_ptr[1:])
        mark("ptr_scan")
    if return_stats:
        per_pair = stats.get("bytes_per_pair", 20)
        own_pairs = rank_lo[rank + 1] - rank_lo[rank]
        comm_marks = stats.pop("_comm_marks", [])
        stats.update({"peer_bytes": per_pair * (P - own_pairs) + 4 * (N - n_loc), "packed": packed_ok, "chunks": K})
        if marks:
            torch.cuda.synchronize(dev)
            stats["phase_ms"] = {marks[k + 1][0]: marks[k][1].elapsed_time(marks[k + 1][1]) for k in range(len(marks) - 1)}
            if comm_marks:
                # the exchanges run on the communication stream, under the kernels of the neighbouring chunks
                stats["phase_ms"]["nccl_chunks_on_comm_stream"] = [b.elapsed_time(e) for b, e in comm_marks]
        return edge, neighbor_ptr, shifts, stats
    return edge, neighbor_ptr, shifts
