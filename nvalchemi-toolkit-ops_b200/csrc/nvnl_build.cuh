// nvnl_build.cuh — grid selection, atom->cell hash, look-back scan and counting-sort scatter.
//
// Replaces (different algorithm, same role): cell_list.py:102-369 / batch_cell_list.py:102-377
// (_cell_list_construct_bin_size, _count_atoms_per_bin, torch.cumsum, _bin_atoms).
// Differences that matter:
//   * ideal grid (cell width >= cutoff, no 1000-cell cap) bounded only by #cells <= #atoms;
//   * non-periodic dimensions are gridded over the atoms' bounding slab instead of being clamped
//     into the unit cell;
//   * the sort moves POSITIONS (float4 {x,y,z,orig_index} runs per cell), not just indices, so the
//     sweep streams contiguous 16-byte records (TMA-able) instead of chasing an index list;
//   * one atomic per atom in total (the count pass hands out the in-cell rank).
#pragma once
#include "nvnl_common.cuh"

namespace nvnl {

// ------------------------------------------------------------------------------------------------
// Workspace layout (identical on host and device; everything 256-byte aligned)
// ------------------------------------------------------------------------------------------------
struct WsLayout {
    size_t ctrl, sys, bbox, cell_count, cell_start, atom_cell, atom_rank, atom_ashift, sorted, sorted_ashift,
        cursor, scan_status0, scan_status1, masks, deferred, ptr_sorted, row_ref, split, huge, rows, total;
    long long max_cells;  // N + S (upper bound on the number of cells, see k_grid)
    long long rows_cap;   // entries of the temporary row buffer (single-sweep COO path, nvnl_rows.cuh)
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline WsLayout make_layout(long long n, long long s, int rec_bytes, long long rows_per_atom = kRowsPerAtom,
                                                long long rows_slack = kRowsSlackEntries) {
    WsLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t at = o;
        o = align_up(o + bytes, 256);
        return at;
    };
    L.max_cells = n + s;
    L.ctrl = take(sizeof(Ctrl));
    L.sys = take(sizeof(SysParams) * (size_t)s);
    L.bbox = take(sizeof(long long) * 6 * (size_t)s);
    L.cell_count = take(sizeof(int) * (size_t)(L.max_cells + 2));
    L.cell_start = take(sizeof(int) * (size_t)(L.max_cells + 2));
    L.atom_cell = take(sizeof(int) * (size_t)n);
    L.atom_rank = take(sizeof(int) * (size_t)n);
    L.atom_ashift = take(sizeof(int4) * (size_t)n);
    L.sorted = take((size_t)rec_bytes * (size_t)n);
    L.sorted_ashift = take(sizeof(int4) * (size_t)n);
    L.cursor = take(sizeof(int) * (size_t)n);
    L.scan_status0 = take(sizeof(unsigned long long) * (size_t)((L.max_cells + 2) / kScanTile + 2));
    L.scan_status1 = take(sizeof(unsigned long long) * (size_t)((n + 1) / kScanTile + 2));
    L.masks = take(sizeof(unsigned) * 32 * (size_t)n);   // one hit mask per (atom, 32-candidate chunk)
    // (cell, first target) work items the fast kernel leaves to the general kernel: <= #cells + N/32 entries
    L.deferred = take(sizeof(int2) * (size_t)(L.max_cells + 2 + n / 32 + 1));
    L.ptr_sorted = take(sizeof(int) * (size_t)(n + 4));     // neighbor_ptr gathered into cell-sorted atom order
    // single-sweep COO path: row_ref[i] = (first entry << 2) | header kind, or -1; rows = compact rows in sweep
    // order.  Budget: kRowsPerAtom entries per atom + one reservation block per resident warp; more pairs than that
    // (very large cutoffs) make the query fall back to the two-pass path.
    L.row_ref = take(sizeof(int) * (size_t)n);
    // parts of cells with many targets: (cell + 1, first target) — a cell of more than 64 targets is cut into parts of
    // 32, so there are fewer than n / 16 + 1 of them
    L.split = take(sizeof(int2) * (size_t)(n / 16 + 2));
    // parts (20 targets) of single-cell systems that take the six-mask-word launch: (cell, first target)
    L.huge = take(sizeof(int2) * (size_t)(n / 16 + s + 2));
    L.rows_cap = rows_per_atom * n + rows_slack;
    if (L.rows_cap < 1) L.rows_cap = 1;
    if (L.rows_cap > (1LL << 29) - 1) L.rows_cap = (1LL << 29) - 1;   // row_ref = first entry << 2 | header kind
    L.rows = take(sizeof(int) * (size_t)L.rows_cap);
    L.total = o;
    return L;
}

__device__ __forceinline__ long long order_key(double d) {
    long long k = __double_as_longlong(d);
    return k >= 0 ? k : (k ^ 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double order_unkey(long long k) {
    return __longlong_as_double(k >= 0 ? k : (k ^ 0x7fffffffffffffffLL));
}

// ------------------------------------------------------------------------------------------------
// k_init: zero the queues / counters / scan status / per-atom cursors and derive the per-system
// inverse cell.  One launch replaces the reference's five zero_() calls.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_init(unsigned char* __restrict__ ws, WsLayout L, long long n, int num_systems,
                       const T* __restrict__ cell, const unsigned char* __restrict__ pbc,
                       const int* __restrict__ batch_ptr, int* __restrict__ num_neighbors) {
    pdl_enter();
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    Ctrl* ctrl = reinterpret_cast<Ctrl*>(ws + L.ctrl);
    if (gid == 0) {
        for (int k = 0; k < 4; ++k) { ctrl->work_counter[k] = 0; ctrl->done[k] = 0; }
        ctrl->unwrapped = 0;
        ctrl->total_cells = 0;
        ctrl->error = 0;
        ctrl->scan_tile[0] = ctrl->scan_tile[1] = 0;
        ctrl->total_pairs = 0ull;
        ctrl->max_count = 0;
        ctrl->n_deferred = 0;
        ctrl->had_deferred = 0;
        ctrl->wide_stencil = 0;
        ctrl->shift_heavy = 0;
        ctrl->split_reserved = ctrl->split_next = ctrl->cells_done = 0;
        ctrl->n_huge = ctrl->huge_next = ctrl->huge_done = ctrl->had_huge = 0;
    }
    {
        int2* split = reinterpret_cast<int2*>(ws + L.split);   // (entry.x == 0: not pushed yet)
        for (long long i = gid; i < n / 16 + 2; i += stride) split[i] = make_int2(0, 0);
    }
    int* cell_count = reinterpret_cast<int*>(ws + L.cell_count);
    for (long long i = gid; i < L.max_cells + 2; i += stride) cell_count[i] = 0;
    int* cursor = reinterpret_cast<int*>(ws + L.cursor);
    for (long long i = gid; i < n; i += stride) cursor[i] = 0;
    if (num_neighbors)
        for (long long i = gid; i < n; i += stride) num_neighbors[i] = 0;
    unsigned long long* st0 = reinterpret_cast<unsigned long long*>(ws + L.scan_status0);
    unsigned long long* st1 = reinterpret_cast<unsigned long long*>(ws + L.scan_status1);
    for (long long i = gid; i < (L.max_cells + 2) / kScanTile + 2; i += stride) st0[i] = 0ull;
    for (long long i = gid; i < (n + 1) / kScanTile + 2; i += stride) st1[i] = 0ull;

    SysParams* sys = reinterpret_cast<SysParams*>(ws + L.sys);
    long long* bbox = reinterpret_cast<long long*>(ws + L.bbox);
    for (long long s = gid; s < num_systems; s += stride) {
        SysParams sp;
        double m[9];
        for (int k = 0; k < 9; ++k) m[k] = (double)cell[s * 9 + k];
        const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
        const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
        if (det == 0.0 || !(det == det)) atomicOr(&ctrl->error, ERR_SINGULAR_CELL);
        const double r = det != 0.0 ? 1.0 / det : 0.0;
        sp.inv[0] = (e * i - f * h) * r; sp.inv[1] = (c * h - b * i) * r; sp.inv[2] = (b * f - c * e) * r;
        sp.inv[3] = (f * g - d * i) * r; sp.inv[4] = (a * i - c * g) * r; sp.inv[5] = (c * d - a * f) * r;
        sp.inv[6] = (d * h - e * g) * r; sp.inv[7] = (b * g - a * h) * r; sp.inv[8] = (a * e - b * d) * r;
        for (int k = 0; k < 9; ++k) sp.cellm[k] = m[k];
        for (int dd = 0; dd < 3; ++dd) {
            const double x = sp.inv[dd], y = sp.inv[3 + dd], z = sp.inv[6 + dd];
            const double len = sqrt(x * x + y * y + z * z);
            sp.face[dd] = len > 0.0 ? 1.0 / len : 0.0;
            sp.pbc[dd] = pbc[s * 3 + dd] ? 1 : 0;
            sp.fmin[dd] = 0.0;
            sp.fscale[dd] = 0.0;
            sp.cpd[dd] = 1;
            sp.R[dd] = 0;
            bbox[s * 6 + dd] = 0x7fffffffffffffffLL;          // running min (ordered key)
            bbox[s * 6 + 3 + dd] = (long long)0x8000000000000000ULL;  // running max
        }
        sp.cell_offset = 0;
        sp.ncells = 1;
        sp.natoms = batch_ptr ? (batch_ptr[s + 1] - batch_ptr[s]) : (num_systems == 1 ? (int)n : 0);
        sp.pad = 0;
        sys[s] = sp;
    }
}

// ------------------------------------------------------------------------------------------------
// k_bbox: fractional bounding slab of every system along its non-periodic dims, plus atoms per
// system when the caller gave only batch_idx.  Blocks whose atoms need neither exit at once.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_bbox(unsigned char* __restrict__ ws, WsLayout L, long long n, int num_systems,
                       const T* __restrict__ pos, const int* __restrict__ batch_idx, int need_counts) {
    pdl_enter();
    SysParams* sys = reinterpret_cast<SysParams*>(ws + L.sys);
    long long* bbox = reinterpret_cast<long long*>(ws + L.bbox);
    Ctrl* ctrl = reinterpret_cast<Ctrl*>(ws + L.ctrl);
    __shared__ long long s_red[6][kSmallBlock / 32];
    __shared__ int s_cnt;
    const long long chunk = num_systems > 1 ? 512 : 2048;   // (batches: more, shorter blocks — the atomics of one system serialise)
    for (long long base = (long long)blockIdx.x * chunk; base < n; base += (long long)gridDim.x * chunk) {
        const long long end = base + chunk < n ? base + chunk : n;
        const int s_first = batch_idx ? batch_idx[base] : 0;
        bool uniform = true;
        long long mn[3] = {0x7fffffffffffffffLL, 0x7fffffffffffffffLL, 0x7fffffffffffffffLL};
        long long mx[3] = {(long long)0x8000000000000000ULL, (long long)0x8000000000000000ULL,
                           (long long)0x8000000000000000ULL};
        int cnt = 0;
        // pass 1: is the block's chunk a single system?
        for (long long i = base + threadIdx.x; i < end; i += blockDim.x) {
            const int s = batch_idx ? batch_idx[i] : 0;
            if (s != s_first) uniform = false;
        }
        uniform = __syncthreads_and(uniform);
        if (uniform) {
            if (s_first < 0 || s_first >= num_systems) {
                if (threadIdx.x == 0) atomicOr(&ctrl->error, ERR_BAD_BATCH_IDX);
                continue;
            }
            const SysParams& sp = sys[s_first];
            const bool any_open = !(sp.pbc[0] && sp.pbc[1] && sp.pbc[2]);
            if (!any_open && !need_counts) continue;
            if (any_open) {
                for (long long i = base + threadIdx.x; i < end; i += blockDim.x) {
                    const double px = (double)pos[3 * i], py = (double)pos[3 * i + 1], pz = (double)pos[3 * i + 2];
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        if (!sp.pbc[d]) {
                            const double fr = px * sp.inv[d] + py * sp.inv[3 + d] + pz * sp.inv[6 + d];
                            const long long k = order_key(fr);
                            mn[d] = k < mn[d] ? k : mn[d];
                            mx[d] = k > mx[d] ? k : mx[d];
                        }
                    }
                }
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    for (int o = 16; o > 0; o >>= 1) {
                        long long a = __shfl_xor_sync(0xffffffffu, mn[d], o);
                        long long b = __shfl_xor_sync(0xffffffffu, mx[d], o);
                        mn[d] = a < mn[d] ? a : mn[d];
                        mx[d] = b > mx[d] ? b : mx[d];
                    }
                    if ((threadIdx.x & 31) == 0) {
                        s_red[d][threadIdx.x >> 5] = mn[d];
                        s_red[3 + d][threadIdx.x >> 5] = mx[d];
                    }
                }
                __syncthreads();
                if (threadIdx.x < 6) {
                    const int d = threadIdx.x;
                    long long v = s_red[d][0];
                    for (int w = 1; w < (int)blockDim.x / 32; ++w) {
                        const long long u = s_red[d][w];
                        v = d < 3 ? (u < v ? u : v) : (u > v ? u : v);
                    }
                    if (d < 3) {
                        if (!sp.pbc[d]) atomicMin(&bbox[s_first * 6 + d], v);
                    } else {
                        if (!sp.pbc[d - 3]) atomicMax(&bbox[s_first * 6 + d], v);
                    }
                }
                __syncthreads();
            }
            if (need_counts && threadIdx.x == 0) atomicAdd(&sys[s_first].natoms, (int)(end - base));
        } else {
            (void)cnt; (void)s_cnt;
            // several systems in the chunk: warps whose 32 atoms belong to ONE system (the usual case, systems of a few
            // hundred atoms) reduce in registers and issue one atomic per value; mixed warps fall back to per-atom atomics
            const int lane = threadIdx.x & 31;
            for (long long i0 = base + (threadIdx.x & ~31); i0 < end; i0 += blockDim.x) {
                const long long i = i0 + lane;
                bool act = i < end;
                int s = act ? batch_idx[i] : -1;
                if (act && (s < 0 || s >= num_systems)) {
                    atomicOr(&ctrl->error, ERR_BAD_BATCH_IDX);
                    act = false;
                    s = -1;
                }
                const int s0 = __shfl_sync(0xffffffffu, s, 0);
                if (__all_sync(0xffffffffu, act && s == s0)) {
                    const SysParams& sp = sys[s0];
                    if (need_counts && lane == 0) atomicAdd(&sys[s0].natoms, 32);
                    if (sp.pbc[0] && sp.pbc[1] && sp.pbc[2]) continue;
                    const double px = (double)pos[3 * i], py = (double)pos[3 * i + 1], pz = (double)pos[3 * i + 2];
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        if (!sp.pbc[d]) {
                            const double fr = px * sp.inv[d] + py * sp.inv[3 + d] + pz * sp.inv[6 + d];
                            long long lo = order_key(fr), hi = lo;
                            for (int o = 16; o > 0; o >>= 1) {
                                const long long u = __shfl_xor_sync(0xffffffffu, lo, o), v = __shfl_xor_sync(0xffffffffu, hi, o);
                                lo = u < lo ? u : lo;
                                hi = v > hi ? v : hi;
                            }
                            if (lane == 0) {
                                atomicMin(&bbox[s0 * 6 + d], lo);
                                atomicMax(&bbox[s0 * 6 + 3 + d], hi);
                            }
                        }
                    }
                    continue;
                }
                if (!act) continue;
                const SysParams& sp = sys[s];
                if (need_counts) atomicAdd(&sys[s].natoms, 1);
                if (sp.pbc[0] && sp.pbc[1] && sp.pbc[2]) continue;
                const double px = (double)pos[3 * i], py = (double)pos[3 * i + 1], pz = (double)pos[3 * i + 2];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    if (!sp.pbc[d]) {
                        const double fr = px * sp.inv[d] + py * sp.inv[3 + d] + pz * sp.inv[6 + d];
                        const long long k = order_key(fr);
                        atomicMin(&bbox[s * 6 + d], k);
                        atomicMax(&bbox[s * 6 + 3 + d], k);
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_grid: cells per dimension, stencil radius and cell offsets of every system (single block).
//   periodic dim     : cpd = max(1, floor(face / rc_eff)),  R = ceil(rc_eff * cpd / face)
//   non-periodic dim : the atoms' bounding slab [fmin, fmax] is cut into cells of width >= rc_eff,
//                      R = 0 if one cell else 1
//   cap              : halve all dims (like the reference's max_nbins loop) until #cells <= #atoms (and <= the
//                      caller's per-system cap, if any), so the total never exceeds N + S and no host sync is
//                      needed to allocate.
// rc_eff = rc * (1 + 1e-3): the cell assignment is done in double, the distance test in the input
// precision; the margin guarantees every pair the fp test accepts lies inside the stencil.
// ------------------------------------------------------------------------------------------------
__global__ void k_grid(unsigned char* __restrict__ ws, WsLayout L, int num_systems, double cutoff, long long cell_cap) {
    pdl_enter();
    SysParams* sys = reinterpret_cast<SysParams*>(ws + L.sys);
    const long long* bbox = reinterpret_cast<const long long*>(ws + L.bbox);
    Ctrl* ctrl = reinterpret_cast<Ctrl*>(ws + L.ctrl);
    __shared__ int s_warp[32];   // (launched with up to 1024 threads: many small systems)
    __shared__ int s_carry;
    __shared__ unsigned long long s_atoms[2];   // atoms in cells at a periodic boundary (estimate), all atoms
    unsigned long long my_bnd = 0ull, my_all = 0ull;
    if (threadIdx.x == 0) { s_carry = 0; s_atoms[0] = s_atoms[1] = 0ull; }
    __syncthreads();
    const double rc = cutoff * (1.0 + 1e-3);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < num_systems; base += blockDim.x) {
        const int s = base + threadIdx.x;
        int ncells = 0;
        if (s < num_systems) {
            SysParams& sp = sys[s];
            long long tot = 1;
            for (int d = 0; d < 3; ++d) {
                int cpd = 1, R = 0;
                if (sp.pbc[d]) {
                    const double q = sp.face[d] / rc;
                    cpd = q >= 1.0 ? (q < 1.0e6 ? (int)q : 1000000) : 1;
                    const double rr = ceil(rc * (double)cpd / sp.face[d]);
                    R = (rr >= 1.0 && rr < 64.0) ? (int)rr : 1;
                    if (!(rr < 64.0)) atomicOr(&ctrl->error, ERR_IMAGE_RANGE);
                    sp.fmin[d] = 0.0;
                    sp.fscale[d] = (double)cpd;
                } else {
                    const long long kmn = bbox[s * 6 + d], kmx = bbox[s * 6 + 3 + d];
                    double fmin = 0.0, fmax = 0.0;
                    if (kmn <= kmx) { fmin = order_unkey(kmn); fmax = order_unkey(kmx); }
                    // capped mode (reference-shaped cache): grid the unit cell like the reference does, atoms outside
                    // are clamped into the edge cells (cell_list.py:227-232) — the binning is then a function of
                    // (cell, cells_per_dimension) alone and cell_list_needs_rebuild needs no hidden state
                    if (cell_cap > 0) { fmin = 0.0; fmax = 1.0; }
                    const double ext = (fmax - fmin) * sp.face[d];
                    const double q = ext / rc;
                    cpd = q >= 1.0 ? (q < 1.0e6 ? (int)q : 1000000) : 1;
                    R = cpd > 1 ? 1 : 0;
                    sp.fmin[d] = fmin;
                    sp.fscale[d] = (fmax > fmin) ? (double)cpd / (fmax - fmin) : 0.0;
                }
                sp.cpd[d] = cpd;
                sp.R[d] = R;
                if (R > 1) ctrl->wide_stencil = 1;
                tot *= cpd;
            }
            // caller-imposed cap (the reference-shaped cache holds cell_cap cells per system, cell_list.py:131-150)
            long long cap = sp.natoms > 1 ? sp.natoms : 1;
            if (cell_cap > 0 && cell_cap < cap) cap = cell_cap;
            while (tot > cap) {
                tot = 1;
                for (int d = 0; d < 3; ++d) {
                    const int old = sp.cpd[d];
                    const int nw = old / 2 > 1 ? old / 2 : 1;
                    if (sp.pbc[d]) {
                        sp.fscale[d] = (double)nw;
                    } else {
                        sp.fscale[d] = old > 0 ? sp.fscale[d] * (double)nw / (double)old : 0.0;
                        if (nw == 1) sp.R[d] = 0;
                    }
                    sp.cpd[d] = nw;
                    tot *= nw;
                }
            }
            ncells = (int)tot;
            sp.ncells = ncells;
            // share of the system's cells whose stencil crosses a periodic boundary (their neighbor rows carry shifts)
            long long interior = 1;
            for (int d = 0; d < 3; ++d) interior *= sp.pbc[d] ? (sp.cpd[d] > 2 ? sp.cpd[d] - 2 : 0) : sp.cpd[d];
            const long long na = sp.natoms > 0 ? sp.natoms : 0;
            my_bnd += (unsigned long long)(na * (tot - interior) / (tot > 0 ? tot : 1));
            my_all += (unsigned long long)na;
        }
        // block-wide exclusive scan of ncells with carry
        int incl = warp_incl_scan(ncells, lane);
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int v = lane < (int)blockDim.x / 32 ? s_warp[lane] : 0;
            v = warp_incl_scan(v, lane);
            if (lane < (int)blockDim.x / 32) s_warp[lane] = v;
        }
        __syncthreads();
        const int warp_off = wid > 0 ? s_warp[wid - 1] : 0;
        const int carry = s_carry;
        if (s < num_systems) sys[s].cell_offset = carry + warp_off + incl - ncells;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = carry + warp_off + incl;
        __syncthreads();
    }
    // (one shared-memory atomic per warp: 64-bit shared atomics serialise)
    for (int o = 16; o > 0; o >>= 1) {
        my_bnd += __shfl_down_sync(0xffffffffu, my_bnd, o);
        my_all += __shfl_down_sync(0xffffffffu, my_all, o);
    }
    if (lane == 0) {
        atomicAdd(&s_atoms[0], my_bnd);
        atomicAdd(&s_atoms[1], my_all);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ctrl->total_cells = s_carry;
        ctrl->shift_heavy = 2ull * s_atoms[0] > s_atoms[1] ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------------
// k_hash: atom -> (global cell id, rank inside the cell, periodic image of the atom).
// Four atoms per thread: three coalesced 16-byte loads cover four xyz triples (float path).
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void hash_one(const SysParams& sp, double px, double py, double pz, int& gcell,
                                         int4& ash, int& err) {
    int cc[3];
    int sh[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double fr = px * sp.inv[d] + py * sp.inv[3 + d] + pz * sp.inv[6 + d];
        if (sp.pbc[d]) {
            double c = floor(fr * (double)sp.cpd[d]);
            if (!(c > -1.0e9 && c < 1.0e9)) { c = 0.0; err |= ERR_IMAGE_RANGE; }
            int q, r;
            divmod_floor((int)c, sp.cpd[d], q, r);
            if (q < -1000000 || q > 1000000) err |= ERR_IMAGE_RANGE;
            sh[d] = q;
            cc[d] = r;
        } else {
            double c = floor((fr - sp.fmin[d]) * sp.fscale[d]);
            int ci = c > 0.0 ? (c < (double)(sp.cpd[d] - 1) ? (int)c : sp.cpd[d] - 1) : 0;
            sh[d] = 0;
            cc[d] = ci;
        }
    }
    gcell = sp.cell_offset + cc[0] + sp.cpd[0] * (cc[1] + sp.cpd[1] * cc[2]);
    ash = make_int4(sh[0], sh[1], sh[2], 0);
}

template <typename T, bool VEC>
__device__ __forceinline__ void load4(const T* __restrict__ pos, long long i0, long long n, T (&p)[12]) {
    if (VEC && i0 + 4 <= n) {
        if (sizeof(T) == 4) {
            const float4* v = reinterpret_cast<const float4*>(pos + 3 * i0);
            float4 a = __ldg(v), b = __ldg(v + 1), c = __ldg(v + 2);
            p[0] = a.x; p[1] = a.y; p[2] = a.z; p[3] = a.w; p[4] = b.x; p[5] = b.y; p[6] = b.z; p[7] = b.w;
            p[8] = c.x; p[9] = c.y; p[10] = c.z; p[11] = c.w;
        } else {
            const double2* v = reinterpret_cast<const double2*>(pos + 3 * i0);
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                double2 a = __ldg(v + k);
                p[2 * k] = a.x;
                p[2 * k + 1] = a.y;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < 12; ++k) p[k] = (3 * i0 + k < 3 * n) ? pos[3 * i0 + k] : (T)0;
    }
}

template <typename T, bool VEC>
__global__ void k_hash(unsigned char* __restrict__ ws, WsLayout L, long long n, int num_systems,
                       const T* __restrict__ pos, const int* __restrict__ batch_idx) {
    pdl_enter();
    const SysParams* sys = reinterpret_cast<const SysParams*>(ws + L.sys);
    Ctrl* ctrl = reinterpret_cast<Ctrl*>(ws + L.ctrl);
    int* cell_count = reinterpret_cast<int*>(ws + L.cell_count);
    int* atom_cell = reinterpret_cast<int*>(ws + L.atom_cell);
    int* atom_rank = reinterpret_cast<int*>(ws + L.atom_rank);
    int4* atom_ashift = reinterpret_cast<int4*>(ws + L.atom_ashift);
    const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= n) return;
    T p[12];
    load4<T, VEC>(pos, i0, n, p);
    int err = 0, unwrapped = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long long i = i0 + k;
        if (i < n) {
            int s = batch_idx ? batch_idx[i] : 0;
            if (s < 0 || s >= num_systems) { err |= ERR_BAD_BATCH_IDX; s = 0; }
            int gcell;
            int4 ash;
            hash_one<T>(sys[s], (double)p[3 * k], (double)p[3 * k + 1], (double)p[3 * k + 2], gcell, ash, err);
            const int rank = atomicAdd(&cell_count[gcell], 1);
            atom_cell[i] = gcell;
            atom_rank[i] = rank;
            atom_ashift[i] = ash;
            unwrapped |= (ash.x | ash.y | ash.z);
        }
    }
    if (unwrapped) atomicOr(&ctrl->unwrapped, 1);
    if (err) atomicOr(&ctrl->error, err);
}

// ------------------------------------------------------------------------------------------------
// k_scan: single-pass exclusive prefix sum with decoupled look-back (dynamic tile ids).
//   out[i] = sum(in[0..i)) for i in [0, n];  out has n+1 entries.  status must be zero on entry.
//   total64 (optional) receives the 64-bit sum (overflow check for int32 CSR pointers).
// Replaces torch.cumsum (cell_list.py:869-871, neighbor_utils.py:432-435).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScanThreads) k_scan(const int* __restrict__ in, int* __restrict__ out, long long n,
                                                       unsigned long long* __restrict__ status,
                                                       int* __restrict__ tile_counter,
                                                       unsigned long long* __restrict__ total64,
                                                       int* __restrict__ max_out, const int* __restrict__ live_ptr) {
    pdl_enter();
    __shared__ int s_tile;
    __shared__ int s_warp[kScanThreads / 32];
    __shared__ int s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1);
    __syncthreads();
    const int tile = s_tile;
    // inputs past *live_ptr are known to be zero and their prefixes are never read (cell histogram: only the first
    // total_cells entries are live, the array is sized for the N + S upper bound): those tiles retire at once
    if (live_ptr && (long long)tile * kScanTile > (long long)(*live_ptr) + 1) return;
    const long long base = (long long)tile * kScanTile + (long long)threadIdx.x * kScanItems;
    int v[kScanItems];
    int tsum = 0, tmax = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        tsum += v[k];
        tmax = v[k] > tmax ? v[k] : tmax;
    }
    if (max_out) {
        tmax = __reduce_max_sync(0xffffffffu, tmax);
        if ((threadIdx.x & 31) == 0 && tmax > 0) atomicMax(max_out, tmax);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = warp_incl_scan(tsum, lane);
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = lane < kScanThreads / 32 ? s_warp[lane] : 0;
        w = warp_incl_scan(w, lane);
        if (lane < kScanThreads / 32) s_warp[lane] = w;
    }
    __syncthreads();
    const int block_agg = s_warp[kScanThreads / 32 - 1];
    if (wid == 0) {
        // decoupled look-back, one warp wide: lane l inspects predecessor tile t - l; the window slides back by 32 tiles
        // until it contains an inclusive prefix (flag 2).  Flag and value share one 64-bit word (single store/load).
        volatile unsigned long long* st = status;
        if (lane == 0) st[tile] = ((tile == 0 ? 2ull : 1ull) << 32) | (unsigned)block_agg;
        int prefix = 0;
        if (tile > 0) {
            int t = tile - 1;
            for (;;) {
                const int idx = t - lane;
                const unsigned long long w = idx >= 0 ? st[idx] : (2ull << 32);   // before tile 0: inclusive prefix 0
                const unsigned flag = (unsigned)(w >> 32);
                const unsigned inc = __ballot_sync(0xffffffffu, flag == 2u);
                const unsigned none = __ballot_sync(0xffffffffu, flag == 0u);
                const int first = inc ? __ffs((int)inc) - 1 : 31;                    // nearest inclusive predecessor
                const unsigned need = first == 31 ? 0xffffffffu : ((2u << first) - 1u);
                if (none & need) continue;                                           // some aggregate not posted yet
                prefix += __reduce_add_sync(0xffffffffu, lane <= first ? (int)(unsigned)w : 0);
                if (inc) break;
                t -= 32;
            }
            if (lane == 0) st[tile] = (2ull << 32) | (unsigned)(prefix + block_agg);
        }
        if (lane == 0) {
            s_prefix = prefix;
            if (total64 && block_agg) atomicAdd(total64, (unsigned long long)(unsigned)block_agg);
        }
    }
    __syncthreads();
    int run = s_prefix + (wid > 0 ? s_warp[wid - 1] : 0) + incl - tsum;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k <= n) out[base + k] = run;
        run += v[k];
    }
}

// ------------------------------------------------------------------------------------------------
// k_scatter: counting-sort scatter of positions into per-cell runs of 16-byte records.
// ------------------------------------------------------------------------------------------------
template <typename T, bool VEC>
__global__ void k_scatter(unsigned char* __restrict__ ws, WsLayout L, long long n, const T* __restrict__ pos) {
    pdl_enter();
    const Ctrl* ctrl = reinterpret_cast<const Ctrl*>(ws + L.ctrl);
    const int* cell_start = reinterpret_cast<const int*>(ws + L.cell_start);
    const int* atom_cell = reinterpret_cast<const int*>(ws + L.atom_cell);
    const int* atom_rank = reinterpret_cast<const int*>(ws + L.atom_rank);
    const int4* atom_ashift = reinterpret_cast<const int4*>(ws + L.atom_ashift);
    Rec<T>* sorted = reinterpret_cast<Rec<T>*>(ws + L.sorted);
    int4* sorted_ashift = reinterpret_cast<int4*>(ws + L.sorted_ashift);
    const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= n) return;
    const bool unwrapped = ctrl->unwrapped != 0;
    T p[12];
    load4<T, VEC>(pos, i0, n, p);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long long i = i0 + k;
        if (i < n) {
            const int dest = cell_start[atom_cell[i]] + atom_rank[i];
            Rec<T> r;
            r.x = p[3 * k]; r.y = p[3 * k + 1]; r.z = p[3 * k + 2]; r.j = (int)i;
            sorted[dest] = r;
            if (unwrapped) sorted_ashift[dest] = atom_ashift[i];
        }
    }
}

}  // namespace nvnl
