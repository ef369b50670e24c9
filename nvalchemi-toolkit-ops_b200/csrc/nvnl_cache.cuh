// nvnl_cache.cuh — the reference-shaped cell-list cache and the rebuild checks.
//
//   k_export_cache       : fills the reference's 7-tensor cache (neighbor_utils.py:494-539) from the workspace, for
//                          callers that inspect it (values describe THIS implementation's grid)
//   k_refresh_positions  : re-gathers current positions into the cell-sorted records WITHOUT re-binning — the
//                          "query a stale cell list with moved atoms" workflow of query_cell_list
//                          (cell_list.py:1108-1192, docs/userguide/components/neighborlist.md:421-500)
//   k_cells_changed      : cell_list_needs_rebuild (rebuild_detection.py:36-121): any atom in another cell?
//   k_moved_beyond       : neighbor_list_needs_rebuild (rebuild_detection.py:168-217): any |r - r_ref| > skin?
#pragma once
#include "nvnl_build.cuh"

namespace nvnl {

template <typename T>
__global__ void k_export_cache(const unsigned char* __restrict__ ws, WsLayout L, long long n, int num_systems,
                               const int* __restrict__ batch_idx, int* __restrict__ cpd_out, int* __restrict__ radius_out,
                               int* __restrict__ atom_shifts, int* __restrict__ atom_cell, int* __restrict__ cell_count_out,
                               int* __restrict__ cell_start_out, long long cache_cells, int* __restrict__ cell_atom_list) {
    const SysParams* sys = reinterpret_cast<const SysParams*>(ws + L.sys);
    const int* cell_start = reinterpret_cast<const int*>(ws + L.cell_start);
    const int* a_cell = reinterpret_cast<const int*>(ws + L.atom_cell);
    const int4* a_shift = reinterpret_cast<const int4*>(ws + L.atom_ashift);
    const Rec<T>* sorted = reinterpret_cast<const Rec<T>*>(ws + L.sorted);
    const Ctrl* ctrl = reinterpret_cast<const Ctrl*>(ws + L.ctrl);
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long s = gid; s < num_systems; s += stride)
        for (int d = 0; d < 3; ++d) {
            if (cpd_out) cpd_out[3 * s + d] = sys[s].cpd[d];
            if (radius_out) radius_out[3 * s + d] = sys[s].R[d];
        }
    for (long long i = gid; i < n; i += stride) {
        int s = batch_idx ? batch_idx[i] : 0;
        s = s < 0 ? 0 : (s >= num_systems ? num_systems - 1 : s);   // (k_hash reported the error)
        const SysParams& sp = sys[s];
        const int local = a_cell[i] - sp.cell_offset;
        if (atom_cell) {
            atom_cell[3 * i] = local % sp.cpd[0];
            atom_cell[3 * i + 1] = (local / sp.cpd[0]) % sp.cpd[1];
            atom_cell[3 * i + 2] = local / (sp.cpd[0] * sp.cpd[1]);
        }
        if (atom_shifts) {
            const int4 a = a_shift[i];
            atom_shifts[3 * i] = a.x; atom_shifts[3 * i + 1] = a.y; atom_shifts[3 * i + 2] = a.z;
        }
        if (cell_atom_list) cell_atom_list[i] = sorted[i].j;
    }
    const int total_cells = ctrl->total_cells;
    for (long long c = gid; c < cache_cells; c += stride) {
        const bool in = c < total_cells;
        const int st = in ? cell_start[c] : 0;
        if (cell_start_out) cell_start_out[c] = st;
        if (cell_count_out) cell_count_out[c] = in ? cell_start[c + 1] - st : 0;
    }
}

// ------------------------------------------------------------------------------------------------
// Import: rebuild the workspace from the reference-shaped cache tensors (the VALUES of the seven tensors
// build_cell_list exported, cell_list.py:725-749) and the CURRENT positions, so that query_cell_list needs no hidden
// state: the cache tensors may be cloned, moved, serialised or traced by torch.compile in between.
//   k_import_sys   (one block)  grid of every system from cells_per_dimension; stencil radius = max(cached radius,
//                               radius the query cutoff needs on that grid); cell offsets; total number of cells
//   k_import_atoms              sorted records {position of cell_atom_list[k], cell_atom_list[k]}, periodic images,
//                               global cell ids, cell_start / cell_count
// ------------------------------------------------------------------------------------------------
__global__ void k_import_sys(unsigned char* __restrict__ ws, WsLayout L, int num_systems, double cutoff,
                             const int* __restrict__ cpd_in, const int* __restrict__ radius_in, long long cache_cells) {
    SysParams* sys = reinterpret_cast<SysParams*>(ws + L.sys);
    Ctrl* ctrl = reinterpret_cast<Ctrl*>(ws + L.ctrl);
    __shared__ int s_warp[kSmallBlock / 32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
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
                int cpd = cpd_in[3 * s + d];
                if (cpd < 1 || cpd > 1000000) { cpd = 1; atomicOr(&ctrl->error, ERR_BAD_CACHE); }
                int R = radius_in ? radius_in[3 * s + d] : 0;
                if (R < 0 || R >= 64) { R = 1; atomicOr(&ctrl->error, ERR_BAD_CACHE); }
                int need;
                if (sp.pbc[d]) {
                    const double rr = ceil(rc * (double)cpd / sp.face[d]);
                    need = (rr >= 1.0 && rr < 64.0) ? (int)rr : 1;
                    if (!(rr < 64.0)) atomicOr(&ctrl->error, ERR_IMAGE_RANGE);
                } else {
                    // capped builds grid the unit cell along open dimensions too (atoms outside sit in the edge cells)
                    const double rr = ceil(rc * (double)cpd / sp.face[d]);
                    need = (rr >= 1.0 && rr < 64.0) ? (int)rr : 1;
                    need = need < cpd - 1 ? need : cpd - 1;
                }
                sp.cpd[d] = cpd;
                sp.R[d] = R > need ? R : need;
                if (sp.R[d] > 1) ctrl->wide_stencil = 1;
                sp.fmin[d] = 0.0;
                sp.fscale[d] = sp.pbc[d] ? (double)cpd : 0.0;
                tot *= cpd;
            }
            if (tot > 2000000000LL) { tot = 1; atomicOr(&ctrl->error, ERR_BAD_CACHE); }
            ncells = (int)tot;
            sp.ncells = ncells;
        }
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
    if (threadIdx.x == 0) {
        int total = s_carry;
        if ((long long)total > cache_cells || (long long)total > L.max_cells) {
            total = 0;                                    // nothing is swept; the host raises on the error bit
            atomicOr(&ctrl->error, ERR_BAD_CACHE);
        }
        ctrl->total_cells = total;
    }
}

template <typename T>
__global__ void k_import_atoms(unsigned char* __restrict__ ws, WsLayout L, long long n, int num_systems,
                               const T* __restrict__ pos, const int* __restrict__ batch_idx,
                               const int* __restrict__ atom_shifts, const int* __restrict__ atom_cell_map,
                               const int* __restrict__ cell_count_in, const int* __restrict__ cell_start_in,
                               const int* __restrict__ cell_atom_list) {
    const SysParams* sys = reinterpret_cast<const SysParams*>(ws + L.sys);
    Ctrl* ctrl = reinterpret_cast<Ctrl*>(ws + L.ctrl);
    int* cell_count = reinterpret_cast<int*>(ws + L.cell_count);
    int* cell_start = reinterpret_cast<int*>(ws + L.cell_start);
    int* atom_cell = reinterpret_cast<int*>(ws + L.atom_cell);
    int4* atom_ashift = reinterpret_cast<int4*>(ws + L.atom_ashift);
    Rec<T>* sorted = reinterpret_cast<Rec<T>*>(ws + L.sorted);
    int4* sorted_ashift = reinterpret_cast<int4*>(ws + L.sorted_ashift);
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int total_cells = ctrl->total_cells;
    int err = 0, unwrapped = 0;
    for (long long k = gid; k < n; k += stride) {
        int i = cell_atom_list[k];
        if (i < 0 || i >= n) { i = 0; err |= ERR_BAD_CACHE; }
        Rec<T> r;
        r.x = pos[3 * (long long)i]; r.y = pos[3 * (long long)i + 1]; r.z = pos[3 * (long long)i + 2]; r.j = i;
        sorted[k] = r;
        const int4 a = make_int4(atom_shifts[3 * (long long)i], atom_shifts[3 * (long long)i + 1], atom_shifts[3 * (long long)i + 2], 0);
        sorted_ashift[k] = a;
        unwrapped |= (a.x | a.y | a.z);
    }
    for (long long i = gid; i < n; i += stride) {
        int s = batch_idx ? batch_idx[i] : 0;
        if (s < 0 || s >= num_systems) { s = 0; err |= ERR_BAD_BATCH_IDX; }
        const SysParams& sp = sys[s];
        int cx = atom_cell_map[3 * i], cy = atom_cell_map[3 * i + 1], cz = atom_cell_map[3 * i + 2];
        if (cx < 0 || cx >= sp.cpd[0] || cy < 0 || cy >= sp.cpd[1] || cz < 0 || cz >= sp.cpd[2]) { cx = cy = cz = 0; err |= ERR_BAD_CACHE; }
        atom_cell[i] = sp.cell_offset + cx + sp.cpd[0] * (cy + sp.cpd[1] * cz);
        atom_ashift[i] = make_int4(atom_shifts[3 * i], atom_shifts[3 * i + 1], atom_shifts[3 * i + 2], 0);
    }
    for (long long c = gid; c <= total_cells; c += stride) {
        if (c == total_cells) { cell_start[c] = (int)n; cell_count[c] = 0; break; }
        const int st = cell_start_in[c], cn = cell_count_in[c];
        if (st < 0 || cn < 0 || (long long)st + cn > n) err |= ERR_BAD_CACHE;
        cell_start[c] = st < 0 ? 0 : (st > n ? (int)n : st);
        cell_count[c] = cn;
    }
    if (unwrapped) atomicOr(&ctrl->unwrapped, 1);
    if (err) atomicOr(&ctrl->error, err);
}

// cell_list_needs_rebuild from the cache VALUES (rebuild_detection.py:36-121): re-hash the current positions on the grid
// (cell, pbc, cells_per_dimension) exactly like k_hash does for a capped build and compare with atom_to_cell_mapping.
template <typename T>
__global__ void k_cells_changed_cache(long long n, int num_systems, const T* __restrict__ pos, const T* __restrict__ cell,
                                      const unsigned char* __restrict__ pbc, const int* __restrict__ batch_idx,
                                      const int* __restrict__ cpd_in, const int* __restrict__ atom_cell_map,
                                      int* __restrict__ flag) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int s = batch_idx ? batch_idx[i] : 0;
    if (s < 0 || s >= num_systems) s = 0;
    double m[9];
    for (int k = 0; k < 9; ++k) m[k] = (double)cell[(long long)s * 9 + k];
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], q = m[8];
    const double det = a * (e * q - f * h) - b * (d * q - f * g) + c * (d * h - e * g);
    const double r = det != 0.0 ? 1.0 / det : 0.0;
    double inv[9];
    inv[0] = (e * q - f * h) * r; inv[1] = (c * h - b * q) * r; inv[2] = (b * f - c * e) * r;
    inv[3] = (f * g - d * q) * r; inv[4] = (a * q - c * g) * r; inv[5] = (c * d - a * f) * r;
    inv[6] = (d * h - e * g) * r; inv[7] = (b * g - a * h) * r; inv[8] = (a * e - b * d) * r;
    const double px = (double)pos[3 * i], py = (double)pos[3 * i + 1], pz = (double)pos[3 * i + 2];
    bool changed = false;
    for (int dd = 0; dd < 3; ++dd) {
        const int cpd = cpd_in[3 * s + dd] > 0 ? cpd_in[3 * s + dd] : 1;
        const double fr = px * inv[dd] + py * inv[3 + dd] + pz * inv[6 + dd];
        double cf = floor(fr * (double)cpd);
        int ci;
        if (pbc[3 * s + dd]) {
            if (!(cf > -1.0e9 && cf < 1.0e9)) cf = 0.0;
            int qq, rr;
            divmod_floor((int)cf, cpd, qq, rr);
            ci = rr;
        } else {
            ci = cf > 0.0 ? (cf < (double)(cpd - 1) ? (int)cf : cpd - 1) : 0;
        }
        changed = changed || (ci != atom_cell_map[3 * i + dd]);
    }
    if (changed) *flag = 1;
}

template <typename T>
__global__ void k_refresh_positions(unsigned char* __restrict__ ws, WsLayout L, long long n, const T* __restrict__ pos) {
    Rec<T>* sorted = reinterpret_cast<Rec<T>*>(ws + L.sorted);
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    Rec<T> r = sorted[k];
    const long long j = r.j;
    r.x = pos[3 * j]; r.y = pos[3 * j + 1]; r.z = pos[3 * j + 2];
    sorted[k] = r;
}

template <typename T>
__global__ void k_cells_changed(const unsigned char* __restrict__ ws, WsLayout L, long long n, int num_systems,
                                const T* __restrict__ pos, const int* __restrict__ batch_idx, int* __restrict__ flag) {
    const SysParams* sys = reinterpret_cast<const SysParams*>(ws + L.sys);
    const int* a_cell = reinterpret_cast<const int*>(ws + L.atom_cell);
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (*reinterpret_cast<volatile int*>(flag)) return;  // early out once flagged (reference :100-102)
    int s = batch_idx ? batch_idx[i] : 0;
    if (s < 0 || s >= num_systems) s = 0;
    int gcell, err = 0;
    int4 ash;
    hash_one<T>(sys[s], (double)pos[3 * i], (double)pos[3 * i + 1], (double)pos[3 * i + 2], gcell, ash, err);
    if (gcell != a_cell[i]) *flag = 1;
}

template <typename T>
__global__ void k_moved_beyond(const T* __restrict__ ref, const T* __restrict__ cur, long long n, T threshold,
                               int* __restrict__ flag) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (*reinterpret_cast<volatile int*>(flag)) return;
    const T dx = cur[3 * i] - ref[3 * i], dy = cur[3 * i + 1] - ref[3 * i + 1], dz = cur[3 * i + 2] - ref[3 * i + 2];
    const T len = sqrt(dx * dx + dy * dy + dz * dz);  // wp.length: sqrt(dot(v, v))
    if (len > threshold) *flag = 1;
}

}  // namespace nvnl
