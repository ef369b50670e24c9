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
        const int s = batch_idx ? batch_idx[i] : 0;
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
