// nvnl_sweep.cuh — the stencil sweep: count pass, COO fill pass, padded-matrix fill pass.
//
// Replaces cell_list.py:372-556 / batch_cell_list.py:380-569 (_cell_list_build_neighbor_matrix),
// neighbor_utils.py:106-147 (_update_neighbor_matrix_pbc) and, for COO output,
// neighbor_utils.py:362-441 (get_neighbor_list_from_neighbor_matrix) — with a different algorithm:
//
//   * persistent CTAs pull target cells from a device-side queue;
//   * the target cell's stencil (every periodic IMAGE of every neighbor cell, so small boxes and
//     search radii > 1 are covered) is sorted by image shift and its runs of 16-byte records are
//     concatenated in shared memory by 1-D TMA bulk copies (cp.async.bulk + mbarrier);
//   * FULL stencil, one warp per target atom, lanes over candidates: every row is produced by
//     exactly one warp, so there are no global atomics and row contents are written once;
//   * hits are compacted with __ballot_sync + popc prefix into a per-warp staging row and flushed
//     as coalesced stores (COO rows at neighbor_ptr[i], or matrix rows incl. the padding, which
//     fuses the reference's fill_()/zero_() memsets into the same pass).
//
// The result is the same set {(i, j, s)} the reference produces (SURVEY.md §8a "semantics"):
// ||r_j - r_i + s·cell||^2 < rc^2 evaluated in the input precision with the reference's operation
// order, s = 0 in non-periodic dims, (i, i, 0) excluded.
#pragma once
#include "nvnl_build.cuh"

namespace nvnl {

template <typename T>
struct SweepArgs {
    unsigned char* ws;
    WsLayout L;
    const int* batch_idx;
    int num_systems;
    long long n;
    T cutoff_sq;
    int* num_neighbors;       // COUNT: out.  FILL_MATRIX: out.
    const int* neighbor_ptr;  // FILL_COO: in (exclusive scan of num_neighbors, N+1 entries)
    int* out_i;               // FILL_COO: edge_index row 0
    int* out_j;               // FILL_COO: edge_index row 1
    int* out_shifts;          // FILL_COO: [P,3];  FILL_MATRIX: [N,M,3]
    int* neighbor_matrix;     // FILL_MATRIX: [N,M]
    int max_neighbors;
    int fill_value;
    int index_offset;         // added to every atom index written to COO outputs (rank sharding)
    int queue;                // which Ctrl::work_counter this launch uses
    int pad;                  // FILL_MATRIX: write fill_value / zero shifts into the unused slots of every row
    int keep_deferred;        // COUNT of the single-sweep COO path: leave the deferred list for the FILL launch (wrapped inputs)
};

constexpr int kKeyEmpty = 0x7fffffff;
constexpr int kZeroPack = 128 | (128 << 8) | (128 << 16);

__device__ __forceinline__ int pack_key(int csx, int csy, int csz) {
    const int p = (csx + 128) | ((csy + 128) << 8) | ((csz + 128) << 16);
    return p == kZeroPack ? 0 : p + 1;
}
__device__ __forceinline__ void unpack_key(int key, int& csx, int& csy, int& csz) {
    if (key == 0) { csx = csy = csz = 0; return; }
    const int p = key - 1;
    csx = (p & 255) - 128;
    csy = ((p >> 8) & 255) - 128;
    csz = ((p >> 16) & 255) - 128;
}

struct SweepSmem {
    // image / piece tables (indices: position in the shift-sorted order)
    int ustart[kMaxImg], ucnt[kMaxImg], ukey[kMaxImg];
    int sstart[kMaxImg], scnt[kMaxImg], skey[kMaxImg];
    int seg_begin[kMaxImg + 1], seg_end[kMaxImg], seg_key[kMaxImg];
    int sfirst[kMaxImg], sdst[kMaxImg];  // aliasing plan: first image of the same cell, its tile offset
    int nseg, total, item, more, pc_img, pc_off, aliased;
    unsigned long long mbar;
};

// ---- flush of one warp's staged hits -----------------------------------------------------------
template <typename T, int MODE>
__device__ __forceinline__ void flush_staged(const SweepArgs<T>& a, int lane, int i, size_t row_base, int written,
                                             int staged, int nz, const int* __restrict__ row_j,
                                             const int* __restrict__ row_s) {
    if (MODE == MODE_FILL_COO) {
        const size_t p0 = row_base + (size_t)written;
        const int iv = i + a.index_offset;
        for (int k = lane; k < staged; k += 32) {
            a.out_i[p0 + k] = iv;
            a.out_j[p0 + k] = row_j[k] + a.index_offset;
        }
        int* sh = a.out_shifts + 3 * p0;
        if (staged == nz) {
            for (int e = lane; e < 3 * staged; e += 32) sh[e] = 0;
        } else {
            for (int e = lane; e < 3 * staged; e += 32) {
                const int k = e / 3, c = e - 3 * k;
                sh[e] = k < nz ? 0 : row_s[c * kRowCap + k];
            }
        }
    } else if (MODE == MODE_FILL_MATRIX) {
        const int M = a.max_neighbors;
        int room = M - written;
        room = room < 0 ? 0 : room;
        const int ns = staged < room ? staged : room;
        const size_t p0 = (size_t)i * (size_t)M + (size_t)written;
        for (int k = lane; k < ns; k += 32) a.neighbor_matrix[p0 + k] = row_j[k];
        int* sh = a.out_shifts + 3 * p0;
        for (int e = lane; e < 3 * ns; e += 32) {
            const int k = e / 3, c = e - 3 * k;
            sh[e] = k < nz ? 0 : row_s[c * kRowCap + k];
        }
    }
}

// pad matrix row i from `used` to M (fill_value / zero shifts) and store the neighbor count
template <typename T>
__device__ __forceinline__ void finish_matrix_row(const SweepArgs<T>& a, int lane, int i, int total) {
    const int M = a.max_neighbors;
    const int used = total < M ? total : M;
    const size_t p0 = (size_t)i * (size_t)M;
    if (a.pad) {
        for (int k = used + lane; k < M; k += 32) a.neighbor_matrix[p0 + k] = a.fill_value;
        int* sh = a.out_shifts + 3 * p0;
        for (int e = 3 * used + lane; e < 3 * M; e += 32) sh[e] = 0;
    }
    if (lane == 0) a.num_neighbors[i] = total;
}

// ---- one target atom against one staged tile ---------------------------------------------------
// Returns the number of hits found in this tile (COUNT) / appended (FILL).
template <typename T, int MODE, bool HALF, bool UNW, bool FMA>
__device__ __forceinline__ int sweep_target(const SweepArgs<T>& a, const SweepSmem& sm, const Rec<T>* __restrict__ cand,
                                            const int4* __restrict__ cand_ash, const T* __restrict__ cm,
                                            const int* __restrict__ pbc, const Rec<T>& ti, const int4& ai,
                                            int lane, size_t row_base, int written_in, int* __restrict__ row_j,
                                            int* __restrict__ row_s) {
    using A = Arith<T>;
    const T xi = ti.x, yi = ti.y, zi = ti.z;
    const int i = ti.j;
    const T rc2 = a.cutoff_sq;
    const unsigned lt_mask = (1u << lane) - 1u;
    int cnt = 0;            // COUNT: per-lane hit counter
    int staged = 0, nz = 0; // FILL: warp-uniform staging state
    int written = written_in;
    const int nseg = sm.nseg;
    for (int sg = 0; sg < nseg; ++sg) {
        const int b = sm.seg_begin[sg], e = sm.seg_end[sg];
        const int key = sm.seg_key[sg];
        int csx, csy, csz;
        unpack_key(key, csx, csy, csz);
        const bool zero = (key == 0);
        T Sx = (T)0, Sy = (T)0, Sz = (T)0;
        if (!UNW && !zero) shift_vector<T, FMA>(cm, csx, csy, csz, Sx, Sy, Sz);
        const bool seg_lexpos = csx > 0 || (csx == 0 && (csy > 0 || (csy == 0 && csz > 0)));
#pragma unroll 2
        for (int c0 = b; c0 < e; c0 += 32) {
            const int c = c0 + lane;
            const bool valid = c < e;
            const int cc = valid ? c : b;
            const Rec<T> r = cand[cc];
            int sx = csx, sy = csy, sz = csz;
            T dx, dy, dz;
            bool szero = zero;
            if (UNW) {
                const int4 aj = cand_ash[cc];
                sx = pbc[0] ? csx + ai.x - aj.x : 0;
                sy = pbc[1] ? csy + ai.y - aj.y : 0;
                sz = pbc[2] ? csz + ai.z - aj.z : 0;
                szero = (sx | sy | sz) == 0;
                T lx, ly, lz;
                shift_vector<T, FMA>(cm, sx, sy, sz, lx, ly, lz);
                dx = A::add(A::sub(r.x, xi), lx);
                dy = A::add(A::sub(r.y, yi), ly);
                dz = A::add(A::sub(r.z, zi), lz);
            } else if (zero) {
                // (r_j - r_i) + 0 == r_j - r_i bit for bit (up to the sign of zero, which squares away)
                dx = A::sub(r.x, xi);
                dy = A::sub(r.y, yi);
                dz = A::sub(r.z, zi);
            } else {
                dx = A::add(A::sub(r.x, xi), Sx);
                dy = A::add(A::sub(r.y, yi), Sy);
                dz = A::add(A::sub(r.z, zi), Sz);
            }
            const T d2 = dist2<T, FMA>(dx, dy, dz);
            bool hit = valid && (d2 < rc2);
            hit = hit && !(szero && r.j == i);  // (i, i, 0) is not a pair
            if (HALF) {
                const bool lexpos =
                    UNW ? (sx > 0 || (sx == 0 && (sy > 0 || (sy == 0 && sz > 0)))) : seg_lexpos;
                hit = hit && (i < r.j || (i == r.j && lexpos));
            }
            if (MODE == MODE_COUNT) {
                cnt += hit ? 1 : 0;
            } else {
                const unsigned mask = __ballot_sync(0xffffffffu, hit);
                if (mask) {
                    const int pos = staged + __popc(mask & lt_mask);
                    if (hit) {
                        row_j[pos] = r.j;
                        if (UNW || !zero) {
                            row_s[pos] = sx;
                            row_s[kRowCap + pos] = sy;
                            row_s[2 * kRowCap + pos] = sz;
                        }
                    }
                    const int h = __popc(mask);
                    staged += h;
                    if (!UNW && zero) nz += h;
                    if (staged > kRowCap - 32) {
                        __syncwarp();
                        flush_staged<T, MODE>(a, lane, i, row_base, written, staged, nz, row_j, row_s);
                        written += staged;
                        staged = 0;
                        nz = 0;
                        __syncwarp();
                    }
                }
            }
        }
    }
    if (MODE == MODE_COUNT) {
        return __reduce_add_sync(0xffffffffu, cnt);
    } else {
        if (staged) {
            __syncwarp();
            flush_staged<T, MODE>(a, lane, i, row_base, written, staged, nz, row_j, row_s);
            written += staged;
            __syncwarp();
        }
        return written - written_in;
    }
}

// ------------------------------------------------------------------------------------------------
// k_sweep
// ------------------------------------------------------------------------------------------------
template <typename T, int MODE, bool HALF, bool FMA>
__global__ void __launch_bounds__(kSweepThreads, 3) k_sweep(const SweepArgs<T> a) {
    pdl_enter();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Rec<T>* cand = reinterpret_cast<Rec<T>*>(smem_raw);
    int* rows = reinterpret_cast<int*>(smem_raw + kSweepCandBytes);
    SweepSmem& sm = *reinterpret_cast<SweepSmem*>(smem_raw + kSweepCandBytes + kSweepWarps * kRowCap * 4 * sizeof(int));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int* row_j = rows + warp * (4 * kRowCap);
    int* row_s = row_j + kRowCap;

    Ctrl* ctrl = reinterpret_cast<Ctrl*>(a.ws + a.L.ctrl);
    const SysParams* sys = reinterpret_cast<const SysParams*>(a.ws + a.L.sys);
    const int* cell_count = reinterpret_cast<const int*>(a.ws + a.L.cell_count);
    const int* cell_start = reinterpret_cast<const int*>(a.ws + a.L.cell_start);
    const Rec<T>* sorted = reinterpret_cast<const Rec<T>*>(a.ws + a.L.sorted);
    const int4* sorted_ashift = reinterpret_cast<const int4*>(a.ws + a.L.sorted_ashift);
    int* cursor = reinterpret_cast<int*>(a.ws + a.L.cursor);

    const bool unwrapped = ctrl->unwrapped != 0;
    // work source: the (cell, first target) items the fast kernels (nvnl_fast.cuh) deferred — too many images,
    // candidates or target atoms for their single shared-memory tile
    const int2* deferred = reinterpret_cast<const int2*>(a.ws + a.L.deferred);
    const int total_items = ctrl->n_deferred;
    // unwrapped inputs also stage each candidate's periodic image (int4): half the record capacity
    const int cap = unwrapped ? (kSweepCandBytes / 2) / (int)sizeof(Rec<T>) : kSweepCandBytes / (int)sizeof(Rec<T>);
    int4* cand_ash = reinterpret_cast<int4*>(smem_raw + kSweepCandBytes / 2);

    if (tid == 0) {
        mbar_init(reinterpret_cast<uint64_t*>(&sm.mbar), 1);
        mbar_fence_init();
    }
    if (MODE == MODE_COUNT) {
        // re-arm the look-back scan that turns the counts into neighbor_ptr (runs after this kernel)
        unsigned long long* st1 = reinterpret_cast<unsigned long long*>(a.ws + a.L.scan_status1);
        const long long nst = (a.n + 1) / kScanTile + 2;
        for (long long k = (long long)blockIdx.x * blockDim.x + tid; k < nst; k += (long long)gridDim.x * blockDim.x)
            st1[k] = 0ull;
        if (blockIdx.x == 0 && tid == 0) {
            ctrl->scan_tile[1] = 0;
            ctrl->total_pairs = 0ull;
            ctrl->max_count = 0;
        }
    }
    __syncthreads();
    uint32_t phase = 0;

    for (;;) {
        if (tid == 0) sm.item = atomicAdd(&ctrl->work_counter[a.queue], 1);
        __syncthreads();
        const int item = sm.item;
        if (item >= total_items) break;
        const int2 it = deferred[item];
        const int g = it.x;
        const int ncell_atoms = cell_count[g];
        // targets of this work item: a kDeferTargets slice of the cell
        const int t_begin = it.y;
        const int ntarget = t_begin + kDeferTargets < ncell_atoms ? t_begin + kDeferTargets : ncell_atoms;
        const int home_start = cell_start[g];
        if (ntarget == 0) {
            __syncthreads();
            continue;
        }
        // ---- system of this cell and its grid ----
        const int j0 = sorted[home_start].j;
        int s = a.batch_idx ? a.batch_idx[j0] : 0;
        s = s < 0 ? 0 : (s >= a.num_systems ? a.num_systems - 1 : s);   // (k_hash reported the error)
        const SysParams& sp = sys[s];
        const int cpd0 = sp.cpd[0], cpd1 = sp.cpd[1], cpd2 = sp.cpd[2];
        const int R0 = sp.R[0], R1 = sp.R[1], R2 = sp.R[2];
        int pbc[3] = {sp.pbc[0], sp.pbc[1], sp.pbc[2]};
        const int coff = sp.cell_offset;
        T cm[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) cm[k] = (T)sp.cellm[k];
        const int local = g - coff;
        const int cx = local % cpd0, cy = (local / cpd0) % cpd1, cz = local / (cpd0 * cpd1);
        const int nx = 2 * R0 + 1, ny = 2 * R1 + 1, nz_ = 2 * R2 + 1;
        const int nimg = nx * ny * nz_;
        const bool multi_batch = nimg > kMaxImg;
        bool multi = multi_batch;
        bool first_tile = true;

        for (int img_base = 0; img_base < nimg; img_base += kMaxImg) {
            const int nb = (nimg - img_base) < kMaxImg ? (nimg - img_base) : kMaxImg;
            // ---- enumerate the images of this batch ----
            bool has_shift = false;
            if (tid < kMaxImg) {
                int st = 0, cn = 0, key = kKeyEmpty;
                if (tid < nb) {
                    const int m = img_base + tid;
                    const int dx = m % nx - R0, dy = (m / nx) % ny - R1, dz = m / (nx * ny) - R2;
                    int tx = cx + dx, ty = cy + dy, tz = cz + dz;
                    bool ok = true;
                    int csx = 0, csy = 0, csz = 0;
                    if (pbc[0]) divmod_floor(tx, cpd0, csx, tx); else ok = ok && tx >= 0 && tx < cpd0;
                    if (pbc[1]) divmod_floor(ty, cpd1, csy, ty); else ok = ok && ty >= 0 && ty < cpd1;
                    if (pbc[2]) divmod_floor(tz, cpd2, csz, tz); else ok = ok && tz >= 0 && tz < cpd2;
                    if (ok) {
                        const int gc = coff + tx + cpd0 * (ty + cpd1 * tz);
                        cn = cell_count[gc];
                        st = cell_start[gc];
                        if (cn > 0) key = pack_key(csx, csy, csz);
                    }
                }
                sm.ustart[tid] = st;
                sm.ucnt[tid] = cn;
                sm.ukey[tid] = key;
                has_shift = (key != 0 && key != kKeyEmpty);
            }
            const int any_shift = __syncthreads_or(has_shift ? 1 : 0);
            // ---- order images by shift so that equal shifts form one contiguous segment ----
            if (tid < kMaxImg) {
                int rank = tid;
                if (any_shift) {
                    const int kmine = sm.ukey[tid];
                    rank = 0;
                    for (int m = 0; m < kMaxImg; ++m) {
                        const int km = sm.ukey[m];
                        rank += (km < kmine || (km == kmine && m < tid)) ? 1 : 0;
                    }
                }
                sm.sstart[rank] = sm.ustart[tid];
                sm.scnt[rank] = sm.ucnt[tid];
                sm.skey[rank] = sm.ukey[tid];
            }
            if (tid == 0) { sm.pc_img = 0; sm.pc_off = 0; sm.aliased = 0; }
            __syncthreads();
            // Small periodic boxes repeat the same cell under several shifts (box < 2 rc: 27+ images of one cell).
            // Find, for every image, the first image of the same cell: such a batch can be staged once per DISTINCT
            // cell, with one segment per image aliasing the shared records.
            if (!multi_batch && tid < kMaxImg) {
                const int st = sm.sstart[tid], cn = sm.scnt[tid];
                int f = tid;
                if (cn > 0)
                    for (int q = 0; q < tid; ++q)
                        if (sm.scnt[q] > 0 && sm.sstart[q] == st) { f = q; break; }
                sm.sfirst[tid] = f;
            }
            __syncthreads();
            if (!multi_batch && warp == 0) {
                int uc[4], cn4[4];
                int lsum = 0, lall = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int p = lane * 4 + q;
                    cn4[q] = sm.scnt[p];
                    uc[q] = (cn4[q] > 0 && sm.sfirst[p] == p) ? cn4[q] : 0;
                    lsum += uc[q];
                    lall += cn4[q];
                }
                const int incl = warp_incl_scan(lsum, lane);
                const int total_unique = __shfl_sync(0xffffffffu, incl, 31);
                const int total_all = __reduce_add_sync(0xffffffffu, lall);
                // only worth it when the plain concatenation would not fit one tile but the distinct cells do
                if (total_all > cap && total_unique <= cap && total_unique > 0) {
                    int run = incl - lsum;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        sm.sdst[lane * 4 + q] = run;
                        run += uc[q];
                    }
                    __syncwarp();
                    // one segment per non-empty image
                    int lseg = 0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) lseg += cn4[q] > 0 ? 1 : 0;
                    const int sincl = warp_incl_scan(lseg, lane);
                    int si = sincl - lseg;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int p = lane * 4 + q;
                        if (cn4[q] > 0) {
                            const int b = sm.sdst[sm.sfirst[p]];
                            sm.seg_begin[si] = b;
                            sm.seg_end[si] = b + cn4[q];
                            sm.seg_key[si] = sm.skey[p];
                            ++si;
                        }
                    }
                    if (lane == 31) {
                        sm.nseg = sincl;
                        sm.total = total_unique;
                        sm.more = 0;
                        sm.aliased = 1;
                    }
                    if (lane == 0) {
                        const uint32_t bytes =
                            (uint32_t)total_unique * (uint32_t)(sizeof(Rec<T>) + (unwrapped ? sizeof(int4) : 0));
                        mbar_arrive_expect_tx(reinterpret_cast<uint64_t*>(&sm.mbar), bytes);
                    }
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int p = lane * 4 + q;
                        if (uc[q] > 0) {
                            tma_load_1d(cand + sm.sdst[p], sorted + sm.sstart[p], (uint32_t)uc[q] * (uint32_t)sizeof(Rec<T>),
                                        reinterpret_cast<uint64_t*>(&sm.mbar));
                            if (unwrapped)
                                tma_load_1d(cand_ash + sm.sdst[p], sorted_ashift + sm.sstart[p],
                                            (uint32_t)uc[q] * (uint32_t)sizeof(int4), reinterpret_cast<uint64_t*>(&sm.mbar));
                        }
                    }
                }
            }
            __syncthreads();
            const bool aliased = sm.aliased != 0;

            // ---- tiles of this batch ----
            for (;;) {
                // plan (warp 0): which pieces of which images go into this tile, and where
                if (warp == 0 && !aliased) {
                    const int pc_img = sm.pc_img, pc_off = sm.pc_off;
                    int avail[4], take[4], dst[4], src[4], keyv[4];
                    int lsum = 0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int p = lane * 4 + q;
                        const int cn = sm.scnt[p];
                        avail[q] = p < pc_img ? 0 : (p == pc_img ? cn - pc_off : cn);
                        src[q] = sm.sstart[p] + (p == pc_img ? pc_off : 0);
                        keyv[q] = sm.skey[p];
                        lsum += avail[q];
                    }
                    const int incl = warp_incl_scan(lsum, lane);
                    int run = incl - lsum;
                    int my_take = 0;
                    int new_img = kMaxImg, new_off = 0;  // first entry left (partly) unconsumed
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        int t = cap - run;
                        t = t < 0 ? 0 : (t > avail[q] ? avail[q] : t);
                        take[q] = t;
                        dst[q] = run < cap ? run : cap;
                        if (t < avail[q] && new_img == kMaxImg) {
                            new_img = lane * 4 + q;
                            new_off = (lane * 4 + q == pc_img ? pc_off : 0) + t;
                        }
                        run += avail[q];
                        my_take += t;
                    }
                    const int total = __reduce_add_sync(0xffffffffu, my_take);
                    // cursor: the lowest unconsumed entry over the warp
                    const int min_img = __reduce_min_sync(0xffffffffu, new_img);
                    // segments: runs of equal key among the taken pieces (taken pieces are contiguous
                    // in the sorted order, empties sort last)
                    {
                        const int pk3 = __shfl_up_sync(0xffffffffu, keyv[3], 1);
                        const int pt3 = __shfl_up_sync(0xffffffffu, take[3], 1);
                        int heads[4];
                        int lheads = 0;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int prev_key = q == 0 ? pk3 : keyv[q - 1];
                            const int prev_take = q == 0 ? (lane == 0 ? 0 : pt3) : take[q - 1];
                            heads[q] = (take[q] > 0 && (prev_take == 0 || prev_key != keyv[q])) ? 1 : 0;
                            lheads += heads[q];
                        }
                        const int hincl = warp_incl_scan(lheads, lane);
                        int segi = hincl - lheads;
                        if (any_shift) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if (heads[q]) {
                                    sm.seg_begin[segi] = dst[q];
                                    sm.seg_key[segi] = keyv[q];
                                    ++segi;
                                }
                            }
                        }
                        const int nseg = __shfl_sync(0xffffffffu, hincl, 31);
                        __syncwarp();
                        if (new_img == min_img && min_img < kMaxImg) {
                            sm.pc_img = new_img;  // exactly one lane owns entry min_img
                            sm.pc_off = new_off;
                        }
                        if (lane == 0) {
                            if (any_shift) {
                                sm.nseg = nseg;
                                sm.seg_begin[nseg] = total;
                            } else {
                                sm.nseg = total > 0 ? 1 : 0;
                                sm.seg_begin[0] = 0;
                                sm.seg_begin[1] = total;
                                sm.seg_key[0] = 0;
                            }
                            sm.total = total;
                            sm.more = (min_img < kMaxImg) ? 1 : 0;
                        }
                        __syncwarp();
                        const int ns_ = any_shift ? nseg : (total > 0 ? 1 : 0);
                        for (int sgi = lane; sgi < ns_; sgi += 32) sm.seg_end[sgi] = sm.seg_begin[sgi + 1];
                    }
                    // TMA: concatenate the pieces in shared memory
                    if (total > 0) {
                        if (lane == 0) {
                            const uint32_t bytes =
                                (uint32_t)total * (uint32_t)(sizeof(Rec<T>) + (unwrapped ? sizeof(int4) : 0));
                            mbar_arrive_expect_tx(reinterpret_cast<uint64_t*>(&sm.mbar), bytes);
                        }
                        __syncwarp();
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (take[q] > 0) {
                                tma_load_1d(cand + dst[q], sorted + src[q], (uint32_t)take[q] * (uint32_t)sizeof(Rec<T>),
                                            reinterpret_cast<uint64_t*>(&sm.mbar));
                                if (unwrapped)
                                    tma_load_1d(cand_ash + dst[q], sorted_ashift + src[q],
                                                (uint32_t)take[q] * (uint32_t)sizeof(int4),
                                                reinterpret_cast<uint64_t*>(&sm.mbar));
                            }
                        }
                    }
                }
                __syncthreads();
                const int total = sm.total;
                const bool more = sm.more != 0;
                if (first_tile) {
                    multi = multi_batch || more;
                    first_tile = false;
                }
                if (total > 0) {
                    mbar_wait(reinterpret_cast<uint64_t*>(&sm.mbar), phase);
                    phase ^= 1u;
                    // ---- sweep: one warp per target atom ----
                    for (int t = t_begin + warp; t < ntarget; t += kSweepWarps) {
                        const Rec<T> ti = sorted[home_start + t];
                        int4 ai = make_int4(0, 0, 0, 0);
                        if (unwrapped) ai = sorted_ashift[home_start + t];
                        const int i = ti.j;
                        int written = 0;
                        if (multi) written = cursor[i];  // running total over tiles (all modes)
                        size_t row_base = 0;
                        if (MODE == MODE_FILL_COO) row_base = (size_t)a.neighbor_ptr[i];
                        int found;
                        if (unwrapped)
                            found = sweep_target<T, MODE, HALF, true, FMA>(a, sm, cand, cand_ash, cm, pbc, ti, ai, lane,
                                                                           row_base, written, row_j, row_s);
                        else
                            found = sweep_target<T, MODE, HALF, false, FMA>(a, sm, cand, cand_ash, cm, pbc, ti, ai, lane,
                                                                            row_base, written, row_j, row_s);
                        if (multi) {
                            if (lane == 0) cursor[i] = written + found;
                        } else if (MODE == MODE_COUNT) {
                            if (lane == 0) a.num_neighbors[i] = found;
                        } else if (MODE == MODE_FILL_MATRIX) {
                            finish_matrix_row<T>(a, lane, i, found);
                        }
                    }
                }
                __syncthreads();
                if (!more) break;
            }
        }
        // multi-tile rows: finalize once every tile has been swept, and re-arm the running totals
        if (multi) {
            for (int t = t_begin + warp; t < ntarget; t += kSweepWarps) {
                const int i = sorted[home_start + t].j;
                const int tot = cursor[i];
                __syncwarp();
                if (lane == 0) cursor[i] = 0;
                if (MODE == MODE_COUNT) {
                    if (lane == 0) a.num_neighbors[i] = tot;
                } else if (MODE == MODE_FILL_MATRIX) {
                    finish_matrix_row<T>(a, lane, i, tot);
                }
            }
        }
    }
    // the last CTA to drain the queue re-arms it for the next launch on this workspace
    if (tid == 0) {
        __threadfence();
        const int d = atomicAdd(&ctrl->done[a.queue], 1);
        if (d == (int)gridDim.x - 1) {
            ctrl->work_counter[a.queue] = 0;
            ctrl->done[a.queue] = 0;
            // consumed: the next fast launch rebuilds the list.  The single-sweep COO path has no fast FILL launch, so
            // its COUNT launch keeps the list for the general FILL launch (unwrapped inputs take the two-pass path).
            if (!(a.keep_deferred && ctrl->unwrapped == 0)) ctrl->n_deferred = 0;
        }
    }
}

}  // namespace nvnl
