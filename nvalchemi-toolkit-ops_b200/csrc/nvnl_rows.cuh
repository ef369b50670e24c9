// nvnl_rows.cuh — single-sweep COO path for fp32 inputs inside the primary periodic image (round 2 rewrite).
//
//   k_rows      (cell order)   stencil sweep.  A consumer warp takes FOUR target atoms per trip (two packed f32x2
//                              pairs) and walks the staged stencil tile in 32-candidate chunks; every lane keeps, per
//                              target, a TRANSPOSED hit mask in registers (bit v = "my candidate of chunk v hit"):
//                              per chunk and target the compaction costs one FSETP and one predicated LOP — no
//                              ballot, no popc, no shared-memory list.  Per trip: ONE packed warp scan pair of the
//                              four popcounts, ONE reservation in the temporary row buffer, then every lane walks
//                              the set bits of its masks (tile slot -> original atom index) into four compact rows.
//   k_scan                     neighbor_ptr (unchanged).
//   k_rows_out  (index order)  a warp prefetches the temporary rows of 32 consecutive atoms into shared memory with
//                              one TMA bulk copy per lane (rows are 16-byte aligned and padded to 4 entries), then
//                              streams edge_index / shifts strictly sequentially.
//
// Tile staging (producer warp): variable-size tiles in a byte ring with a descriptor ring (more small tiles in flight
// for big boxes, whole small systems for 3-cell boxes).  Images of the stencil are ordered by periodic shift; every
// shift SEGMENT starts on a 32-slot boundary and its tail is padded with far-away sentinel records, so a chunk never
// straddles two shifts (warp-uniform shift vector, no per-lane selection) and no chunk needs a validity mask.
//
// Work distribution beyond "one cell = one tile": a cell with more than 64 targets (small systems that are one or a few
// cells) is pushed to a device-side split list as parts of 32 targets and swept by whichever CTAs drain the cell queue
// first; a single-cell periodic system whose 27 images do not fit 64 chunks is staged ONCE with the image segments
// aliasing the same records (six mask words per lane and target) — that variant is the second instantiation
// k_rows<HUGE>, launched only for workloads that have such systems, so the common kernel keeps its register budget.
// k_rows<PAIR> (nvnl_pair.cuh) feeds a pair consumer instead of writing rows: the neighbor list is never materialised.
//
// Cells the lean kernel cannot take (> 32 images, > 64 chunks that cannot be aliased) go to the general kernel; unwrapped
// inputs and fp64 use the two-pass path.  Replaces (different algorithm, same result): cell_list.py:372-556 +
// neighbor_utils.py:106-147, 362-441.
#pragma once
#include "nvnl_fast.cuh"

namespace nvnl {

// CTA shape (tunable at build time for profiling: -DNVNL_ROWS_CONS=.. -DNVNL_ROWS_MINB=.. -DNVNL_ROWS_RING_KB=..)
#ifndef NVNL_ROWS_CONS
#define NVNL_ROWS_CONS 5
#endif
#ifndef NVNL_ROWS_MINB
#define NVNL_ROWS_MINB 4
#endif
#ifndef NVNL_ROWS_RING_KB
#define NVNL_ROWS_RING_KB 40
#endif
constexpr int kRowsCons = NVNL_ROWS_CONS;             // consumer warps per CTA
constexpr int kRowsThreads = (kRowsCons + 1) * 32;    // + producer warp
constexpr int kRowsMinBlocks = NVNL_ROWS_MINB;        // CTAs per SM
constexpr int kRowsDesc = 6;                          // tiles in flight per CTA (descriptor ring)
constexpr int kRowsRingBytes = NVNL_ROWS_RING_KB * 1024;   // staged records of the tiles in flight
constexpr int kRowsZeroBytes = 4 * 1024;              // shared-memory zero block behind the fused zero-fill (TMA bulk stores)
constexpr int kRowsMaxVC = 64;                        // 32-candidate chunks per tile (two mask words per lane and target)
constexpr int kRowsMaxVCAlias = 192;                  // ... of an aliased tile (single-cell system: every image = the same records)
constexpr int kRowsHugeWords = kRowsMaxVCAlias / 32;  // mask words per lane and target of such a tile
constexpr int kRowsMaxSeg = 32;                       // shift segments per tile (one per image at most)
constexpr int kRowsSplitAbove = 64;                   // cells with more targets are swept in parts, by several CTAs
constexpr int kRowsSplitPart = 32;                    // targets per part
constexpr int kRowsHugePart = 20;                     // ... of a huge (aliased) tile: one trip per consumer warp, the sweep of
                                                      // 100+ chunks is a long serial chain and there are few such tiles
constexpr int kRowsBlock = 2048;                      // temp-buffer entries a warp reserves per cursor bump
constexpr int kRowsSegShift = 27;                     // boundary rows: entry = atom | segment << 27 (atoms < 2^27)
constexpr float kRowsFar = 3.0e18f;                   // sentinel coordinate: squares stay finite, never within any cutoff

// Per-tile descriptor, written by the producer warp, read by the consumers.
struct RowsDesc {
    int item;                         // global cell id, -1 = end of work
    int ntarget, home_slot;           // targets of the cell and the tile slot of the first one
    int nseg, nvc;                    // shift segments, 32-candidate chunks (interior tiles: even)
    int data_off;                     // byte offset of the tile in the ring
    int shifted;                      // some segment has a non-zero shift
    int next_target;                  // consumer claim counter
    int footprint;                    // producer-private: ring bytes held by this tile
    int aliased;                      // every segment is an image of the SAME staged records (single-cell system)
    int pad0[2];
    int seg_vc[kRowsMaxSeg + 1];      // first (virtual) chunk of every segment; the mask bits count virtual chunks
    int seg_key[kRowsMaxSeg];         // packed integer shift of the segment
    float segS[kRowsMaxSeg][3];       // its lattice vector s·cell
    unsigned char vc_seg[kRowsMaxVCAlias];  // virtual chunk -> segment
    unsigned char vc_phys[kRowsMaxVCAlias]; // virtual chunk -> staged chunk (aliased tiles)
};

struct RowsSmem {
    RowsDesc desc[kRowsDesc];
    int e_st[32], e_cn[32], e_key[32], e_tag[32];                           // producer scratch (shift sort)
    unsigned long long full[kRowsDesc], empty[kRowsDesc];                   // mbarriers of the descriptor ring
    alignas(128) unsigned char zeros[kRowsZeroBytes];                       // source of the TMA zero-fill stores
};

// PAIR mode keeps the fp64 lattice vector of every shift segment of every tile in flight behind RowsSmem
constexpr size_t kRowsPairS64Bytes = sizeof(double) * 3 * kRowsMaxSeg * kRowsDesc;
constexpr size_t rows_smem_bytes(bool pair = false) {
    return (size_t)kRowsRingBytes + sizeof(RowsSmem) + (pair ? kRowsPairS64Bytes : 0);
}

struct RowsArgs {
    unsigned char* ws;
    WsLayout L;
    const int* batch_idx;
    int num_systems;
    long long n;
    float cutoff_sq;
    int* num_neighbors;
    int* prezero;             // optional: buffer the kernel zero-fills while it sweeps (the shifts output, sized by the caller's guess)
    long long prezero_ints;
    // PAIR mode (nvnl_pair.cuh): the sweep feeds a pair consumer instead of writing rows
    const double* q_sorted;   // charges in cell-sorted order
    double* pair_energies;    // [n]
    double* pair_forces;      // [n, 3]
    double pair_cutoff, pair_alpha;
};

// the k_rows barriers are polled with a short sleep between tries: a spinning warp would otherwise take issue slots
// from the sweeping warps of its scheduler
template <int SLEEP_NS = 64>
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
#ifdef NVNL_MBAR_SUSPEND
    // variant: let the hardware suspend the warp inside try_wait (time-limit hint) instead of polling with naps
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "NVNL_SWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@!p bra NVNL_SWAIT_%=;\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(NVNL_MBAR_SUSPEND) : "memory");
    return;
#endif
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra NVNL_BDONE_%=;\n\t"
        "NVNL_BWAIT_%=:\n\t"
        "nanosleep.u32 %2;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra NVNL_BWAIT_%=;\n\t"
        "NVNL_BDONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity), "n"(SLEEP_NS) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void sts_rec(uint32_t addr, float x, float y, float z, int j) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(x), "f"(y), "f"(z), "r"(j) : "memory");
}
__device__ __forceinline__ int lds_b32(uint32_t addr) {
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// 1-D TMA bulk copy shared -> global (bulk-group completion).  dst/src 16-byte aligned, bytes a positive multiple of 16.
__device__ __forceinline__ void tma_store_1d(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// m |= bit where d < rc2: one FSETP and one predicated LOP
__device__ __forceinline__ void rows_acc(unsigned& m, float d, float rc2, unsigned bit) {
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(m) : "f"(d), "f"(rc2), "r"(bit));
}

// packed target coordinates of a trip: pairs (0,1) and (2,3)
struct RowsTargets {
    f32x2_t X01, Y01, Z01, X23, Y23, Z23;
    int i[4];
};

// squared distances of one candidate to a packed target pair; op order of the reference: (r_j - r_i) + s·cell, then
// dx*dx (+) dy*dy (+) dz*dz
template <bool FMA, bool SHIFTED>
__device__ __forceinline__ f32x2_t rows_d2(float x, float y, float z, f32x2_t XI, f32x2_t YI, f32x2_t ZI, float Sx, float Sy,
                                           float Sz) {
    f32x2_t dx = sub2(pack2(x, x), XI), dy = sub2(pack2(y, y), YI), dz = sub2(pack2(z, z), ZI);
    if (SHIFTED) {
        dx = add2(dx, pack2(Sx, Sx));
        dy = add2(dy, pack2(Sy, Sy));
        dz = add2(dz, pack2(Sz, Sz));
    }
    f32x2_t d2;
    if (FMA) {
        d2 = mul2(dx, dx);
        d2 = fma2(dy, dy, d2);
        d2 = fma2(dz, dz, d2);
    } else {
        d2 = add2(mul2(dx, dx), mul2(dy, dy));
        d2 = add2(d2, mul2(dz, dz));
    }
    return d2;
}

// one 32-candidate chunk against the four targets of the trip; `bit` = this chunk's bit in the mask word
template <bool HALF, bool FMA, bool SHIFTED>
__device__ __forceinline__ void rows_chunk4(uint32_t addr, const RowsTargets& t, float Sx, float Sy, float Sz, float rc2,
                                            unsigned bit, bool lexpos, unsigned (&m)[4]) {
    float x, y, z;
    int j;
    lds_rec(addr, x, y, z, j);
    const f32x2_t a = rows_d2<FMA, SHIFTED>(x, y, z, t.X01, t.Y01, t.Z01, Sx, Sy, Sz);
    const f32x2_t b = rows_d2<FMA, SHIFTED>(x, y, z, t.X23, t.Y23, t.Z23, Sx, Sy, Sz);
    float d[4];
    unpack2(a, d[0], d[1]);
    unpack2(b, d[2], d[3]);
    if (!HALF) {
#pragma unroll
        for (int k = 0; k < 4; ++k) rows_acc(m[k], d[k], rc2, bit);
    } else {
        // half fill keeps (i, j, s) only for i < j, or i == j with a lexicographically positive shift
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool keep = t.i[k] < j || (t.i[k] == j && lexpos);
            if (d[k] < rc2 && keep) m[k] |= bit;
        }
    }
}

// chunks [v0, v1) of the zero-shift home segment, two per trip (interior tiles pad their chunk count to an even number)
template <bool HALF, bool FMA>
__device__ __forceinline__ void rows_run_zero(uint32_t addr, int v0, int v1, const RowsTargets& t, float rc2, unsigned (&m)[4]) {
    constexpr uint32_t RS = sizeof(Rec<float>);
    unsigned bit = 1u << (v0 & 31);
    int v = v0;
#pragma unroll 1
    for (; v + 2 <= v1; v += 2) {
        rows_chunk4<HALF, FMA, false>(addr, t, 0.f, 0.f, 0.f, rc2, bit, false, m);
        rows_chunk4<HALF, FMA, false>(addr + 32u * RS, t, 0.f, 0.f, 0.f, rc2, bit << 1, false, m);
        addr += 64u * RS;
        bit <<= 2;
    }
    if (v < v1) rows_chunk4<HALF, FMA, false>(addr, t, 0.f, 0.f, 0.f, rc2, bit, false, m);
}

template <bool HALF, bool FMA>
__device__ __forceinline__ void rows_run_shift(uint32_t addr, int v0, int v1, const RowsTargets& t, float Sx, float Sy,
                                               float Sz, float rc2, bool lexpos, unsigned (&m)[4]) {
    constexpr uint32_t RS = sizeof(Rec<float>);
    unsigned bit = 1u << (v0 & 31);
#pragma unroll 1
    for (int v = v0; v < v1; ++v) {
        rows_chunk4<HALF, FMA, true>(addr, t, Sx, Sy, Sz, rc2, bit, lexpos, m);
        addr += 32u * RS;
        bit <<= 1;
    }
}

// Sweep of the four targets over (virtual) chunks [32 w, 32 w + 32) of the staged tile: m = their hit masks
// (bit = chunk - 32 w).  One call per mask word, so that the masks stay in registers; words past 0 only exist for tiles
// of more than 1024 slots.
template <bool HALF, bool FMA, bool ALIAS>
__device__ __forceinline__ void rows_sweep_word(const RowsDesc& d, uint32_t tile_addr, const RowsTargets& t, float rc2,
                                                int lane, int w, unsigned (&m)[4]) {
    constexpr uint32_t RS = sizeof(Rec<float>);
    const int nvc = d.nvc;
    const uint32_t lane_addr = tile_addr + (uint32_t)lane * RS;
    const int wb = w << 5, we = nvc < wb + 32 ? nvc : wb + 32;
    if (!d.shifted) {
        rows_run_zero<HALF, FMA>(lane_addr + (uint32_t)wb * (32u * RS), wb, we, t, rc2, m);
        return;
    }
    const int nseg = d.nseg;
#pragma unroll 1
    for (int s = 0; s < nseg; ++s) {
        int v0 = d.seg_vc[s], v1 = d.seg_vc[s + 1];
        v0 = v0 > wb ? v0 : wb;
        v1 = v1 < we ? v1 : we;
        if (v0 >= v1) continue;
        // (aliased tiles: the segment's records are the staged chunks 0 ..; else virtual chunk == staged chunk)
        const uint32_t addr = lane_addr + (uint32_t)(ALIAS ? v0 - d.seg_vc[s] : v0) * (32u * RS);   // ALIAS: see rows_trip
        const int key = d.seg_key[s];
        if (key == 0) {
            rows_run_zero<HALF, FMA>(addr, v0, v1, t, rc2, m);
        } else {
            bool lexpos = false;
            if (HALF) {
                int csx, csy, csz;
                unpack_key(key, csx, csy, csz);
                lexpos = csx > 0 || (csx == 0 && (csy > 0 || (csy == 0 && csz > 0)));
            }
            rows_run_shift<HALF, FMA>(addr, v0, v1, t, d.segS[s][0], d.segS[s][1], d.segS[s][2], rc2, lexpos, m);
        }
    }
}

// Every lane walks the set bits of its own masks of word w, leading bit first (entry r of the lane goes to slot r of
// its piece of the row, p[k]).  Branch-free: a lane that has run out of bits keeps executing with its load and store
// predicated off (the loop runs to the longest list of the warp).  SHIFTED rows carry the entry's shift segment in the
// top bits.  Advances p[k] past the entries written.
template <bool SHIFTED, bool ALIAS>
__device__ __forceinline__ void rows_gather_word(const RowsDesc& d, uint32_t tile_addr, const unsigned (&mw)[4], int w,
                                                 int lane, int* __restrict__ (&p)[4]) {
    constexpr uint32_t RS = sizeof(Rec<float>);
    unsigned m[4];
    int mx = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        m[k] = mw[k];
        const int c = __popc(m[k]);
        mx = c > mx ? c : mx;
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    // address of .j of this lane's candidate of chunk 0 of the word; chunk b is b * 512 bytes above
    const uint32_t bot_addr = tile_addr + (uint32_t)lane * RS + 12u + (uint32_t)(w * 32) * (32u * RS);
    const uint32_t seg_bot = smem_u32(&d.vc_seg[0]) + (uint32_t)(w * 32);
    const uint32_t phys_bot = smem_u32(&d.vc_phys[0]) + (uint32_t)(w * 32);
    const uint32_t lane_j = tile_addr + (uint32_t)lane * RS + 12u;
#pragma unroll 2
    for (int r = 0; r < mx; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned mk = m[k];
            int b;                                                  // leading set bit, -1 when the lane has none left
            asm("bfind.u32 %0, %1;" : "=r"(b) : "r"(mk));
            unsigned one;
            asm("shl.b32 %0, %1, %2;" : "=r"(one) : "r"(1u), "r"(b));   // (clamped shift: 0 for b = -1)
            m[k] = mk ^ one;
            const int b0 = b < 0 ? 0 : b;                           // a lane without bits reads chunk 0 and stores nothing
            int j;
            if (ALIAS) {
                int pc;   // the staged chunk behind virtual chunk 32 w + b0
                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(pc) : "r"(phys_bot + (uint32_t)b0));
                j = lds_b32(lane_j + (uint32_t)pc * (32u * RS));
            } else {
                j = lds_b32(bot_addr + (uint32_t)b0 * (32u * RS));
            }
            if (SHIFTED) {
                int sg;
                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(sg) : "r"(seg_bot + (uint32_t)b0));
                j |= sg << kRowsSegShift;
            }
            if (mk != 0u) p[k][r] = j;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) p[k] += __popc(mw[k]);
}

// Per-warp allocator state of the temporary row buffer.
struct RowsAlloc {
    int pos, end;   // entries (the buffer holds fewer than 2^29)
};

// Row format in the temporary buffer (all starts are multiples of 4 entries = 16 bytes; TMA-able):
//   interior cell:        [entries ... padded to 4]                                   row_ref = start << 2
//   cell at a boundary:   [8 or 32 packed image keys][entries = atom | segment << 27]  row_ref = start << 2 | 1 or 2
// Epilogue of a trip: popcounts, ONE packed scan pair, ONE reservation, then every lane walks its own set bits.
// W = mask words per lane and target (1: tiles of <= 32 chunks, 2: <= 64, kRowsHugeWords: aliased single-cell tiles);
// s0 = tile slot of the trip's first target.
template <int W>
__device__ __forceinline__ void rows_emit4(long long rows_cap, int* __restrict__ num_neighbors, const RowsDesc& d, Ctrl* ctrl,
                                           uint32_t tile_addr, unsigned (&m)[W][4], int s0, int nt, int lane, RowsAlloc& al,
                                           int* __restrict__ rows, int* __restrict__ row_ref) {
    constexpr uint32_t RS = sizeof(Rec<float>);
    int n[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        n[k] = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            if (k >= nt) m[w][k] = 0u;   // targets past the end of the cell
            n[k] += __popc(m[w][k]);
        }
    }
    // a lane holds <= 192 hits per target, a row < 65536 entries: two 16-bit fields per scan word
    const int pa = n[0] | (n[1] << 16), pb = n[2] | (n[3] << 16);
    int ia = pa, ib = pb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int ua = __shfl_up_sync(0xffffffffu, ia, o), ub = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { ia += ua; ib += ub; }
    }
    const int ta = __shfl_sync(0xffffffffu, ia, 31), tb = __shfl_sync(0xffffffffu, ib, 31);
    const int ea = ia - pa, eb = ib - pb;
    const int cnt[4] = {ta & 0xffff, (int)((unsigned)ta >> 16), tb & 0xffff, (int)((unsigned)tb >> 16)};
    const int ex[4] = {ea & 0xffff, (int)((unsigned)ea >> 16), eb & 0xffff, (int)((unsigned)eb >> 16)};
    const bool shifted = d.shifted != 0;
    const int hdr = shifted ? (d.nseg <= 8 ? 8 : 32) : 0;
    int start[4];
    int need = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        start[k] = need;
        if (k < nt) need += hdr + ((cnt[k] + 3) & ~3);
    }
    bool ok = true;
    if (al.pos + need > al.end) {
        const int sz = need > kRowsBlock ? need : kRowsBlock;
        unsigned long long b = 0ull;
        if (lane == 0) b = atomicAdd(&ctrl->rows_cursor, (unsigned long long)sz);
        b = __shfl_sync(0xffffffffu, b, 0);
        if ((long long)b + sz > rows_cap) {
            // temporary buffer exhausted: the host re-runs the query on the two-pass path (nvnl_status.rows_overflow)
            if (lane == 0) ctrl->rows_overflow = 1;
            al.pos = al.end = 0;
            ok = false;
        } else {
            al.pos = (int)b;
            al.end = (int)b + sz;
        }
    }
    if (ok) {
        const int base = al.pos;
        al.pos += need;
        int* __restrict__ p[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            start[k] += base;
            p[k] = rows + start[k] + hdr + ex[k];
        }
        if (shifted) {
            // the stencil's packed image keys in front of every row: the shift of an entry is a table lookup later
            if (lane < hdr) {
                const int key = lane < d.nseg ? d.seg_key[lane] : 0;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < nt) rows[start[k] + lane] = key;
            }
        }
        if (W > 2) {
            // (aliased tiles — and only they — take the path with more than two mask words; they always carry shifts)
#pragma unroll
            for (int w = 0; w < W; ++w) rows_gather_word<true, true>(d, tile_addr, m[w], w, lane, p);
        } else if (shifted) {
#pragma unroll
            for (int w = 0; w < W; ++w) rows_gather_word<true, false>(d, tile_addr, m[w], w, lane, p);
        } else {
#pragma unroll
            for (int w = 0; w < W; ++w) rows_gather_word<false, false>(d, tile_addr, m[w], w, lane, p);
        }
        // padding entries of every row (read by the bulk copies of the output kernel, never used)
        if (lane < 3) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < nt && cnt[k] + lane < ((cnt[k] + 3) & ~3)) rows[start[k] + hdr + cnt[k] + lane] = -1;
        }
    }
    if (lane < nt) {
        const int i = lds_b32(tile_addr + (uint32_t)(s0 + lane) * RS + 12u);   // original index of target `lane`
        const int c = lane == 0 ? cnt[0] : (lane == 1 ? cnt[1] : (lane == 2 ? cnt[2] : cnt[3]));
        const int s = lane == 0 ? start[0] : (lane == 1 ? start[1] : (lane == 2 ? start[2] : start[3]));
        if (ok) row_ref[i] = (s << 2) | (hdr == 0 ? 0 : (hdr == 8 ? 1 : 2));
        num_neighbors[i] = c;
    }
}

// One trip of a consumer warp: four targets (tile slots s0 .. s0 + nt - 1) against the staged tile, then their rows.
// Four targets share every candidate load; two packed pairs share every FP instruction.
template <bool HALF, bool FMA, int W>
__device__ __forceinline__ void rows_trip(long long rows_cap, int* __restrict__ num_neighbors, const RowsDesc& ds, Ctrl* ctrl,
                                          uint32_t tile_addr, int s0, int nt, float rc2, int lane, RowsAlloc& al,
                                          int* __restrict__ rows, int* __restrict__ row_ref) {
    constexpr uint32_t RS = sizeof(Rec<float>);
    RowsTargets tg;
    {
        // targets past the end of the cell repeat the last one (their results are dropped)
        float x[4], y[4], z[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            lds_rec(tile_addr + (uint32_t)(s0 + (k < nt ? k : nt - 1)) * RS, x[k], y[k], z[k], tg.i[k]);
        // x + (-0) == x bit for bit: one packed add gives each target pair a home in an aligned register pair
        const f32x2_t nz = pack2(-0.0f, -0.0f);
        tg.X01 = add2(pack2(x[0], x[1]), nz); tg.Y01 = add2(pack2(y[0], y[1]), nz); tg.Z01 = add2(pack2(z[0], z[1]), nz);
        tg.X23 = add2(pack2(x[2], x[3]), nz); tg.Y23 = add2(pack2(y[2], y[3]), nz); tg.Z23 = add2(pack2(z[2], z[3]), nz);
    }
    unsigned m[W][4];
#pragma unroll
    for (int w = 0; w < W; ++w) {
#pragma unroll
        for (int k = 0; k < 4; ++k) m[w][k] = 0u;
        if (w == 0 || (w << 5) < ds.nvc) rows_sweep_word<HALF, FMA, (W > 2)>(ds, tile_addr, tg, rc2, lane, w, m[w]);
    }
    if (!HALF) {
        // (i, i, 0) is not a pair: the targets sit in the zero-shift home segment, whose virtual chunks are its staged
        // chunks (chunk = slot / 32)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int sl = s0 + (k < nt ? k : nt - 1);
            // (branch-free on purpose: a conditional on the word index makes the compiler index m[][] dynamically,
            //  which moves the masks to local memory)
            const unsigned bit = lane == (sl & 31) ? 1u << ((sl >> 5) & 31) : 0u;
#pragma unroll
            for (int w = 0; w < W; ++w) m[w][k] &= ~(W == 1 ? bit : (bit & (0u - (unsigned)((sl >> 10) == w))));
        }
    }
    rows_emit4<W>(rows_cap, num_neighbors, ds, ctrl, tile_addr, m, s0, nt, lane, al, rows, row_ref);
}

}  // namespace nvnl
#include "nvnl_pair.cuh"
namespace nvnl {

// One trip in PAIR mode: the same masks, then the pair consumer instead of the rows.
template <bool FMA, int W>
__device__ __forceinline__ void pair_trip(const RowsArgs& a, const RowsDesc& ds, const double (*S64)[3], uint32_t tile_addr,
                                          int s0, int nt, float rc2, int lane) {
    constexpr uint32_t RS = sizeof(Rec<float>);
    RowsTargets tg;
    {
        float x[4], y[4], z[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            lds_rec(tile_addr + (uint32_t)(s0 + (k < nt ? k : nt - 1)) * RS, x[k], y[k], z[k], tg.i[k]);
        const f32x2_t nz = pack2(-0.0f, -0.0f);
        tg.X01 = add2(pack2(x[0], x[1]), nz); tg.Y01 = add2(pack2(y[0], y[1]), nz); tg.Z01 = add2(pack2(z[0], z[1]), nz);
        tg.X23 = add2(pack2(x[2], x[3]), nz); tg.Y23 = add2(pack2(y[2], y[3]), nz); tg.Z23 = add2(pack2(z[2], z[3]), nz);
    }
    unsigned m[W][4];
#pragma unroll
    for (int w = 0; w < W; ++w) {
#pragma unroll
        for (int k = 0; k < 4; ++k) m[w][k] = 0u;
        if (w == 0 || (w << 5) < ds.nvc) rows_sweep_word<false, FMA, false>(ds, tile_addr, tg, rc2, lane, w, m[w]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int sl = s0 + (k < nt ? k : nt - 1);
        const unsigned bit = lane == (sl & 31) ? 1u << ((sl >> 5) & 31) : 0u;
#pragma unroll
        for (int w = 0; w < W; ++w) m[w][k] &= ~(W == 1 ? bit : (bit & (0u - (unsigned)((sl >> 10) == w))));
    }
    pair_consume4<W>(a, ds, S64, tile_addr, tile_addr + (uint32_t)ds.nvc * (32u * RS), m, s0, nt, lane);
}

// ------------------------------------------------------------------------------------------------
// k_rows: warp-specialised persistent kernel.  The LAST warp is the producer (the warp scheduler favours the highest
// warp id among eligible warps, and the producer's serial per-tile latency bounds the whole CTA): queue -> stencil
// images -> shift sort -> aligned segment layout -> descriptor + sentinels -> TMA bulk copies (x-adjacent cells are one
// copy) into the byte ring.  Warps 0..kRowsCons-1 (consumers): claim four targets of the current tile at a time,
// sweep, emit rows.  No CTA-wide barrier in the steady state.
// ------------------------------------------------------------------------------------------------
// HUGE = the second launch: parts of single-cell systems whose images need more than kRowsMaxVC chunks (aliased tiles,
// six mask words per lane and target — a register budget the common kernel should not pay for), popped from the huge
// list the first launch filled.
// PAIR = the sweep feeds the pair consumer of nvnl_pair.cuh (charges staged next to the records, fp64 shift vectors per
// segment, no rows): cells it cannot take are reported like deferred cells and the host falls back to list + consumer.
template <bool HALF, bool FMA, bool HUGE, bool PAIR = false>
__global__ void __launch_bounds__(kRowsThreads, HUGE ? 3 : (PAIR ? 3 : kRowsMinBlocks)) k_rows(const RowsArgs a) {
    pdl_enter();
    using T = float;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    RowsSmem& sm = *reinterpret_cast<RowsSmem*>(smem_raw + (size_t)kRowsRingBytes);
    double (*const segS64)[kRowsMaxSeg][3] =
        reinterpret_cast<double (*)[kRowsMaxSeg][3]>(smem_raw + (size_t)kRowsRingBytes + sizeof(RowsSmem));   // PAIR only
    const uint32_t ring_addr = smem_u32(smem_raw);
    constexpr uint32_t RS = sizeof(Rec<T>);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Ctrl* ctrl = reinterpret_cast<Ctrl*>(a.ws + a.L.ctrl);

    if (!HUGE) {
        // re-arm the look-back scan that turns the counts into neighbor_ptr (runs after the count kernels)
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
    // unwrapped inputs are served by the two-pass kernels launched next to this one
    const bool active = ctrl->unwrapped == 0;
    // Zero-fill of the caller's shifts buffer, fused into the sweep: the sweep is issue-bound and leaves HBM idle, the
    // zero-fill is pure HBM writes.  Every CTA owns one slice; its producer thread writes it with TMA bulk STORES from a
    // zeroed shared-memory block — the copy engine moves the bytes, no SM instruction or LSU slot is spent on them —
    // a quota per published tile (paced over the kernel's lifetime), the rest when the queue is drained.
    // (A separate memset kernel does not do: next to this kernel's large CTAs it only runs if the SM already has the
    // large shared-memory carve-out, otherwise the two serialise — profiles/r2_zero_overlap.txt.)
    // Workloads whose rows mostly carry shifts (small periodic boxes; Ctrl::shift_heavy, set by k_grid) are not
    // pre-zeroed: the output kernel writes their shifts densely, zeros included, in one pass.
    int zpos = 0, zend = 0, zquota = 0;    // in 16-byte units (the buffer holds fewer than 2^31 of them)
    if (!HUGE && a.prezero && ctrl->shift_heavy == 0) {
        const long long n16 = a.prezero_ints >> 2;
        const long long slice = (n16 + gridDim.x - 1) / gridDim.x;
        long long lo = (long long)blockIdx.x * slice, hi = lo + slice < n16 ? lo + slice : n16;
        if (lo > hi) lo = hi;
        zpos = (int)lo; zend = (int)hi;
        if (blockIdx.x == 0 && tid < (int)(a.prezero_ints & 3)) a.prezero[(n16 << 2) + tid] = 0;
    }
    for (int k = tid; k < kRowsZeroBytes / 16; k += kRowsThreads) reinterpret_cast<int4*>(sm.zeros)[k] = make_int4(0, 0, 0, 0);
    fence_proxy_async_smem();   // the zero block is read by the async proxy (TMA stores)
    if (tid == 0) {
        for (int st = 0; st < kRowsDesc; ++st) {
            mbar_init(reinterpret_cast<uint64_t*>(&sm.full[st]), 1);
            mbar_init(reinterpret_cast<uint64_t*>(&sm.empty[st]), kRowsCons);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kRowsCons) {
        // =========================== producer ===========================
        const SysParams* sys = reinterpret_cast<const SysParams*>(a.ws + a.L.sys);
        const int* cell_start = reinterpret_cast<const int*>(a.ws + a.L.cell_start);
        const Rec<T>* sorted = reinterpret_cast<const Rec<T>*>(a.ws + a.L.sorted);
        int2* deferred = reinterpret_cast<int2*>(a.ws + a.L.deferred);
        int* row_ref = reinterpret_cast<int*>(a.ws + a.L.row_ref);
        const int total_cells = active ? ctrl->total_cells : 0;
        // zero-fill quota per published tile: the slice spread over the tiles this CTA is expected to produce
        // (pacing knob for experiments: quota x 0.7 / 1.5 / 2 were all slower than the even pace, profiles/r2_variants_d.txt)
#ifndef NVNL_ZQUOTA_NUM
#define NVNL_ZQUOTA_NUM 1
#define NVNL_ZQUOTA_DEN 1
#endif
        zquota = (int)(((long long)(zend - zpos) * NVNL_ZQUOTA_NUM / NVNL_ZQUOTA_DEN / (total_cells / (int)gridDim.x + 1) + kRowsZeroBytes / 16) /
                       (kRowsZeroBytes / 16) * (kRowsZeroBytes / 16));
        // descriptor ring + byte ring (FIFO): tiles [oldest, nprod) are live
        int nprod = 0, oldest = 0;
        int head = 0, used = 0;
        int g_next = 0;
        if (!HUGE && lane == 0) g_next = atomicAdd(&ctrl->work_counter[0], 1);   // (the second launch pops the huge list only)
        // cells with many targets (small systems that are one or a few cells) are not swept by one CTA: whoever meets
        // such a cell pushes it to the split list as parts of kRowsSplitPart targets; once the cell queue is drained
        // every CTA pops parts until all cells have been classified and the list is empty
#define NVNL_SPLIT_PTR reinterpret_cast<int2*>(a.ws + a.L.split)
#define NVNL_SPLIT_CAP ((int)(a.n / 16 + 2))
        bool phase2 = false;
        int my_cells = 0;
        // grid of the current system (reloaded only when the system changes)
        int s_cur = -1, cpd0 = 1, cpd1 = 1, cpd2 = 1, R0 = 0, R1 = 0, R2 = 0, pb0 = 0, pb1 = 0, pb2 = 0, coff = 0, c01 = 1;
        bool one_cell = false;
        float rcp0 = 1.f, rcp01 = 1.f;   // reciprocals for the cell-coordinate division (corrected exactly below)
        bool r111 = false;
        auto zero_some = [&](int quota) {   // lane 0: TMA bulk stores of zeros, up to `quota` bytes of the slice
            int e = zpos + quota;
            e = (e < zend && e >= zpos) ? e : zend;
            unsigned char* const zbase = reinterpret_cast<unsigned char*>(a.prezero);
            while (zpos < e) {
                const int nb = e - zpos < kRowsZeroBytes / 16 ? e - zpos : kRowsZeroBytes / 16;
                tma_store_1d(zbase + (size_t)zpos * 16, sm.zeros, (uint32_t)nb * 16u);
                zpos += nb;
            }
            tma_store_commit();
        };
        for (;;) {
            // ---- 1. prepare the next cell entirely in registers (dependent global loads, image enumeration, shift
            //         sort, aligned layout) BEFORE waiting for ring space ----
            bool have = false;
            int g = 0, ntarget = 0, st = 0, cn = 0, key = kKeyEmpty, aoff = 0, total = 0, nseg = 1, nvc = 0;
            int home_slot = 0, si = 0, seg_len = 0, seg_vcb = 0, seg_nch = 0, grp_cn = 0;
            int tgt0 = 0, tgt1 = 0;          // the targets of the cell this work item covers
            unsigned shiftmask = 0u, hm = 1u;
            bool head_lane = false, alias = false;
            T Sx = (T)0, Sy = (T)0, Sz = (T)0;
            double S64x = 0.0, S64y = 0.0, S64z = 0.0;   // PAIR: the segment's lattice vector in fp64
            for (;;) {
                bool part = false;           // the item is a part of a cell, popped from the split list
                if (HUGE) {
                    int t = 0;
                    if (lane == 0) t = atomicAdd(&ctrl->huge_next, 1);
                    t = __shfl_sync(0xffffffffu, t, 0);
                    if (t >= ctrl->n_huge) break;                     // (the list is final: the first launch has completed)
                    const int2 it = reinterpret_cast<const int2*>(a.ws + a.L.huge)[t];
                    g = it.x; tgt0 = it.y; part = true;
                } else if (!phase2) {
                    g = __shfl_sync(0xffffffffu, g_next, 0);
                    if (g >= total_cells) {
                        // cell queue drained: publish how many cells this CTA has classified, then serve the split list
                        phase2 = true;
                        if (lane == 0) {
                            __threadfence();
                            atomicAdd(&ctrl->cells_done, my_cells);
                        }
                        continue;
                    }
                    if (lane == 0) g_next = atomicAdd(&ctrl->work_counter[0], 1);
                    ++my_cells;
                } else {
                    int ix = 0, iy = 0;
                    if (lane == 0) {
                        const int t = atomicAdd(&ctrl->split_next, 1);
                        volatile int* vdone = &ctrl->cells_done;
                        volatile int* vres = &ctrl->split_reserved;
                        for (;;) {
                            if (t >= NVNL_SPLIT_CAP) break;
                            const unsigned long long w = *reinterpret_cast<volatile unsigned long long*>(&NVNL_SPLIT_PTR[t]);
                            ix = (int)(unsigned)w; iy = (int)(w >> 32);
                            if (ix != 0) break;                       // the part has been pushed
                            if (*vdone >= total_cells) {              // every cell classified: the list is final
                                __threadfence();
                                if (t >= *vres) break;                // ... and entry t does not exist
                            }
                            __nanosleep(200);
                        }
                    }
                    ix = __shfl_sync(0xffffffffu, ix, 0);
                    iy = __shfl_sync(0xffffffffu, iy, 0);
                    if (ix == 0) break;                               // nothing left
                    g = ix - 1; tgt0 = iy; part = true;
                }
                const int home_start = cell_start[g];
                ntarget = cell_start[g + 1] - home_start;
                if (ntarget == 0) continue;
                int s = 0;
                if (a.num_systems > 1) {
                    s = a.batch_idx[sorted[home_start].j];
                    s = s < 0 ? 0 : (s >= a.num_systems ? a.num_systems - 1 : s);   // (k_hash reported the error)
                }
                const SysParams& sp = sys[s];
                if (s != s_cur) {
                    s_cur = s;
                    cpd0 = sp.cpd[0]; cpd1 = sp.cpd[1]; cpd2 = sp.cpd[2];
                    R0 = sp.R[0]; R1 = sp.R[1]; R2 = sp.R[2];
                    pb0 = sp.pbc[0]; pb1 = sp.pbc[1]; pb2 = sp.pbc[2];
                    coff = sp.cell_offset;
                    c01 = cpd0 * cpd1;
                    rcp0 = 1.0f / (float)cpd0;
                    rcp01 = 1.0f / (float)c01;
                    r111 = R0 == 1 && R1 == 1 && R2 == 1;   // the common 3 x 3 x 3 stencil
                    one_cell = c01 * cpd2 == 1;
                }
                // cell coordinates: float-reciprocal quotient, corrected to the exact one
                int cx, cy, cz;
                {
                    const int local = g - coff;
                    int q = (int)((float)local * rcp01), r = local - q * c01;
                    while (r < 0) { --q; r += c01; }
                    while (r >= c01) { ++q; r -= c01; }
                    cz = q;
                    int q2 = (int)((float)r * rcp0), r2 = r - q2 * cpd0;
                    while (r2 < 0) { --q2; r2 += cpd0; }
                    while (r2 >= cpd0) { ++q2; r2 -= cpd0; }
                    cy = q2; cx = r2;
                }
                auto defer_cell = [&]() {
                    // leave the cell to the general kernel as work items of kDeferTargets target atoms; its atoms have
                    // no temporary row
                    const int nitems = (ntarget + kDeferTargets - 1) / kDeferTargets;
                    int base = 0;
                    if (lane == 0) {
                        base = atomicAdd(&ctrl->n_deferred, nitems);
                        ctrl->had_deferred = 1;
                    }
                    base = __shfl_sync(0xffffffffu, base, 0);
                    for (int k = lane; k < nitems; k += 32) deferred[base + k] = make_int2(g, k * kDeferTargets);
                    for (int k = lane; k < ntarget; k += 32) row_ref[sorted[home_start + k].j] = -1;
                };
                auto split_cell = [&]() -> bool {
                    // a cell with many targets (met in the cell queue): push it as parts instead of sweeping it here
                    if (part || ntarget <= kRowsSplitAbove) return false;
                    const int nparts = (ntarget + kRowsSplitPart - 1) / kRowsSplitPart;
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&ctrl->split_reserved, nparts);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    for (int k = lane; k < nparts; k += 32)
                        if (base + k < NVNL_SPLIT_CAP)
                            *reinterpret_cast<volatile unsigned long long*>(&NVNL_SPLIT_PTR[base + k]) =
                                (unsigned long long)(unsigned)(g + 1) | ((unsigned long long)(unsigned)(k * kRowsSplitPart) << 32);
                    __threadfence();     // (every lane's entries are visible before this CTA reports its cells as classified)
                    __syncwarp();
                    return true;
                };
                alias = false;
                bool ok_item = false;
                do {     // (a `continue` in this block leaves it with ok_item == false)
                if (r111 && cx >= 1 && cx <= cpd0 - 2 && cy >= 1 && cy <= cpd1 - 2 && cz >= 1 && cz <= cpd2 - 2) {
                    // ---- interior cell (no wrap, every stencil cell in range): the stencil is nine runs of three
                    //      x-adjacent cells, contiguous in the sorted array — nine lanes, two loads each, nine copies
                    st = 0; cn = 0;
                    if (lane < 9) {
                        const int gc0 = g + (lane % 3 - 1) * cpd0 + (lane / 3 - 1) * c01 - 1;
                        st = cell_start[gc0];
                        cn = cell_start[gc0 + 3] - st;
                    }
                    const int incl = warp_incl_scan(cn, lane);
                    aoff = incl - cn;
                    total = __shfl_sync(0xffffffffu, incl, 31);
                    nvc = ((total + 63) >> 6) << 1;    // chunk count padded to an even number (two chunks per trip)
                    if (nvc > kRowsMaxVC || (PAIR && nvc * 32 * (int)(RS + 8) > kRowsRingBytes)) { defer_cell(); continue; }
                    seg_len = total; seg_vcb = 0; seg_nch = nvc; si = 0; nseg = 1; hm = 1u; head_lane = lane == 0;
                    shiftmask = 0u; key = 0; grp_cn = cn;
                    home_slot = __shfl_sync(0xffffffffu, aoff, 4) + (home_start - __shfl_sync(0xffffffffu, st, 4));
                    ok_item = true;
                    break;
                }
                const int nx = 2 * R0 + 1, ny = 2 * R1 + 1, nzz = 2 * R2 + 1;
                const int nimg = nx * ny * nzz;
                bool ok = nimg <= 32;
                int tag = 0;
                st = 0; cn = 0; key = kKeyEmpty;
                if (ok && lane < nimg) {
                    int dx, dy, dz;
                    if (r111) { dx = lane % 3 - 1; dy = (lane / 3) % 3 - 1; dz = lane / 9 - 1; }
                    else { dx = lane % nx - R0; dy = (lane / nx) % ny - R1; dz = lane / (nx * ny) - R2; }
                    int tx = cx + dx, ty = cy + dy, tz = cz + dz;
                    bool in = true;
                    int csx = 0, csy = 0, csz = 0;
                    if (r111) {
                        // |d| <= 1 in every dimension: one conditional wrap instead of a division
                        if (pb0) { if (tx < 0) { tx += cpd0; csx = -1; } else if (tx >= cpd0) { tx -= cpd0; csx = 1; } }
                        else in = in && tx >= 0 && tx < cpd0;
                        if (pb1) { if (ty < 0) { ty += cpd1; csy = -1; } else if (ty >= cpd1) { ty -= cpd1; csy = 1; } }
                        else in = in && ty >= 0 && ty < cpd1;
                        if (pb2) { if (tz < 0) { tz += cpd2; csz = -1; } else if (tz >= cpd2) { tz -= cpd2; csz = 1; } }
                        else in = in && tz >= 0 && tz < cpd2;
                    } else {
                        if (pb0) divmod_floor(tx, cpd0, csx, tx); else in = in && tx >= 0 && tx < cpd0;
                        if (pb1) divmod_floor(ty, cpd1, csy, ty); else in = in && ty >= 0 && ty < cpd1;
                        if (pb2) divmod_floor(tz, cpd2, csz, tz); else in = in && tz >= 0 && tz < cpd2;
                    }
                    if (in) {
                        const int gc = coff + tx + cpd0 * (ty + cpd1 * tz);
                        st = cell_start[gc];
                        cn = cell_start[gc + 1] - st;
                        if (cn > 0) key = pack_key(csx, csy, csz);
                    }
                    tag = (dx == 0 && dy == 0 && dz == 0) ? 1 : 0;
                }
                shiftmask = __ballot_sync(0xffffffffu, key != 0 && key != kKeyEmpty);
                if (ok && shiftmask) {
                    // order the images by shift: equal shifts become one contiguous segment, zero shift first
                    int rank = 0;
                    for (int t = 0; t < 32; ++t) {
                        const int kt = __shfl_sync(0xffffffffu, key, t);
                        rank += (kt < key || (kt == key && t < lane)) ? 1 : 0;
                    }
                    sm.e_st[rank] = st; sm.e_cn[rank] = cn; sm.e_key[rank] = key; sm.e_tag[rank] = tag;
                    __syncwarp();
                    st = sm.e_st[lane]; cn = sm.e_cn[lane]; key = sm.e_key[lane]; tag = sm.e_tag[lane];
                    __syncwarp();
                }
                const int incl = warp_incl_scan(cn, lane);
                const int off = incl - cn;            // dense offset of this image
                total = __shfl_sync(0xffffffffu, incl, 31);
                nseg = 1;
                hm = 1u;
                head_lane = lane == 0;
                // a system that is ONE cell: every image is the same run of records -> stage it once and let the image
                // segments alias it (virtual chunks = image * chunks-per-image; up to kRowsMaxVCAlias of them)
                // (only when the images do not fit kRowsMaxVC chunks staged one by one; smaller tiles take the regular path)
                alias = ok && shiftmask != 0u && one_cell;
                if (alias) {
                    const unsigned nonempty0 = __ballot_sync(0xffffffffu, cn > 0);
                    const int c0 = __shfl_sync(0xffffffffu, cn, 0);
                    if (__popc(nonempty0) * ((c0 + 31) >> 5) <= kRowsMaxVC) alias = false;
                }
                if (PAIR && alias) { defer_cell(); continue; }   // (the pair consumer leaves these to the host's fallback)
                if (alias) {
                    const unsigned nonempty = __ballot_sync(0xffffffffu, cn > 0);
                    const int cn1 = __shfl_sync(0xffffffffu, cn, 0), st1 = __shfl_sync(0xffffffffu, st, 0);   // (zero shift first)
                    const int nch = (cn1 + 31) >> 5;
                    nseg = __popc(nonempty);
                    hm = nonempty;
                    head_lane = cn > 0;
                    si = __popc(nonempty & (0xffffffffu >> (31 - lane))) - 1;
                    seg_len = cn > 0 ? cn1 : 0;
                    seg_nch = cn > 0 ? nch : 0;
                    seg_vcb = si * nch;
                    nvc = nseg * nch;
                    total = cn1;                       // records staged
                    aoff = 0;
                    st = st1;
                    grp_cn = lane == 0 ? cn1 : 0;
                    home_slot = home_start - st1;
                    if (head_lane) {
                        int csx, csy, csz;
                        unpack_key(key, csx, csy, csz);
                        T cm[9];
#pragma unroll
                        for (int k = 0; k < 9; ++k) cm[k] = (T)sp.cellm[k];
                        shift_vector<T, FMA>(cm, csx, csy, csz, Sx, Sy, Sz);
                    }
                    if (nvc > kRowsMaxVCAlias || nch * 32 * (int)RS > kRowsRingBytes) { defer_cell(); continue; }
                    if (!HUGE && nvc > kRowsMaxVC) {
                        // more than two mask words: the cell goes to the second launch, as parts of kRowsSplitPart targets
                        const int nparts = (ntarget + kRowsHugePart - 1) / kRowsHugePart;
                        int base = 0;
                        if (lane == 0) {
                            base = atomicAdd(&ctrl->n_huge, nparts);
                            ctrl->had_huge = 1;
                        }
                        base = __shfl_sync(0xffffffffu, base, 0);
                        int2* huge = reinterpret_cast<int2*>(a.ws + a.L.huge);
                        for (int k = lane; k < nparts; k += 32) huge[base + k] = make_int2(g, k * kRowsHugePart);
                        continue;
                    }
                    ok_item = true;
                    break;
                }
                if (!shiftmask) {
                    // interior cell: one zero-shift segment, chunk count padded to an even number (two chunks per trip)
                    aoff = off;
                    nvc = ((total + 63) >> 6) << 1;
                    seg_len = total; seg_vcb = 0; seg_nch = nvc; si = 0;
                } else {
                    // segments = runs of equal shift among the non-empty images; each starts on a 32-slot boundary
                    const int pk = __shfl_up_sync(0xffffffffu, key, 1);
                    head_lane = cn > 0 && (lane == 0 || pk != key);
                    hm = __ballot_sync(0xffffffffu, head_lane);
                    nseg = __popc(hm);
                    const unsigned le = 0xffffffffu >> (31 - lane);                 // lanes <= me
                    const int hl = 31 - __clz((int)(hm & le) | 1);                  // head lane of my segment
                    const unsigned gt = lane == 31 ? 0u : (0xffffffffu << (lane + 1));
                    const unsigned nh = hm & gt;                                    // heads after me
                    const int nhl = nh ? __ffs((int)nh) - 1 : 31;
                    const int off_next = __shfl_sync(0xffffffffu, off, nhl);
                    seg_len = head_lane ? ((nh ? off_next : total) - off) : 0;
                    seg_nch = (seg_len + 31) >> 5;
                    const int vincl = warp_incl_scan(seg_nch, lane);
                    seg_vcb = vincl - seg_nch;
                    nvc = __shfl_sync(0xffffffffu, vincl, 31);
                    si = __popc(hm & le) - 1;
                    const int seg_off = __shfl_sync(0xffffffffu, off, hl);
                    const int seg_slot = __shfl_sync(0xffffffffu, seg_vcb, hl) << 5;
                    aoff = seg_slot + (off - seg_off);
                    if (head_lane) {
                        int csx, csy, csz;
                        unpack_key(key, csx, csy, csz);
                        T cm[9];
#pragma unroll
                        for (int k = 0; k < 9; ++k) cm[k] = (T)sp.cellm[k];
                        shift_vector<T, FMA>(cm, csx, csy, csz, Sx, Sy, Sz);
                        if (PAIR) {
                            S64x = (double)csx * sp.cellm[0] + (double)csy * sp.cellm[3] + (double)csz * sp.cellm[6];
                            S64y = (double)csx * sp.cellm[1] + (double)csy * sp.cellm[4] + (double)csz * sp.cellm[7];
                            S64z = (double)csx * sp.cellm[2] + (double)csy * sp.cellm[5] + (double)csz * sp.cellm[8];
                        }
                    }
                }
                ok = ok && nvc <= kRowsMaxVC && nseg <= kRowsMaxSeg && !(PAIR && nvc * 32 * (int)(RS + 8) > kRowsRingBytes);
                if (!ok) { defer_cell(); continue; }
                const unsigned tagm = __ballot_sync(0xffffffffu, tag != 0);
                const int home_lane = __ffs(tagm) - 1;
                home_slot = __shfl_sync(0xffffffffu, aoff, home_lane);
                // copy groups: images whose source runs AND destinations are contiguous (x-adjacent cells of one
                // stencil row) go as ONE bulk copy, issued by the first lane of the group
                {
                    const int pe_src = __shfl_up_sync(0xffffffffu, st + cn, 1);
                    const int pe_dst = __shfl_up_sync(0xffffffffu, aoff + cn, 1);
                    const bool cont = lane > 0 && st == pe_src && aoff == pe_dst;
                    const unsigned gh = __ballot_sync(0xffffffffu, !cont);                 // group heads
                    const unsigned gt2 = lane == 31 ? 0u : (0xffffffffu << (lane + 1));
                    const unsigned ngh = gh & gt2;
                    const int last = ngh ? __ffs((int)ngh) - 2 : 31;                        // last lane of my group
                    const int incl_last = __shfl_sync(0xffffffffu, incl, last);
                    grp_cn = cont ? 0 : incl_last - off;
                }
                ok_item = true;
                } while (0);
                if (!ok_item) continue;      // empty, deferred: next work item
                if (split_cell()) continue;  // pushed as parts
                if (!part) tgt0 = 0;
                {
                    const int plen = HUGE ? kRowsHugePart : kRowsSplitPart;
                    tgt1 = part && tgt0 + plen < ntarget ? tgt0 + plen : ntarget;
                }
                have = true;
                break;
            }
            // ---- 2. descriptor slot + ring space (FIFO release), then publish the tables
            //         and issue the copies ----
            const int dslot = nprod % kRowsDesc;
            const int bytes = have ? (alias ? ((total + 31) >> 5) : nvc) * 32 * (int)(PAIR ? RS + 8 : RS) : 0;   // PAIR: + charges
            int foot, data_off;
            for (;;) {
                foot = bytes; data_off = head;
                if (head + bytes > kRowsRingBytes) { foot += kRowsRingBytes - head; data_off = 0; }   // wrap: the tail fragment is charged to this tile
                if (nprod - oldest < kRowsDesc && used + foot <= kRowsRingBytes) break;
                const int os = oldest % kRowsDesc;
                uint64_t* eb = reinterpret_cast<uint64_t*>(&sm.empty[os]);
                const uint32_t ep = (uint32_t)((oldest / kRowsDesc) & 1);
                mbar_wait_backoff<400>(eb, ep);   // (the ring is full: the producer is tiles ahead, long naps cost nothing)
                used -= sm.desc[os].footprint;
                ++oldest;
                if (oldest == nprod) { head = 0; used = 0; }   // ring drained: start over at its beginning
            }
            RowsDesc& ds = sm.desc[dslot];
            if (!have) {
                if (lane == 0) {
                    ds.item = -1;
                    ds.footprint = 0;
                    mbar_arrive(reinterpret_cast<uint64_t*>(&sm.full[dslot]));
                }
                break;
            }
            used += foot;
            head = data_off + bytes;
            const uint32_t tile_addr = ring_addr + (uint32_t)data_off;
            if (head_lane) {
                ds.seg_vc[si] = seg_vcb;
                ds.seg_key[si] = shiftmask ? key : 0;
                ds.segS[si][0] = Sx; ds.segS[si][1] = Sy; ds.segS[si][2] = Sz;
                if (PAIR) { segS64[dslot][si][0] = S64x; segS64[dslot][si][1] = S64y; segS64[dslot][si][2] = S64z; }
            }
            if (PAIR) {
                // the charges of the staged records, image by image (copied by the lanes: runs of 8-byte values are not
                // 16-byte aligned, so they cannot ride the bulk copies)
                const uint32_t q_addr = tile_addr + (uint32_t)nvc * (32u * RS);
                unsigned live = __ballot_sync(0xffffffffu, cn > 0);
                while (live) {
                    const int l = __ffs((int)live) - 1;
                    live &= live - 1u;
                    const int s_ = __shfl_sync(0xffffffffu, st, l), c_ = __shfl_sync(0xffffffffu, cn, l);
                    const int a_ = __shfl_sync(0xffffffffu, aoff, l);
                    for (int t = lane; t < c_; t += 32) sts_f64(q_addr + (uint32_t)(a_ + t) * 8u, a.q_sorted[s_ + t]);
                }
            }
            if (lane == 0) {
                ds.seg_vc[nseg] = nvc;
                ds.item = g; ds.ntarget = tgt1; ds.home_slot = home_slot; ds.nseg = nseg; ds.nvc = nvc;
                ds.data_off = data_off; ds.shifted = shiftmask ? 1 : 0; ds.next_target = tgt0; ds.footprint = foot;
                ds.aliased = alias ? 1 : 0;
            }
            // sentinel records behind every segment (up to its padded end) and the chunk -> segment table
            if (alias) {
                const int nch = (total + 31) >> 5;
                for (int q = total + lane; q < (nch << 5); q += 32)
                    sts_rec(tile_addr + (uint32_t)q * RS, kRowsFar, kRowsFar, kRowsFar, -1);
                for (int v = lane; v < nvc; v += 32) {
                    const int sg = v / nch;
                    ds.vc_seg[v] = (unsigned char)sg;
                    ds.vc_phys[v] = (unsigned char)(v - sg * nch);
                }
            } else {
                int myseg0 = 0, myseg1 = 0;
                unsigned rest = hm;
                while (rest) {
                    const int h = __ffs((int)rest) - 1;
                    rest &= rest - 1u;
                    const int vb = __shfl_sync(0xffffffffu, seg_vcb, h), nch = __shfl_sync(0xffffffffu, seg_nch, h);
                    const int len = __shfl_sync(0xffffffffu, seg_len, h), sidx = __shfl_sync(0xffffffffu, si, h);
                    if (lane >= vb && lane < vb + nch) myseg0 = sidx;
                    if (lane + 32 >= vb && lane + 32 < vb + nch) myseg1 = sidx;
                    const int gb = (vb << 5) + len, ge = (vb + nch) << 5;     // < 64 slots
                    for (int q = gb + lane; q < ge; q += 32)
                        sts_rec(tile_addr + (uint32_t)q * RS, kRowsFar, kRowsFar, kRowsFar, -1);
                }
                if (shiftmask) {
                    ds.vc_seg[lane] = (unsigned char)myseg0;
                    ds.vc_seg[lane + 32] = (unsigned char)myseg1;
                }
            }
            const uint32_t tx = (uint32_t)total * RS;
            __syncwarp();  // every lane's table / sentinel writes precede lane 0's release-arrive below
            if (lane == 0) mbar_arrive_expect_tx(reinterpret_cast<uint64_t*>(&sm.full[dslot]), tx);
            __syncwarp();
            if (grp_cn > 0)
                tma_load_1d(smem_raw + data_off + (size_t)aoff * RS, sorted + st, (uint32_t)grp_cn * RS,
                            reinterpret_cast<uint64_t*>(&sm.full[dslot]));
            ++nprod;
            if (lane == 0 && zpos < zend) zero_some(zquota);
        }
        if (lane == 0 && zpos < zend) zero_some(zend - zpos);
        if (lane == 0 && zend > 0) tma_store_wait_all();
        // the last CTA to drain the queue re-arms it for the next launch on this workspace
        if (HUGE) {
            if (lane == 0) {
                __threadfence();
                if (atomicAdd(&ctrl->huge_done, 1) == (int)gridDim.x - 1) {
                    ctrl->huge_next = 0;
                    ctrl->huge_done = 0;
                }
            }
        } else {
            int dn = 0;
            if (lane == 0) {
                __threadfence();
                dn = atomicAdd(&ctrl->done[0], 1);
            }
            dn = __shfl_sync(0xffffffffu, dn, 0);
            if (dn == (int)gridDim.x - 1) {
                // (every other CTA has left its pop loop: the split list can be cleared for the next query)
                const int used_parts = *reinterpret_cast<volatile int*>(&ctrl->split_reserved);
                for (int k = lane; k < used_parts && k < NVNL_SPLIT_CAP; k += 32) NVNL_SPLIT_PTR[k] = make_int2(0, 0);
                __syncwarp();
                if (lane == 0) {
                    ctrl->work_counter[0] = 0;
                    ctrl->done[0] = 0;
                    ctrl->split_reserved = 0;
                    ctrl->split_next = 0;
                    ctrl->cells_done = 0;
                }
            }
        }
#undef NVNL_SPLIT_PTR
#undef NVNL_SPLIT_CAP
    } else {
        // =========================== consumers ===========================
        int* rows = reinterpret_cast<int*>(a.ws + a.L.rows);
        int* row_ref = reinterpret_cast<int*>(a.ws + a.L.row_ref);
        RowsAlloc al;
        al.pos = al.end = 0;
        const float rc2 = a.cutoff_sq;
        const long long rows_cap = a.L.rows_cap;
        for (int ncons = 0;; ++ncons) {
            const int dslot = ncons % kRowsDesc;
            mbar_wait_backoff(reinterpret_cast<uint64_t*>(&sm.full[dslot]), (uint32_t)((ncons / kRowsDesc) & 1));
            RowsDesc& ds = sm.desc[dslot];
            if (ds.item < 0) break;
            const uint32_t tile_addr = ring_addr + (uint32_t)ds.data_off;
            const int ntarget = ds.ntarget, home_slot = ds.home_slot;
            const int words = ds.nvc <= 32 ? 1 : 2;   // mask words per lane and target (first launch)
            for (;;) {
                int t0 = 0;
                if (lane == 0) t0 = atomicAdd(&ds.next_target, 4);
                t0 = __shfl_sync(0xffffffffu, t0, 0);
                if (t0 >= ntarget) break;
                const int nt = ntarget - t0 < 4 ? ntarget - t0 : 4;
                if (PAIR) {
                    if (words == 1) pair_trip<FMA, 1>(a, ds, segS64[dslot], tile_addr, home_slot + t0, nt, rc2, lane);
                    else pair_trip<FMA, 2>(a, ds, segS64[dslot], tile_addr, home_slot + t0, nt, rc2, lane);
                } else if (HUGE)
                    rows_trip<HALF, FMA, kRowsHugeWords>(rows_cap, a.num_neighbors, ds, ctrl, tile_addr, home_slot + t0, nt, rc2,
                                                         lane, al, rows, row_ref);
                else if (words == 1)
                    rows_trip<HALF, FMA, 1>(rows_cap, a.num_neighbors, ds, ctrl, tile_addr, home_slot + t0, nt, rc2, lane, al,
                                            rows, row_ref);
                else
                    rows_trip<HALF, FMA, 2>(rows_cap, a.num_neighbors, ds, ctrl, tile_addr, home_slot + t0, nt, rc2, lane, al,
                                            rows, row_ref);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(reinterpret_cast<uint64_t*>(&sm.empty[dslot]));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_rows_out: the final arrays in atom-index order.  (A two-stage software pipeline per warp — bulk copies of batch k + 1 in
// flight while batch k is written, row references loaded one group ahead — was measured and changed nothing: 0.297 vs
// 0.292 ms, profiles/r2_variants_c.txt; the kernel sits at what DRAM gives this read/write mix.)
// One warp per 32 consecutive atoms: every lane issues ONE TMA bulk
// copy that brings its atom's temporary row into the warp's shared-memory staging buffer (as many rows per batch as
// fit), then the warp writes out_i (= i), out_j and, for rows of cells at a periodic boundary, the shifts unpacked from
// the image keys in front of the row — all strictly sequential in the large.
// Atoms with row_ref < 0 were handled by the general kernel, which writes their rows itself.
// ------------------------------------------------------------------------------------------------
constexpr int kOutWarps = 8;
#ifndef NVNL_OUT_CAP
#define NVNL_OUT_CAP 2048
#endif
constexpr int kOutCap = NVNL_OUT_CAP;   // staging entries per warp (8 KB)

__device__ __forceinline__ void warp_fill(int* __restrict__ dst, int n, int value, int lane) {
    int head = (int)(((16u - (unsigned)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u) >> 2);
    head = head < n ? head : n;
    if (lane < head) dst[lane] = value;
    int4* v = reinterpret_cast<int4*>(dst + head);
    const int nv = (n - head) >> 2;
    const int4 val = make_int4(value, value, value, value);
    for (int q = lane; q < nv; q += 32) v[q] = val;
    const int tail = (n - head) & 3;
    if (lane < tail) dst[head + 4 * nv + lane] = value;
}

struct RowsOutSmem {
    alignas(128) int buf[kOutWarps][kOutCap];
    unsigned long long bar[kOutWarps];
};

// writes one row (already in shared or global memory at `row`, header of `hdr` keys in front of it): the general case,
// kept out of line so that the hot loop of k_rows_out stays small
__device__ __noinline__ void rows_out_row(const int* row, int hdr, int cnt, int iv, int index_offset, int shifts_zeroed,
                                          int* __restrict__ oi, int* __restrict__ oj, int* __restrict__ sh, int lane) {
    if (hdr == 0) {
        for (int k = lane; k < cnt; k += 32) {
            oj[k] = row[k] + index_offset;
            oi[k] = iv;
        }
        if (!shifts_zeroed) warp_fill(sh, 3 * cnt, 0, lane);
    } else if (!shifts_zeroed) {
        // every component of every shift, zeros included, as coalesced stores: element e = lane + 32 u of a group of 32
        // pairs (96 ints) belongs to pair e / 3, component e % 3
        const int keyl = lane < hdr ? row[lane - hdr] : 0;
        int q[3], c8[3];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int e = lane + 32 * u;
            q[u] = e / 3;
            c8[u] = 8 * (e - 3 * q[u]);
        }
        for (int k0 = 0; k0 < cnt; k0 += 32) {
            const int k = k0 + lane;
            const int e = k < cnt ? row[k] : 0;
            if (k < cnt) {
                oj[k] = (e & ((1 << kRowsSegShift) - 1)) + index_offset;
                oi[k] = iv;
            }
            const int key = __shfl_sync(0xffffffffu, keyl, ((unsigned)e >> kRowsSegShift) & 31);
            const int pk = key == 0 ? kZeroPack : key - 1;           // bytes = shift component + 128
            int* __restrict__ shg = sh + 3 * k0;
            const int nel = 3 * (cnt - k0);
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int pq = __shfl_sync(0xffffffffu, pk, q[u]);
                if (lane + 32 * u < nel) shg[lane + 32 * u] = ((pq >> c8[u]) & 255) - 128;
            }
        }
    } else {
        const int keyl = lane < hdr ? row[lane - hdr] : 0;
        for (int k0 = 0; k0 < cnt; k0 += 32) {
            const int k = k0 + lane;
            const int e = k < cnt ? row[k] : 0;
            const int key = __shfl_sync(0xffffffffu, keyl, ((unsigned)e >> kRowsSegShift) & 31);
            if (k < cnt) {
                oj[k] = (e & ((1 << kRowsSegShift) - 1)) + index_offset;
                oi[k] = iv;
                if (key != 0) {
                    int csx, csy, csz;
                    unpack_key(key, csx, csy, csz);
                    sh[3 * k] = csx;
                    sh[3 * k + 1] = csy;
                    sh[3 * k + 2] = csz;
                }
            }
        }
    }
}

// SPEC: launched BEFORE the host knows the pair count, into buffers sized from a guess — the kernel reads the count
// itself, places row 1 of edge_index behind it, and does nothing if the guess was too small or the query was served by
// the two-pass kernels (the host then repeats the fill the regular way).
template <bool SPEC>
__global__ void __launch_bounds__(kOutWarps * 32) k_rows_out(const unsigned char* __restrict__ ws, WsLayout L, long long n,
                                                             const int* __restrict__ neighbor_ptr, int* __restrict__ out_i,
                                                             int* __restrict__ out_j_arg, int* __restrict__ out_shifts,
                                                             int index_offset, int shifts_zeroed, long long spec_cap, int group) {
    // group = atoms a warp takes per pass (32, or 8 for small inputs: more, shorter dependent chains per warp)
    pdl_enter();
    extern __shared__ __align__(128) unsigned char out_smem_raw[];
    RowsOutSmem& sm = *reinterpret_cast<RowsOutSmem*>(out_smem_raw);
    int* __restrict__ out_j = out_j_arg;
    if (SPEC) {
        const Ctrl* ctrl = reinterpret_cast<const Ctrl*>(ws + L.ctrl);
        const unsigned long long total = ctrl->total_pairs;
        if (ctrl->unwrapped != 0 || ctrl->rows_overflow != 0 || total > (unsigned long long)spec_cap) return;
        out_j = out_i + total;
    }
    // (the sweep did not pre-zero the shifts of a shift-heavy workload, whatever buffer it was handed)
    if (reinterpret_cast<const Ctrl*>(ws + L.ctrl)->shift_heavy != 0) shifts_zeroed = 0;
    const int* __restrict__ rows = reinterpret_cast<const int*>(ws + L.rows);
    const int* __restrict__ row_ref = reinterpret_cast<const int*>(ws + L.row_ref);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t* bar = reinterpret_cast<uint64_t*>(&sm.bar[warp]);
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncwarp();
    int* buf = sm.buf[warp];
    uint32_t parity = 0;
    const long long nwarps = (long long)gridDim.x * kOutWarps;
    for (long long base = ((long long)blockIdx.x * kOutWarps + warp) * group; base < n; base += nwarps * group) {
        const long long il = base + lane;
        const bool mine = lane < group && il < n;
        const int ref = mine ? row_ref[il] : -1;
        const int p = neighbor_ptr[mine ? il : (base < n ? base : n)];
        const int cnt = mine ? neighbor_ptr[il + 1] - p : 0;
        const int na = n - base < group ? (int)(n - base) : group;
        const int flag = ref & 3;
        const int hdr = ref < 0 ? 0 : (flag == 0 ? 0 : (flag == 1 ? 8 : 32));
        const int clen = (ref < 0 || cnt == 0) ? 0 : hdr + ((cnt + 3) & ~3);      // entries to copy
        const int* src = rows + (ref >> 2);
        int s = 0;
        while (s < na) {
            const int v = lane >= s ? clen : 0;
            const int incl = warp_incl_scan(v, lane);
            const unsigned fit = __ballot_sync(0xffffffffu, lane >= s && lane < na && incl <= kOutCap);
            // lanes [s, e) fit one batch (incl is monotone: the fitting lanes are a prefix)
            const int e = s + __popc(fit);
            if (e == s) {
                // a single row longer than the staging buffer: straight from global memory
                const int ref_t = __shfl_sync(0xffffffffu, ref, s), p_t = __shfl_sync(0xffffffffu, p, s);
                const int cnt_t = __shfl_sync(0xffffffffu, cnt, s), hdr_t = __shfl_sync(0xffffffffu, hdr, s);
                if (ref_t >= 0 && cnt_t > 0)
                    rows_out_row(rows + (ref_t >> 2) + hdr_t, hdr_t, cnt_t, (int)(base + s) + index_offset, index_offset,
                                 shifts_zeroed, out_i + (size_t)p_t, out_j + (size_t)p_t, out_shifts + 3 * (size_t)p_t, lane);
                ++s;
                continue;
            }
            const int tot = __shfl_sync(0xffffffffu, incl, e - 1);
            if (tot > 0) {
                fence_proxy_async_smem();
                if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)tot * 4u);
                __syncwarp();
                if (lane >= s && lane < e && v > 0) tma_load_1d(buf + (incl - v), src, (uint32_t)v * 4u, bar);
                mbar_wait(bar, parity);
                parity ^= 1u;
                // per-row metadata in one word: row length (16 bits) | header kind (2 bits) | offset in the staging buffer
                const int meta = cnt | ((ref & 3) << 16) | ((incl - v) << 18);
                const uint32_t buf_addr = smem_u32(buf) + (uint32_t)lane * 4u;
#pragma unroll 1
                for (int t = s; t < e; ++t) {
                    const int m_t = __shfl_sync(0xffffffffu, meta, t), p_t = __shfl_sync(0xffffffffu, p, t);
                    const int cnt_t = m_t & 0xffff, kind_t = (m_t >> 16) & 3, off_t = (int)((unsigned)m_t >> 18);
                    if (cnt_t == 0 || __shfl_sync(0xffffffffu, ref, t) < 0) continue;
                    const int iv = (int)(base + t) + index_offset;
                    if (kind_t == 0 && cnt_t <= 96 && shifts_zeroed) {
                        // the common case, unrolled: interior row, shifts already zero
                        int* __restrict__ oj = out_j + (size_t)p_t + lane;
                        int* __restrict__ oi = out_i + (size_t)p_t + lane;
                        const uint32_t ra = buf_addr + (uint32_t)off_t * 4u;
#pragma unroll
                        for (int u = 0; u < 3; ++u) {
                            if (lane + 32 * u < cnt_t) {
                                oj[32 * u] = lds_b32(ra + 128u * u) + index_offset;
                                oi[32 * u] = iv;
                            }
                        }
                    } else {
                        const int hdr_t = kind_t == 0 ? 0 : (kind_t == 1 ? 8 : 32);
                        rows_out_row(buf + off_t + hdr_t, hdr_t, cnt_t, iv, index_offset, shifts_zeroed, out_i + (size_t)p_t,
                                     out_j + (size_t)p_t, out_shifts + 3 * (size_t)p_t, lane);
                    }
                }
                __syncwarp();   // every lane is done with the buffer before the next batch lands in it
            }
            s = e;
        }
    }
}

// resets the per-query state of the control block
__global__ void k_query_reset(unsigned char* __restrict__ ws, WsLayout L, long long n, int with_rows) {
    pdl_enter();
    Ctrl* ctrl = reinterpret_cast<Ctrl*>(ws + L.ctrl);
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid == 0) {
        ctrl->n_deferred = 0;
        ctrl->rows_cursor = 0ull;
        ctrl->rows_overflow = 0;
        ctrl->split_next = 0;
        ctrl->cells_done = 0;
        ctrl->n_huge = 0;
        ctrl->huge_next = 0;
        ctrl->huge_done = 0;
        // (split_reserved counts entries the last sweep may not have cleared if it was aborted; k_rows clears and resets it)
    }
    (void)n; (void)with_rows;
}

}  // namespace nvnl
