// nvnl_rows.cuh — single-sweep COO path for fp32 inputs inside the primary periodic image.
//
// Why: the two-pass COO path (nvnl_fast.cuh: count -> hit masks -> scan -> mask expansion) was issue-bound in BOTH
// sweeps (330 + 571 warp-instructions per atom, profiles/r1_final_1m_ncu.txt) and its fill pass wrote 360-byte rows
// at random places of the output (cell-ordered sweep over randomly labelled atoms: 0.72 ms write-pattern floor,
// profiles/r1_microbench_write_pattern.txt).  This path computes every distance once AND compacts once:
//
//   k_rows      (cell order)   stencil sweep; every lane APPENDS the tile index of its own hits to a lane-private
//                              list in shared memory (predicated 16-bit store + pointer bump: 2 instructions per
//                              target per 32 candidates, no ballot, no popc); per target one warp scan of the 32
//                              list lengths, then the lists are gathered (tile index -> original atom index) into a
//                              compact row of a TEMPORARY buffer handed out in blocks by one global cursor
//                              (rows of a warp's targets are contiguous -> sequential DRAM writes).
//                              Emits num_neighbors[i] and row_ref[i] (row location).
//   k_scan                     neighbor_ptr (unchanged).
//   k_rows_out  (index order)  streams the final arrays: rows are read from the temporary buffer (random 360-byte
//                              reads) and edge_index / shifts are written strictly sequentially.
//
// Cells the lean kernel cannot take (too many images / candidates / targets) go to the general kernel exactly as in
// the two-pass path; unwrapped inputs and fp64 use the two-pass path.
// Replaces (different algorithm, same result): cell_list.py:372-556 + neighbor_utils.py:106-147, 362-441.
#pragma once
#include "nvnl_fast.cuh"

namespace nvnl {

constexpr int kRowsCons = 8;                          // consumer warps per CTA
constexpr int kRowsThreads = (kRowsCons + 1) * 32;    // + producer warp
constexpr int kRowsStages = 3;                        // default TMA ring depth (consumers may run up to two cells apart)
constexpr int kRowsMaxSeg = 8;                        // image segments per stencil the lean kernel takes (3 bits in a list entry)
constexpr int kRowsBlock = 2048;                      // temp-buffer entries a warp reserves per cursor bump

template <int STAGES>
struct RowsSmem {
    FastStage<float> stage[STAGES];
    int e_st[32], e_cn[32], e_key[32], e_tag[32];                           // producer scratch (shift sort)
    // lane-private hit lists, two targets per warp: slot s of lane l is byte s * 32 + l; an entry is
    // chunk | segment << 5 (the candidate's tile index is chunk * 32 + l); a lane cannot have more hits than chunks
    alignas(16) unsigned char lists[kRowsCons][2][32 * 32];
    unsigned long long full[STAGES], empty[STAGES];                         // mbarriers of the ring
};

template <int STAGES>
constexpr size_t rows_smem_bytes() { return (size_t)STAGES * kFastStageBytes + sizeof(RowsSmem<STAGES>); }

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
};

__device__ __forceinline__ void sts_u8(uint32_t addr, int v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ int lds_u8(uint32_t addr) {
    int v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
// the k_rows barriers are polled with a short sleep between tries: a spinning warp would otherwise take issue slots
// from the sweeping warps of its scheduler (ncu: 13 % of the issued instructions were the bare try_wait loop)
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra NVNL_BDONE_%=;\n\t"
        "NVNL_BWAIT_%=:\n\t"
        "nanosleep.u32 64;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra NVNL_BWAIT_%=;\n\t"
        "NVNL_BDONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// squared distances of one candidate to the two packed targets
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

// lanes that hit append `val` = chunk | segment << 5 to their private lists (predicated 8-bit store + pointer
// bump).  EXCL: the targets themselves (tile indices selfA / selfB, zero shift, d = 0) are not recorded — (i, i, 0) is
// not a pair; half fill needs no such test (i < j fails).
template <bool HALF, bool TAIL, bool EXCL>
__device__ __forceinline__ void rows_append(f32x2_t d2, int c, int val, int j, int iA, int iB, float rc2, bool lexpos,
                                            bool valid, int selfA, int selfB, uint32_t& la, uint32_t& lb) {
    float dA, dB;
    unpack2(d2, dA, dB);
    bool hA = dA < rc2, hB = dB < rc2;
    if (TAIL) { hA = hA && valid; hB = hB && valid; }
    if (EXCL) { hA = hA && c != selfA; hB = hB && c != selfB; }
    if (HALF) {
        hA = hA && (iA < j || (iA == j && lexpos));
        hB = hB && (iB < j || (iB == j && lexpos));
    }
    if (hA) { sts_u8(la, val); la += 32u; }
    if (hB) { sts_u8(lb, val); lb += 32u; }
}

// one 32-candidate chunk against two targets (boundary cells: single chunks and chunks with image shifts)
template <bool HALF, bool FMA, bool SHIFTED, bool TAIL>
__device__ __forceinline__ void rows_chunk2(uint32_t addr, int c, int sg, f32x2_t XI, f32x2_t YI, f32x2_t ZI, int iA, int iB,
                                            float Sx, float Sy, float Sz, float rc2, bool lexpos, bool valid, int selfA,
                                            int selfB, uint32_t& la, uint32_t& lb) {
    float x, y, z;
    int j;
    lds_rec(addr, x, y, z, j);
    const f32x2_t d2 = rows_d2<FMA, SHIFTED>(x, y, z, XI, YI, ZI, Sx, Sy, Sz);
    rows_append<HALF, TAIL, !HALF>(d2, c, (c >> 5) | (sg << 5), j, iA, iB, rc2, lexpos, valid, selfA, selfB, la, lb);
}

// four zero-shift chunks: all loads first, then the four independent FP chains, then the appends (ILP inside the warp).
// MASKED: the group may run past the end of the tile — candidates with index >= limit are ignored (whatever the
// stage buffer holds there is read but never recorded).
template <bool HALF, bool FMA, bool MASKED, bool EXCL>
__device__ __forceinline__ void rows_chunk2x4(uint32_t addr, int c, int limit, f32x2_t XI, f32x2_t YI, f32x2_t ZI, int iA,
                                              int iB, float rc2, int selfA, int selfB, uint32_t& la, uint32_t& lb) {
    constexpr uint32_t RS = sizeof(Rec<float>);
    float x[4], y[4], z[4];
    int j[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) lds_rec(addr + (uint32_t)u * 32u * RS, x[u], y[u], z[u], j[u]);
    f32x2_t d2[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) d2[u] = rows_d2<FMA, false>(x[u], y[u], z[u], XI, YI, ZI, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 4; ++u)
        rows_append<HALF, MASKED, EXCL>(d2[u], c + 32 * u, (c >> 5) + u, j[u], iA, iB, rc2, false, c + 32 * u < limit, selfA,
                                        selfB, la, lb);
}

// four chunks that each lie entirely inside ONE image segment (warp-uniform shift vector per chunk): same structure as
// rows_chunk2x4 plus the three packed adds of the shift.  Experiment (UNI), see k_rows.
template <bool HALF, bool FMA>
__device__ __forceinline__ void rows_chunk2x4_shift(const FastStage<float>& sm, uint32_t addr, int c, int ck, f32x2_t XI,
                                                    f32x2_t YI, f32x2_t ZI, int iA, int iB, float rc2, int selfA, int selfB,
                                                    uint32_t& la, uint32_t& lb) {
    constexpr uint32_t RS = sizeof(Rec<float>);
    float x[4], y[4], z[4];
    int j[4], sgv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) lds_rec(addr + (uint32_t)u * 32u * RS, x[u], y[u], z[u], j[u]);
    f32x2_t d2[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        sgv[u] = sm.chunk_seg[ck + u];
        d2[u] = rows_d2<FMA, true>(x[u], y[u], z[u], XI, YI, ZI, sm.segS[3 * sgv[u]], sm.segS[3 * sgv[u] + 1],
                                   sm.segS[3 * sgv[u] + 2]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        bool lexpos = false;
        if (HALF) {
            int csx, csy, csz;
            unpack_key(sm.seg_key[sgv[u]], csx, csy, csz);
            lexpos = csx > 0 || (csx == 0 && (csy > 0 || (csy == 0 && csz > 0)));
        }
        rows_append<HALF, false, !HALF>(d2[u], c + 32 * u, (ck + u) | (sgv[u] << 5), j[u], iA, iB, rc2, lexpos, true, selfA,
                                        selfB, la, lb);
    }
}

// Sweep of two targets over the staged tile.  la / lb: per-lane append pointers.
//   interior cell (one zero-shift segment): groups of four chunks, the last group masked;
//   cell at a periodic boundary: full zero-shift chunks in groups of four, then every other chunk with its shift vector
//   picked per lane (a chunk may straddle segments).
// Only the group(s) holding the targets themselves pay for the self-exclusion test.
template <bool HALF, bool FMA, bool UNI>
__device__ __forceinline__ void rows_sweep2(const FastStage<float>& sm, uint32_t cand_addr, f32x2_t XI, f32x2_t YI,
                                            f32x2_t ZI, int iA, int iB, float rc2, int lane, int selfA, int selfB,
                                            uint32_t& la, uint32_t& lb) {
    constexpr uint32_t RS = sizeof(Rec<float>);
    constexpr bool EX = !HALF;
    const int total = sm.total, nchunks = sm.nchunks, nseg = sm.nseg;
    const bool zero0 = sm.seg_key[0] == 0;
    const int gA = selfA >> 7, gB = selfB >> 7;
    uint32_t addr = cand_addr + (uint32_t)lane * RS;
    int c = lane;
    if (nseg == 1 && zero0) {
        const int ngroups = total >> 7;
#pragma unroll 1
        for (int g = 0; g < ngroups; ++g) {
            if (EX && (g == gA || g == gB))
                rows_chunk2x4<HALF, FMA, false, true>(addr, c, 0, XI, YI, ZI, iA, iB, rc2, selfA, selfB, la, lb);
            else
                rows_chunk2x4<HALF, FMA, false, false>(addr, c, 0, XI, YI, ZI, iA, iB, rc2, selfA, selfB, la, lb);
            addr += 128 * RS;
            c += 128;
        }
        if (total & 127) rows_chunk2x4<HALF, FMA, true, EX>(addr, c, total, XI, YI, ZI, iA, iB, rc2, selfA, selfB, la, lb);
        return;
    }
    const int zend = zero0 ? sm.seg_begin[1] : 0;
    const int nzfull = zend >> 5;
    int ck = 0;
#pragma unroll 1
    for (; ck + 4 <= nzfull; ck += 4) {
        const int g = ck >> 2;
        if (EX && (g == gA || g == gB))
            rows_chunk2x4<HALF, FMA, false, true>(addr, c, 0, XI, YI, ZI, iA, iB, rc2, selfA, selfB, la, lb);
        else
            rows_chunk2x4<HALF, FMA, false, false>(addr, c, 0, XI, YI, ZI, iA, iB, rc2, selfA, selfB, la, lb);
        addr += 128 * RS;
        c += 128;
    }
#pragma unroll 1
    for (; ck < nzfull; ++ck) {
        rows_chunk2<HALF, FMA, false, false>(addr, c, 0, XI, YI, ZI, iA, iB, 0, 0, 0, rc2, false, true, selfA, selfB, la, lb);
        addr += 32 * RS;
        c += 32;
    }
    const unsigned um = UNI ? (unsigned)sm.qrow[0] : 0u;  // chunks lying entirely inside one segment (producer)
#pragma unroll 1
    for (; ck < nchunks; ++ck) {
        if (UNI && ck + 4 <= nchunks && ((um >> ck) & 15u) == 15u) {
            rows_chunk2x4_shift<HALF, FMA>(sm, addr, c, ck, XI, YI, ZI, iA, iB, rc2, selfA, selfB, la, lb);
            addr += 128 * RS;
            c += 128;
            ck += 3;
            continue;
        }
        int sg = sm.chunk_seg[ck];
        while (sg + 1 < nseg && c >= sm.seg_begin[sg + 1]) ++sg;
        bool lexpos = false;
        if (HALF) {
            int csx, csy, csz;
            unpack_key(sm.seg_key[sg], csx, csy, csz);
            lexpos = csx > 0 || (csx == 0 && (csy > 0 || (csy == 0 && csz > 0)));
        }
        rows_chunk2<HALF, FMA, true, true>(addr, c, sg, XI, YI, ZI, iA, iB, sm.segS[3 * sg], sm.segS[3 * sg + 1],
                                           sm.segS[3 * sg + 2], rc2, lexpos, c < total, selfA, selfB, la, lb);
        addr += 32 * RS;
        c += 32;
    }
}

// Per-warp allocator state of the temporary row buffer.
struct RowsAlloc {
    long long pos, end;
};

// Epilogue of a target pair: drop the self entries, ONE packed warp scan of both targets' list lengths, ONE row
// reservation, then the lists are gathered (tile index -> original atom index) into the two compact rows.
// PAD: every row starts on a 128-byte boundary and is padded to a multiple of 32 entries (the output kernel then reads
// whole lines; experiment, see k_rows).
template <bool HALF, bool PAD>
__device__ __forceinline__ void rows_emit2(const RowsArgs& a, const FastStage<float>& sm, Ctrl* ctrl, uint32_t cand_addr,
                                           uint32_t lbaseA, uint32_t lbaseB, uint32_t la, uint32_t lb, int selfA, int selfB,
                                           int iA, int iB, bool two, int lane, bool shifted, RowsAlloc& al,
                                           int* __restrict__ rows, int* __restrict__ row_ref) {
    constexpr uint32_t RS = sizeof(Rec<float>);
    const int nA = (int)((la - lbaseA) >> 5);
    const int nB = two ? (int)((lb - lbaseB) >> 5) : 0;
    // list lengths are <= 32 per lane, their sums <= 1024: both scans fit one 32-bit word
    const int packed = nA | (nB << 16);
    const int incl = warp_incl_scan(packed, lane);
    const int tot = __shfl_sync(0xffffffffu, incl, 31);
    const int cntA = tot & 0xffff, cntB = tot >> 16;
    const int excl = incl - packed;
    const int exA = excl & 0xffff, exB = excl >> 16;
    const int maxn = __reduce_max_sync(0xffffffffu, nA > nB ? nA : nB);
    // row of a cell at a periodic boundary: [kRowsMaxSeg packed image keys][entries = atom | segment << 28]
    const int hdr = shifted ? (PAD ? 32 : kRowsMaxSeg) : 0;
    const int lenA = PAD ? ((cntA + 31) & ~31) : cntA, lenB = PAD ? ((cntB + 31) & ~31) : cntB;
    const int need = lenA + lenB + 2 * hdr;
    bool ok = true;
    if (al.pos + need > al.end) {
        const long long sz = need > kRowsBlock ? need : kRowsBlock;
        unsigned long long b = 0ull;
        if (lane == 0) b = atomicAdd(&ctrl->rows_cursor, (unsigned long long)sz);
        b = __shfl_sync(0xffffffffu, b, 0);
        if ((long long)b + sz > a.L.rows_cap) {
            // temporary buffer exhausted: the host re-runs the query on the two-pass path (nvnl_status.rows_overflow)
            if (lane == 0) ctrl->rows_overflow = 1;
            al.pos = al.end = 0;
            ok = false;
        } else {
            al.pos = (long long)b;
            al.end = (long long)b + sz;
        }
    }
    int startA = 0, startB = 0;
    if (ok) {
        startA = (int)al.pos;
        startB = startA + hdr + lenA;
        al.pos += need;
        int* __restrict__ rowA = rows + startA + hdr + exA;
        int* __restrict__ rowB = rows + startB + hdr + exB;
        // two slots of both lists per trip: four independent index -> atom gathers in flight.  Slots past a list's
        // length hold stale (valid) entries — the lists are zero-initialised — and are read but not stored.
        // entry -> record address: tile index = chunk * 32 + lane
        const uint32_t cand_lane = cand_addr + (uint32_t)lane * RS;
        if (!shifted) {
#pragma unroll 1
            for (int s = 0; s < maxn; s += 2) {
                const uint32_t o = (uint32_t)s * 32u;
                // & 31: a stale slot may hold an entry (with segment bits) of an earlier boundary cell
                const int cA0 = lds_u8(lbaseA + o) & 31, cA1 = lds_u8(lbaseA + o + 32u) & 31;
                const int cB0 = lds_u8(lbaseB + o) & 31, cB1 = lds_u8(lbaseB + o + 32u) & 31;
                const int jA0 = lds_rec_j<float>(cand_lane + (uint32_t)cA0 * (32u * RS));
                const int jA1 = lds_rec_j<float>(cand_lane + (uint32_t)cA1 * (32u * RS));
                const int jB0 = lds_rec_j<float>(cand_lane + (uint32_t)cB0 * (32u * RS));
                const int jB1 = lds_rec_j<float>(cand_lane + (uint32_t)cB1 * (32u * RS));
                if (s < nA) rowA[s] = jA0;
                if (s + 1 < nA) rowA[s + 1] = jA1;
                if (s < nB) rowB[s] = jB0;
                if (s + 1 < nB) rowB[s + 1] = jB1;
            }
        } else {
            // cell at a periodic boundary: the row starts with the stencil's packed image keys, every entry carries
            // its segment in the top bits (atom indices < 2^28 on this path)
            if (lane < kRowsMaxSeg) {
                const int key = lane < sm.nseg ? sm.seg_key[lane] : 0;
                rows[startA + (PAD ? hdr - kRowsMaxSeg : 0) + lane] = key;   // the keys sit right in front of the entries
                rows[startB + (PAD ? hdr - kRowsMaxSeg : 0) + lane] = key;
            }
#pragma unroll 1
            for (int s = 0; s < maxn; s += 2) {
                const uint32_t o = (uint32_t)s * 32u;
                const int vA0 = lds_u8(lbaseA + o), vA1 = lds_u8(lbaseA + o + 32u);
                const int vB0 = lds_u8(lbaseB + o), vB1 = lds_u8(lbaseB + o + 32u);
                const int jA0 = lds_rec_j<float>(cand_lane + (uint32_t)(vA0 & 31) * (32u * RS));
                const int jA1 = lds_rec_j<float>(cand_lane + (uint32_t)(vA1 & 31) * (32u * RS));
                const int jB0 = lds_rec_j<float>(cand_lane + (uint32_t)(vB0 & 31) * (32u * RS));
                const int jB1 = lds_rec_j<float>(cand_lane + (uint32_t)(vB1 & 31) * (32u * RS));
                if (s < nA) rowA[s] = jA0 | ((vA0 >> 5) << 28);
                if (s + 1 < nA) rowA[s + 1] = jA1 | ((vA1 >> 5) << 28);
                if (s < nB) rowB[s] = jB0 | ((vB0 >> 5) << 28);
                if (s + 1 < nB) rowB[s + 1] = jB1 | ((vB1 >> 5) << 28);
            }
        }
    }
    if (lane < 2 && (lane == 0 || two)) {
        const int i = lane ? iB : iA;
        if (ok) row_ref[i] = (((lane ? startB : startA) + (PAD && shifted ? hdr - kRowsMaxSeg : 0)) << 1) | (shifted ? 1 : 0);
        a.num_neighbors[i] = lane ? cntB : cntA;
    }
}

// ------------------------------------------------------------------------------------------------
// k_rows: warp-specialised persistent kernel (producer identical in role to k_fast's: queue -> stencil images ->
// shift sort -> segment/chunk tables -> TMA bulk copies into a 2-stage ring).
// ------------------------------------------------------------------------------------------------
// STAGES / MINB: ring depth and CTAs per SM.  <3, 3> (72 KB, <= 72 registers) is the measured default; <2, 4> (54 KB,
// 56 registers, no spills) trades ring depth for 36 instead of 27 resident warps — compiled, selectable with
// NVNL_ROWS_CONFIG bit 0, not yet measured.  PAD (bit 1): 128-byte aligned, padded temporary rows.  UNI (bit 2): chunks
// of boundary cells that lie inside one image segment are swept in groups of four with a warp-uniform shift vector.
template <bool HALF, bool FMA, int STAGES, int MINB, bool PAD, bool UNI>
__global__ void __launch_bounds__(kRowsThreads, MINB) k_rows(const RowsArgs a) {
    using T = float;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int kStageBytes = kFastStageBytes;
    RowsSmem<STAGES>& sm = *reinterpret_cast<RowsSmem<STAGES>*>(smem_raw + (size_t)STAGES * kStageBytes);
    const uint32_t smem_base = smem_u32(smem_raw);
    constexpr uint32_t RS = sizeof(Rec<T>);
    constexpr int cap = kCandBytes / (int)RS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Ctrl* ctrl = reinterpret_cast<Ctrl*>(a.ws + a.L.ctrl);

    {
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
    // stale list slots are read (never stored) by the epilogue: they must hold valid tile indices from the start
    for (int k = tid; k < (int)(sizeof(sm.lists) / sizeof(unsigned)); k += kRowsThreads)
        reinterpret_cast<unsigned*>(&sm.lists[0][0][0])[k] = 0u;
    // unwrapped inputs are served by the two-pass kernels launched next to this one
    const bool active = ctrl->unwrapped == 0;
    if (tid == 0) {
        for (int st = 0; st < STAGES; ++st) {
            mbar_init(reinterpret_cast<uint64_t*>(&sm.full[st]), 1);
            mbar_init(reinterpret_cast<uint64_t*>(&sm.empty[st]), kRowsCons);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == 0) {
        // =========================== producer ===========================
        const SysParams* sys = reinterpret_cast<const SysParams*>(a.ws + a.L.sys);
        const int* cell_start = reinterpret_cast<const int*>(a.ws + a.L.cell_start);
        const Rec<T>* sorted = reinterpret_cast<const Rec<T>*>(a.ws + a.L.sorted);
        int2* deferred = reinterpret_cast<int2*>(a.ws + a.L.deferred);
        const int total_cells = active ? ctrl->total_cells : 0;
        int stage = 0;
        uint32_t ephase = 1;  // a fresh mbarrier passes a wait on the opposite parity: the ring starts empty
        int g_next = 0;
        if (lane == 0) g_next = atomicAdd(&ctrl->work_counter[0], 1);
        // Zero-fill of the caller's shifts buffer, fused into the sweep: the sweep is issue-bound and leaves HBM idle, the
        // zero-fill is pure HBM writes.  Every CTA owns one slice; its producer warp writes a quota of it per published
        // cell (paced over the kernel's lifetime) and the rest when the queue is drained.  A separate memset kernel does
        // not do: next to this kernel's 72 KB CTAs it only runs if the SM already has the large shared-memory carve-out,
        // otherwise the two serialise (profiles/r2_zero_overlap.txt).
        int4* const zbase = reinterpret_cast<int4*>(a.prezero);
        long long zpos = 0, zend = 0, zquota = 0;
        if (a.prezero) {
            const long long n16 = a.prezero_ints >> 2;
            const long long slice = (n16 + gridDim.x - 1) / gridDim.x;
            zpos = (long long)blockIdx.x * slice;
            zend = zpos + slice < n16 ? zpos + slice : n16;
            if (zpos > zend) zpos = zend;
            const long long cells_here = total_cells / (int)gridDim.x > 0 ? total_cells / (int)gridDim.x : 1;
            zquota = (zend - zpos) / cells_here + 32;
            if (blockIdx.x == 0 && lane < (int)(a.prezero_ints & 3)) a.prezero[(n16 << 2) + lane] = 0;
        }
        auto zero_some = [&](long long quota) {
            long long e = zpos + quota;
            e = e < zend ? e : zend;
            const int4 z4 = make_int4(0, 0, 0, 0);
            long long k = zpos + lane;
            for (; k + 96 < e; k += 128) {
                zbase[k] = z4; zbase[k + 32] = z4; zbase[k + 64] = z4; zbase[k + 96] = z4;
            }
            for (; k < e; k += 32) zbase[k] = z4;
            zpos = e;
        };
        for (;;) {
            // ---- 1. prepare the next cell entirely in registers (dependent global loads, image enumeration, shift
            //         sort) BEFORE waiting for a ring stage: after the consumers release a stage only the table
            //         writes and the TMA issue remain on the critical path ----
            bool have = false;
            int g = 0, home_start = 0, ntarget = 0, st = 0, cn = 0, key = kKeyEmpty, off = 0, total = 0, nseg = 1;
            int home_off = 0, si = 0;
            unsigned shiftmask = 0u;
            bool head = false;
            T Sx = (T)0, Sy = (T)0, Sz = (T)0;
            for (;;) {
                g = __shfl_sync(0xffffffffu, g_next, 0);
                if (g >= total_cells) break;
                if (lane == 0) g_next = atomicAdd(&ctrl->work_counter[0], 1);
                home_start = cell_start[g];
                ntarget = cell_start[g + 1] - home_start;
                if (ntarget == 0) continue;
                int s = 0;
                if (a.num_systems > 1) s = a.batch_idx[sorted[home_start].j];
                const SysParams& sp = sys[s];
                const int cpd0 = sp.cpd[0], cpd1 = sp.cpd[1], cpd2 = sp.cpd[2];
                const int R0 = sp.R[0], R1 = sp.R[1], R2 = sp.R[2];
                const int nx = 2 * R0 + 1, ny = 2 * R1 + 1, nzz = 2 * R2 + 1;
                const int nimg = nx * ny * nzz;
                bool ok = nimg <= 32 && ntarget <= kFastMaxTargets;
                int tag = 0;
                st = 0; cn = 0; key = kKeyEmpty;
                const int coff = sp.cell_offset;
                if (ok && lane < nimg) {
                    const int local = g - coff;
                    const int cx = local % cpd0, cy = (local / cpd0) % cpd1, cz = local / (cpd0 * cpd1);
                    const int dx = lane % nx - R0, dy = (lane / nx) % ny - R1, dz = lane / (nx * ny) - R2;
                    int tx = cx + dx, ty = cy + dy, tz = cz + dz;
                    bool in = true;
                    int csx = 0, csy = 0, csz = 0;
                    if (sp.pbc[0]) divmod_floor(tx, cpd0, csx, tx); else in = in && tx >= 0 && tx < cpd0;
                    if (sp.pbc[1]) divmod_floor(ty, cpd1, csy, ty); else in = in && ty >= 0 && ty < cpd1;
                    if (sp.pbc[2]) divmod_floor(tz, cpd2, csz, tz); else in = in && tz >= 0 && tz < cpd2;
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
                off = incl - cn;
                total = __shfl_sync(0xffffffffu, incl, 31);
                ok = ok && total <= cap;
                // segments = runs of equal shift among the non-empty images (sorted by shift above)
                nseg = 1;
                head = false;
                unsigned hm = 0u;
                if (shiftmask) {
                    const int pk = __shfl_up_sync(0xffffffffu, key, 1);
                    head = cn > 0 && (lane == 0 || pk != key);
                    hm = __ballot_sync(0xffffffffu, head);
                    nseg = __popc(hm);
                    ok = ok && nseg <= kRowsMaxSeg;  // a list entry has 3 bits for the segment (boxes < 3 cells wide: general kernel)
                }
                if (!ok) {
                    // leave the cell to the general kernel as work items of kDeferTargets target atoms
                    const int nitems = (ntarget + kDeferTargets - 1) / kDeferTargets;
                    int base = 0;
                    if (lane == 0) {
                        base = atomicAdd(&ctrl->n_deferred, nitems);
                        ctrl->had_deferred = 1;
                    }
                    base = __shfl_sync(0xffffffffu, base, 0);
                    for (int k = lane; k < nitems; k += 32) deferred[base + k] = make_int2(g, k * kDeferTargets);
                    continue;
                }
                if (shiftmask) {
                    if (head) {
                        si = __popc(hm & ((1u << lane) - 1u));
                        int csx, csy, csz;
                        unpack_key(key, csx, csy, csz);
                        T cm[9];
#pragma unroll
                        for (int k = 0; k < 9; ++k) cm[k] = (T)sp.cellm[k];
                        shift_vector<T, FMA>(cm, csx, csy, csz, Sx, Sy, Sz);
                    }
                }
                const unsigned tagm = __ballot_sync(0xffffffffu, tag != 0);
                const int home_lane = __ffs(tagm) - 1;
                home_off = __shfl_sync(0xffffffffu, off, home_lane);
                have = true;
                break;
            }
            if (zpos < zend) zero_some(have ? zquota : zend - zpos);
            // ---- 2. wait until the consumers have released this ring stage, then publish the tables and issue the copies ----
            mbar_wait_backoff(reinterpret_cast<uint64_t*>(&sm.empty[stage]), ephase);
            FastStage<T>& sg = sm.stage[stage];
            if (!have) {
                if (lane == 0) {
                    sg.item = -1;
                    mbar_arrive(reinterpret_cast<uint64_t*>(&sm.full[stage]));
                }
                break;
            }
            Rec<T>* cand = reinterpret_cast<Rec<T>*>(smem_raw + (size_t)stage * kStageBytes);
            if (shiftmask) {
                if (head) {
                    sg.seg_begin[si] = off;
                    sg.seg_key[si] = key;
                    sg.segS[3 * si] = Sx; sg.segS[3 * si + 1] = Sy; sg.segS[3 * si + 2] = Sz;
                }
                if (lane == 0) sg.seg_begin[nseg] = total;
            } else {
                if (lane == 0) { sg.seg_begin[0] = 0; sg.seg_begin[1] = total; sg.seg_key[0] = 0; }
                if (lane < 3) sg.segS[lane] = (T)0;
            }
            __syncwarp();
            const int nchunks = (total + 31) >> 5;
            bool uni_chunk = false;
            if (lane < nchunks) {
                int sgi = 0;
                while (sgi + 1 < nseg && (lane << 5) >= sg.seg_begin[sgi + 1]) ++sgi;
                sg.chunk_seg[lane] = sgi;
                if (UNI) {
                    const int last = (lane << 5) + 31;
                    uni_chunk = last < total && (sgi + 1 >= nseg || last < sg.seg_begin[sgi + 1]);
                }
            }
            if (UNI) {
                const unsigned umask = __ballot_sync(0xffffffffu, uni_chunk);
                if (lane == 0) sg.qrow[0] = (int)umask;  // (qrow is unused on this path)
            }
            if (lane == 0) {
                sg.nchunks = nchunks;
                sg.item = g; sg.ntarget = ntarget; sg.home_start = home_start; sg.home_off = home_off;
                sg.nseg = nseg; sg.total = total; sg.next_target = 0;
            }
            const uint32_t tx = (uint32_t)total * RS;
            __syncwarp();  // every lane's table writes precede lane 0's release-arrive below
            if (lane == 0) mbar_arrive_expect_tx(reinterpret_cast<uint64_t*>(&sm.full[stage]), tx);
            __syncwarp();
            if (cn > 0)
                tma_load_1d(cand + off, sorted + st, (uint32_t)cn * RS, reinterpret_cast<uint64_t*>(&sm.full[stage]));
            if (++stage == STAGES) { stage = 0; ephase ^= 1u; }
        }
        // the last CTA to drain the queue re-arms it for the next launch on this workspace
        if (lane == 0) {
            __threadfence();
            const int d = atomicAdd(&ctrl->done[0], 1);
            if (d == (int)gridDim.x - 1) {
                ctrl->work_counter[0] = 0;
                ctrl->done[0] = 0;
            }
        }
    } else {
        // =========================== consumers ===========================
        const int cw = warp - 1;
        int* rows = reinterpret_cast<int*>(a.ws + a.L.rows);
        int* row_ref = reinterpret_cast<int*>(a.ws + a.L.row_ref);
        const uint32_t lbaseA = smem_u32(&sm.lists[cw][0][0]) + (uint32_t)lane;
        const uint32_t lbaseB = smem_u32(&sm.lists[cw][1][0]) + (uint32_t)lane;
        RowsAlloc al;
        al.pos = al.end = 0;
        const float rc2 = a.cutoff_sq;
        int stage = 0;
        uint32_t fphase = 0;
        for (;;) {
            mbar_wait_backoff(reinterpret_cast<uint64_t*>(&sm.full[stage]), fphase);
            FastStage<T>& sg = sm.stage[stage];
            if (sg.item < 0) break;
            const uint32_t cand_addr = smem_base + (uint32_t)stage * kStageBytes;
            const int ntarget = sg.ntarget, home_off = sg.home_off;
            const bool shifted = sg.nseg > 1 || sg.seg_key[0] != 0;
            for (;;) {
                // two targets per trip: they share every candidate load and every FP instruction (f32x2)
                int t = 0;
                if (lane == 0) t = atomicAdd(&sg.next_target, 2);
                t = __shfl_sync(0xffffffffu, t, 0);
                if (t >= ntarget) break;
                const bool two = t + 1 < ntarget;
                const int selfA = home_off + t, selfB = two ? selfA + 1 : selfA;
                float xa, ya, za, xb, yb, zb;
                int iA, iB;
                lds_rec(cand_addr + (uint32_t)selfA * RS, xa, ya, za, iA);
                lds_rec(cand_addr + (uint32_t)selfB * RS, xb, yb, zb, iB);
                // x + (-0) == x bit for bit: one packed add gives each target pair a home in an aligned register pair
                const f32x2_t nz = pack2(-0.0f, -0.0f);
                const f32x2_t XI = add2(pack2(xa, xb), nz), YI = add2(pack2(ya, yb), nz), ZI = add2(pack2(za, zb), nz);
                uint32_t la = lbaseA, lb = lbaseB;
                rows_sweep2<HALF, FMA, UNI>(sg, cand_addr, XI, YI, ZI, iA, iB, rc2, lane, selfA, selfB, la, lb);
                rows_emit2<HALF, PAD>(a, sg, ctrl, cand_addr, lbaseA, lbaseB, la, lb, selfA, selfB, iA, iB, two, lane, shifted, al,
                                 rows, row_ref);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(reinterpret_cast<uint64_t*>(&sm.empty[stage]));
            if (++stage == STAGES) { stage = 0; fphase ^= 1u; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_rows_out: the final arrays in atom-index order.  One warp per 32 consecutive atoms; rows come from the temporary
// buffer (row_ref), everything written is sequential in the large: out_i (= i), out_j, shifts (zeros unless the row's
// cell touched a periodic boundary, then the packed image keys stored behind the row are unpacked).
// Atoms with row_ref < 0 were handled by the general kernel, which writes their rows itself.
// ------------------------------------------------------------------------------------------------
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

// SPEC (experiment): launched BEFORE the host knows the pair count, into buffers sized from a guess — the kernel reads
// the count itself, places row 1 of edge_index behind it, and does nothing if the guess was too small or the query was
// served by the two-pass kernels (the host then repeats the fill the regular way).
template <bool SPEC>
__global__ void __launch_bounds__(256, 5) k_rows_out(const unsigned char* __restrict__ ws, WsLayout L, long long n,
                                                     const int* __restrict__ neighbor_ptr, int* __restrict__ out_i,
                                                     int* __restrict__ out_j_arg, int* __restrict__ out_shifts,
                                                     int index_offset, int shifts_zeroed, long long spec_cap) {
    int* __restrict__ out_j = out_j_arg;
    if (SPEC) {
        const Ctrl* ctrl = reinterpret_cast<const Ctrl*>(ws + L.ctrl);
        const unsigned long long total = ctrl->total_pairs;
        if (ctrl->unwrapped != 0 || ctrl->rows_overflow != 0 || total > (unsigned long long)spec_cap) return;
        out_j = out_i + total;
    }
    const int* __restrict__ rows = reinterpret_cast<const int*>(ws + L.rows);
    const int* __restrict__ row_ref = reinterpret_cast<const int*>(ws + L.row_ref);
    const int lane = threadIdx.x & 31;
    const long long base = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
    if (base >= n) return;
    const long long il = base + lane;
    const int ref_l = il < n ? row_ref[il] : -1;
    const int p_l = neighbor_ptr[il < n ? il : n];
    const int pe_l = neighbor_ptr[il + 1 < n ? il + 1 : n];
    const int na = n - base < 32 ? (int)(n - base) : 32;
    // software pipeline: the row of atom t + 1 is in flight while atom t is written.  entries of a row start behind
    // its header (boundary rows: kRowsMaxSeg image keys)
    int ref = __shfl_sync(0xffffffffu, ref_l, 0);
    int p = __shfl_sync(0xffffffffu, p_l, 0);
    int cnt = __shfl_sync(0xffffffffu, pe_l, 0) - p;
    int v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int k = lane + 32 * u;
        v[u] = (ref >= 0 && k < cnt) ? rows[(ref >> 1) + ((ref & 1) ? kRowsMaxSeg : 0) + k] : 0;
    }
    for (int t = 0; t < na; ++t) {
        int nref = -1, np = 0, ncnt = 0;
        int nv[4] = {0, 0, 0, 0};
        if (t + 1 < na) {
            nref = __shfl_sync(0xffffffffu, ref_l, t + 1);
            np = __shfl_sync(0xffffffffu, p_l, t + 1);
            ncnt = __shfl_sync(0xffffffffu, pe_l, t + 1) - np;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = lane + 32 * u;
                nv[u] = (nref >= 0 && k < ncnt) ? rows[(nref >> 1) + ((nref & 1) ? kRowsMaxSeg : 0) + k] : 0;
            }
        }
        if (ref >= 0 && cnt > 0) {
            const int iv = (int)(base + t) + index_offset;
            int* __restrict__ oj = out_j + (size_t)p;
            int* sh = out_shifts + 3 * (size_t)p;
            int* __restrict__ oi = out_i + (size_t)p;
            if (!(ref & 1)) {
                const int* __restrict__ row = rows + (ref >> 1);
                int* __restrict__ ojl = oj + lane;   // per-lane bases: the four stores below differ by immediates only
                int* __restrict__ oil = oi + lane;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (lane + 32 * u < cnt) {
                        ojl[32 * u] = v[u] + index_offset;
                        oil[32 * u] = iv;
                    }
                }
                for (int k = 128 + lane; k < cnt; k += 32) {  // rows longer than 128 (rare)
                    oj[k] = row[k] + index_offset;
                    oi[k] = iv;
                }
                if (!shifts_zeroed) warp_fill(sh, 3 * cnt, 0, lane);
            } else {
                const int* __restrict__ hdr = rows + (ref >> 1);
                const int* __restrict__ row = hdr + kRowsMaxSeg;
                const int keyl = lane < kRowsMaxSeg ? hdr[lane] : 0;
                if (!shifts_zeroed) {
                    warp_fill(sh, 3 * cnt, 0, lane);
                    __syncwarp();  // the zeros of other lanes precede the image shifts written below
                }
                auto emit = [&](int e, int k) {
                    const int key = __shfl_sync(0xffffffffu, keyl, ((unsigned)e >> 28) & (kRowsMaxSeg - 1));
                    if (k < cnt) {
                        oj[k] = (e & 0x0fffffff) + index_offset;
                        oi[k] = iv;
                        if (key != 0) {
                            int csx, csy, csz;
                            unpack_key(key, csx, csy, csz);
                            sh[3 * k] = csx;
                            sh[3 * k + 1] = csy;
                            sh[3 * k + 2] = csz;
                        }
                    }
                };
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (32 * u < cnt) emit(v[u], 32 * u + lane);
                for (int k0 = 128; k0 < cnt; k0 += 32) emit(k0 + lane < cnt ? row[k0 + lane] : 0, k0 + lane);
            }
        }
        ref = nref; p = np; cnt = ncnt;
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = nv[u];
    }
}

// resets the per-query state of the control block and marks every atom "no temporary row"
__global__ void k_query_reset(unsigned char* __restrict__ ws, WsLayout L, long long n, int with_rows) {
    Ctrl* ctrl = reinterpret_cast<Ctrl*>(ws + L.ctrl);
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid == 0) {
        ctrl->n_deferred = 0;
        ctrl->rows_cursor = 0ull;
        ctrl->rows_overflow = 0;
    }
    if (with_rows) {
        int* row_ref = reinterpret_cast<int*>(ws + L.row_ref);
        for (long long i = gid; i < n; i += (long long)gridDim.x * blockDim.x) row_ref[i] = -1;
    }
}

}  // namespace nvnl
