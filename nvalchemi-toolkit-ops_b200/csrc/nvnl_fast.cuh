// nvnl_fast.cuh — the lean sweep for the common case (atoms inside the primary image, <= 32 cell
// images per stencil, <= one shared-memory tile of candidates).  Everything else is appended to a
// deferred-cell list and handled by the general kernel in nvnl_sweep.cuh.
//
// Why a second kernel: ncu on the first implementation (profiles/r1_baseline_*.txt) showed both
// sweeps issue-bound at 588 / 1351 warp-instructions per atom.  This version
//   * does all per-cell work (image enumeration, shift sort, segment table, TMA issue) in warp 0 only;
//   * addresses shared memory with 32-bit shared-space pointers and immediate offsets;
//   * specialises the inner loop for zero-shift segments (interior cells: 3 FADD + FMUL + 2 FFMA +
//     FSETP + VOTE per 32 candidates);
//   * computes every distance ONCE: the count pass stores one 32-bit hit mask per (atom, 32-candidate
//     chunk) — 128 B per atom — and the COO fill pass only expands masks (popc prefix) into rows.
//   * the matrix pass fuses mask computation, row write and padding (no global masks).
#pragma once
#include "nvnl_sweep.cuh"

namespace nvnl {

constexpr int kFastCons = 6;                          // consumer warps per CTA
constexpr int kFastThreads = (kFastCons + 1) * 32;    // + one producer warp (warp 0)
constexpr int kFastStages = 2;                        // TMA ring depth
constexpr int kFastMaxTargets = 64;                   // target atoms per cell the fast kernel takes (else: general kernel)
constexpr int kFastSlackBytes = 1024;                 // the tail chunk of the last segment may read past the staged data
constexpr int kFastStageBytes = kCandBytes + kFastSlackBytes;
// unwrapped variant: every candidate's periodic image (int4) is staged behind the records of the stage
constexpr int kFastStageBytesU = 2 * (kCandBytes + kFastSlackBytes);

enum FastMode { FAST_COUNT = 0, FAST_FILL_COO = 1, FAST_MATRIX = 2 };

// Per-stage tables written by the producer warp, read by the consumer warps.
template <typename T>
struct FastStage {
    T segS[32 * 3];
    int seg_begin[33], seg_key[32];
    int chunk_seg[32];  // segment of the first candidate of each dense 32-candidate chunk
    int item, ntarget, home_off, home_start, nseg, total, nchunks, next_target;
    int qrow[kFastMaxTargets];  // FILL: neighbor_ptr (row start) of every target
    T cm[9];                    // unwrapped variant: cell matrix and periodicity of the cell's system
    int pbc[3];
};

template <typename T, int MODE>
struct FastSmem {
    FastStage<T> stage[kFastStages];
    int e_st[32], e_cn[32], e_key[32], e_tag[32];                // producer scratch (shift sort)
    alignas(16) unsigned maskbuf[kFastCons][2][32];              // per consumer warp: hit masks of the current row(s)
    int pre[kFastCons][32];                                      // per consumer warp: inclusive popc prefix per chunk
    // FILL only: the cell's hit masks (TMA from global); other modes keep a token array so 5 CTAs fit per SM
    alignas(16) unsigned smasks[kFastStages][MODE == 1 ? kFastMaxTargets * 32 : 4];
    unsigned long long full[kFastStages], empty[kFastStages];    // mbarriers of the ring
};

template <typename T, int MODE, bool UNW>
constexpr size_t fast_smem_bytes() {
    return (size_t)kFastStages * (UNW ? kFastStageBytesU : kFastStageBytes) + sizeof(FastSmem<T, MODE>);
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- shared-memory accessors on 32-bit shared-space addresses -----------------------------------
__device__ __forceinline__ void lds_rec(uint32_t addr, float& x, float& y, float& z, int& j) {
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=r"(j) : "r"(addr));
}
__device__ __forceinline__ void lds_rec(uint32_t addr, double& x, double& y, double& z, int& j) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
    asm volatile("ld.shared.f64 %0, [%1+16];" : "=d"(z) : "r"(addr));
    asm volatile("ld.shared.b32 %0, [%1+24];" : "=r"(j) : "r"(addr));
}
template <typename T>
__device__ __forceinline__ int lds_rec_j(uint32_t addr) {
    int j;
    if (sizeof(T) == 4)
        asm volatile("ld.shared.b32 %0, [%1+12];" : "=r"(j) : "r"(addr));
    else
        asm volatile("ld.shared.b32 %0, [%1+24];" : "=r"(j) : "r"(addr));
    return j;
}

__device__ __forceinline__ void sts_v4(uint32_t addr, unsigned a, unsigned b, unsigned c, unsigned d) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// one 32-candidate chunk: returns the hit ballot
template <typename T, bool HALF, bool FMA, bool SHIFTED, bool TAIL>
__device__ __forceinline__ unsigned chunk_mask(uint32_t addr, T xi, T yi, T zi, int i, T Sx, T Sy, T Sz, T rc2,
                                               bool seg_lexpos, bool valid) {
    using A = Arith<T>;
    T x, y, z;
    int j;
    lds_rec(addr, x, y, z, j);
    T dx = A::sub(x, xi), dy = A::sub(y, yi), dz = A::sub(z, zi);
    if (SHIFTED) {
        dx = A::add(dx, Sx);
        dy = A::add(dy, Sy);
        dz = A::add(dz, Sz);
    }
    const T d2 = dist2<T, FMA>(dx, dy, dz);
    bool hit = d2 < rc2;
    if (TAIL) hit = hit && valid;
    if (HALF) hit = hit && (i < j || (i == j && seg_lexpos));
    return __ballot_sync(0xffffffffu, hit);
}

// Phase 1: hit masks of one target atom against the staged stencil; chunk ck's ballot -> mb[ck].
// Chunks are DENSE over the concatenated tile (candidate c lives in chunk c >> 5, bit c & 31), so a tile of
// <= 1024 candidates never needs more than 32 mask words.  Chunks that lie entirely inside the leading zero-shift
// segment take the lean path; every other chunk picks its shift vector per lane (a chunk may straddle segments).
template <typename T, bool HALF, bool FMA>
__device__ __forceinline__ void fast_masks(const FastStage<T>& sm, uint32_t cand_addr, T xi, T yi, T zi, int i, T rc2,
                                           int lane, unsigned* __restrict__ mb) {
    constexpr uint32_t RS = sizeof(Rec<T>);
    using A = Arith<T>;
    const bool l0 = lane == 0;
    const int total = sm.total;
    const int nchunks = sm.nchunks;
    const int zend = sm.seg_key[0] == 0 ? sm.seg_begin[1] : 0;  // candidates [0, zend) have zero shift
    const int nzfull = zend >> 5;                                // chunks entirely inside the zero-shift segment
    uint32_t addr = cand_addr + (uint32_t)lane * RS;
    int ck = 0;
#pragma unroll 1
    for (; ck + 4 <= nzfull; ck += 4) {
        const unsigned m0 = chunk_mask<T, HALF, FMA, false, false>(addr, xi, yi, zi, i, 0, 0, 0, rc2, false, true);
        const unsigned m1 = chunk_mask<T, HALF, FMA, false, false>(addr + 32 * RS, xi, yi, zi, i, 0, 0, 0, rc2, false, true);
        const unsigned m2 = chunk_mask<T, HALF, FMA, false, false>(addr + 64 * RS, xi, yi, zi, i, 0, 0, 0, rc2, false, true);
        const unsigned m3 = chunk_mask<T, HALF, FMA, false, false>(addr + 96 * RS, xi, yi, zi, i, 0, 0, 0, rc2, false, true);
        if (l0) sts_v4(smem_u32(mb + ck), m0, m1, m2, m3);
        addr += 128 * RS;
    }
#pragma unroll 1
    for (; ck < nzfull; ++ck) {
        const unsigned m0 = chunk_mask<T, HALF, FMA, false, false>(addr, xi, yi, zi, i, 0, 0, 0, rc2, false, true);
        if (l0) mb[ck] = m0;
        addr += 32 * RS;
    }
    if (sm.nseg == 1 && ck < nchunks) {
        // interior cell: the only other chunk is the partial tail of the zero-shift segment
        const unsigned m0 = chunk_mask<T, HALF, FMA, false, true>(addr, xi, yi, zi, i, 0, 0, 0, rc2, false,
                                                                  (ck << 5) + lane < total);
        if (l0) mb[ck] = m0;
        ++ck;
    }
    // remaining chunks: shifted segments, segment boundaries, the tail
#pragma unroll 1
    for (; ck < nchunks; ++ck) {
        const int c = (ck << 5) + lane;
        int sg = sm.chunk_seg[ck];
        while (sg + 1 < sm.nseg && c >= sm.seg_begin[sg + 1]) ++sg;  // per lane; usually 0 or 1 step
        const T Sx = sm.segS[3 * sg], Sy = sm.segS[3 * sg + 1], Sz = sm.segS[3 * sg + 2];
        T x, y, z;
        int j;
        lds_rec(addr, x, y, z, j);
        // (r_j - r_i) + 0 == r_j - r_i bit for bit, so zero-shift lanes can share the shifted arithmetic
        const T dx = A::add(A::sub(x, xi), Sx), dy = A::add(A::sub(y, yi), Sy), dz = A::add(A::sub(z, zi), Sz);
        const T d2 = dist2<T, FMA>(dx, dy, dz);
        bool hit = (d2 < rc2) && (c < total);
        if (HALF) {
            int csx, csy, csz;
            unpack_key(sm.seg_key[sg], csx, csy, csz);
            const bool lexpos = csx > 0 || (csx == 0 && (csy > 0 || (csy == 0 && csz > 0)));
            hit = hit && (i < j || (i == j && lexpos));
        }
        const unsigned m0 = __ballot_sync(0xffffffffu, hit);
        if (l0) mb[ck] = m0;
        addr += 32 * RS;
    }
}

// ---- packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2 take a scalar broadcast operand) -------------------
// Two target atoms A, B share one candidate load; every step of the distance test is one packed instruction for
// both.  Each half is an IEEE round-to-nearest fp32 operation, so the results equal the scalar path bit for bit.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float a, float b) {
    f32x2_t r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2_t sub2(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// one 32-candidate chunk against two targets: hit ballots mA, mB
template <bool HALF, bool FMA, bool SHIFTED, bool TAIL>
__device__ __forceinline__ void chunk_mask2(uint32_t addr, f32x2_t XI, f32x2_t YI, f32x2_t ZI, int iA, int iB, float Sx,
                                            float Sy, float Sz, float rc2, bool lexpos, bool valid, unsigned& mA,
                                            unsigned& mB) {
    float x, y, z;
    int j;
    lds_rec(addr, x, y, z, j);
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
    float dA, dB;
    unpack2(d2, dA, dB);
    bool hA = dA < rc2, hB = dB < rc2;
    if (TAIL) { hA = hA && valid; hB = hB && valid; }
    if (HALF) {
        hA = hA && (iA < j || (iA == j && lexpos));
        hB = hB && (iB < j || (iB == j && lexpos));
    }
    mA = __ballot_sync(0xffffffffu, hA);
    mB = __ballot_sync(0xffffffffu, hB);
}

// Phase 1 for two targets at once (fp32): same chunk structure as fast_masks.
template <bool HALF, bool FMA>
__device__ __forceinline__ void fast_masks2(const FastStage<float>& sm, uint32_t cand_addr, float xa, float ya, float za, int iA,
                                            float xb, float yb, float zb, int iB, float rc2, int lane,
                                            unsigned* __restrict__ mbA, unsigned* __restrict__ mbB) {
    constexpr uint32_t RS = sizeof(Rec<float>);
    const bool l0 = lane == 0;
    const int total = sm.total, nchunks = sm.nchunks;
    const int zend = sm.seg_key[0] == 0 ? sm.seg_begin[1] : 0;
    const int nzfull = zend >> 5;
    const f32x2_t XI = pack2(xa, xb), YI = pack2(ya, yb), ZI = pack2(za, zb);
    uint32_t addr = cand_addr + (uint32_t)lane * RS;
    int ck = 0;
#pragma unroll 1
    for (; ck + 4 <= nzfull; ck += 4) {
        unsigned a0, b0, a1, b1, a2, b2, a3, b3;
        chunk_mask2<HALF, FMA, false, false>(addr, XI, YI, ZI, iA, iB, 0, 0, 0, rc2, false, true, a0, b0);
        chunk_mask2<HALF, FMA, false, false>(addr + 32 * RS, XI, YI, ZI, iA, iB, 0, 0, 0, rc2, false, true, a1, b1);
        chunk_mask2<HALF, FMA, false, false>(addr + 64 * RS, XI, YI, ZI, iA, iB, 0, 0, 0, rc2, false, true, a2, b2);
        chunk_mask2<HALF, FMA, false, false>(addr + 96 * RS, XI, YI, ZI, iA, iB, 0, 0, 0, rc2, false, true, a3, b3);
        if (l0) {
            sts_v4(smem_u32(mbA + ck), a0, a1, a2, a3);
            sts_v4(smem_u32(mbB + ck), b0, b1, b2, b3);
        }
        addr += 128 * RS;
    }
#pragma unroll 1
    for (; ck < nzfull; ++ck) {
        unsigned a0, b0;
        chunk_mask2<HALF, FMA, false, false>(addr, XI, YI, ZI, iA, iB, 0, 0, 0, rc2, false, true, a0, b0);
        if (l0) { mbA[ck] = a0; mbB[ck] = b0; }
        addr += 32 * RS;
    }
    if (sm.nseg == 1 && ck < nchunks) {
        unsigned a0, b0;
        chunk_mask2<HALF, FMA, false, true>(addr, XI, YI, ZI, iA, iB, 0, 0, 0, rc2, false, (ck << 5) + lane < total, a0, b0);
        if (l0) { mbA[ck] = a0; mbB[ck] = b0; }
        ++ck;
    }
#pragma unroll 1
    for (; ck < nchunks; ++ck) {
        const int c = (ck << 5) + lane;
        int sg = sm.chunk_seg[ck];
        while (sg + 1 < sm.nseg && c >= sm.seg_begin[sg + 1]) ++sg;
        bool lexpos = false;
        if (HALF) {
            int csx, csy, csz;
            unpack_key(sm.seg_key[sg], csx, csy, csz);
            lexpos = csx > 0 || (csx == 0 && (csy > 0 || (csy == 0 && csz > 0)));
        }
        unsigned a0, b0;
        chunk_mask2<HALF, FMA, true, true>(addr, XI, YI, ZI, iA, iB, sm.segS[3 * sg], sm.segS[3 * sg + 1], sm.segS[3 * sg + 2],
                                           rc2, lexpos, c < total, a0, b0);
        if (l0) { mbA[ck] = a0; mbB[ck] = b0; }
        addr += 32 * RS;
    }
}
// double precision has no packed form: two scalar sweeps
template <bool HALF, bool FMA>
__device__ __forceinline__ void fast_masks2(const FastStage<double>& sm, uint32_t cand_addr, double xa, double ya, double za,
                                            int iA, double xb, double yb, double zb, int iB, double rc2, int lane,
                                            unsigned* __restrict__ mbA, unsigned* __restrict__ mbB) {
    fast_masks<double, HALF, FMA>(sm, cand_addr, xa, ya, za, iA, rc2, lane, mbA);
    fast_masks<double, HALF, FMA>(sm, cand_addr, xb, yb, zb, iB, rc2, lane, mbB);
}

__device__ __forceinline__ int4 lds_int4(uint32_t addr) {
    int4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// Phase 1 for inputs with atoms outside the primary periodic image: the integer shift of a pair is
// cs(segment) + a_i - a_j (cell_list.py:506-523), so it — and its lattice vector — is evaluated per lane.
template <typename T, bool HALF, bool FMA>
__device__ __forceinline__ void fast_masks_unw(const FastStage<T>& sm, uint32_t cand_addr, uint32_t ash_addr, T xi, T yi,
                                               T zi, int i, int4 ai, const T* __restrict__ cm, T rc2, int lane,
                                               unsigned* __restrict__ mb) {
    constexpr uint32_t RS = sizeof(Rec<T>);
    using A = Arith<T>;
    const int total = sm.total, nchunks = sm.nchunks, nseg = sm.nseg;
    const int p0 = sm.pbc[0], p1 = sm.pbc[1], p2 = sm.pbc[2];
#pragma unroll 1
    for (int ck = 0; ck < nchunks; ++ck) {
        const int c = (ck << 5) + lane;
        int sg = sm.chunk_seg[ck];
        while (sg + 1 < nseg && c >= sm.seg_begin[sg + 1]) ++sg;
        int csx, csy, csz;
        unpack_key(sm.seg_key[sg], csx, csy, csz);
        T x, y, z;
        int j;
        lds_rec(cand_addr + (uint32_t)c * RS, x, y, z, j);
        const int4 aj = lds_int4(ash_addr + (uint32_t)c * 16u);
        const int sx = p0 ? csx + ai.x - aj.x : 0, sy = p1 ? csy + ai.y - aj.y : 0, sz = p2 ? csz + ai.z - aj.z : 0;
        T Sx, Sy, Sz;
        shift_vector<T, FMA>(cm, sx, sy, sz, Sx, Sy, Sz);
        const T dx = A::add(A::sub(x, xi), Sx), dy = A::add(A::sub(y, yi), Sy), dz = A::add(A::sub(z, zi), Sz);
        const T d2 = dist2<T, FMA>(dx, dy, dz);
        bool hit = (d2 < rc2) && (c < total);
        if (HALF) {
            const bool lexpos = sx > 0 || (sx == 0 && (sy > 0 || (sy == 0 && sz > 0)));
            hit = hit && (i < j || (i == j && lexpos));
        }
        const unsigned m0 = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) mb[ck] = m0;
    }
}

// position of the r-th (0-based) set bit of m (r < popc(m)): popc halving, ~33 SASS instructions
__device__ __forceinline__ int nth_set_bit(unsigned m, int r) {
    int bit = 0, t;
    t = __popc(m & 0xFFFFu); if (r >= t) { r -= t; bit = 16; m >>= 16; }
    t = __popc(m & 0xFFu);   if (r >= t) { r -= t; bit += 8; m >>= 8; }
    t = __popc(m & 0xFu);    if (r >= t) { r -= t; bit += 4; m >>= 4; }
    t = __popc(m & 0x3u);    if (r >= t) { r -= t; bit += 2; m >>= 2; }
    t = (int)(m & 1u);       if (r >= t) { bit += 1; }
    return bit;
}

// Phase 2 (balanced): every lane produces its own output entries k = lane, lane+32, ...:
//   chunk of entry k  = first chunk whose inclusive popc prefix exceeds k (5-step search in shared memory),
//   bit inside it     = nth_set_bit(mask[chunk], k - prefix[chunk-1]).
// ~90 hits / 32 lanes = 3 iterations per atom instead of a serial walk bounded by the densest chunk (~22 bits).
template <typename T, bool COO, bool UNW>
__device__ __forceinline__ int fast_expand2(const SweepArgs<T>& a, const FastStage<T>& sm, uint32_t cand_addr,
                                            unsigned mymask, int lane, int i, size_t p0, int limit, int* __restrict__ out_j,
                                            int* __restrict__ out_sh, unsigned* __restrict__ mb, int* __restrict__ pre,
                                            uint32_t ash_addr, int4 ai) {
    constexpr uint32_t RS = sizeof(Rec<T>);
    const int pc = __popc(mymask);
    const int incl = warp_incl_scan(pc, lane);
    const int cnt = __shfl_sync(0xffffffffu, incl, 31);
    const int nstore = cnt < limit ? cnt : limit;
    mb[lane] = mymask;
    pre[lane] = incl;
    // hits of the leading zero-shift segment (candidates [0, zend)) occupy the first nzero row slots
    int nzero = 0;
    if (!UNW) {
        const int zend = sm.seg_key[0] == 0 ? sm.seg_begin[1] : 0;
        const int zc = zend >> 5, zb = zend & 31;
        const int before = zc > 0 ? __shfl_sync(0xffffffffu, incl, (zc - 1) & 31) : 0;
        const unsigned mz = __shfl_sync(0xffffffffu, mymask, zc & 31);
        nzero = before + (zb ? __popc(mz & ((1u << zb) - 1u)) : 0);
    }
    __syncwarp();
    const int off_idx = COO ? a.index_offset : 0;
    const int iv = i + off_idx;
    int* sh = out_sh + 3 * p0;
    // four independent entries per lane per trip: the search / nth-bit chains interleave (ILP at low occupancy)
    for (int k0 = lane; k0 < nstore; k0 += 128) {
        int ckv[4], cv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = k0 + 32 * u;
            const int kk = k < nstore ? k : 0;
            int ck = (pre[15] <= kk) ? 16 : 0;
            ck += (pre[ck + 7] <= kk) ? 8 : 0;
            ck += (pre[ck + 3] <= kk) ? 4 : 0;
            ck += (pre[ck + 1] <= kk) ? 2 : 0;
            ck += (pre[ck] <= kk) ? 1 : 0;
            const int r = kk - (ck ? pre[ck - 1] : 0);
            const int bit = nth_set_bit(mb[ck], r);
            ckv[u] = ck;
            cv[u] = (ck << 5) + bit;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = k0 + 32 * u;
            if (k < nstore) {
                const int j = lds_rec_j<T>(cand_addr + (uint32_t)cv[u] * RS);
                out_j[p0 + k] = j + off_idx;
                if (COO) a.out_i[p0 + k] = iv;
                if (k >= nzero) {
                    int sg = sm.chunk_seg[ckv[u]];
                    while (sg + 1 < sm.nseg && cv[u] >= sm.seg_begin[sg + 1]) ++sg;
                    int csx, csy, csz;
                    unpack_key(sm.seg_key[sg], csx, csy, csz);
                    if (UNW) {
                        const int4 aj = lds_int4(ash_addr + (uint32_t)cv[u] * 16u);
                        csx = sm.pbc[0] ? csx + ai.x - aj.x : 0;
                        csy = sm.pbc[1] ? csy + ai.y - aj.y : 0;
                        csz = sm.pbc[2] ? csz + ai.z - aj.z : 0;
                    }
                    sh[3 * k] = csx;
                    sh[3 * k + 1] = csy;
                    sh[3 * k + 2] = csz;
                }
            }
        }
    }
    const int nz = nzero < nstore ? nzero : nstore;
    for (int e = lane; e < 3 * nz; e += 32) sh[e] = 0;
    __syncwarp();
    return cnt;
}

// neighbor_ptr in cell-sorted atom order: ptr_sorted[k] = neighbor_ptr[sorted[k].j].  Lets the FILL producer read the
// row pointers of a cell's targets as one coalesced load instead of a dependent gather on its critical path.
template <typename T>
__global__ void k_gather_ptr(const unsigned char* __restrict__ ws, WsLayout L, long long n,
                             const int* __restrict__ neighbor_ptr, int* __restrict__ ptr_sorted) {
    pdl_enter();
    const Rec<T>* sorted = reinterpret_cast<const Rec<T>*>(ws + L.sorted);
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) ptr_sorted[k] = neighbor_ptr[sorted[k].j];
}

// ------------------------------------------------------------------------------------------------
// k_fast: warp-specialised persistent kernel.
//   warp 0 (producer): pulls target cells from the device queue, enumerates the stencil images, sorts them by
//     shift, builds the segment/chunk tables of a ring stage and issues the TMA bulk copies that concatenate
//     the stencil's runs in that stage's shared-memory buffer (full[stage] mbarrier, expect_tx).
//   warps 1..kFastCons (consumers): wait on full[stage], claim target atoms of the cell one at a time
//     (shared-memory counter), sweep / expand, and release the stage through empty[stage].
// No CTA-wide barrier in the steady state: the setup latency of cell k+1 hides behind the sweep of cell k.
// ------------------------------------------------------------------------------------------------
template <typename T, int MODE, bool HALF, bool FMA, bool UNW>
__global__ void __launch_bounds__(kFastThreads, UNW ? (MODE == 1 ? 2 : 3) : (MODE == 1 ? 4 : 5))
k_fast(const SweepArgs<T> a) {
    pdl_enter();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int kStageBytes = UNW ? kFastStageBytesU : kFastStageBytes;
    constexpr int kAshOffset = kCandBytes + kFastSlackBytes;  // unwrapped variant: periodic images behind the records
    FastSmem<T, MODE>& sm = *reinterpret_cast<FastSmem<T, MODE>*>(smem_raw + (size_t)kFastStages * kStageBytes);
    const uint32_t smem_base = smem_u32(smem_raw);
    constexpr uint32_t RS = sizeof(Rec<T>);
    constexpr int cap = kCandBytes / (int)RS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Ctrl* ctrl = reinterpret_cast<Ctrl*>(a.ws + a.L.ctrl);

    if (MODE == FAST_COUNT) {
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
    // two variants of this kernel are launched back to back; the one matching the input (all atoms inside the primary
    // periodic image, or not) does the work, the other retires at once
    const bool active = (ctrl->unwrapped != 0) == UNW;
    if (tid == 0) {
        for (int st = 0; st < kFastStages; ++st) {
            mbar_init(reinterpret_cast<uint64_t*>(&sm.full[st]), 1);
            mbar_init(reinterpret_cast<uint64_t*>(&sm.empty[st]), kFastCons);
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
        const int4* sorted_ashift = reinterpret_cast<const int4*>(a.ws + a.L.sorted_ashift);
        const int* ptr_sorted = reinterpret_cast<const int*>(a.ws + a.L.ptr_sorted);
        int stage = 0;
        uint32_t ephase = 1;  // a fresh mbarrier passes a wait on the opposite parity: the ring starts empty
        // the queue index of the NEXT cell is always in flight (one hop off the critical path)
        int g_next = 0;
        if (lane == 0) g_next = atomicAdd(&ctrl->work_counter[a.queue], 1);
        for (;;) {
            mbar_wait(reinterpret_cast<uint64_t*>(&sm.empty[stage]), ephase);
            FastStage<T>& sg = sm.stage[stage];
            Rec<T>* cand = reinterpret_cast<Rec<T>*>(smem_raw + (size_t)stage * kStageBytes);
            int4* cand_ash = reinterpret_cast<int4*>(smem_raw + (size_t)stage * kStageBytes + kAshOffset);
            bool done = false;
            for (;;) {
                const int g = __shfl_sync(0xffffffffu, g_next, 0);
                if (g >= total_cells) {
                    done = true;
                    break;
                }
                if (lane == 0) g_next = atomicAdd(&ctrl->work_counter[a.queue], 1);
                // hop 1: the cell's run (count = start[g+1] - start[g]; cell_start has one entry past the last cell)
                const int home_start = cell_start[g];
                const int ntarget = cell_start[g + 1] - home_start;
                if (ntarget == 0) continue;
                int s = 0;
                if (a.num_systems > 1) {
                    s = a.batch_idx[sorted[home_start].j];
                    s = s < 0 ? 0 : (s >= a.num_systems ? a.num_systems - 1 : s);   // (k_hash reported the error)
                }
                const SysParams& sp = sys[s];
                const int cpd0 = sp.cpd[0], cpd1 = sp.cpd[1], cpd2 = sp.cpd[2];
                const int R0 = sp.R[0], R1 = sp.R[1], R2 = sp.R[2];
                const int nx = 2 * R0 + 1, ny = 2 * R1 + 1, nzz = 2 * R2 + 1;
                const int nimg = nx * ny * nzz;
                bool ok = nimg <= 32 && ntarget <= kFastMaxTargets;
                int st = 0, cn = 0, key = kKeyEmpty, tag = 0;
                const int coff = sp.cell_offset;
                // hop 2 (all independent): the stencil images' runs and, for FILL, the targets' row pointers
                int q0 = 0, q1 = 0;
                if (MODE == FAST_FILL_COO && ok) {
                    if (lane < ntarget) q0 = ptr_sorted[home_start + lane];
                    if (lane + 32 < ntarget) q1 = ptr_sorted[home_start + lane + 32];
                }
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
                const unsigned shiftmask = __ballot_sync(0xffffffffu, key != 0 && key != kKeyEmpty);
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
                const int off = incl - cn;
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                ok = ok && total <= cap;
                int nseg = 1;
                if (ok) {
                    if (shiftmask) {
                        const int pk = __shfl_up_sync(0xffffffffu, key, 1);
                        const bool head = cn > 0 && (lane == 0 || pk != key);
                        const unsigned hm = __ballot_sync(0xffffffffu, head);
                        nseg = __popc(hm);
                        if (head) {
                            const int si = __popc(hm & ((1u << lane) - 1u));
                            sg.seg_begin[si] = off;
                            sg.seg_key[si] = key;
                            int csx, csy, csz;
                            unpack_key(key, csx, csy, csz);
                            T cm[9];
#pragma unroll
                            for (int k = 0; k < 9; ++k) cm[k] = (T)sp.cellm[k];
                            T Sx, Sy, Sz;
                            shift_vector<T, FMA>(cm, csx, csy, csz, Sx, Sy, Sz);
                            sg.segS[3 * si] = Sx; sg.segS[3 * si + 1] = Sy; sg.segS[3 * si + 2] = Sz;
                        }
                        if (lane == 0) sg.seg_begin[nseg] = total;
                    } else if (lane == 0) {
                        sg.seg_begin[0] = 0; sg.seg_begin[1] = total; sg.seg_key[0] = 0;
                    }
                    __syncwarp();
                    // dense chunk table: segment of the first candidate of chunk `lane`
                    const int nchunks = (total + 31) >> 5;
                    if (lane < nchunks) {
                        int sgi = 0;
                        while (sgi + 1 < nseg && (lane << 5) >= sg.seg_begin[sgi + 1]) ++sgi;
                        sg.chunk_seg[lane] = sgi;
                    }
                    if (!shiftmask && lane < 3) sg.segS[lane] = (T)0;  // single zero-shift segment
                    if (lane == 0) sg.nchunks = nchunks;
                }
                if (!ok) {
                    // too many images / candidates / targets for one tile: leave the cell to the general kernel, as
                    // work items of kDeferTargets target atoms each (a whole small periodic system can be one "cell")
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
                const unsigned tagm = __ballot_sync(0xffffffffu, tag != 0);
                const int home_lane = __ffs(tagm) - 1;
                const int home_off = __shfl_sync(0xffffffffu, off, home_lane);
                if (lane == 0) {
                    sg.item = g; sg.ntarget = ntarget; sg.home_start = home_start; sg.home_off = home_off;
                    sg.nseg = nseg; sg.total = total; sg.next_target = 0;
                }
                uint32_t tx = (uint32_t)total * RS;
                if (UNW) {
                    tx += (uint32_t)total * 16u;
                    if (lane < 9) sg.cm[lane] = (T)sp.cellm[lane];
                    if (lane < 3) sg.pbc[lane] = sp.pbc[lane];
                }
                if (MODE == FAST_FILL_COO) {
                    // the consumers of a FILL stage touch no global memory but their stores: the producer brings in the
                    // targets' row pointers (pre-gathered into sorted order) and TMA-copies the cell's mask block
                    if (lane < ntarget) sg.qrow[lane] = q0;
                    if (lane + 32 < ntarget) sg.qrow[lane + 32] = q1;
                    tx += (uint32_t)ntarget * 128u;
                }
                __syncwarp();  // every lane's table writes precede lane 0's release-arrive below
                if (lane == 0) {
                    mbar_arrive_expect_tx(reinterpret_cast<uint64_t*>(&sm.full[stage]), tx);
                    if (MODE == FAST_FILL_COO)
                        tma_load_1d(sm.smasks[stage], reinterpret_cast<const unsigned*>(a.ws + a.L.masks) + (size_t)home_start * 32,
                                    (uint32_t)ntarget * 128u, reinterpret_cast<uint64_t*>(&sm.full[stage]));
                }
                __syncwarp();
                if (cn > 0) {
                    tma_load_1d(cand + off, sorted + st, (uint32_t)cn * RS, reinterpret_cast<uint64_t*>(&sm.full[stage]));
                    if (UNW)
                        tma_load_1d(cand_ash + off, sorted_ashift + st, (uint32_t)cn * 16u,
                                    reinterpret_cast<uint64_t*>(&sm.full[stage]));
                }
                break;
            }
            if (done) {
                if (lane == 0) {
                    sg.item = -1;
                    mbar_arrive(reinterpret_cast<uint64_t*>(&sm.full[stage]));
                }
                break;
            }
            if (++stage == kFastStages) { stage = 0; ephase ^= 1u; }
        }
        // the last CTA to drain the queue re-arms it for the next launch on this workspace
        if (lane == 0) {
            __threadfence();
            const int d = atomicAdd(&ctrl->done[a.queue], 1);
            if (d == (int)gridDim.x - 1) {
                ctrl->work_counter[a.queue] = 0;
                ctrl->done[a.queue] = 0;
            }
        }
    } else {
        // =========================== consumers ===========================
        const int cw = warp - 1;
        unsigned* masks = reinterpret_cast<unsigned*>(a.ws + a.L.masks);
        unsigned* mb = sm.maskbuf[cw][0];
        unsigned* mb2 = sm.maskbuf[cw][1];
        int* pre = sm.pre[cw];
        int stage = 0;
        uint32_t fphase = 0;
        for (;;) {
            mbar_wait(reinterpret_cast<uint64_t*>(&sm.full[stage]), fphase);
            FastStage<T>& sg = sm.stage[stage];
            if (sg.item < 0) break;
            const uint32_t cand_addr = smem_base + (uint32_t)stage * kStageBytes;
            const uint32_t ash_addr = cand_addr + kAshOffset;
            T cm[9];
            if (UNW) {
#pragma unroll
                for (int k = 0; k < 9; ++k) cm[k] = sg.cm[k];
            }
            const int ntarget = sg.ntarget, home_off = sg.home_off, home_start = sg.home_start;
            const int nchunks = sg.nchunks;
            if (MODE == FAST_FILL_COO) {
                const unsigned* __restrict__ smk = sm.smasks[stage];
                for (;;) {
                    int t = 0;
                    if (lane == 0) t = atomicAdd(&sg.next_target, 1);
                    t = __shfl_sync(0xffffffffu, t, 0);
                    if (t >= ntarget) break;
                    int4 ai = make_int4(0, 0, 0, 0);
                    if (UNW) ai = lds_int4(ash_addr + (uint32_t)(home_off + t) * 16u);
                    fast_expand2<T, true, UNW>(a, sg, cand_addr, smk[t * 32 + lane], lane,
                                               lds_rec_j<T>(cand_addr + (uint32_t)(home_off + t) * RS), (size_t)sg.qrow[t],
                                               0x7fffffff, a.out_j, a.out_shifts, mb, pre, ash_addr, ai);
                }
            } else {
                for (;;) {
                    // two targets per trip: they share every candidate load and, in fp32, every FP instruction (f32x2)
                    int t = 0;
                    if (lane == 0) t = atomicAdd(&sg.next_target, 2);
                    t = __shfl_sync(0xffffffffu, t, 0);
                    if (t >= ntarget) break;
                    const bool two = !UNW && (t + 1 < ntarget);
                    const int ntrip = (t + 1 < ntarget) ? 2 : 1;
                    T xs[2], ys[2], zs[2];
                    int is[2];
                    lds_rec(cand_addr + (uint32_t)(home_off + t) * RS, xs[0], ys[0], zs[0], is[0]);
                    xs[1] = xs[0]; ys[1] = ys[0]; zs[1] = zs[0]; is[1] = is[0];
                    if (ntrip == 2) lds_rec(cand_addr + (uint32_t)(home_off + t + 1) * RS, xs[1], ys[1], zs[1], is[1]);
                    if (two) {
                        fast_masks2<HALF, FMA>(sg, cand_addr, xs[0], ys[0], zs[0], is[0], xs[1], ys[1], zs[1], is[1], a.cutoff_sq,
                                               lane, mb, mb2);
                        __syncwarp();
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        if (u >= ntrip) break;
                        const int tt = t + u;
                        const int self = home_off + tt;
                        const int i = is[u];
                        unsigned* mbu = u ? mb2 : mb;
                        int4 ai = make_int4(0, 0, 0, 0);
                        if (!two) {
                            if (UNW) {
                                ai = lds_int4(ash_addr + (uint32_t)self * 16u);
                                fast_masks_unw<T, HALF, FMA>(sg, cand_addr, ash_addr, xs[u], ys[u], zs[u], i, ai, cm, a.cutoff_sq, lane, mbu);
                            } else {
                                fast_masks<T, HALF, FMA>(sg, cand_addr, xs[u], ys[u], zs[u], i, a.cutoff_sq, lane, mbu);
                            }
                            __syncwarp();
                        }
                        unsigned mymask = lane < nchunks ? mbu[lane] : 0u;
                        if (!HALF && lane == (self >> 5)) mymask &= ~(1u << (self & 31));  // (i, i, 0) is not a pair
                        __syncwarp();
                        if (MODE == FAST_COUNT) {
                            masks[(size_t)(home_start + tt) * 32 + lane] = mymask;
                            const int cnt = __reduce_add_sync(0xffffffffu, __popc(mymask));
                            if (lane == 0) a.num_neighbors[i] = cnt;
                        } else {
                            const size_t p0 = (size_t)i * (size_t)a.max_neighbors;
                            // expansion scratch: reuse this target's own mask buffer (its words are already in registers)
                            const int cnt = fast_expand2<T, false, UNW>(a, sg, cand_addr, mymask, lane, i, p0, a.max_neighbors,
                                                                        a.neighbor_matrix, a.out_shifts, mbu, pre, ash_addr, ai);
                            finish_matrix_row<T>(a, lane, i, cnt);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(reinterpret_cast<uint64_t*>(&sm.empty[stage]));
            if (++stage == kFastStages) { stage = 0; fphase ^= 1u; }
        }
    }
}

}  // namespace nvnl
