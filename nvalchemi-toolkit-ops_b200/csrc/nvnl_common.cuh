// nvnl_common.cuh — shared definitions for the B200 (sm_100a) cell-list neighbor-list kernels.
//
// Path replaced (reference, read-only study):  nvalchemiops/neighborlist/cell_list.py:35-556,
// batch_cell_list.py:35-569, neighbor_utils.py:106-147.  Nothing here is a translation of the
// Warp kernels: the grid, the sort layout (float4 runs), the full-stencil warp-per-atom sweep and
// the ballot compaction are a different algorithm that produces the same neighbor set.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nvnl {

constexpr int kSweepThreads = 256;                 // 8 warps per CTA
constexpr int kSweepWarps = kSweepThreads / 32;
constexpr int kCandBytes = 16384;                  // fast kernel: staged candidate records per ring stage (bytes)
constexpr int kSweepCandBytes = 32768;             // general kernel: one tile (records, or records + periodic images)
constexpr int kRowCap = 160;                       // per-warp staged hits before a flush
constexpr int kDeferTargets = 32;                  // target atoms per deferred work item of the general kernel
constexpr int kMaxImg = 128;                       // cell images handled per batch (5^3 = 125 fits)
constexpr int kScanItems = 8;                      // items per thread in the look-back scan
constexpr int kScanThreads = 1024;                 // 8192-element tiles: the serial look-back chain is 4x shorter
constexpr int kScanTile = kScanItems * kScanThreads;
constexpr int kRowsPerAtom = 160;                  // single-sweep COO path: temporary row entries budgeted per atom
constexpr int kRowsSlackEntries = 148 * 4 * 8 * 2048;  // + one 2048-entry reservation block per resident consumer warp
constexpr int kSmallBlock = 256;                   // block size of the per-system / bounding-box kernels

enum SweepMode { MODE_COUNT = 0, MODE_FILL_COO = 1, MODE_FILL_MATRIX = 2 };

enum ErrorBits {
    ERR_IMAGE_RANGE = 1,      // periodic image index / search radius outside the supported range
    ERR_BAD_BATCH_IDX = 2,    // batch_idx outside [0, num_systems)
    ERR_SINGULAR_CELL = 4,    // cell matrix not invertible
    ERR_BAD_CACHE = 8,        // imported cell-list cache tensors are inconsistent (not produced by a build of this library)
};

// Per-system grid description, written by k_sys_init / k_grid (device side only).
struct SysParams {
    double inv[9];     // inverse of the cell matrix (row-major): frac_d = sum_k p_k * inv[3k + d]
    double cellm[9];   // cell matrix rows = lattice vectors
    double face[3];    // distance between opposite faces along each lattice direction
    double fmin[3];    // non-periodic dims: lower fractional bound of the atoms' bounding slab
    double fscale[3];  // non-periodic dims: cells per unit fractional length
    int cpd[3];        // cells per dimension
    int R[3];          // stencil radius (in cells) per dimension
    int pbc[3];
    int cell_offset;   // first global cell id of this system
    int ncells;
    int natoms;
    int pad;
};

// Device control block (lives at the start of the workspace).
struct Ctrl {
    int work_counter[4];             // persistent-CTA work queues (one per sweep launch kind)
    int done[4];                     // CTAs that drained their queue (the last one re-arms it)
    int unwrapped;                   // != 0 when some atom lies outside the primary periodic image
    int total_cells;
    int error;                       // ErrorBits
    int scan_tile[2];                // dynamic tile ids for the two look-back scans
    int n_deferred;                  // work items the fast kernels left to the general kernel (consumed per launch)
    int had_deferred;                // sticky since the last build: some cell was deferred
    int max_count;                   // max over atoms of num_neighbors (overflow check vs max_neighbors)
    unsigned long long total_pairs;  // 64-bit sum of num_neighbors (overflow check for int32 ptr)
    unsigned long long rows_cursor;  // single-sweep COO path: next free entry of the temporary row buffer
    int rows_overflow;               // single-sweep COO path: the temporary row buffer was too small (host falls back)
    int wide_stencil;                // some system searches more than one cell per side (periodic shifts may exceed +-1)
    int shift_heavy;                 // most atoms sit in cells at a periodic boundary (small boxes): their rows carry shifts,
                                     // so the single-sweep path writes the shifts output densely instead of pre-zeroing it
    // single-sweep COO path: cells with many targets are swept in parts by several CTAs (nvnl_rows.cuh)
    int split_reserved;              // entries of the part list handed out to pushers
    int split_next;                  // next entry to pop
    int cells_done;                  // cells whose parts (if any) have been pushed; == total_cells ends the pop loop
    // ... and single-cell systems whose 27 images need more than 64 chunks go to a second launch (k_rows<HUGE>)
    int n_huge;                      // entries of the huge list (parts of such cells)
    int huge_next, huge_done;        // pop counter / CTAs that drained it
    int had_huge;                    // sticky since the last build: some cell went to the huge list
};

// Programmatic dependent launch (PDL): kernels of the hot chain are launched with the programmatic-stream-serialization
// attribute, so the NEXT kernel's launch latency overlaps this kernel's execution.  First statement of every such
// kernel: let the dependents launch, then wait until the predecessor grid has completed and flushed its writes
// (a no-op when the kernel was launched without the attribute).
__device__ __forceinline__ void pdl_enter() {
#ifndef NVNL_NO_PDL
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

// Sorted candidate record: position + original atom index. 16 B (float) / 32 B (double) so that a
// cell's run is a 16-byte-aligned, 16-byte-granular block — a legal cp.async.bulk (TMA 1-D) source.
template <typename T> struct Rec;
template <> struct __align__(16) Rec<float> { float x, y, z; int j; };
template <> struct __align__(16) Rec<double> { double x, y, z; int j; int pad; };

// ---- arithmetic with a pinned operation order ------------------------------------------------
// The reference evaluates  dr = (r_j - r_i) + s·cell ;  d2 = dr·dr ;  d2 < rc*rc  in the input
// precision (cell_list.py:531-545).  With Warp's default fuse_fp the GPU code is mul + fma chains.
// Intrinsics (not operators) are used so nvcc can neither re-associate nor contract differently.
template <typename T> struct Arith;
template <> struct Arith<float> {
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
};
template <> struct Arith<double> {
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
};

// s·cell (row vector times matrix): r = row0*s0; r = fma(row1, s1, r); r = fma(row2, s2, r)
template <typename T, bool FMA>
__device__ __forceinline__ void shift_vector(const T* __restrict__ cm, int sx, int sy, int sz, T& Sx, T& Sy, T& Sz) {
    using A = Arith<T>;
    const T f0 = (T)sx, f1 = (T)sy, f2 = (T)sz;
    T r[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        T v = A::mul(cm[k], f0);
        if (FMA) {
            v = A::fma(cm[3 + k], f1, v);
            v = A::fma(cm[6 + k], f2, v);
        } else {
            v = A::add(v, A::mul(cm[3 + k], f1));
            v = A::add(v, A::mul(cm[6 + k], f2));
        }
        r[k] = v;
    }
    Sx = r[0]; Sy = r[1]; Sz = r[2];
}

template <typename T, bool FMA>
__device__ __forceinline__ T dist2(T dx, T dy, T dz) {
    using A = Arith<T>;
    if (FMA) {
        T d2 = A::mul(dx, dx);
        d2 = A::fma(dy, dy, d2);
        return A::fma(dz, dz, d2);
    } else {
        T d2 = A::add(A::mul(dx, dx), A::mul(dy, dy));
        return A::add(d2, A::mul(dz, dz));
    }
}

// floor division with non-negative remainder (same contract as the reference's wpdivmod,
// math/math.py:40-48), b > 0.
__device__ __forceinline__ void divmod_floor(int a, int b, int& q, int& r) {
    q = a / b;
    r = a - q * b;
    if (r < 0) { q -= 1; r += b; }
}

// ---- mbarrier / TMA (cp.async.bulk) wrappers ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // try_wait with a suspend-time hint: a waiting warp sleeps in hardware instead of spinning through issue slots
    // (ncu showed ~8 % of the count kernel's issued instructions in the bare try_wait/branch/yield loop)
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "NVNL_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra NVNL_DONE_%=;\n\t"
        "bra NVNL_WAIT_%=;\n\t"
        "NVNL_DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// dst/src 16-byte aligned, bytes a positive multiple of 16.
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

}  // namespace nvnl
