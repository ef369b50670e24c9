// nvnl_api.cu — C ABI (include/nvalchemi_nl_b200.h) over the sm_100a kernels.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>

#include "../../include/nvalchemi_nl_b200.h"
#include "nvnl_cache.cuh"
#include "nvnl_fast.cuh"
#include "nvnl_rows.cuh"

using namespace nvnl;

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
    if (e != cudaSuccess)
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    else
        snprintf(g_err, sizeof(g_err), "%s", what);
    return code;
}

#define NVNL_CHECK_LAUNCH(what)                                   \
    do {                                                          \
        g_launches.fetch_add(1, std::memory_order_relaxed);       \
        cudaError_t e__ = cudaGetLastError();                     \
        if (e__ != cudaSuccess) return fail(-2, what, e__);       \
    } while (0)

// Launch with the programmatic-stream-serialization attribute (PDL): the kernel may be scheduled while its predecessor
// in the stream still runs; every kernel launched this way starts with pdl_enter(), which waits for the predecessor
// grid to complete and flush.  NVNL_NO_PDL=1 in the environment (read once) turns the attribute off.
inline bool pdl_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("NVNL_NO_PDL");
        return !(e && e[0] && e[0] != '0');
    }();
    return on;
}
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);   // errors surface in NVNL_CHECK_LAUNCH
}

// temporary-row budget of the single-sweep COO path (nvnl_set_rows_budget; tests shrink it to force the fallback)
std::atomic<long long> g_rows_per_atom{kRowsPerAtom};
std::atomic<long long> g_rows_slack{kRowsSlackEntries};

WsLayout mk_layout(long long n, long long s, int rec) {
    // the slack of the temporary row buffer is one 2048-entry reservation block per consumer warp that can get work: no
    // more warps than trips (every trip has at least one target), so small inputs do not pay the 39 MB of a full machine
    long long slack = g_rows_slack.load();
    const long long by_atoms = (n + 64) * 2048;
    if (slack > by_atoms) slack = by_atoms;
    return make_layout(n, s, rec, g_rows_per_atom.load(), slack);
}

int rec_bytes(int dtype) { return dtype == NVNL_F64 ? (int)sizeof(Rec<double>) : (int)sizeof(Rec<float>); }

constexpr int kMaxDevices = 64;

int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

int sm_count() {
    static int n[kMaxDevices] = {0};
    const int dev = current_device();
    if (n[dev] == 0) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        n[dev] = v > 0 ? v : 148;
    }
    return n[dev];
}

constexpr size_t kSweepSmemBytes = (size_t)kSweepCandBytes + (size_t)kSweepWarps * kRowCap * 4 * sizeof(int) + sizeof(SweepSmem);

template <typename T, int MODE, bool HALF, bool FMA>
int launch_sweep_t(const SweepArgs<T>& a, cudaStream_t st) {
    auto kern = k_sweep<T, MODE, HALF, FMA>;
    static int bps[kMaxDevices] = {0};  // per instantiation and device (function attributes are per device)
    int& blocks_per_sm = bps[current_device()];
    if (blocks_per_sm == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSweepSmemBytes);
        if (e != cudaSuccess) return fail(-2, "cudaFuncSetAttribute(k_sweep)", e);
        int b = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, kSweepThreads, kSweepSmemBytes);
        if (e != cudaSuccess) return fail(-2, "occupancy(k_sweep)", e);
        blocks_per_sm = b > 0 ? b : 1;
    }
    long long grid = (long long)sm_count() * blocks_per_sm;
    const long long max_items = a.L.max_cells;
    if (grid > max_items) grid = max_items > 0 ? max_items : 1;
    launch_pdl(kern, (unsigned)grid, kSweepThreads, kSweepSmemBytes, st, a);
    NVNL_CHECK_LAUNCH("k_sweep");
    return 0;
}

template <typename T, int MODE, bool HALF, bool FMA, bool UNW>
int launch_fast_t(const SweepArgs<T>& a, cudaStream_t st) {
    auto kern = k_fast<T, MODE, HALF, FMA, UNW>;
    constexpr size_t smem = fast_smem_bytes<T, MODE, UNW>();
    static int bps[kMaxDevices] = {0};
    int& blocks_per_sm = bps[current_device()];
    if (blocks_per_sm == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(-2, "cudaFuncSetAttribute(k_fast)", e);
        int b = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, kFastThreads, smem);
        if (e != cudaSuccess) return fail(-2, "occupancy(k_fast)", e);
        blocks_per_sm = b > 0 ? b : 1;
    }
    long long grid = (long long)sm_count() * blocks_per_sm;
    const long long max_items = a.L.max_cells;
    if (grid > max_items) grid = max_items > 0 ? max_items : 1;
    launch_pdl(kern, (unsigned)grid, kFastThreads, smem, st, a);
    NVNL_CHECK_LAUNCH("k_fast");
    return 0;
}

template <bool HALF, bool FMA, bool HUGE, bool PAIR = false>
int launch_rows_t(const RowsArgs& a, cudaStream_t st) {
    auto kern = k_rows<HALF, FMA, HUGE, PAIR>;
    constexpr size_t smem = rows_smem_bytes(PAIR);
    static std::atomic<int> bps[kMaxDevices];
    int blocks_per_sm = bps[current_device()].load(std::memory_order_relaxed);
    if (blocks_per_sm == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(-2, "cudaFuncSetAttribute(k_rows)", e);
        int b = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, kRowsThreads, smem);
        if (e != cudaSuccess) return fail(-2, "occupancy(k_rows)", e);
        blocks_per_sm = b > 0 ? b : 1;
        bps[current_device()].store(blocks_per_sm, std::memory_order_relaxed);
    }
    long long grid = (long long)sm_count() * blocks_per_sm;
    const long long max_items = a.L.max_cells;
    if (grid > max_items) grid = max_items > 0 ? max_items : 1;
    launch_pdl(kern, (unsigned)grid, kRowsThreads, smem, st, a);
    NVNL_CHECK_LAUNCH("k_rows");
    return 0;
}

template <bool SPEC>
int launch_rows_out(const unsigned char* ws, const WsLayout& L, long long n, const int* neighbor_ptr, int* out_i, int* out_j,
                    int* out_shifts, int index_offset, int shifts_zeroed, long long spec_cap, cudaStream_t st) {
    auto kern = k_rows_out<SPEC>;
    constexpr size_t smem = sizeof(RowsOutSmem);
    static std::atomic<int> bps[kMaxDevices];
    int blocks_per_sm = bps[current_device()].load(std::memory_order_relaxed);
    if (blocks_per_sm == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(-2, "cudaFuncSetAttribute(k_rows_out)", e);
        int b = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, kOutWarps * 32, smem);
        if (e != cudaSuccess) return fail(-2, "occupancy(k_rows_out)", e);
        blocks_per_sm = b > 0 ? b : 1;
        bps[current_device()].store(blocks_per_sm, std::memory_order_relaxed);
    }
    // small inputs: 8 atoms per warp pass instead of 32 (the per-warp chain row references -> bulk copies -> rows is
    // latency, and there are too few warps to hide it)
    const int group = n < 262144 ? 8 : 32;
    long long grid = (n + kOutWarps * group - 1) / (kOutWarps * group);
    const long long cap = (long long)sm_count() * blocks_per_sm * 4;   // a few waves: the tail is short
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    launch_pdl(kern, (unsigned)grid, kOutWarps * 32, smem, st, ws, L, n, neighbor_ptr, out_i, out_j, out_shifts, index_offset,
                                                      shifts_zeroed, spec_cap, group);
    NVNL_CHECK_LAUNCH("k_rows_out");
    return 0;
}

// Start of every query: empty deferred list, empty temporary row buffer (and, for the single-sweep COO path,
// row_ref = -1 for every atom).
int launch_query_reset(unsigned char* ws, const WsLayout& L, long long n, int with_rows, cudaStream_t st) {
    launch_pdl(k_query_reset, 1, 32, 0, st, ws, L, n, with_rows);
    NVNL_CHECK_LAUNCH("k_query_reset");
    return 0;
}

// Every query is three launches: the lean kernel in its wrapped and unwrapped variant (the one that does not match the
// input retires at once) takes the cells it can and lists the rest; the general kernel takes the listed work items.
// hint < 0: nothing known about the workspace state (launch everything).  hint >= 0 (from nvnl_status after the
// count pass): bit 0 = atoms outside the primary image, bit 1 = some cell was deferred -> only the kernels that have
// work are launched.
template <typename T, int MODE, bool HALF, bool FMA>
int launch_pair(SweepArgs<T> a, int hint, cudaStream_t st) {
    a.queue = MODE;  // fast: queues 0..2 (the wrapped and the unwrapped variant run back to back; one of them retires at once)
    int rc = 0;
    if (hint < 0 || !(hint & 1)) rc = launch_fast_t<T, MODE, HALF, FMA, false>(a, st);
    if (rc) return rc;
    if (hint < 0 || (hint & 1)) rc = launch_fast_t<T, MODE, HALF, FMA, true>(a, st);
    if (rc) return rc;
    a.queue = 3;     // general
    if (hint < 0 || (hint & 2)) rc = launch_sweep_t<T, MODE, HALF, FMA>(a, st);
    return rc;
}

template <typename T, int MODE>
int launch_sweep(const SweepArgs<T>& a, int half_fill, int fma, cudaStream_t st, int hint = -1) {
    if (half_fill) {
        return fma ? launch_pair<T, MODE, true, true>(a, hint, st) : launch_pair<T, MODE, true, false>(a, hint, st);
    }
    return fma ? launch_pair<T, MODE, false, true>(a, hint, st) : launch_pair<T, MODE, false, false>(a, hint, st);
}

template <typename T>
int build_t(const T* pos, long long n, const T* cell, const uint8_t* pbc, const int* batch_idx, const int* batch_ptr,
            int ns, double cutoff, long long max_cells, unsigned char* ws, cudaStream_t st) {
    const WsLayout L = mk_layout(n, ns, (int)sizeof(Rec<T>));
    const int sms = sm_count();
    {
        long long work = L.max_cells + 2 > n ? L.max_cells + 2 : n;
        long long blocks = (work + 255) / 256;
        if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
        if (blocks < 1) blocks = 1;
        launch_pdl(k_init<T>, (unsigned)blocks, 256, 0, st, ws, L, n, ns, cell, pbc, batch_ptr, nullptr);
        NVNL_CHECK_LAUNCH("k_init");
    }
    {
        const int need_counts = (batch_ptr == nullptr && ns > 1) ? 1 : 0;
        const long long chunk = ns > 1 ? 512 : 2048;
        long long blocks = (n + chunk - 1) / chunk;
        if (blocks > (long long)sms * 4) blocks = (long long)sms * 4;
        if (blocks < 1) blocks = 1;
        launch_pdl(k_bbox<T>, (unsigned)blocks, kSmallBlock, 0, st, ws, L, n, ns, pos, batch_idx, need_counts);
        NVNL_CHECK_LAUNCH("k_bbox");
    }
    launch_pdl(k_grid, 1, ns > kSmallBlock ? 1024 : kSmallBlock, 0, st, ws, L, ns, cutoff, max_cells > 0 ? (max_cells / ns > 0 ? max_cells / ns : 1) : 0);
    NVNL_CHECK_LAUNCH("k_grid");
    const bool vec = (reinterpret_cast<uintptr_t>(pos) % 16) == 0;
    {
        const long long threads = (n + 3) / 4;
        const unsigned blocks = (unsigned)((threads + 255) / 256);
        if (vec)
            launch_pdl(k_hash<T, true>, blocks, 256, 0, st, ws, L, n, ns, pos, batch_idx);
        else
            launch_pdl(k_hash<T, false>, blocks, 256, 0, st, ws, L, n, ns, pos, batch_idx);
        NVNL_CHECK_LAUNCH("k_hash");
    }
    {
        Ctrl* ctrl = reinterpret_cast<Ctrl*>(ws + L.ctrl);
        const long long cnt = L.max_cells + 1;  // out has cnt + 1 entries
        const unsigned blocks = (unsigned)((cnt + 1 + kScanTile - 1) / kScanTile);
        launch_pdl(k_scan, blocks, kScanThreads, 0, st, reinterpret_cast<const int*>(ws + L.cell_count),
                                               reinterpret_cast<int*>(ws + L.cell_start), cnt,
                                               reinterpret_cast<unsigned long long*>(ws + L.scan_status0),
                                               &ctrl->scan_tile[0], nullptr, nullptr, &ctrl->total_cells);
        NVNL_CHECK_LAUNCH("k_scan(cells)");
    }
    {
        const long long threads = (n + 3) / 4;
        const unsigned blocks = (unsigned)((threads + 255) / 256);
        if (vec)
            launch_pdl(k_scatter<T, true>, blocks, 256, 0, st, ws, L, n, pos);
        else
            launch_pdl(k_scatter<T, false>, blocks, 256, 0, st, ws, L, n, pos);
        NVNL_CHECK_LAUNCH("k_scatter");
    }
    return 0;
}

template <typename T>
int import_t(const T* pos, long long n, const T* cell, const uint8_t* pbc, const int* batch_idx, int ns, double cutoff,
             const int* cpd, const int* radius, const int* atom_shifts, const int* atom_cell_map, const int* cell_count,
             const int* cell_start, long long cache_cells, const int* cell_atom_list, unsigned char* ws, cudaStream_t st) {
    const WsLayout L = mk_layout(n, ns, (int)sizeof(Rec<T>));
    const int sms = sm_count();
    {
        long long work = L.max_cells + 2 > n ? L.max_cells + 2 : n;
        long long blocks = (work + 255) / 256;
        if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
        if (blocks < 1) blocks = 1;
        launch_pdl(k_init<T>, (unsigned)blocks, 256, 0, st, ws, L, n, ns, cell, pbc, nullptr, nullptr);
        NVNL_CHECK_LAUNCH("k_init");
    }
    k_import_sys<<<1, kSmallBlock, 0, st>>>(ws, L, ns, cutoff, cpd, radius, cache_cells);
    NVNL_CHECK_LAUNCH("k_import_sys");
    {
        long long work = n > cache_cells ? n : cache_cells;
        long long blocks = (work + 255) / 256;
        if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
        if (blocks < 1) blocks = 1;
        k_import_atoms<T><<<(unsigned)blocks, 256, 0, st>>>(ws, L, n, ns, pos, batch_idx, atom_shifts, atom_cell_map, cell_count,
                                                            cell_start, cell_atom_list);
        NVNL_CHECK_LAUNCH("k_import_atoms");
    }
    return 0;
}

template <typename T>
SweepArgs<T> base_args(unsigned char* ws, long long n, int ns, const int* batch_idx, double cutoff_sq) {
    SweepArgs<T> a;
    memset(&a, 0, sizeof(a));
    a.ws = ws;
    a.L = mk_layout(n, ns, (int)sizeof(Rec<T>));
    a.batch_idx = batch_idx;
    a.num_systems = ns;
    a.n = n;
    a.cutoff_sq = (T)cutoff_sq;
    return a;
}

template <typename T>
int count_t(unsigned char* ws, long long n, int ns, const int* batch_idx, double cutoff_sq, int half_fill, int fma,
            int* num_neighbors, int* neighbor_ptr, cudaStream_t st) {
    SweepArgs<T> a = base_args<T>(ws, n, ns, batch_idx, cutoff_sq);
    a.num_neighbors = num_neighbors;
    a.queue = 0;
    int rc = launch_query_reset(ws, a.L, n, 0, st);
    if (rc) return rc;
    rc = launch_sweep<T, MODE_COUNT>(a, half_fill, fma, st);
    if (rc) return rc;
    if (neighbor_ptr) {
        Ctrl* ctrl = reinterpret_cast<Ctrl*>(ws + a.L.ctrl);
        const unsigned blocks = (unsigned)((n + 1 + kScanTile - 1) / kScanTile);
        launch_pdl(k_scan, blocks, kScanThreads, 0, st, num_neighbors, neighbor_ptr, n,
                                               reinterpret_cast<unsigned long long*>(ws + a.L.scan_status1),
                                               &ctrl->scan_tile[1], &ctrl->total_pairs, &ctrl->max_count, nullptr);
        NVNL_CHECK_LAUNCH("k_scan(neighbors)");
    }
    return 0;
}

template <bool HALF, bool FMA>
int count_rows_t(unsigned char* ws, long long n, int ns, const int* batch_idx, double cutoff_sq, int* num_neighbors,
                 int* neighbor_ptr, int* prezero, long long prezero_ints, int hint, cudaStream_t st) {
    SweepArgs<float> a = base_args<float>(ws, n, ns, batch_idx, cutoff_sq);
    a.num_neighbors = num_neighbors;
    int rc = launch_query_reset(ws, a.L, n, 1, st);
    if (rc) return rc;
    RowsArgs r;
    r.ws = ws; r.L = a.L; r.batch_idx = batch_idx; r.num_systems = ns; r.n = n; r.cutoff_sq = (float)cutoff_sq;
    r.num_neighbors = num_neighbors;
    r.prezero = prezero_ints > 0 ? prezero : nullptr;
    r.prezero_ints = prezero_ints > 0 ? prezero_ints : 0;
    r.q_sorted = nullptr; r.pair_energies = nullptr; r.pair_forces = nullptr; r.pair_cutoff = 0.0; r.pair_alpha = 0.0;
    rc = launch_rows_t<HALF, FMA, false>(r, st);                            // wrapped inputs: the single sweep
    if (rc) return rc;
    if (hint < 0 || (hint & 32)) {
        rc = launch_rows_t<HALF, FMA, true>(r, st);                         // single-cell systems with 27 images (else retires)
        if (rc) return rc;
    }
    // hint >= 0 (what nvnl_status reported for an earlier query with this signature): the kernels that would find no
    // work are not launched; the caller checks nvnl_status afterwards and repeats the call with hint = -1 if it was wrong
    if (hint < 0 || (hint & 1)) {
        a.queue = 0;
        rc = launch_fast_t<float, FAST_COUNT, HALF, FMA, true>(a, st);      // unwrapped inputs: two-pass count (else retires)
        if (rc) return rc;
    }
    if (hint < 0 || (hint & 2)) {
        a.queue = 3;
        a.keep_deferred = 1;
        rc = launch_sweep_t<float, MODE_COUNT, HALF, FMA>(a, st);           // whatever the lean kernel deferred
        if (rc) return rc;
    }
    if (neighbor_ptr) {
        Ctrl* ctrl = reinterpret_cast<Ctrl*>(ws + a.L.ctrl);
        const unsigned blocks = (unsigned)((n + 1 + kScanTile - 1) / kScanTile);
        launch_pdl(k_scan, blocks, kScanThreads, 0, st, num_neighbors, neighbor_ptr, n,
                                               reinterpret_cast<unsigned long long*>(ws + a.L.scan_status1),
                                               &ctrl->scan_tile[1], &ctrl->total_pairs, &ctrl->max_count, nullptr);
        NVNL_CHECK_LAUNCH("k_scan(neighbors)");
    }
    return 0;
}

template <bool HALF, bool FMA>
int fill_rows_t(unsigned char* ws, long long n, int ns, const int* batch_idx, double cutoff_sq, const int* neighbor_ptr,
                int* edge_index, long long row_stride, int* shifts, int index_offset, int hint, cudaStream_t st) {
    SweepArgs<float> a = base_args<float>(ws, n, ns, batch_idx, cutoff_sq);
    a.neighbor_ptr = neighbor_ptr; a.out_i = edge_index; a.out_j = edge_index + row_stride; a.out_shifts = shifts;
    a.index_offset = index_offset;
    if (hint & 1) {
        // unwrapped input: the count ran on the two-pass kernels (hit masks), so does the fill
        a.queue = 1;
        launch_pdl(k_gather_ptr<float>, (unsigned)((n + 255) / 256), 256, 0, st, ws, a.L, n, neighbor_ptr,
                                                                       reinterpret_cast<int*>(ws + a.L.ptr_sorted));
        NVNL_CHECK_LAUNCH("k_gather_ptr");
        return launch_pair<float, MODE_FILL_COO, HALF, FMA>(a, hint, st);
    }
    if (!(hint & 8)) {  // bit 3: nvnl_fill_rows_speculative already wrote the rows of the lean kernel
        const int rc = launch_rows_out<false>(ws, a.L, n, neighbor_ptr, a.out_i, a.out_j, a.out_shifts, index_offset,
                                              (hint & 4) ? 1 : 0, 0, st);
        if (rc) return rc;
    }
    if (hint & 2) {
        a.queue = 3;  // the deferred list of the count stage was kept for this launch
        return launch_sweep_t<float, MODE_FILL_COO, HALF, FMA>(a, st);
    }
    return 0;
}

__global__ void k_get_grid(const unsigned char* __restrict__ ws, WsLayout L, int ns, int* __restrict__ cpd,
                           int* __restrict__ radius) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= ns) return;
    const SysParams* sys = reinterpret_cast<const SysParams*>(ws + L.sys);
    for (int d = 0; d < 3; ++d) {
        if (cpd) cpd[3 * s + d] = sys[s].cpd[d];
        if (radius) radius[3 * s + d] = sys[s].R[d];
    }
}

// ---- multi-GPU re-assembly: ranks exchange only what cannot be recomputed --------------------------------------
// A pair's source atom follows from neighbor_ptr and its periodic shift fits one byte when every shift component is in
// {-1, 0, 1} (wrapped inputs, search radius 1), so a rank sends 4 B (target atom) + 1 B (packed shift) per pair instead
// of 20 B; the receiver expands the foreign ranges.
//   packed byte = (sx + 1) | (sy + 1) << 2 | (sz + 1) << 4;   k_pack_shifts sets *bad when a component is outside {-1,0,1}
__global__ void k_pack_shifts(const int* __restrict__ shifts, long long n_pairs, unsigned char* __restrict__ packed,
                              int* __restrict__ bad) {
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    int err = 0;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += nthreads) {
        const int sx = shifts[3 * p] + 1, sy = shifts[3 * p + 1] + 1, sz = shifts[3 * p + 2] + 1;
        err |= (sx | sy | sz) & ~3;
        err |= (sx == 3) | (sy == 3) | (sz == 3);
        packed[p] = (unsigned char)((sx & 3) | ((sy & 3) << 2) | ((sz & 3) << 4));
    }
    if (err && bad) atomicOr(bad, 1);
}

// One-word form of the exchange (atom indices below 2^26): the packed shift rides in bits 26..31 of the pair's target
// word, so a rank sends 4 B per pair and ONE array.  *bad is set when a component is outside {-1,0,1} or a target >= 2^26.
constexpr int kWordShift = 26;
constexpr unsigned kWordMask = (1u << kWordShift) - 1u;
__global__ void k_pack_shifts_word(const int* __restrict__ shifts, long long n_pairs, int* __restrict__ targets, int* __restrict__ bad) {
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    int err = 0;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += nthreads) {
        const int sx = shifts[3 * p] + 1, sy = shifts[3 * p + 1] + 1, sz = shifts[3 * p + 2] + 1;
        err |= (sx | sy | sz) & ~3;
        err |= (sx == 3) | (sy == 3) | (sz == 3);
        const unsigned t = (unsigned)targets[p];
        err |= (t & ~kWordMask) != 0u;
        targets[p] = (int)((t & kWordMask) | ((unsigned)((sx & 3) | ((sy & 3) << 2) | ((sz & 3) << 4)) << kWordShift));
    }
    if (err && bad) atomicOr(bad, 1);
}

// out_i / shifts of every atom OUTSIDE [atom_lo, atom_hi) from neighbor_ptr and the gathered packed shifts (the rank's
// own range was written by its fill kernels).  One warp per 32 atoms, rows written sequentially.
__global__ void __launch_bounds__(256) k_expand_gathered(const int* __restrict__ neighbor_ptr, long long n_atoms,
                                                         long long atom_lo, long long atom_hi,
                                                         const unsigned char* __restrict__ packed, int* __restrict__ out_i,
                                                         int* __restrict__ shifts) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long base = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; base < n_atoms;
         base += nwarps * 32) {
        if (base >= atom_lo && base + 32 <= atom_hi) continue;
        const long long il = base + lane;
        const int p_l = neighbor_ptr[il < n_atoms ? il : n_atoms];
        const int e_l = neighbor_ptr[il + 1 < n_atoms ? il + 1 : n_atoms];
        const int na = n_atoms - base < 32 ? (int)(n_atoms - base) : 32;
        for (int t = 0; t < na; ++t) {
            const long long i = base + t;
            if (i >= atom_lo && i < atom_hi) continue;
            const int p = __shfl_sync(0xffffffffu, p_l, t), cnt = __shfl_sync(0xffffffffu, e_l, t) - p;
            for (int k = lane; k < cnt; k += 32) out_i[(size_t)p + k] = (int)i;
            int* __restrict__ sh = shifts + 3 * (size_t)p;
            const unsigned char* __restrict__ pk = packed + (size_t)p;
            for (int e = lane; e < 3 * cnt; e += 32) {
                const int k = e / 3, c = e - 3 * k;
                sh[e] = (int)((pk[k] >> (2 * c)) & 3) - 1;
            }
        }
    }
}

// Re-assembly after the padded all-gather of the packed exchange: rank g's targets / packed shifts sit at
// gathered_dst[g * pmax + k] / gathered_packed[g * pmax + k] (k = pair index inside the rank's range).  Writes out_j for
// every pair, and out_i / shifts for the pairs of the OTHER ranks (the rank's own were written by its fill kernels).
// (Chunked exchange: one launch per chunk; a rank's atoms of that chunk are [atom_lo[g], atom_hi[g]), the atoms between
// atom_hi[g] and atom_lo[g + 1] belong to other chunks and are skipped.)
struct ExpandRanks {
    int world, rank;
    long long atom_lo[17];    // atoms of rank g: [atom_lo[g], atom_hi[g]); atom_lo ascending, atom_lo[world] = end of the last range
    long long atom_hi[17];
    long long pair_lo[17];    // first pair of rank g's range (= neighbor_ptr[atom_lo[g]])
};
// WORD: gathered_dst holds target | packed shift << 26 and gathered_packed is not read.
template <bool WORD>
__global__ void __launch_bounds__(256) k_expand_padded(const int* __restrict__ neighbor_ptr, long long n_atoms, ExpandRanks R,
                                                       long long pmax, const int* __restrict__ gathered_dst,
                                                       const unsigned char* __restrict__ gathered_packed,
                                                       int* __restrict__ out_i, int* __restrict__ out_j, int* __restrict__ shifts) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long base = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; base < n_atoms;
         base += nwarps * 32) {
        const long long il = base + lane;
        const int p_l = neighbor_ptr[il < n_atoms ? il : n_atoms];
        const int e_l = neighbor_ptr[il + 1 < n_atoms ? il + 1 : n_atoms];
        const int na = n_atoms - base < 32 ? (int)(n_atoms - base) : 32;
        int g = 0;
        while (g + 1 < R.world && base >= R.atom_lo[g + 1]) ++g;          // rank of the group's first atom
        // the whole group lies between two ranks' ranges of this chunk (or outside all of them): nothing to do
        if (base + 32 <= R.atom_lo[0] || (base >= R.atom_hi[g] && (g + 1 >= R.world || base + 32 <= R.atom_lo[g + 1]))) continue;
        // element e = lane + 32 u of a group of 32 pairs' shifts (96 ints) belongs to pair e / 3, component e % 3
        int q[3], c2[3];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int e = lane + 32 * u;
            q[u] = e / 3;
            c2[u] = 2 * (e - 3 * q[u]);
        }
        for (int t = 0; t < na; ++t) {
            const long long i = base + t;
            while (g + 1 < R.world && i >= R.atom_lo[g + 1]) ++g;
            if (i < R.atom_lo[g] || i >= R.atom_hi[g]) continue;           // an atom of another chunk (warp-uniform)
            const int p = __shfl_sync(0xffffffffu, p_l, t), cnt = __shfl_sync(0xffffffffu, e_l, t) - p;
            const long long src = (long long)g * pmax + ((long long)p - R.pair_lo[g]);
            const int* __restrict__ dj = gathered_dst + src;
            const unsigned char* __restrict__ pk = gathered_packed + src;
            const bool foreign = g != R.rank;
            // four chunks of 32 pairs per pass: all loads of a pass are issued before its stores (memory-level parallelism;
            // a row has ~90 pairs, so one pass usually covers it)
            for (int k0 = 0; k0 < cnt; k0 += 128) {
                int dv[4], pv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int k = k0 + 32 * u + lane;
                    dv[u] = k < cnt ? dj[k] : 0;
                    if (WORD) {
                        pv[u] = (int)((unsigned)dv[u] >> kWordShift);
                        dv[u] = (int)((unsigned)dv[u] & kWordMask);
                    } else {
                        pv[u] = (foreign && k < cnt) ? (int)pk[k] : 0;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int kb = k0 + 32 * u;
                    if (kb >= cnt) break;
                    const int k = kb + lane;
                    if (k < cnt) out_j[(size_t)p + k] = dv[u];
                    if (!foreign) continue;
                    if (k < cnt) out_i[(size_t)p + k] = (int)i;
                    int* __restrict__ sh = shifts + 3 * ((size_t)p + kb);
                    const int nel = 3 * (cnt - kb);
#pragma unroll
                    for (int w = 0; w < 3; ++w) {
                        const int pq = __shfl_sync(0xffffffffu, pv[u], q[w]);
                        if (lane + 32 * w < nel) sh[lane + 32 * w] = ((pq >> c2[w]) & 3) - 1;
                    }
                }
            }
        }
    }
}

}  // namespace

extern "C" {

int nvnl_abi_version(void) { return NVNL_ABI_VERSION; }
const char* nvnl_last_error(void) { return g_err; }
int64_t nvnl_launch_count(void) { return (int64_t)g_launches.load(); }

void nvnl_set_rows_budget(int64_t entries_per_atom, int64_t slack_entries) {
    g_rows_per_atom.store(entries_per_atom >= 0 ? entries_per_atom : kRowsPerAtom);
    g_rows_slack.store(slack_entries >= 0 ? slack_entries : kRowsSlackEntries);
}

size_t nvnl_workspace_bytes(int64_t n_atoms, int64_t n_systems, int dtype) {
    if (n_atoms < 0 || n_systems < 0) return 0;
    return mk_layout(n_atoms, n_systems > 0 ? n_systems : 1, rec_bytes(dtype)).total;
}

int nvnl_build(const void* positions, int dtype, int64_t n_atoms, const void* cell, const uint8_t* pbc,
               const int32_t* batch_idx, const int32_t* batch_ptr, int32_t n_systems, double cutoff, int64_t max_cells,
               void* workspace, size_t workspace_bytes, void* stream) {
    if (n_atoms <= 0 || n_systems <= 0) return fail(-1, "nvnl_build: n_atoms and n_systems must be positive");
    if (n_atoms > 2000000000LL) return fail(-1, "nvnl_build: n_atoms exceeds the int32 index range");
    if (!(cutoff > 0.0)) return fail(-1, "nvnl_build: cutoff must be positive");
    if (max_cells < 0 || (max_cells > 0 && max_cells < n_systems)) return fail(-1, "nvnl_build: max_cells must be 0 (no cap) or at least n_systems");
    if (!positions || !cell || !pbc || !workspace) return fail(-1, "nvnl_build: null pointer");
    if (n_systems > 1 && !batch_idx) return fail(-1, "nvnl_build: batch_idx is required for n_systems > 1");
    if (dtype != NVNL_F32 && dtype != NVNL_F64) return fail(-1, "nvnl_build: unsupported dtype");
    if (workspace_bytes < nvnl_workspace_bytes(n_atoms, n_systems, dtype))
        return fail(-1, "nvnl_build: workspace too small");
    if (reinterpret_cast<uintptr_t>(workspace) % 256) return fail(-1, "nvnl_build: workspace must be 256-byte aligned");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    if (dtype == NVNL_F32)
        return build_t<float>(static_cast<const float*>(positions), n_atoms, static_cast<const float*>(cell), pbc,
                              batch_idx, batch_ptr, n_systems, cutoff, max_cells, ws, st);
    return build_t<double>(static_cast<const double*>(positions), n_atoms, static_cast<const double*>(cell), pbc,
                           batch_idx, batch_ptr, n_systems, cutoff, max_cells, ws, st);
}

int nvnl_count(void* workspace, int dtype, int64_t n_atoms, int32_t n_systems, const int32_t* batch_idx,
               double cutoff_sq, int half_fill, int fma, int32_t* num_neighbors, int32_t* neighbor_ptr, void* stream) {
    if (!workspace || !num_neighbors || n_atoms <= 0) return fail(-1, "nvnl_count: bad arguments");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    if (dtype == NVNL_F32)
        return count_t<float>(ws, n_atoms, n_systems, batch_idx, cutoff_sq, half_fill, fma, num_neighbors, neighbor_ptr, st);
    if (dtype == NVNL_F64)
        return count_t<double>(ws, n_atoms, n_systems, batch_idx, cutoff_sq, half_fill, fma, num_neighbors, neighbor_ptr, st);
    return fail(-1, "nvnl_count: unsupported dtype");
}

int nvnl_status(void* workspace, int dtype, int64_t n_atoms, int32_t n_systems, int64_t* total_pairs,
                int32_t* max_count, int32_t* total_cells, int32_t* error_bits, int32_t* unwrapped, int32_t* had_deferred,
                int32_t* rows_overflow, void* stream) {
    if (!workspace) return fail(-1, "nvnl_status: null workspace");
    const WsLayout L = mk_layout(n_atoms, n_systems, rec_bytes(dtype));
    Ctrl h;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemcpyAsync(&h, static_cast<unsigned char*>(workspace) + L.ctrl, sizeof(Ctrl),
                                    cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return fail(-2, "nvnl_status: memcpy", e);
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(-2, "nvnl_status: sync", e);
    if (total_pairs) *total_pairs = (int64_t)h.total_pairs;
    if (max_count) *max_count = h.max_count;
    if (total_cells) *total_cells = h.total_cells;
    if (error_bits) *error_bits = h.error;
    if (unwrapped) *unwrapped = (h.unwrapped ? 1 : 0) | (h.wide_stencil ? 2 : 0);
    if (had_deferred) *had_deferred = (h.had_deferred ? 1 : 0) | (h.had_huge ? 2 : 0);
    if (rows_overflow) *rows_overflow = h.rows_overflow;
    return 0;
}

int nvnl_fill_coo(void* workspace, int dtype, int64_t n_atoms, int32_t n_systems, const int32_t* batch_idx,
                  double cutoff_sq, int half_fill, int fma, const int32_t* neighbor_ptr, int32_t* edge_index,
                  int64_t num_pairs, int64_t row_stride, int32_t* shifts, int32_t index_offset, int32_t launch_hint,
                  void* stream) {
    if (!workspace || !neighbor_ptr || n_atoms <= 0) return fail(-1, "nvnl_fill_coo: bad arguments");
    if (row_stride <= 0) row_stride = num_pairs;
    if (num_pairs < 0 || num_pairs > 2147483647LL) return fail(-1, "nvnl_fill_coo: num_pairs outside int32 range");
    if (num_pairs == 0) return 0;
    if (!edge_index || !shifts) return fail(-1, "nvnl_fill_coo: null output");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    if (dtype == NVNL_F32) {
        SweepArgs<float> a = base_args<float>(ws, n_atoms, n_systems, batch_idx, cutoff_sq);
        a.neighbor_ptr = neighbor_ptr; a.out_i = edge_index; a.out_j = edge_index + row_stride; a.out_shifts = shifts;
        a.index_offset = index_offset; a.queue = 1;
        launch_pdl(k_gather_ptr<float>, (unsigned)((n_atoms + 255) / 256), 256, 0, st, ws, a.L, n_atoms, neighbor_ptr,
                                                                              reinterpret_cast<int*>(ws + a.L.ptr_sorted));
        NVNL_CHECK_LAUNCH("k_gather_ptr");
        return launch_sweep<float, MODE_FILL_COO>(a, half_fill, fma, st, launch_hint);
    }
    if (dtype == NVNL_F64) {
        SweepArgs<double> a = base_args<double>(ws, n_atoms, n_systems, batch_idx, cutoff_sq);
        a.neighbor_ptr = neighbor_ptr; a.out_i = edge_index; a.out_j = edge_index + row_stride; a.out_shifts = shifts;
        a.index_offset = index_offset; a.queue = 1;
        launch_pdl(k_gather_ptr<double>, (unsigned)((n_atoms + 255) / 256), 256, 0, st, ws, a.L, n_atoms, neighbor_ptr,
                                                                               reinterpret_cast<int*>(ws + a.L.ptr_sorted));
        NVNL_CHECK_LAUNCH("k_gather_ptr");
        return launch_sweep<double, MODE_FILL_COO>(a, half_fill, fma, st, launch_hint);
    }
    return fail(-1, "nvnl_fill_coo: unsupported dtype");
}

int nvnl_count_rows(void* workspace, int dtype, int64_t n_atoms, int32_t n_systems, const int32_t* batch_idx,
                    double cutoff_sq, int half_fill, int fma, int32_t* num_neighbors, int32_t* neighbor_ptr,
                    int32_t* prezero, int64_t prezero_ints, int32_t launch_hint, void* stream) {
    if (!workspace || !num_neighbors || n_atoms <= 0) return fail(-1, "nvnl_count_rows: bad arguments");
    if (dtype != NVNL_F32) return fail(-1, "nvnl_count_rows: the single-sweep path is fp32 only (use nvnl_count)");
    if (n_atoms >= (1LL << 27)) return fail(-1, "nvnl_count_rows: the single-sweep path takes fewer than 2^27 atoms (use nvnl_count)");
    if (prezero_ints < 0 || (prezero_ints > 0 && (!prezero || reinterpret_cast<uintptr_t>(prezero) % 16)))
        return fail(-1, "nvnl_count_rows: prezero must be a 16-byte aligned device pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    if (half_fill)
        return fma ? count_rows_t<true, true>(ws, n_atoms, n_systems, batch_idx, cutoff_sq, num_neighbors, neighbor_ptr, prezero, prezero_ints, launch_hint, st)
                   : count_rows_t<true, false>(ws, n_atoms, n_systems, batch_idx, cutoff_sq, num_neighbors, neighbor_ptr, prezero, prezero_ints, launch_hint, st);
    return fma ? count_rows_t<false, true>(ws, n_atoms, n_systems, batch_idx, cutoff_sq, num_neighbors, neighbor_ptr, prezero, prezero_ints, launch_hint, st)
               : count_rows_t<false, false>(ws, n_atoms, n_systems, batch_idx, cutoff_sq, num_neighbors, neighbor_ptr, prezero, prezero_ints, launch_hint, st);
}

int nvnl_fill_rows(void* workspace, int dtype, int64_t n_atoms, int32_t n_systems, const int32_t* batch_idx,
                   double cutoff_sq, int half_fill, int fma, const int32_t* neighbor_ptr, int32_t* edge_index,
                   int64_t num_pairs, int64_t row_stride, int32_t* shifts, int32_t index_offset, int32_t launch_hint,
                   void* stream) {
    if (!workspace || !neighbor_ptr || n_atoms <= 0) return fail(-1, "nvnl_fill_rows: bad arguments");
    if (row_stride <= 0) row_stride = num_pairs;
    if (dtype != NVNL_F32) return fail(-1, "nvnl_fill_rows: the single-sweep path is fp32 only (use nvnl_fill_coo)");
    if (launch_hint < 0) return fail(-1, "nvnl_fill_rows: launch_hint from nvnl_status is required");
    if (num_pairs < 0 || num_pairs > 2147483647LL) return fail(-1, "nvnl_fill_rows: num_pairs outside int32 range");
    if (num_pairs == 0) return 0;
    if (!edge_index || !shifts) return fail(-1, "nvnl_fill_rows: null output");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    if (half_fill)
        return fma ? fill_rows_t<true, true>(ws, n_atoms, n_systems, batch_idx, cutoff_sq, neighbor_ptr, edge_index, row_stride,
                                             shifts, index_offset, launch_hint, st)
                   : fill_rows_t<true, false>(ws, n_atoms, n_systems, batch_idx, cutoff_sq, neighbor_ptr, edge_index, row_stride,
                                              shifts, index_offset, launch_hint, st);
    return fma ? fill_rows_t<false, true>(ws, n_atoms, n_systems, batch_idx, cutoff_sq, neighbor_ptr, edge_index, row_stride,
                                          shifts, index_offset, launch_hint, st)
               : fill_rows_t<false, false>(ws, n_atoms, n_systems, batch_idx, cutoff_sq, neighbor_ptr, edge_index, row_stride,
                                           shifts, index_offset, launch_hint, st);
}

int nvnl_fill_rows_speculative(void* workspace, int dtype, int64_t n_atoms, int32_t n_systems, const int32_t* neighbor_ptr,
                               int32_t* edge_buffer, int64_t capacity_pairs, int32_t* shifts_zeroed, int32_t index_offset,
                               void* stream) {
    if (!workspace || !neighbor_ptr || !edge_buffer || !shifts_zeroed || n_atoms <= 0)
        return fail(-1, "nvnl_fill_rows_speculative: bad arguments");
    if (dtype != NVNL_F32) return fail(-1, "nvnl_fill_rows_speculative: the single-sweep path is fp32 only");
    if (capacity_pairs <= 0 || capacity_pairs > 2147483647LL) return fail(-1, "nvnl_fill_rows_speculative: capacity outside int32 range");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    const WsLayout L = mk_layout(n_atoms, n_systems, rec_bytes(dtype));
    return launch_rows_out<true>(ws, L, n_atoms, neighbor_ptr, edge_buffer, nullptr, shifts_zeroed, index_offset, 1,
                                 capacity_pairs, st);
}

int nvnl_coulomb_fused(void* workspace, int dtype, int64_t n_atoms, int32_t n_systems, const int32_t* batch_idx,
                       double cutoff_sq, int fma, const double* charges, double cutoff, double alpha, double* energies,
                       double* forces, void* stream) {
    if (!workspace || !charges || !energies || !forces || n_atoms <= 0) return fail(-1, "nvnl_coulomb_fused: bad arguments");
    if (dtype != NVNL_F32) return fail(-1, "nvnl_coulomb_fused: the fused sweep is fp32-position only");
    if (n_atoms >= (1LL << 27)) return fail(-1, "nvnl_coulomb_fused: atom indices must be below 2^27");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    SweepArgs<float> a = base_args<float>(ws, n_atoms, n_systems, batch_idx, cutoff_sq);
    int rc = launch_query_reset(ws, a.L, n_atoms, 1, st);
    if (rc) return rc;
    // charges in cell-sorted order; they live in the hit-mask region of the workspace, which this path does not use
    double* q_sorted = reinterpret_cast<double*>(ws + a.L.masks);
    {
        long long blocks = (n_atoms + 255) / 256;
        if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
        launch_pdl(k_gather_q, (unsigned)blocks, 256, 0, st, ws, a.L, n_atoms, charges, q_sorted);
        NVNL_CHECK_LAUNCH("k_gather_q");
    }
    RowsArgs r;
    r.ws = ws; r.L = a.L; r.batch_idx = batch_idx; r.num_systems = n_systems; r.n = n_atoms; r.cutoff_sq = (float)cutoff_sq;
    r.num_neighbors = nullptr; r.prezero = nullptr; r.prezero_ints = 0;
    r.q_sorted = q_sorted; r.pair_energies = energies; r.pair_forces = forces; r.pair_cutoff = cutoff; r.pair_alpha = alpha;
    return fma ? launch_rows_t<false, true, false, true>(r, st) : launch_rows_t<false, false, false, true>(r, st);
}

int nvnl_coulomb_list(const void* positions, int dtype, int64_t n_atoms, const void* cell, int32_t n_systems,
                      const int32_t* batch_idx, const double* charges, double cutoff, double alpha, const int32_t* neighbor_ptr,
                      const int32_t* neighbors, const int32_t* shifts, int32_t max_neighbors, int32_t fill_value,
                      double* energies, double* forces, void* stream) {
    if (!positions || !cell || !charges || !energies || !forces || n_atoms <= 0 || n_systems <= 0)
        return fail(-1, "nvnl_coulomb_list: bad arguments");
    if (!neighbor_ptr && max_neighbors < 0) return fail(-1, "nvnl_coulomb_list: matrix format needs max_neighbors >= 0");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(energies, 0, sizeof(double) * (size_t)n_atoms, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(forces, 0, sizeof(double) * 3 * (size_t)n_atoms, st);
    if (e != cudaSuccess) return fail(-2, "nvnl_coulomb_list: memset", e);
    if ((!neighbor_ptr && max_neighbors == 0) || !neighbors || !shifts) return 0;
    long long blocks = (n_atoms * 32 + 255) / 256;
    if (blocks > (long long)sm_count() * 32) blocks = (long long)sm_count() * 32;
    if (dtype == NVNL_F32) {
        k_coulomb_list<float><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const float*>(positions), charges,
                                                                 static_cast<const float*>(cell), batch_idx, n_systems, n_atoms,
                                                                 neighbor_ptr, neighbors, shifts, max_neighbors, fill_value,
                                                                 cutoff, alpha, energies, forces);
    } else if (dtype == NVNL_F64) {
        k_coulomb_list<double><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const double*>(positions), charges,
                                                                  static_cast<const double*>(cell), batch_idx, n_systems, n_atoms,
                                                                  neighbor_ptr, neighbors, shifts, max_neighbors, fill_value,
                                                                  cutoff, alpha, energies, forces);
    } else {
        return fail(-1, "nvnl_coulomb_list: unsupported dtype");
    }
    NVNL_CHECK_LAUNCH("k_coulomb_list");
    return 0;
}

int nvnl_fill_matrix(void* workspace, int dtype, int64_t n_atoms, int32_t n_systems, const int32_t* batch_idx,
                     double cutoff_sq, int half_fill, int fma, int32_t* neighbor_matrix, int32_t* neighbor_matrix_shifts,
                     int32_t* num_neighbors, int32_t max_neighbors, int32_t fill_value, int32_t pad_rows, void* stream) {
    if (!workspace || !num_neighbors || n_atoms <= 0 || max_neighbors < 0)
        return fail(-1, "nvnl_fill_matrix: bad arguments");
    if (max_neighbors > 0 && (!neighbor_matrix || !neighbor_matrix_shifts))
        return fail(-1, "nvnl_fill_matrix: null output");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    if (dtype == NVNL_F32 || dtype == NVNL_F64) {
        const int rc = launch_query_reset(ws, mk_layout(n_atoms, n_systems, rec_bytes(dtype)), n_atoms, 0, st);
        if (rc) return rc;
    }
    if (dtype == NVNL_F32) {
        SweepArgs<float> a = base_args<float>(ws, n_atoms, n_systems, batch_idx, cutoff_sq);
        a.neighbor_matrix = neighbor_matrix; a.out_shifts = neighbor_matrix_shifts; a.num_neighbors = num_neighbors;
        a.max_neighbors = max_neighbors; a.fill_value = fill_value; a.queue = 2; a.pad = pad_rows ? 1 : 0;
        return launch_sweep<float, MODE_FILL_MATRIX>(a, half_fill, fma, st);
    }
    if (dtype == NVNL_F64) {
        SweepArgs<double> a = base_args<double>(ws, n_atoms, n_systems, batch_idx, cutoff_sq);
        a.neighbor_matrix = neighbor_matrix; a.out_shifts = neighbor_matrix_shifts; a.num_neighbors = num_neighbors;
        a.max_neighbors = max_neighbors; a.fill_value = fill_value; a.queue = 2; a.pad = pad_rows ? 1 : 0;
        return launch_sweep<double, MODE_FILL_MATRIX>(a, half_fill, fma, st);
    }
    return fail(-1, "nvnl_fill_matrix: unsupported dtype");
}

int nvnl_get_grid(void* workspace, int dtype, int64_t n_atoms, int32_t n_systems, int32_t* cells_per_dimension,
                  int32_t* neighbor_search_radius, void* stream) {
    if (!workspace || n_systems <= 0) return fail(-1, "nvnl_get_grid: bad arguments");
    const WsLayout L = mk_layout(n_atoms, n_systems, rec_bytes(dtype));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    k_get_grid<<<(n_systems + 127) / 128, 128, 0, st>>>(static_cast<const unsigned char*>(workspace), L, n_systems,
                                                       cells_per_dimension, neighbor_search_radius);
    NVNL_CHECK_LAUNCH("k_get_grid");
    return 0;
}

int nvnl_export_cache(void* workspace, int dtype, int64_t n_atoms, int32_t n_systems, const int32_t* batch_idx,
                      int32_t* cells_per_dimension, int32_t* neighbor_search_radius, int32_t* atom_periodic_shifts,
                      int32_t* atom_to_cell_mapping, int32_t* atoms_per_cell_count, int32_t* cell_atom_start_indices,
                      int64_t cache_cells, int32_t* cell_atom_list, void* stream) {
    if (!workspace || n_atoms <= 0 || n_systems <= 0 || cache_cells < 0) return fail(-1, "nvnl_export_cache: bad arguments");
    const WsLayout L = mk_layout(n_atoms, n_systems, rec_bytes(dtype));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const unsigned char* ws = static_cast<const unsigned char*>(workspace);
    long long work = n_atoms > cache_cells ? n_atoms : cache_cells;
    unsigned blocks = (unsigned)((work + 255) / 256);
    if (blocks > (unsigned)sm_count() * 16) blocks = (unsigned)sm_count() * 16;
    if (dtype == NVNL_F32)
        k_export_cache<float><<<blocks, 256, 0, st>>>(ws, L, n_atoms, n_systems, batch_idx, cells_per_dimension,
                                                     neighbor_search_radius, atom_periodic_shifts, atom_to_cell_mapping,
                                                     atoms_per_cell_count, cell_atom_start_indices, cache_cells, cell_atom_list);
    else if (dtype == NVNL_F64)
        k_export_cache<double><<<blocks, 256, 0, st>>>(ws, L, n_atoms, n_systems, batch_idx, cells_per_dimension,
                                                      neighbor_search_radius, atom_periodic_shifts, atom_to_cell_mapping,
                                                      atoms_per_cell_count, cell_atom_start_indices, cache_cells, cell_atom_list);
    else
        return fail(-1, "nvnl_export_cache: unsupported dtype");
    NVNL_CHECK_LAUNCH("k_export_cache");
    return 0;
}

int nvnl_import_cache(const void* positions, int dtype, int64_t n_atoms, const void* cell, const uint8_t* pbc,
                      const int32_t* batch_idx, int32_t n_systems, double cutoff, const int32_t* cells_per_dimension,
                      const int32_t* neighbor_search_radius, const int32_t* atom_periodic_shifts,
                      const int32_t* atom_to_cell_mapping, const int32_t* atoms_per_cell_count,
                      const int32_t* cell_atom_start_indices, int64_t cache_cells, const int32_t* cell_atom_list,
                      void* workspace, size_t workspace_bytes, void* stream) {
    if (n_atoms <= 0 || n_systems <= 0) return fail(-1, "nvnl_import_cache: n_atoms and n_systems must be positive");
    if (!(cutoff > 0.0)) return fail(-1, "nvnl_import_cache: cutoff must be positive");
    if (!positions || !cell || !pbc || !workspace || !cells_per_dimension || !atom_periodic_shifts || !atom_to_cell_mapping ||
        !atoms_per_cell_count || !cell_atom_start_indices || !cell_atom_list)
        return fail(-1, "nvnl_import_cache: null pointer");
    if (n_systems > 1 && !batch_idx) return fail(-1, "nvnl_import_cache: batch_idx is required for n_systems > 1");
    if (cache_cells < n_systems) return fail(-1, "nvnl_import_cache: the cache holds fewer cells than systems");
    if (dtype != NVNL_F32 && dtype != NVNL_F64) return fail(-1, "nvnl_import_cache: unsupported dtype");
    if (workspace_bytes < nvnl_workspace_bytes(n_atoms, n_systems, dtype)) return fail(-1, "nvnl_import_cache: workspace too small");
    if (reinterpret_cast<uintptr_t>(workspace) % 256) return fail(-1, "nvnl_import_cache: workspace must be 256-byte aligned");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    if (dtype == NVNL_F32)
        return import_t<float>(static_cast<const float*>(positions), n_atoms, static_cast<const float*>(cell), pbc, batch_idx,
                               n_systems, cutoff, cells_per_dimension, neighbor_search_radius, atom_periodic_shifts,
                               atom_to_cell_mapping, atoms_per_cell_count, cell_atom_start_indices, cache_cells, cell_atom_list, ws, st);
    return import_t<double>(static_cast<const double*>(positions), n_atoms, static_cast<const double*>(cell), pbc, batch_idx,
                            n_systems, cutoff, cells_per_dimension, neighbor_search_radius, atom_periodic_shifts,
                            atom_to_cell_mapping, atoms_per_cell_count, cell_atom_start_indices, cache_cells, cell_atom_list, ws, st);
}

int nvnl_refresh_positions(void* workspace, int dtype, int64_t n_atoms, int32_t n_systems, const void* positions,
                           void* stream) {
    if (!workspace || !positions || n_atoms <= 0) return fail(-1, "nvnl_refresh_positions: bad arguments");
    const WsLayout L = mk_layout(n_atoms, n_systems, rec_bytes(dtype));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    const unsigned blocks = (unsigned)((n_atoms + 255) / 256);
    if (dtype == NVNL_F32)
        k_refresh_positions<float><<<blocks, 256, 0, st>>>(ws, L, n_atoms, static_cast<const float*>(positions));
    else if (dtype == NVNL_F64)
        k_refresh_positions<double><<<blocks, 256, 0, st>>>(ws, L, n_atoms, static_cast<const double*>(positions));
    else
        return fail(-1, "nvnl_refresh_positions: unsupported dtype");
    NVNL_CHECK_LAUNCH("k_refresh_positions");
    return 0;
}

int nvnl_cells_changed(void* workspace, int dtype, int64_t n_atoms, int32_t n_systems, const void* positions,
                       const int32_t* batch_idx, int32_t* flag, void* stream) {
    if (!workspace || !positions || !flag || n_atoms <= 0) return fail(-1, "nvnl_cells_changed: bad arguments");
    const WsLayout L = mk_layout(n_atoms, n_systems, rec_bytes(dtype));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const unsigned char* ws = static_cast<const unsigned char*>(workspace);
    cudaError_t e = cudaMemsetAsync(flag, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return fail(-2, "nvnl_cells_changed: memset", e);
    const unsigned blocks = (unsigned)((n_atoms + 255) / 256);
    if (dtype == NVNL_F32)
        k_cells_changed<float><<<blocks, 256, 0, st>>>(ws, L, n_atoms, n_systems, static_cast<const float*>(positions), batch_idx, flag);
    else if (dtype == NVNL_F64)
        k_cells_changed<double><<<blocks, 256, 0, st>>>(ws, L, n_atoms, n_systems, static_cast<const double*>(positions), batch_idx, flag);
    else
        return fail(-1, "nvnl_cells_changed: unsupported dtype");
    NVNL_CHECK_LAUNCH("k_cells_changed");
    return 0;
}

int nvnl_cells_changed_cache(const void* positions, int dtype, int64_t n_atoms, const void* cell, const uint8_t* pbc,
                             const int32_t* batch_idx, int32_t n_systems, const int32_t* cells_per_dimension,
                             const int32_t* atom_to_cell_mapping, int32_t* flag, void* stream) {
    if (!positions || !cell || !pbc || !cells_per_dimension || !atom_to_cell_mapping || !flag || n_atoms <= 0 || n_systems <= 0)
        return fail(-1, "nvnl_cells_changed_cache: bad arguments");
    if (n_systems > 1 && !batch_idx) return fail(-1, "nvnl_cells_changed_cache: batch_idx is required for n_systems > 1");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(flag, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return fail(-2, "nvnl_cells_changed_cache: memset", e);
    const unsigned blocks = (unsigned)((n_atoms + 255) / 256);
    if (dtype == NVNL_F32)
        k_cells_changed_cache<float><<<blocks, 256, 0, st>>>(n_atoms, n_systems, static_cast<const float*>(positions),
                                                            static_cast<const float*>(cell), pbc, batch_idx, cells_per_dimension,
                                                            atom_to_cell_mapping, flag);
    else if (dtype == NVNL_F64)
        k_cells_changed_cache<double><<<blocks, 256, 0, st>>>(n_atoms, n_systems, static_cast<const double*>(positions),
                                                             static_cast<const double*>(cell), pbc, batch_idx, cells_per_dimension,
                                                             atom_to_cell_mapping, flag);
    else
        return fail(-1, "nvnl_cells_changed_cache: unsupported dtype");
    NVNL_CHECK_LAUNCH("k_cells_changed_cache");
    return 0;
}

int nvnl_moved_beyond(const void* reference_positions, const void* current_positions, int dtype, int64_t n_atoms,
                      double threshold, int32_t* flag, void* stream) {
    if (!reference_positions || !current_positions || !flag || n_atoms <= 0) return fail(-1, "nvnl_moved_beyond: bad arguments");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(flag, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return fail(-2, "nvnl_moved_beyond: memset", e);
    const unsigned blocks = (unsigned)((n_atoms + 255) / 256);
    if (dtype == NVNL_F32)
        k_moved_beyond<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(reference_positions),
                                                     static_cast<const float*>(current_positions), n_atoms, (float)threshold, flag);
    else if (dtype == NVNL_F64)
        k_moved_beyond<double><<<blocks, 256, 0, st>>>(static_cast<const double*>(reference_positions),
                                                      static_cast<const double*>(current_positions), n_atoms, threshold, flag);
    else
        return fail(-1, "nvnl_moved_beyond: unsupported dtype");
    NVNL_CHECK_LAUNCH("k_moved_beyond");
    return 0;
}

int nvnl_pack_shifts(const int32_t* shifts, int64_t n_pairs, uint8_t* packed, int32_t* bad_flag, void* stream) {
    if (n_pairs < 0) return fail(-1, "nvnl_pack_shifts: negative pair count");
    if (n_pairs == 0) return 0;
    if (!shifts || !packed) return fail(-1, "nvnl_pack_shifts: null pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    long long blocks = (n_pairs + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    k_pack_shifts<<<(unsigned)blocks, 256, 0, st>>>(shifts, n_pairs, packed, bad_flag);
    NVNL_CHECK_LAUNCH("k_pack_shifts");
    return 0;
}

int nvnl_pack_shifts_word(const int32_t* shifts, int64_t n_pairs, int32_t* targets, int32_t* bad_flag, void* stream) {
    if (n_pairs < 0) return fail(-1, "nvnl_pack_shifts_word: negative pair count");
    if (n_pairs == 0) return 0;
    if (!shifts || !targets) return fail(-1, "nvnl_pack_shifts_word: null pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    long long blocks = (n_pairs + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    k_pack_shifts_word<<<(unsigned)blocks, 256, 0, st>>>(shifts, n_pairs, targets, bad_flag);
    NVNL_CHECK_LAUNCH("k_pack_shifts_word");
    return 0;
}

int nvnl_expand_gathered(const int32_t* neighbor_ptr, int64_t n_atoms, int64_t atom_lo, int64_t atom_hi,
                         const uint8_t* packed_shifts, int32_t* out_i, int32_t* shifts, void* stream) {
    if (n_atoms < 0 || atom_lo < 0 || atom_hi < atom_lo || atom_hi > n_atoms) return fail(-1, "nvnl_expand_gathered: bad atom range");
    if (n_atoms == 0) return 0;
    if (!neighbor_ptr || !packed_shifts || !out_i || !shifts) return fail(-1, "nvnl_expand_gathered: null pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    long long blocks = (n_atoms + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    k_expand_gathered<<<(unsigned)blocks, 256, 0, st>>>(neighbor_ptr, n_atoms, atom_lo, atom_hi, packed_shifts, out_i, shifts);
    NVNL_CHECK_LAUNCH("k_expand_gathered");
    return 0;
}

static int expand_padded_launch(const char* who, const int32_t* neighbor_ptr, int64_t n_atoms, const ExpandRanks& R, int64_t pmax,
                                const int32_t* gathered_dst, const uint8_t* gathered_packed, int32_t* out_i, int32_t* out_j,
                                int32_t* shifts, void* stream) {
    if (n_atoms <= 0) return 0;
    if (!neighbor_ptr || !gathered_dst || !out_i || !out_j || !shifts) {
        char msg[96];
        snprintf(msg, sizeof(msg), "%s: null pointer", who);
        return fail(-1, msg);
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    long long blocks = (n_atoms + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (gathered_packed)
        k_expand_padded<false><<<(unsigned)blocks, 256, 0, st>>>(neighbor_ptr, n_atoms, R, pmax, gathered_dst, gathered_packed,
                                                                 out_i, out_j, shifts);
    else   // one-word exchange: the packed shifts ride in the top bits of gathered_dst
        k_expand_padded<true><<<(unsigned)blocks, 256, 0, st>>>(neighbor_ptr, n_atoms, R, pmax, gathered_dst, nullptr, out_i, out_j,
                                                                shifts);
    NVNL_CHECK_LAUNCH("k_expand_padded");
    return 0;
}

int nvnl_expand_padded(const int32_t* neighbor_ptr, int64_t n_atoms, int32_t world, int32_t rank, const int64_t* atom_bounds,
                       const int64_t* pair_bounds, int64_t pmax, const int32_t* gathered_dst, const uint8_t* gathered_packed,
                       int32_t* out_i, int32_t* out_j, int32_t* shifts, void* stream) {
    if (world < 1 || world > 16 || rank < 0 || rank >= world || !atom_bounds || !pair_bounds || pmax < 0)
        return fail(-1, "nvnl_expand_padded: bad rank layout (1 <= world <= 16)");
    if (n_atoms <= 0) return 0;
    ExpandRanks R;
    R.world = world; R.rank = rank;
    for (int g = 0; g <= 16; ++g) {
        R.atom_lo[g] = atom_bounds[g <= world ? g : world];
        R.atom_hi[g] = atom_bounds[g + 1 <= world ? g + 1 : world];
        R.pair_lo[g] = pair_bounds[g <= world ? g : world];
    }
    if (R.atom_lo[0] != 0 || R.atom_lo[world] != n_atoms) return fail(-1, "nvnl_expand_padded: atom bounds do not cover the atoms");
    return expand_padded_launch("nvnl_expand_padded", neighbor_ptr, n_atoms, R, pmax, gathered_dst, gathered_packed, out_i, out_j,
                                shifts, stream);
}

int nvnl_expand_padded_ranges(const int32_t* neighbor_ptr, int64_t n_atoms, int32_t world, int32_t rank, const int64_t* atom_begin,
                              const int64_t* atom_end, const int64_t* pair_begin, int64_t pmax, const int32_t* gathered_dst,
                              const uint8_t* gathered_packed, int32_t* out_i, int32_t* out_j, int32_t* shifts, void* stream) {
    if (world < 1 || world > 16 || rank < 0 || rank >= world || !atom_begin || !atom_end || !pair_begin || pmax < 0)
        return fail(-1, "nvnl_expand_padded_ranges: bad rank layout (1 <= world <= 16)");
    if (n_atoms <= 0) return 0;
    ExpandRanks R;
    R.world = world; R.rank = rank;
    long long prev = 0;
    for (int g = 0; g < world; ++g) {
        if (atom_begin[g] < prev || atom_end[g] < atom_begin[g] || atom_end[g] > n_atoms || pair_begin[g] < 0)
            return fail(-1, "nvnl_expand_padded_ranges: atom ranges must be ascending, disjoint and inside [0, n_atoms)");
        R.atom_lo[g] = atom_begin[g]; R.atom_hi[g] = atom_end[g]; R.pair_lo[g] = pair_begin[g];
        prev = atom_end[g];
    }
    for (int g = world; g <= 16; ++g) { R.atom_lo[g] = prev; R.atom_hi[g] = prev; R.pair_lo[g] = 0; }
    return expand_padded_launch("nvnl_expand_padded_ranges", neighbor_ptr, n_atoms, R, pmax, gathered_dst, gathered_packed, out_i,
                                out_j, shifts, stream);
}

}  // extern "C"
