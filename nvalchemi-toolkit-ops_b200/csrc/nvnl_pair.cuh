// nvnl_pair.cuh — SURVEY.md §8f rank 2: the stencil sweep fused with a pair consumer, so that the neighbor list is never
// written (it is 20 B/pair of output traffic, the dominant cost of the list path).
//
// Consumer: real-space Coulomb / Ewald energies and forces, the reference's
//   nvalchemiops/interactions/electrostatics/coulomb.py:206-292 (_coulomb_energy_forces_kernel, list format),
//   :352-428 (matrix format), :1540-1700 (coulomb_energy_forces) —
//   r_ij = r_i - r_j - s·cell,  r = |r_ij|,  skipped when r >= cutoff or r < 1e-10,
//   E_i += 1/2 q_i q_j erfc(alpha r) / r,
//   f_ij = 1/2 q_i q_j [erfc(alpha r)/r^3 + 2 alpha/sqrt(pi) exp(-alpha^2 r^2)/r^2] r_ij,  F_i += f_ij,  F_j -= f_ij
// all in float64 (the reference converts its inputs, coulomb.py:1626-1628).
//
//   pair_consume4  epilogue of a k_rows<.., PAIR> trip: every lane walks the hit masks of the four targets, evaluates
//                  the pair term in fp64 from the staged records + staged charges, one warp reduction per target, one
//                  store per output — no list, no atomics (the full list is symmetric: the mirrored entry (j, i, -s)
//                  contributes -f_ji = f_ij to F_i, so F_i = 2 * sum over row i).
//   k_gather_q     charges in cell-sorted order (the producer stages them next to the records).
//   k_coulomb_list the reference's list / matrix consumer for an EXISTING neighbor list (any list, also half lists:
//                  F_j gets its reaction by fp64 atomics) — the unfused path, and the fallback of the fused op.
#pragma once

namespace nvnl {

constexpr double kTwoOverSqrtPi = 1.1283791670955126;   // coulomb.py:266

// One directed entry (i, j, s): d = r_i - r_j - s·cell.  Adds the entry's energy to e and its force on i to (fx, fy, fz).
// The reference's formulas (coulomb.py:247-276) with its erfc (nvalchemiops/math/math.py:52-93, wp_erfc: the Abramowitz &
// Stegun 7.1.26 polynomial, |error| <= 1.5e-7 — NOT the exact erfc, on purpose: the consumer has to reproduce the
// reference's numbers), arranged for fp64 throughput: ONE reciprocal square root, ONE division and ONE exp per entry
// (exp(-(alpha r)^2) serves the polynomial and the force term; 1/r and t = 1/(1 + p alpha r) share the division).  The
// skip test `r >= cutoff` is decided on r^2 away from the knife edge and with the correctly rounded sqrt within 1e-12 of
// it, so it is the reference's decision bit for bit.
__device__ __forceinline__ void pair_coulomb(double qiqj, double dx, double dy, double dz, double cutoff, double alpha,
                                             double& e, double& fx, double& fy, double& fz) {
    const double r2 = dx * dx + dy * dy + dz * dz;
    const double c2 = cutoff * cutoff;
    if (r2 > c2 * (1.0 + 1e-12) || r2 < 1e-20 * (1.0 - 1e-12)) return;
    if (r2 > c2 * (1.0 - 1e-12) || r2 < 1e-20 * (1.0 + 1e-12)) {
        const double r_exact = sqrt(r2);
        if (r_exact >= cutoff || r_exact < 1e-10) return;
    }
    const double pre = 0.5 * qiqj;
    double inv_r = rsqrt(r2);
    inv_r = inv_r * (1.5 - 0.5 * r2 * inv_r * inv_r);        // one Newton step: full fp64 accuracy
    const double r = r2 * inv_r;
    double fm;
    if (alpha > 0.0) {
        const double ar = alpha * r;
        const double u = 1.0 + 0.3275911 * ar;
        const double t = 1.0 / u;
        const double poly = t * (0.254829592 + t * (-0.284496736 + t * (1.421413741 + t * (-1.453152027 + t * 1.061405429))));
        const double exp_t = exp(-(ar * ar));
        const double erfc_t = poly * exp_t;                  // (ar >= 0: the reference's x >= 0 branch)
        e += pre * erfc_t * inv_r;
        fm = pre * inv_r * inv_r * (erfc_t * inv_r + kTwoOverSqrtPi * alpha * exp_t);
    } else {
        e += pre * inv_r;
        fm = pre * inv_r * inv_r * inv_r;
    }
    fx += fm * dx;
    fy += fm * dy;
    fz += fm * dz;
}

__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

// Epilogue of a PAIR trip.  m = hit masks of the four targets (transposed: bit v of lane l = candidate 32 v + l), S64 = the
// fp64 lattice vectors of the tile's shift segments, q_addr = the staged charges (one double per tile slot).
template <int W>
__device__ __forceinline__ void pair_consume4(const RowsArgs& a, const RowsDesc& d, const double (*S64)[3], uint32_t tile_addr,
                                              uint32_t q_addr, unsigned (&m)[W][4], int s0, int nt, int lane) {
    constexpr uint32_t RS = sizeof(Rec<float>);
    const bool shifted = d.shifted != 0;
    const double cutoff = a.pair_cutoff, alpha = a.pair_alpha;
#pragma unroll 1
    for (int k = 0; k < nt; ++k) {
        float xf, yf, zf;
        int i;
        lds_rec(tile_addr + (uint32_t)(s0 + k) * RS, xf, yf, zf, i);
        const double xi = (double)xf, yi = (double)yf, zi = (double)zf;
        const double qi = lds_f64(q_addr + (uint32_t)(s0 + k) * 8u);
        double e = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            unsigned mk = m[w][0];
            while (mk) {
                const int b = __ffs((int)mk) - 1;
                mk &= mk - 1u;
                const int v = (w << 5) + b;
                const uint32_t slot = ((uint32_t)v << 5) + (uint32_t)lane;
                float xj, yj, zj;
                int j;
                lds_rec(tile_addr + slot * RS, xj, yj, zj, j);
                const double qj = lds_f64(q_addr + slot * 8u);
                double dx = xi - (double)xj, dy = yi - (double)yj, dz = zi - (double)zj;
                if (shifted) {
                    const int sg = d.vc_seg[v];
                    dx -= S64[sg][0];
                    dy -= S64[sg][1];
                    dz -= S64[sg][2];
                }
                pair_coulomb(qi * qj, dx, dy, dz, cutoff, alpha, e, fx, fy, fz);
            }
            // rotate the masks so that the next target's are at index 0 (a dynamic index would move them to local memory)
            m[w][0] = m[w][1]; m[w][1] = m[w][2]; m[w][2] = m[w][3];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            e += __shfl_xor_sync(0xffffffffu, e, o);
            fx += __shfl_xor_sync(0xffffffffu, fx, o);
            fy += __shfl_xor_sync(0xffffffffu, fy, o);
            fz += __shfl_xor_sync(0xffffffffu, fz, o);
        }
        if (lane == 0) {
            a.pair_energies[i] = e;
            // the mirrored entries (j, i, -s) of the symmetric list give the same force on i once more
            a.pair_forces[3 * (size_t)i] = 2.0 * fx;
            a.pair_forces[3 * (size_t)i + 1] = 2.0 * fy;
            a.pair_forces[3 * (size_t)i + 2] = 2.0 * fz;
        }
    }
}

// charges gathered into cell-sorted order (what the sweep's producer stages next to the records)
__global__ void k_gather_q(const unsigned char* __restrict__ ws, WsLayout L, long long n, const double* __restrict__ charges,
                           double* __restrict__ q_sorted) {
    pdl_enter();
    const Rec<float>* sorted = reinterpret_cast<const Rec<float>*>(ws + L.sorted);
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        const int j = sorted[k].j;
        q_sorted[k] = (j >= 0 && j < n) ? charges[j] : 0.0;
    }
}

// The consumer over an existing neighbor list (COO with neighbor_ptr, or padded matrix): one warp per atom, lanes over
// the atom's entries.  Reference semantics for ANY list: F_i += f_ij and F_j -= f_ij (fp64 atomics for the reaction).
// energies / forces must be zero on entry.
template <typename T>
__global__ void __launch_bounds__(256) k_coulomb_list(const T* __restrict__ pos, const double* __restrict__ charges,
                                                      const T* __restrict__ cell, const int* __restrict__ batch_idx,
                                                      int num_systems, long long n, const int* __restrict__ neighbor_ptr,
                                                      const int* __restrict__ idx_j, const int* __restrict__ shifts,
                                                      int max_neighbors, int fill_value, double cutoff, double alpha,
                                                      double* __restrict__ energies, double* __restrict__ forces) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long i = warp; i < n; i += nwarps) {
        int s = 0;
        if (batch_idx && num_systems > 1) {
            s = batch_idx[i];
            s = s < 0 ? 0 : (s >= num_systems ? num_systems - 1 : s);
        }
        double c[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) c[k] = (double)cell[(size_t)s * 9 + k];
        const double xi = (double)pos[3 * i], yi = (double)pos[3 * i + 1], zi = (double)pos[3 * i + 2];
        const double qi = charges[i];
        long long e0, e1;
        if (neighbor_ptr) { e0 = neighbor_ptr[i]; e1 = neighbor_ptr[i + 1]; }
        else { e0 = i * (long long)max_neighbors; e1 = e0 + max_neighbors; }
        double e = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
        for (long long p = e0 + lane; p < e1; p += 32) {
            const int j = idx_j[p];
            if (j < 0 || j >= n || (!neighbor_ptr && j >= fill_value)) continue;   // matrix padding (coulomb.py:386-388)
            const double s0 = (double)shifts[3 * p], s1 = (double)shifts[3 * p + 1], s2 = (double)shifts[3 * p + 2];
            // shift_vec = cell^T · s  (coulomb.py:243): component d = sum_k s_k cell[k][d]
            const double dx = xi - (double)pos[3 * (size_t)j] - (s0 * c[0] + s1 * c[3] + s2 * c[6]);
            const double dy = yi - (double)pos[3 * (size_t)j + 1] - (s0 * c[1] + s1 * c[4] + s2 * c[7]);
            const double dz = zi - (double)pos[3 * (size_t)j + 2] - (s0 * c[2] + s1 * c[5] + s2 * c[8]);
            double pe = 0.0, px = 0.0, py = 0.0, pz = 0.0;
            pair_coulomb(qi * charges[j], dx, dy, dz, cutoff, alpha, pe, px, py, pz);
            if (px != 0.0 || py != 0.0 || pz != 0.0 || pe != 0.0) {
                e += pe; fx += px; fy += py; fz += pz;
                atomicAdd(&forces[3 * (size_t)j], -px);
                atomicAdd(&forces[3 * (size_t)j + 1], -py);
                atomicAdd(&forces[3 * (size_t)j + 2], -pz);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            e += __shfl_xor_sync(0xffffffffu, e, o);
            fx += __shfl_xor_sync(0xffffffffu, fx, o);
            fy += __shfl_xor_sync(0xffffffffu, fy, o);
            fz += __shfl_xor_sync(0xffffffffu, fz, o);
        }
        if (lane == 0) {
            atomicAdd(&energies[i], e);
            atomicAdd(&forces[3 * (size_t)i], fx);
            atomicAdd(&forces[3 * (size_t)i + 1], fy);
            atomicAdd(&forces[3 * (size_t)i + 2], fz);
        }
    }
}

}  // namespace nvnl
