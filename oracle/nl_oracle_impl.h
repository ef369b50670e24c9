/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported / linked by the product path.
 *
 * CPU restatement of the reference's cell-list neighbor-list algorithm
 * (NVIDIA/nvalchemi-toolkit-ops @ v0.2.0, Warp DSL kernels), written from the
 * behaviour of the reference kernels.  Every function names the reference
 * file:line it follows (paths relative to the reference repo root).
 *
 * This header is a "template": it is included twice by nl_oracle.c with
 *   REAL = float  / SUF = f32   and   REAL = double / SUF = f64.
 *
 * Parity status: the reference itself (Warp) cannot run in this sandbox, so the
 * restatement is pinned against the reference's own known-answer tests
 * (test/neighborlist/test_cell_list.py:391-419 etc., see tests/golden/) and
 * against an independent fp64 image-sum brute force.  The exact floating-point
 * contraction (FMA or not) of Warp's generated code for the distance test is
 * third-party (warp-lang 1.10.1, not vendored): both variants are implemented
 * (fma_mode) — "parity unpinned" for knife-edge pairs within ~1 ulp of rc^2.
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

/* nvalchemiops/math/math.py:40-48  (wpdivmod: floor division + non-negative remainder) */
#ifndef NLO_DIVMOD_DEFINED
#define NLO_DIVMOD_DEFINED
static inline void nlo_divmod(int a, int b, int *q, int *r) {
    int d = a / b; /* C: truncation toward zero == int(a / b) in Warp */
    int m = a % b; /* C remainder */
    if (m < 0) {
        d -= 1;
        m += b;
    }
    *q = d;
    *r = m;
}
static inline int nlo_imax(int a, int b) { return a > b ? a : b; }
static inline int nlo_imin(int a, int b) { return a < b ? a : b; }
#endif

/* 3x3 inverse as adjugate * (1/det) — what wp.inverse(mat33) computes (warp-lang mat.h;
 * third-party, restated).  m and out are row-major. */
static void FN(nlo_inverse3)(const REAL *m, REAL *out) {
    REAL a = m[0], b = m[1], c = m[2];
    REAL d = m[3], e = m[4], f = m[5];
    REAL g = m[6], h = m[7], i = m[8];
    REAL det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    if (det == (REAL)0) {
        for (int k = 0; k < 9; ++k) out[k] = (REAL)0;
        return;
    }
    REAL r = (REAL)1 / det;
    out[0] = (e * i - f * h) * r;
    out[1] = (c * h - b * i) * r;
    out[2] = (b * f - c * e) * r;
    out[3] = (f * g - d * i) * r;
    out[4] = (a * i - c * g) * r;
    out[5] = (c * d - a * f) * r;
    out[6] = (d * h - e * g) * r;
    out[7] = (b * g - a * h) * r;
    out[8] = (a * e - b * d) * r;
}

/* face distance along lattice direction d: 1 / || row d of transpose(inverse(cell)) ||
 * = 1 / || column d of inverse(cell) ||.   cell_list.py:73-78 */
static REAL FN(nlo_face_distance)(const REAL *inv, int d) {
    REAL x = inv[0 * 3 + d], y = inv[1 * 3 + d], z = inv[2 * 3 + d];
    REAL len = (REAL)SQRT(x * x + y * y + z * z);
    return (REAL)1 / len;
}

/* cell_list.py:35-99 (_estimate_cell_list_sizes) and batch_cell_list.py:35-99:
 * number of cells (capped at max_nbins by repeated halving) and neighbor search radius
 * (computed BEFORE halving).  Returns the number of cells. */
int FN(nlo_estimate_cell_list_sizes)(const REAL *cell, const unsigned char *pbc, REAL cell_size,
                                     int max_nbins, int *radius) {
    REAL inv[9];
    FN(nlo_inverse3)(cell, inv);
    int cpd[3];
    for (int i = 0; i < 3; ++i) {
        REAL face = FN(nlo_face_distance)(inv, i);
        cpd[i] = nlo_imax((int)(face / cell_size), 1);
        if (cpd[i] == 1 && !pbc[i])
            radius[i] = 0;
        else
            radius[i] = (int)CEIL(cell_size * (REAL)cpd[i] / face);
    }
    int total = cpd[0] * cpd[1] * cpd[2];
    while (total > max_nbins) {
        for (int i = 0; i < 3; ++i) cpd[i] = nlo_imax(cpd[i] / 2, 1);
        total = cpd[0] * cpd[1] * cpd[2];
    }
    return total;
}

/* cell_list.py:102-163 (_cell_list_construct_bin_size): cells per dimension, halved until
 * the product fits the allocation.  For the batch version (batch_cell_list.py:102-176) the
 * stop criterion is total * num_systems <= max_total_cells; pass num_systems (1 for single). */
void FN(nlo_construct_bin_size)(const REAL *cell, REAL cell_size, int max_cells_allowed, int num_systems,
                                int *cpd) {
    REAL inv[9];
    FN(nlo_inverse3)(cell, inv);
    for (int i = 0; i < 3; ++i) {
        REAL face = FN(nlo_face_distance)(inv, i);
        cpd[i] = nlo_imax((int)(face / cell_size), 1);
    }
    long total = (long)cpd[0] * cpd[1] * cpd[2];
    while (total * num_systems > max_cells_allowed) {
        for (int i = 0; i < 3; ++i) cpd[i] = nlo_imax(cpd[i] / 2, 1);
        total = (long)cpd[0] * cpd[1] * cpd[2];
    }
}

/* atom -> cell hash shared by count and bin passes.
 * cell_list.py:196-232 / 319-350; batch_cell_list.py:222-260.
 * frac = pos * inverse(cell)  (row vector times matrix). */
static void FN(nlo_atom_cell)(const REAL *p, const REAL *inv, const unsigned char *pbc, const int *cpd,
                              int *cc, int *shift) {
    /* warp mul(vec, mat): r = row0*v0; r += row1*v1; r += row2*v2 */
    REAL frac[3];
    for (int j = 0; j < 3; ++j) {
        REAL r = inv[0 * 3 + j] * p[0];
        r += inv[1 * 3 + j] * p[1];
        r += inv[2 * 3 + j] * p[2];
        frac[j] = r;
    }
    for (int d = 0; d < 3; ++d) {
        int c = (int)FLOOR(frac[d] * (REAL)cpd[d]);
        if (pbc[d]) {
            int q, r;
            nlo_divmod(c, cpd[d], &q, &r);
            shift[d] = q;
            cc[d] = r;
        } else {
            shift[d] = 0;
            cc[d] = nlo_imin(nlo_imax(c, 0), cpd[d] - 1);
        }
    }
}

/*
 * build_cell_list for a batch of S systems (S = 1 and batch_idx = NULL reproduces the
 * single-system op).   cell_list.py:725-889, batch_cell_list.py:739-912.
 *   cell [S,9], pbc [S,3], cpd out [S,3], atom_shifts [N,3], atom_cell [N,3],
 *   count [C], start [C], list [N];  C = max_total_cells (allocation size).
 */
void FN(nlo_build_cell_list)(const REAL *pos, int n, const REAL *cell, const unsigned char *pbc,
                             const int *batch_idx, int num_systems, REAL cutoff, int max_total_cells,
                             int *cpd, int *atom_shifts, int *atom_cell, int *count, int *start,
                             int *list) {
    REAL *inv = (REAL *)malloc(sizeof(REAL) * 9 * (size_t)num_systems);
    int *cell_off = (int *)malloc(sizeof(int) * ((size_t)num_systems + 1));
    cell_off[0] = 0;
    for (int s = 0; s < num_systems; ++s) {
        FN(nlo_construct_bin_size)(cell + 9 * s, cutoff, max_total_cells, num_systems, cpd + 3 * s);
        FN(nlo_inverse3)(cell + 9 * s, inv + 9 * s);
        cell_off[s + 1] = cell_off[s] + cpd[3 * s] * cpd[3 * s + 1] * cpd[3 * s + 2];
    }
    for (int c = 0; c < max_total_cells; ++c) count[c] = 0;
    /* pass 1: count  (cell_list.py:166-240) */
    for (int i = 0; i < n; ++i) {
        int s = batch_idx ? batch_idx[i] : 0;
        int cc[3];
        FN(nlo_atom_cell)(pos + 3 * (size_t)i, inv + 9 * s, pbc + 3 * s, cpd + 3 * s, cc,
                          atom_shifts + 3 * (size_t)i);
        int lin = cell_off[s] + cc[0] + cpd[3 * s] * (cc[1] + cpd[3 * s + 1] * cc[2]);
        count[lin] += 1;
    }
    /* exclusive scan  (cell_list.py:869-871) */
    start[0] = 0;
    for (int c = 1; c < max_total_cells; ++c) start[c] = start[c - 1] + count[c - 1];
    for (int c = 0; c < max_total_cells; ++c) count[c] = 0;
    /* pass 2: bin  (cell_list.py:279-369) */
    for (int i = 0; i < n; ++i) {
        int s = batch_idx ? batch_idx[i] : 0;
        int sh[3];
        int *cc = atom_cell + 3 * (size_t)i;
        FN(nlo_atom_cell)(pos + 3 * (size_t)i, inv + 9 * s, pbc + 3 * s, cpd + 3 * s, cc, sh);
        int lin = cell_off[s] + cc[0] + cpd[3 * s] * (cc[1] + cpd[3 * s + 1] * cc[2]);
        int slot = count[lin]++;
        list[start[lin] + slot] = i;
    }
    free(inv);
    free(cell_off);
}

/* neighbor_utils.py:106-147 (_update_neighbor_matrix_pbc): slot = atomic_add(num[i]); store
 * only when slot < max_neighbors; counter keeps counting.  Mirror (j, i, -s) unless half_fill. */
static inline void FN(nlo_insert)(int i, int j, int sx, int sy, int sz, int *nm, int *nm_shifts, int *num,
                                  int max_neighbors, int half_fill) {
    int p = __atomic_fetch_add(&num[i], 1, __ATOMIC_RELAXED);
    if (p < max_neighbors) {
        nm[(size_t)i * max_neighbors + p] = j;
        int *s = nm_shifts + ((size_t)i * max_neighbors + p) * 3;
        s[0] = sx;
        s[1] = sy;
        s[2] = sz;
    }
    if (!half_fill) {
        p = __atomic_fetch_add(&num[j], 1, __ATOMIC_RELAXED);
        if (p < max_neighbors) {
            nm[(size_t)j * max_neighbors + p] = i;
            int *s = nm_shifts + ((size_t)j * max_neighbors + p) * 3;
            s[0] = -sx;
            s[1] = -sy;
            s[2] = -sz;
        }
    }
}

/* The distance predicate of cell_list.py:531-545:
 *   cartesian_shift = (sx,sy,sz) * cell      (row vector times matrix)
 *   dr = neighbor_pos - central_pos + cartesian_shift
 *   dot(dr,dr) < cutoff*cutoff               (strict)
 * fma_mode 0: every multiply and add rounded separately.
 * fma_mode 1: the contraction nvcc/NVRTC (fmad=true, Warp's default fuse_fp) produces for
 *             these expressions: mul, then fma chains.  */
static inline int FN(nlo_within)(const REAL *pi, const REAL *pj, const REAL *cellm, int sx, int sy, int sz,
                                 REAL cutoff_sq, int fma_mode) {
    REAL fs0 = (REAL)sx, fs1 = (REAL)sy, fs2 = (REAL)sz;
    REAL cs[3], dr[3];
    for (int k = 0; k < 3; ++k) {
        if (fma_mode) {
            REAL r = cellm[0 * 3 + k] * fs0;
            r = FMA(cellm[1 * 3 + k], fs1, r);
            r = FMA(cellm[2 * 3 + k], fs2, r);
            cs[k] = r;
        } else {
            REAL r = cellm[0 * 3 + k] * fs0;
            r = r + cellm[1 * 3 + k] * fs1;
            r = r + cellm[2 * 3 + k] * fs2;
            cs[k] = r;
        }
        dr[k] = (pj[k] - pi[k]) + cs[k];
    }
    REAL d2;
    if (fma_mode) {
        d2 = dr[0] * dr[0];
        d2 = FMA(dr[1], dr[1], d2);
        d2 = FMA(dr[2], dr[2], d2);
    } else {
        d2 = dr[0] * dr[0] + dr[1] * dr[1];
        d2 = d2 + dr[2] * dr[2];
    }
    return d2 < cutoff_sq;
}

/*
 * query_cell_list for S systems (S = 1, batch_idx = NULL: single).
 * cell_list.py:372-556 (loop order dx,dy,dz) and batch_cell_list.py:380-569 (dz,dy,dx) — the loop
 * order only changes slot order inside rows, which is unspecified.
 * nm must be pre-filled with fill_value, nm_shifts and num zeroed by the caller
 * (cell_list.py:1358-1373).
 */
typedef struct {
    const REAL *pos; int n; const REAL *cell; const unsigned char *pbc; const int *batch_idx;
    REAL cutoff_sq; const int *cpd; const int *radius; const int *atom_shifts; const int *atom_cell;
    const int *count; const int *start; const int *list; int *nm; int *nm_shifts; int *num;
    int max_neighbors; int half_fill; int fma_mode; const int *cell_off; int *next_chunk; int n_limit;
} FN(nlo_query_args);

static void *FN(nlo_query_worker)(void *vp) {
    FN(nlo_query_args) *a = (FN(nlo_query_args) *)vp;
    const int CH = 256;
    for (;;) {
        int lo = __atomic_fetch_add(a->next_chunk, CH, __ATOMIC_RELAXED);
        if (lo >= a->n_limit) break;
        int hi = nlo_imin(lo + CH, a->n_limit);
        for (int i = lo; i < hi; ++i) {
            const int s = a->batch_idx ? a->batch_idx[i] : 0;
            const REAL *cm = a->cell + 9 * s;
            const unsigned char *pb = a->pbc + 3 * s;
            const int *cp = a->cpd + 3 * s;
            const int *R = a->radius + 3 * s;
            const REAL *pi = a->pos + 3 * (size_t)i;
            const int *ci = a->atom_cell + 3 * (size_t)i;
            const int *shi = a->atom_shifts + 3 * (size_t)i;
            for (int dx = 0; dx <= R[0]; ++dx)
                for (int dy = -R[1]; dy <= R[1]; ++dy)
                    for (int dz = -R[2]; dz <= R[2]; ++dz) {
                        if (!(dx > 0 || (dx == 0 && dy > 0) || (dx == 0 && dy == 0 && dz >= 0))) continue;
                        int tx = ci[0] + dx, ty = ci[1] + dy, tz = ci[2] + dz;
                        if (!pb[0] && (tx < 0 || tx >= cp[0])) continue;
                        if (!pb[1] && (ty < 0 || ty >= cp[1])) continue;
                        if (!pb[2] && (tz < 0 || tz >= cp[2])) continue;
                        int csx, csy, csz, wx, wy, wz;
                        nlo_divmod(tx, cp[0], &csx, &wx);
                        nlo_divmod(ty, cp[1], &csy, &wy);
                        nlo_divmod(tz, cp[2], &csz, &wz);
                        int lin = a->cell_off[s] + wx + cp[0] * (wy + cp[1] * wz);
                        int c0 = a->start[lin], nc = a->count[lin];
                        for (int k = 0; k < nc; ++k) {
                            int j = a->list[c0 + k];
                            const int *shj = a->atom_shifts + 3 * (size_t)j;
                            int sx = pb[0] ? csx + shi[0] - shj[0] : 0;
                            int sy = pb[1] ? csy + shi[1] - shj[1] : 0;
                            int sz = pb[2] ? csz + shi[2] - shj[2] : 0;
                            if (dx == 0 && dy == 0 && dz == 0 && j <= i) continue;
                            if (FN(nlo_within)(pi, a->pos + 3 * (size_t)j, cm, sx, sy, sz, a->cutoff_sq,
                                               a->fma_mode))
                                FN(nlo_insert)(i, j, sx, sy, sz, a->nm, a->nm_shifts, a->num,
                                               a->max_neighbors, a->half_fill);
                        }
                    }
        }
    }
    return NULL;
}

/*
 * query_cell_list for S systems (S = 1, batch_idx = NULL: single).
 * cell_list.py:372-556 (loop order dx,dy,dz) and batch_cell_list.py:380-569 (dz,dy,dx) — the loop
 * order only changes slot order inside rows, which is unspecified.
 * nm must be pre-filled with fill_value, nm_shifts and num zeroed by the caller
 * (cell_list.py:1358-1373).  The reference runs one Warp thread per atom (serial on the Warp CPU
 * device); nthreads > 1 splits the atom range over pthreads (slot order then varies, like on GPU).
 * n_limit > 0 runs only the first n_limit per-atom threads (bounded benchmark sample; output then partial).
 */
void FN(nlo_query_cell_list)(const REAL *pos, int n, const REAL *cell, const unsigned char *pbc,
                             const int *batch_idx, int num_systems, REAL cutoff, const int *cpd,
                             const int *radius, const int *atom_shifts, const int *atom_cell,
                             const int *count, const int *start, const int *list, int *nm, int *nm_shifts,
                             int *num, int max_neighbors, int half_fill, int fma_mode, int nthreads, int n_limit) {
    int *cell_off = (int *)malloc(sizeof(int) * ((size_t)num_systems + 1));
    cell_off[0] = 0;
    for (int s = 0; s < num_systems; ++s)
        cell_off[s + 1] = cell_off[s] + cpd[3 * s] * cpd[3 * s + 1] * cpd[3 * s + 2];
    int next = 0;
    if (n_limit <= 0 || n_limit > n) n_limit = n; /* benchmark sampling: only atoms [0, n_limit) act as "thread i" */
    FN(nlo_query_args) a = {pos, n, cell, pbc, batch_idx,
                            cutoff * cutoff /* squared in kernel precision, cell_list.py:444 */,
                            cpd, radius, atom_shifts, atom_cell, count, start, list, nm, nm_shifts, num,
                            max_neighbors, half_fill, fma_mode, cell_off, &next, n_limit};
    if (nthreads <= 1) {
        FN(nlo_query_worker)(&a);
    } else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
        for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, FN(nlo_query_worker), &a);
        for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
        free(th);
    }
    free(cell_off);
}

/* naive.py:36-92 (_fill_naive_neighbor_matrix, no PBC): all pairs j > i, mirrored when
 * !half_fill (neighbor_utils.py:70-103).  cutoff_sq is passed in (computed in Python double and
 * cast by the caller, naive.py:290). */
void FN(nlo_naive_no_pbc)(const REAL *pos, int n, REAL cutoff_sq, int *nm, int *num, int max_neighbors,
                          int half_fill, int fma_mode) {
    for (int i = 0; i < n; ++i) {
        const REAL *pi = pos + 3 * (size_t)i;
        for (int j = i + 1; j < n; ++j) {
            const REAL *pj = pos + 3 * (size_t)j;
            REAL dx = pi[0] - pj[0], dy = pi[1] - pj[1], dz = pi[2] - pj[2];
            REAL d2;
            if (fma_mode) {
                d2 = dx * dx;
                d2 = FMA(dy, dy, d2);
                d2 = FMA(dz, dz, d2);
            } else {
                d2 = dx * dx + dy * dy;
                d2 = d2 + dz * dz;
            }
            if (d2 < cutoff_sq) {
                int p = num[i]++;
                if (p < max_neighbors) nm[(size_t)i * max_neighbors + p] = j;
                if (!half_fill) {
                    p = num[j]++;
                    if (p < max_neighbors) nm[(size_t)j * max_neighbors + p] = i;
                }
            }
        }
    }
}

/* Independent brute force (NOT a restatement of the reference): every image s in
 * [-K,K]^3 (K_d = 0 in non-periodic dims) of every ordered pair, evaluated with the same
 * predicate.  Emits (i, j, sx, sy, sz) records; returns the number of records (may exceed cap,
 * in which case only the first cap are stored). */
long FN(nlo_brute_force)(const REAL *pos, int n, const REAL *cell, const unsigned char *pbc, REAL cutoff,
                         const int *K, int fma_mode, int *out, long cap) {
    long cnt = 0;
    const REAL cutoff_sq = cutoff * cutoff;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
            for (int sx = -K[0]; sx <= K[0]; ++sx)
                for (int sy = -K[1]; sy <= K[1]; ++sy)
                    for (int sz = -K[2]; sz <= K[2]; ++sz) {
                        if (i == j && sx == 0 && sy == 0 && sz == 0) continue;
                        if ((!pbc[0] && sx) || (!pbc[1] && sy) || (!pbc[2] && sz)) continue;
                        if (FN(nlo_within)(pos + 3 * (size_t)i, pos + 3 * (size_t)j, cell, sx, sy, sz,
                                           cutoff_sq, fma_mode)) {
                            if (cnt < cap) {
                                int *o = out + 5 * cnt;
                                o[0] = i; o[1] = j; o[2] = sx; o[3] = sy; o[4] = sz;
                            }
                            ++cnt;
                        }
                    }
    return cnt;
}

#undef FN
#undef CAT
#undef CAT_
