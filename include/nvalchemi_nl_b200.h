/*
 * nvalchemi_nl_b200.h — C ABI of the B200-native cell-list neighbor-list library
 * (libnvalchemi_nl_b200.so, sm_100a).  Plain pointers and sizes only: no torch, no Warp.
 *
 * The reference (NVIDIA/nvalchemi-toolkit-ops v0.2.0) has no FFI: its "operator ABI" for this path
 * is four torch.library custom ops that mutate pre-allocated tensors
 *   nvalchemiops::build_cell_list        nvalchemiops/neighborlist/cell_list.py:725-889
 *   nvalchemiops::query_cell_list        nvalchemiops/neighborlist/cell_list.py:892-1034
 *   nvalchemiops::batch_build_cell_list  nvalchemiops/neighborlist/batch_cell_list.py:739-912
 *   nvalchemiops::batch_query_cell_list  nvalchemiops/neighborlist/batch_cell_list.py:915-1067
 * plus the COO conversion  neighbor_utils.py:362-441.  Each entry point below names the reference
 * interface it stands in for.  All pointers are DEVICE pointers unless stated otherwise; `stream`
 * is a cudaStream_t passed as void* (NULL = legacy default stream).  Every function returns 0 on
 * success and a negative code on failure (message via nvnl_last_error()).
 *
 * dtype: 0 = float32 positions/cell, 1 = float64.
 * A "workspace" is an opaque device buffer of nvnl_workspace_bytes() bytes (256-byte aligned)
 * that plays the role of the reference's 7-tensor cell-list cache (neighbor_utils.py:494-539).
 */
#ifndef NVALCHEMI_NL_B200_H
#define NVALCHEMI_NL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NVNL_ABI_VERSION 6
#define NVNL_F32 0
#define NVNL_F64 1

/* error bits reported by nvnl_status() */
#define NVNL_ERR_IMAGE_RANGE 1   /* atom more than 1e6 periodic images away, or search radius >= 64 cells */
#define NVNL_ERR_BAD_BATCH_IDX 2 /* batch_idx outside [0, num_systems) */
#define NVNL_ERR_SINGULAR_CELL 4 /* cell matrix not invertible */
#define NVNL_ERR_BAD_CACHE 8     /* nvnl_import_cache: cache tensors inconsistent (not produced by a build of this library) */

int nvnl_abi_version(void);
const char* nvnl_last_error(void);

/* Size of the opaque cell-list cache for n_atoms atoms in n_systems systems.
 * Stands in for estimate_cell_list_sizes + allocate_cell_list
 * (cell_list.py:639-722, neighbor_utils.py:494-539) — but needs no device->host sync: the number
 * of cells is bounded by n_atoms + n_systems by construction. */
size_t nvnl_workspace_bytes(int64_t n_atoms, int64_t n_systems, int dtype);

/* build_cell_list / batch_build_cell_list (cell_list.py:725-889, batch_cell_list.py:739-912):
 * grid selection, atom->cell hash, counting sort of positions into per-cell float4 runs.
 *   positions [n_atoms,3], cell [n_systems,3,3] (rows = lattice vectors), pbc [n_systems,3] (bytes),
 *   batch_idx [n_atoms] int32 or NULL (single system), batch_ptr [n_systems+1] int32 or NULL.
 *   max_cells: 0 = grid bounded only by #cells <= #atoms per system; > 0 = total cell capacity of the caller's
 *   reference-shaped cache (the length of atoms_per_cell_count from estimate_cell_list_sizes / allocate_cell_list,
 *   cell_list.py:639-722): every system's grid is halved until it has at most max_cells / n_systems cells, exactly
 *   the contract of _cell_list_construct_bin_size (cell_list.py:131-150, batch_cell_list.py:162-176). */
int nvnl_build(const void* positions, int dtype, int64_t n_atoms, const void* cell, const uint8_t* pbc,
               const int32_t* batch_idx, const int32_t* batch_ptr, int32_t n_systems, double cutoff, int64_t max_cells,
               void* workspace, size_t workspace_bytes, void* stream);

/* First half of the COO path: per-atom neighbor counts and their exclusive scan.
 * Replaces the num_neighbors side of query_cell_list and the cumsum of
 * get_neighbor_list_from_neighbor_matrix (neighbor_utils.py:432-435).
 *   cutoff_sq: the squared cutoff already rounded to the input precision
 *   num_neighbors [n_atoms] out, neighbor_ptr [n_atoms+1] out (may be NULL to skip the scan).
 *   fma: 1 = mul+fma-chain distance arithmetic (Warp/NVRTC default), 0 = separately rounded. */
int nvnl_count(void* workspace, int dtype, int64_t n_atoms, int32_t n_systems, const int32_t* batch_idx,
               double cutoff_sq, int half_fill, int fma, int32_t* num_neighbors, int32_t* neighbor_ptr,
               void* stream);

/* Device->host read of the control block (synchronizes `stream`): total number of directed pairs
 * found by the last nvnl_count, the largest per-atom count (what assert_max_neighbors checks,
 * neighbor_utils.py:352-359), number of cells, error bits, and whether any atom was outside the
 * primary periodic image (*unwrapped bit 0; bit 1 = some system searches more than one cell per side, i.e. periodic
 * shifts may exceed +-1), and whether any cell was left to the general kernel (*had_deferred bit 0; bit 1 = some
 * single-cell system needed the second single-sweep launch of nvnl_count_rows, its launch_hint bit 5) (both feed
 * nvnl_fill_coo's launch_hint), and whether nvnl_count_rows ran out of temporary row space (then the caller repeats the query with
 * nvnl_count / nvnl_fill_coo).  The one sync of the COO path (the reference has three). Host pointers. */
int nvnl_status(void* workspace, int dtype, int64_t n_atoms, int32_t n_systems, int64_t* total_pairs,
                int32_t* max_count, int32_t* total_cells, int32_t* error_bits, int32_t* unwrapped, int32_t* had_deferred,
                int32_t* rows_overflow, void* stream);

/* Second half of the COO path: writes edge_index [2,num_pairs] (row 0 = source atoms, sorted),
 * shifts [num_pairs,3].  Output identical in content to cell_list(..., return_neighbor_list=True)
 * (cell_list.py:1432-1441) without materialising the padded matrix.
 *   index_offset is added to every atom index written (rank-sharded batches, 0 otherwise).
 *   launch_hint: -1 = unknown (all kernel variants are launched, the idle ones retire at once); otherwise
 *   bit 0 = nvnl_status' `unwrapped`, bit 1 = its `had_deferred`: only the kernels with work are launched. */
int nvnl_fill_coo(void* workspace, int dtype, int64_t n_atoms, int32_t n_systems, const int32_t* batch_idx,
                  double cutoff_sq, int half_fill, int fma, const int32_t* neighbor_ptr, int32_t* edge_index,
                  int64_t num_pairs, int64_t row_stride, int32_t* shifts, int32_t index_offset, int32_t launch_hint,
                  void* stream);

/* Single-sweep COO path (fp32; same outputs as nvnl_count + nvnl_fill_coo, same reference interfaces:
 * query_cell_list cell_list.py:892-1034 + get_neighbor_list_from_neighbor_matrix neighbor_utils.py:362-441).
 * nvnl_count_rows evaluates every distance once and leaves each atom's neighbors as a compact row in a temporary
 * buffer inside the workspace (in sweep order); nvnl_fill_rows then streams edge_index / shifts in atom order.
 * Between the two: nvnl_status (sizes the outputs; launch_hint = bit 0 unwrapped | bit 1 had_deferred is REQUIRED
 * here; if rows_overflow is set the temporary buffer — 160 entries per atom — was too small and the query must be
 * repeated with nvnl_count / nvnl_fill_coo).  Inputs with atoms outside the primary periodic image are detected on
 * the device and served by the two-pass kernels inside these same calls.
 * prezero / prezero_ints (optional, 16-byte aligned): a buffer the sweep kernel zero-fills while it runs — the sweep is
 * issue-bound, the zero_() of the shifts output (cell_list.py:1358-1373) is pure HBM writes.  A caller that can guess
 * the pair count (e.g. from its previous query) passes its shifts buffer here and sets launch_hint bit 2 (value 4) of
 * nvnl_fill_rows: `shifts` is already zero — only rows of cells at a periodic boundary write their image shifts.
 * Workloads whose rows mostly carry shifts (small periodic boxes; decided on the device when the grid is built) are
 * not pre-zeroed whatever buffer is passed: nvnl_fill_rows / _speculative then write every shift component themselves.
 * Cells with more than 64 target atoms are swept in parts of 32 targets by several CTAs (device-side work list).
 * Atom indices must be below 2^27 on this path.
 *   launch_hint: -1 = launch every kernel variant (the ones without work retire at once); >= 0 = the launch_hint
 *   nvnl_status reported for an earlier query of this kind (bit 0: atoms outside the primary image, bit 1: cells left to
 *   the general kernel, bit 5 = value 32: single-cell periodic systems whose 27 images take the second launch with six
 *   mask words): variants that would find no work are not launched.  The caller compares with nvnl_status
 *   afterwards and repeats the call with -1 if the hint missed a bit. */
int nvnl_count_rows(void* workspace, int dtype, int64_t n_atoms, int32_t n_systems, const int32_t* batch_idx,
                    double cutoff_sq, int half_fill, int fma, int32_t* num_neighbors, int32_t* neighbor_ptr,
                    int32_t* prezero, int64_t prezero_ints, int32_t launch_hint, void* stream);
int nvnl_fill_rows(void* workspace, int dtype, int64_t n_atoms, int32_t n_systems, const int32_t* batch_idx,
                   double cutoff_sq, int half_fill, int fma, const int32_t* neighbor_ptr, int32_t* edge_index,
                   int64_t num_pairs, int64_t row_stride, int32_t* shifts, int32_t index_offset, int32_t launch_hint,
                   void* stream);

/* Speculative output launch (nvalchemiops_b200.config.speculative_fill, on by default):
 * the output kernel of the single-sweep path launched BEFORE the size sync, into buffers sized from a guess
 * (edge_buffer: 2 * capacity_pairs int32, shifts_zeroed: 3 * capacity_pairs int32, already zero — e.g. the prezero
 * buffer of nvnl_count_rows).  The kernel reads the pair count P on the device: if P <= capacity_pairs (and the query was
 * not served by the two-pass kernels) it writes edge_index as the [2,P] prefix of edge_buffer and the image shifts, else
 * nothing.  The caller then reads nvnl_status and either takes the prefixes (running nvnl_fill_rows with launch_hint
 * bits 2|3 = 12 set if had_deferred, which then only launches the general kernel) or repeats the fill the regular way. */
int nvnl_fill_rows_speculative(void* workspace, int dtype, int64_t n_atoms, int32_t n_systems, const int32_t* neighbor_ptr,
                               int32_t* edge_buffer, int64_t capacity_pairs, int32_t* shifts_zeroed, int32_t index_offset,
                               void* stream);

/* SURVEY.md §8f rank 2 — the stencil sweep fused with a pair consumer: the neighbor list is never written.
 * Consumer = the reference's real-space Coulomb / Ewald energies and forces
 * (nvalchemiops/interactions/electrostatics/coulomb.py:1540 coulomb_energy_forces; kernels :206-292 list format,
 * :352-428 matrix format): per directed entry (i, j, s) with r = |r_i - r_j - s·cell|, skipped when r >= cutoff or
 * r < 1e-10:  E_i += q_i q_j erfc(alpha r) / (2 r);  f = q_i q_j/2 [erfc(alpha r)/r^3 + 2 alpha/sqrt(pi) exp(-alpha^2 r^2)/r^2] r_ij;
 * F_i += f, F_j -= f.  alpha = 0: plain Coulomb.  Everything in float64, charges/energies/forces are double arrays.
 *
 * nvnl_coulomb_fused: after nvnl_build.  The entries are the ones nvnl_count_rows / the reference's cell_list would list
 * (fp32 predicate d^2 < cutoff_sq, full list); the pair terms are evaluated in fp64 from the staged records while the
 * sweep runs.  energies [n_atoms], forces [n_atoms,3] are written (not accumulated) for every atom the sweep handled.
 * The caller MUST read nvnl_status afterwards: if *unwrapped bit 0 or *had_deferred is set, some atoms were not handled
 * (atoms outside the primary image, stencils wider than one cell, over-full cells) and the result must be recomputed
 * with nvnl_count_rows/nvnl_fill_rows + nvnl_coulomb_list — the Python binding does exactly that.  fp32 positions only.
 *
 * nvnl_coulomb_list: the same consumer over an EXISTING neighbor list — COO (neighbor_ptr [n_atoms+1], neighbors [P] = row 1
 * of edge_index, shifts [P,3]) or, with neighbor_ptr == NULL, the padded matrix (neighbors [n_atoms,max_neighbors],
 * shifts [n_atoms,max_neighbors,3], entries >= fill_value are padding).  Any list (also half lists): the reaction on j
 * is accumulated with fp64 atomics, as the reference does.  cell [n_systems,3,3] in the dtype of positions. */
int nvnl_coulomb_fused(void* workspace, int dtype, int64_t n_atoms, int32_t n_systems, const int32_t* batch_idx,
                       double cutoff_sq, int fma, const double* charges, double cutoff, double alpha, double* energies,
                       double* forces, void* stream);
int nvnl_coulomb_list(const void* positions, int dtype, int64_t n_atoms, const void* cell, int32_t n_systems,
                      const int32_t* batch_idx, const double* charges, double cutoff, double alpha, const int32_t* neighbor_ptr,
                      const int32_t* neighbors, const int32_t* shifts, int32_t max_neighbors, int32_t fill_value,
                      double* energies, double* forces, void* stream);

/* query_cell_list / batch_query_cell_list (cell_list.py:892-1034, batch_cell_list.py:915-1067)
 * fused with the fill_()/zero_() of the outputs (cell_list.py:1358-1373): every slot of
 * neighbor_matrix [n_atoms,max_neighbors], neighbor_matrix_shifts [n_atoms,max_neighbors,3] and
 * num_neighbors [n_atoms] is written exactly once.  num_neighbors keeps counting past
 * max_neighbors (neighbor_utils.py:139-147).  pad_rows = 0 leaves the unused slots untouched — the bare
 * query_cell_list op, which does not reset its outputs either (cell_list.py:892-1034). */
int nvnl_fill_matrix(void* workspace, int dtype, int64_t n_atoms, int32_t n_systems, const int32_t* batch_idx,
                     double cutoff_sq, int half_fill, int fma, int32_t* neighbor_matrix,
                     int32_t* neighbor_matrix_shifts, int32_t* num_neighbors, int32_t max_neighbors,
                     int32_t fill_value, int32_t pad_rows, void* stream);

/* Introspection for tests / rebuild detection: copies the per-system grid (cells per dimension and
 * stencil radius, int32 [n_systems,3] each, device pointers, either may be NULL). */
int nvnl_get_grid(void* workspace, int dtype, int64_t n_atoms, int32_t n_systems, int32_t* cells_per_dimension,
                  int32_t* neighbor_search_radius, void* stream);

/* ---- split build/query workflow and rebuild detection ------------------------------------------ */

/* Fills reference-shaped cache tensors (neighbor_utils.py:494-539) from the workspace for callers that inspect them;
 * any pointer may be NULL.  Values describe THIS implementation's grid.  cache_cells = length of the two per-cell
 * arrays (entries past the number of cells are zeroed). */
int nvnl_export_cache(void* workspace, int dtype, int64_t n_atoms, int32_t n_systems, const int32_t* batch_idx,
                      int32_t* cells_per_dimension, int32_t* neighbor_search_radius, int32_t* atom_periodic_shifts,
                      int32_t* atom_to_cell_mapping, int32_t* atoms_per_cell_count, int32_t* cell_atom_start_indices,
                      int64_t cache_cells, int32_t* cell_atom_list, void* stream);

/* query_cell_list / batch_query_cell_list from the cache VALUES (cell_list.py:892-1034, batch_cell_list.py:915-1067):
 * rebuilds the workspace from the seven reference-shaped cache tensors nvnl_export_cache wrote (possibly cloned, moved
 * or reloaded since) and the CURRENT positions — the stale-cache MD query needs no hidden state.  `cutoff` is the
 * query cutoff (<= the build cutoff); the stencil radius used is max(cached radius, what `cutoff` needs on the cached
 * grid).  Inconsistent tensors set NVNL_ERR_BAD_CACHE (nvnl_status). */
int nvnl_import_cache(const void* positions, int dtype, int64_t n_atoms, const void* cell, const uint8_t* pbc,
                      const int32_t* batch_idx, int32_t n_systems, double cutoff, const int32_t* cells_per_dimension,
                      const int32_t* neighbor_search_radius, const int32_t* atom_periodic_shifts,
                      const int32_t* atom_to_cell_mapping, const int32_t* atoms_per_cell_count,
                      const int32_t* cell_atom_start_indices, int64_t cache_cells, const int32_t* cell_atom_list,
                      void* workspace, size_t workspace_bytes, void* stream);

/* query_cell_list with moved atoms (cell_list.py:1108-1192): re-gathers `positions` into the cell-sorted records
 * without re-binning, so the next nvnl_count / nvnl_fill_* evaluates distances with the current coordinates against
 * the cell assignment of the last nvnl_build (valid while no atom moved more than (cell width - cutoff)/2). */
int nvnl_refresh_positions(void* workspace, int dtype, int64_t n_atoms, int32_t n_systems, const void* positions,
                           void* stream);

/* cell_list_needs_rebuild (rebuild_detection.py:36-121, 258-383): *flag (device int32) = 1 if any atom's cell under
 * the grid of the last nvnl_build differs from its stored cell, else 0. */
int nvnl_cells_changed(void* workspace, int dtype, int64_t n_atoms, int32_t n_systems, const void* positions,
                       const int32_t* batch_idx, int32_t* flag, void* stream);

/* cell_list_needs_rebuild from the cache VALUES (same reference lines; no workspace): re-hashes `positions` on the grid
 * (cell, pbc, cells_per_dimension) of a capped build (nvnl_build with max_cells > 0) and compares with
 * atom_to_cell_mapping [n_atoms,3].  cells_per_dimension [n_systems,3]. */
int nvnl_cells_changed_cache(const void* positions, int dtype, int64_t n_atoms, const void* cell, const uint8_t* pbc,
                             const int32_t* batch_idx, int32_t n_systems, const int32_t* cells_per_dimension,
                             const int32_t* atom_to_cell_mapping, int32_t* flag, void* stream);

/* neighbor_list_needs_rebuild (rebuild_detection.py:168-217, 386-503): *flag = 1 if any atom moved farther than
 * `threshold` from its reference position.  No workspace involved. */
int nvnl_moved_beyond(const void* reference_positions, const void* current_positions, int dtype, int64_t n_atoms,
                      double threshold, int32_t* flag, void* stream);

/* Multi-GPU re-assembly (north_star: batch_ptr-sharded ranks + an NCCL all-gather over NVLink; the reference has no
 * distributed code).  Ranks exchange only what cannot be recomputed: per pair the target atom (4 B, gathered straight
 * into row 1 of the global edge_index) and the periodic shift packed into one byte, per atom num_neighbors; a pair's
 * source atom follows from neighbor_ptr.  Each rank writes its own range of the global arrays with nvnl_fill_rows /
 * nvnl_fill_coo (edge_index = global + own offset, row_stride = global pair count), packs its own shifts with
 * nvnl_pack_shifts, and after the gathers expands the foreign ranges with nvnl_expand_gathered.
 *   packed byte = (sx + 1) | (sy + 1) << 2 | (sz + 1) << 4; *bad_flag (device int32, may be NULL; caller zeroes it) is
 *   set when a component lies outside {-1, 0, 1} (unwrapped inputs, boxes smaller than the cutoff: the caller then
 *   gathers the int32 shifts instead). */
int nvnl_pack_shifts(const int32_t* shifts, int64_t n_pairs, uint8_t* packed, int32_t* bad_flag, void* stream);
/* One-word form of the exchange (atom indices below 2^26): targets[p] |= packed byte << 26, in place on the rank's own
 * staging slot — 4 B per pair and ONE array travel.  *bad_flag is also set when a target index does not fit 26 bits.
 * nvnl_expand_padded / nvnl_expand_padded_ranges take such an array as gathered_dst with gathered_packed = NULL. */
int nvnl_pack_shifts_word(const int32_t* shifts, int64_t n_pairs, int32_t* targets, int32_t* bad_flag, void* stream);
/* out_i (row 0 of edge_index) and shifts [P,3] of every atom outside [atom_lo, atom_hi) from the global
 * neighbor_ptr [n_atoms+1] and the gathered packed shifts [P]. */
int nvnl_expand_gathered(const int32_t* neighbor_ptr, int64_t n_atoms, int64_t atom_lo, int64_t atom_hi,
                         const uint8_t* packed_shifts, int32_t* out_i, int32_t* shifts, void* stream);

/* Re-assembly after a PADDED all-gather (one ncclAllGather per array, in place): rank g's targets / packed shifts sit at
 * gathered_dst[g * pmax + k] / gathered_packed[g * pmax + k], k = pair index inside the rank's range
 * [pair_bounds[g], pair_bounds[g+1]); atoms of rank g are [atom_bounds[g], atom_bounds[g+1]) (world + 1 HOST int64 each,
 * world <= 16).  Writes out_j (row 1 of edge_index) for every pair and out_i / shifts for the pairs of the other ranks.
 * gathered_packed = NULL: one-word exchange, gathered_dst[.] = target | packed shift << 26 (nvnl_pack_shifts_word). */
int nvnl_expand_padded(const int32_t* neighbor_ptr, int64_t n_atoms, int32_t world, int32_t rank, const int64_t* atom_bounds,
                       const int64_t* pair_bounds, int64_t pmax, const int32_t* gathered_dst, const uint8_t* gathered_packed,
                       int32_t* out_i, int32_t* out_j, int32_t* shifts, void* stream);

/* The same for a CHUNKED exchange (the all-gather of chunk k overlaps the fill of chunk k + 1 and the expansion of chunk
 * k - 1): one call per chunk.  Rank g's atoms of this chunk are [atom_begin[g], atom_end[g]) — ascending, disjoint; the atoms
 * in between belong to other chunks and are left alone — and its pairs start at pair_begin[g] (= neighbor_ptr[atom_begin[g]]);
 * gathered_dst / gathered_packed are this chunk's world x pmax staging buffers (world HOST int64 each, world <= 16). */
int nvnl_expand_padded_ranges(const int32_t* neighbor_ptr, int64_t n_atoms, int32_t world, int32_t rank, const int64_t* atom_begin,
                              const int64_t* atom_end, const int64_t* pair_begin, int64_t pmax, const int32_t* gathered_dst,
                              const uint8_t* gathered_packed, int32_t* out_i, int32_t* out_j, int32_t* shifts, void* stream);

/* Size of the temporary row buffer nvnl_count_rows may use: entries_per_atom * n_atoms + slack_entries int32 entries
 * (defaults 160 and 148*4*8*2048; negative = default).  Changes nvnl_workspace_bytes(): set it before sizing a
 * workspace and keep it fixed while that workspace is in use.  Process-wide. */
void nvnl_set_rows_budget(int64_t entries_per_atom, int64_t slack_entries);

/* Number of kernels launched by this library since load (bench.py's gpu_launches). */
int64_t nvnl_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* NVALCHEMI_NL_B200_H */
