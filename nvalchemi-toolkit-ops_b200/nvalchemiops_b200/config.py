"""Run-time switches of the B200 neighbor-list path."""

# Arithmetic of the distance predicate ||r_j - r_i + s.cell||^2 < rc^2 (reference cell_list.py:531-545).
# True  = mul + fma chain — what NVRTC emits for the reference's Warp kernels with Warp's default
#         ``fuse_fp`` (fmad=true);  False = every multiply/add rounded separately.
# Only pairs within ~1 ulp of the cutoff can differ between the two (SURVEY.md §7 "knife-edge").
fma: bool = True

# COO output path for fp32 inputs.
# "rows"  = single sweep: distances and compaction done once into a temporary row buffer, then the final arrays are
#           streamed in atom order (csrc/nvnl_rows.cuh);  falls back to "masks" when the temporary buffer (160 entries
#           per atom) is too small.
# "masks" = two sweeps: count pass stores hit masks, fill pass expands them at neighbor_ptr (csrc/nvnl_fast.cuh).
# fp64 inputs always take "masks".  Both produce the same sets (tests/test_gpu_parity.py runs both).
coo_path: str = "rows"

# Single-sweep COO path: let the sweep kernel zero-fill the shifts output while it runs (TMA bulk stores from its producer
# warps), sized from the pair count of the previous query with the same (device, stream, atoms, systems, cutoff, half_fill)
# signature.  False = always zero after the
# size sync (inside the output kernel).
prezero_shifts: bool = True
prezero_min_pairs: int = 1_000_000   # below this the extra launch + event cost more than the overlap saves

# With a speculative shifts buffer in hand, also launch the output kernel BEFORE the size sync, into an edge_index buffer
# of the guessed size (nvnl_fill_rows_speculative, which reads the pair count on the device); the host then only
# creates the views.  Removes the ~30 us the GPU idles at the sync.  A wrong guess is detected after the sync and the
# output kernel is run again the regular way.
speculative_fill: bool = True

# Debug: validate inputs on the padded-matrix path too.  That path is sync-free (CUDA-graph capturable) and therefore
# does not read the device error word: an out-of-range batch_idx is clamped to a valid system, a search radius of 64+
# cells is truncated, a singular cell yields an empty list — the COO path raises ValueError for all three.  True = one
# host sync per matrix query to raise the same errors.
check_inputs: bool = False

# Multi-GPU sharded batches (neighborlist/distributed.py): every rank splits its own systems into this many chunks and the
# exchange of chunk k (NCCL all-gather on a communication stream) runs under the output kernels of chunk k + 1 and the
# re-assembly of chunk k - 1.  1 = one exchange after all kernels (no overlap).
exchange_chunks: int = 2

# ... and with atom indices below 2^26 the packed shift rides in the top six bits of the pair's target word: 4 B per pair
# and one all-gather per chunk instead of 5 B and two.  False = separate byte array (any number of atoms).
exchange_word: bool = True
