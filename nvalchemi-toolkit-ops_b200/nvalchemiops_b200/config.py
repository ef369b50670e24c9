"""Run-time switches of the B200 neighbor-list path."""

# Arithmetic of the distance predicate ||r_j - r_i + s.cell||^2 < rc^2 (reference cell_list.py:531-545).
# True  = mul + fma chain — what NVRTC emits for the reference's Warp kernels with Warp's default
#         ``fuse_fp`` (fmad=true);  False = every multiply/add rounded separately.
# Only pairs within ~1 ulp of the cutoff can differ between the two (SURVEY.md §7 "knife-edge").
fma: bool = True
