/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see nl_oracle_impl.h).  Instantiates the CPU
 * restatement of the reference cell-list neighbor-list for float and double.
 * Built with -ffp-contract=off so that only the explicit FMA() calls fuse.
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>

#define REAL float
#define SUF f32
#define SQRT sqrtf
#define CEIL ceilf
#define FLOOR floorf
#define FMA fmaf
#include "nl_oracle_impl.h"
#undef REAL
#undef SUF
#undef SQRT
#undef CEIL
#undef FLOOR
#undef FMA

#define REAL double
#define SUF f64
#define SQRT sqrt
#define CEIL ceil
#define FLOOR floor
#define FMA fma
#include "nl_oracle_impl.h"
#undef REAL
#undef SUF
#undef SQRT
#undef CEIL
#undef FLOOR
#undef FMA

int nlo_abi_version(void) { return 1; }
