// kernels_fast.cu -- second compilation of the cooperative ADMM and polish kernels of kernels.cu with the storage mode
// of the common large sparse problem fixed at compile time (tile streams in the lane-row layout, [A; P] in cluster
// pairs, fp32 slices, Jacobi preconditioner): see the note at the top of kernels.cu and fast_mode() there.
// Exports launch_solve_fast / launch_polish_fast / fast_kernels only.
#define OSQP_B200_FAST 1
#include "kernels.cu"
