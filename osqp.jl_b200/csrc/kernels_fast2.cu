// kernels_fast2.cu -- third compilation of the cooperative ADMM and polish kernels of kernels.cu: fixed mode 2, the same
// as kernels_fast.cu but WITHOUT cluster pairs (problems whose [A; P] stream has one column group, n <= 27,648, or
// more than two): see the note at the top of kernels.cu and fast_mode() there.
// Exports launch_solve_fast2 / launch_polish_fast2 / kernels_fast2 only.
#define OSQP_B200_FAST 2
#include "kernels.cu"
