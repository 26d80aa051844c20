// kernels_fast3.cu -- fourth compilation of the cooperative ADMM and polish kernels of kernels.cu: fixed mode 3, tile
// streams in the lane-row layout without cluster pairs, fp32 slices and fp32 value streams in the PCG, and the
// slack-elimination preconditioner (the Lasso config): see the note at the top of kernels.cu and fast_mode() there.
// Exports launch_solve_fast3 / launch_polish_fast3 / kernels_fast3 only.
#define OSQP_B200_FAST 3
#include "kernels.cu"
