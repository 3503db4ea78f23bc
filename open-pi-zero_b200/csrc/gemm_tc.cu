#include "common.cuh"
#include "kernels.h"
int gemm_tc_supported(const LinearArgs &) { return 0; }
int launch_linear_tc(const LinearArgs &, cudaStream_t, const char **err) { if (err) *err = "not built"; return PZ_ERR_INVALID; }
