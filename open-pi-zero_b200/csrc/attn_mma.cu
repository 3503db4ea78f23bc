#include "common.cuh"
#include "kernels.h"
int attn_mma_supported(const AttnArgs &) { return 0; }
int launch_attn_mma(const AttnArgs &, cudaStream_t) { return PZ_ERR_INVALID; }
