// rd_stage_v5.cu — instantiates the rd_stage kernel (rd_stage.cuh): up to 2 species, 4 cells per thread,
// 8 warps per CTA, expression stack depth 2, three CTAs per SM (80 registers).
#include "rd_stage.cuh"

cudaError_t sb2_rd_launch_v5(const Sb2RdArgs& a, cudaStream_t stream) { return rd::launch_mode<2, 2, 8, 2, 3>(a, stream); }
