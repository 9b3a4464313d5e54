// rd_stage_v4.cu — instantiates the rd_stage kernel (rd_stage.cuh): up to 4 species, 2 cells per thread,
// 4 warps per CTA, expression stack depth 8.
#include "rd_stage.cuh"

cudaError_t sb2_rd_launch_v4(const Sb2RdArgs& a, cudaStream_t stream) { return rd::launch_mode<4, 1, 4, 8, 2>(a, stream); }
