// rd_stage_v1.cu — instantiates the rd_stage kernel (rd_stage.cuh): up to 2 species, 4 cells per thread,
// 8 warps per CTA, expression stack depth 4.
#include "rd_stage.cuh"

cudaError_t sb2_rd_launch_v1(const Sb2RdArgs& a, cudaStream_t stream) { return rd::launch_mode<2, 2, 8, 4, 2>(a, stream); }
