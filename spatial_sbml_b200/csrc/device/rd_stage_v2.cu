// rd_stage_v2.cu — instantiates the rd_stage kernel (rd_stage.cuh): up to 2 species, 2 cells per thread,
// 8 warps per CTA, expression stack depth 8.
#include "rd_stage.cuh"

cudaError_t sb2_rd_launch_v2(const Sb2RdArgs& a, cudaStream_t stream) { return rd::launch_mode<2, 1, 8, 8, 2>(a, stream); }
