// rd_stage_v3.cu — instantiates the rd_stage kernel (rd_stage.cuh): up to 4 species, 4 cells per thread,
// 4 warps per CTA, expression stack depth 4.
#include "rd_stage.cuh"

cudaError_t sb2_rd_launch_v3(const Sb2RdArgs& a, cudaStream_t stream) { return rd::launch_mode<4, 2, 4, 4, 2>(a, stream); }
