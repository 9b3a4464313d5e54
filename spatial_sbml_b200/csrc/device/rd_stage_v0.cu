// rd_stage_v0.cu — instantiates the rd_stage kernel (rd_stage.cuh): up to 2 species, 8 cells per thread,
// 4 warps per CTA, expression stack depth 2.
#include "rd_stage.cuh"

cudaError_t sb2_rd_launch_v0(const Sb2RdArgs& a, cudaStream_t stream) { return rd::launch_mode<2, 4, 4, 2, 3>(a, stream); }
