// rd_stage_v6.cu — instantiates the rd_stage kernel (rd_stage.cuh): up to 4 species, 2 cells per thread,
// 8 warps per CTA, expression stack depth 8 (3-D grids with three or four species).
#include "rd_stage.cuh"

cudaError_t sb2_rd_launch_v6(const Sb2RdArgs& a, cudaStream_t stream) { return rd::launch_mode<4, 1, 8, 8, 2>(a, stream); }
