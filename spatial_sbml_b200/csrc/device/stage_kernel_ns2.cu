// stage_kernel_ns2.cu — instantiates the fused RK-stage kernel (stage_kernel.cuh) for up to 2 staged species.
#include "stage_kernel.cuh"

cudaError_t sb2_launch_stage_ns2(const Sb2StageArgs& a, cudaStream_t stream) {
  if (a.np == 2) return launch_depth<2, 2>(a, stream);
  return launch_depth<2, 1>(a, stream);
}
