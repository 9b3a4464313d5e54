// stage_kernel_ns8.cu — instantiates the fused RK-stage kernel (stage_kernel.cuh) for up to 8 staged species.
#include "stage_kernel.cuh"

cudaError_t sb2_launch_stage_ns8(const Sb2StageArgs& a, cudaStream_t stream) {
  return launch_depth<8, 1>(a, stream);
}
