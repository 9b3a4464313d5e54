// stage_kernel.cu — dispatch of the fused RK-stage kernel variants (stage_kernel.cuh, instantiated in
// stage_kernel_ns{2,4,8}.cu) and the stand-alone RK combine kernel.
#include "sb2_internal.h"
#include "sb2_ops.cuh"

cudaError_t sb2_launch_stage_ns2(const Sb2StageArgs& a, cudaStream_t stream);
cudaError_t sb2_launch_stage_ns4(const Sb2StageArgs& a, cudaStream_t stream);
cudaError_t sb2_launch_stage_ns8(const Sb2StageArgs& a, cudaStream_t stream);

namespace {

// ---------------------------------------------------------------------------------------------
// stand-alone RK combine (used when the combine cannot be fused into stage 4, i.e. when advection
// runs between the last stage and the update; spatialsimulator.cpp:1484-1491, 1535-1536)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) combine_kernel(const __grid_constant__ Sb2CombineArgs A) {
  const int x = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int y = blockIdx.y;
  const int z = blockIdx.z;
  if (x >= A.nx) return;
  const int64_t idx = A.first + (int64_t)z * A.PL + (int64_t)y * A.P + x;
  for (int s = 0; s < A.S; ++s) {
    const uchar2 f = *reinterpret_cast<const uchar2*>(A.flags[s] + idx);
    if (!((f.x | f.y) & SB2_F_DOMAIN)) continue;
    double2 u = *reinterpret_cast<const double2*>(A.u[s] + idx);
    const double2 k0 = *reinterpret_cast<const double2*>(A.k[s][0] + idx);
    const double2 k1 = *reinterpret_cast<const double2*>(A.k[s][1] + idx);
    const double2 k2 = *reinterpret_cast<const double2*>(A.k[s][2] + idx);
    const double2 k3 = *reinterpret_cast<const double2*>(A.k[s][3] + idx);
    if (f.x & SB2_F_DOMAIN) u.x = dadd(u.x, ddiv_pos(dmul(A.dt, dadd(dadd(dadd(k0.x, dmul(2.0, k1.x)), dmul(2.0, k2.x)), k3.x)), 6.0));
    if (f.y & SB2_F_DOMAIN) u.y = dadd(u.y, ddiv_pos(dmul(A.dt, dadd(dadd(dadd(k0.y, dmul(2.0, k1.y)), dmul(2.0, k2.y)), k3.y)), 6.0));
    *reinterpret_cast<double2*>(A.u[s] + idx) = u;
  }
}

}  // namespace

// cell pairs per thread of the variant that will run: wide strips need >= 512 cells per row to pay off
int sb2_stage_pairs(int nx, int nspecies) { return (nspecies <= 4 && nx > 256) ? 2 : 1; }

cudaError_t sb2_launch_stage(const Sb2StageArgs& a, cudaStream_t stream, int* launches) {
  if (a.row_end <= a.row_begin) return cudaSuccess;
  if (launches) ++*launches;
  if (a.S <= 2) return sb2_launch_stage_ns2(a, stream);
  if (a.S <= 4) return sb2_launch_stage_ns4(a, stream);
  return sb2_launch_stage_ns8(a, stream);
}

cudaError_t sb2_launch_combine(const Sb2CombineArgs& a, cudaStream_t stream, int* launches) {
  dim3 block(256), grid(((a.nx + 1) / 2 + 255) / 256, a.ny, a.nz);
  if (launches) ++*launches;
  combine_kernel<<<grid, block, 0, stream>>>(a);
  return cudaGetLastError();
}
