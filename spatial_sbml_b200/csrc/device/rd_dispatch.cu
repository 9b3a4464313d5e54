// rd_dispatch.cu — variant table of the rd_stage kernel family and the strip-class map kernel.
#include "sb2_internal.h"
#include <cstdlib>

cudaError_t sb2_rd_launch_v1(const Sb2RdArgs& a, cudaStream_t stream);
cudaError_t sb2_rd_launch_v2(const Sb2RdArgs& a, cudaStream_t stream);
cudaError_t sb2_rd_launch_v3(const Sb2RdArgs& a, cudaStream_t stream);
cudaError_t sb2_rd_launch_v4(const Sb2RdArgs& a, cudaStream_t stream);
cudaError_t sb2_rd_launch_v6(const Sb2RdArgs& a, cudaStream_t stream);

namespace {
// id, max species, cell pairs per thread, warps per CTA, expression stack depth.  Ids 0 (8 cells per thread, 4 warps) and 5
// (three CTAs per SM at 80 registers) were measured slower than 1 at every size and are no longer built (ns 0: not available).
const Sb2RdVariant kVariants[] = {
  {0, 0, 4, 4, 2}, {1, 2, 2, 8, 4}, {2, 2, 1, 8, 8}, {3, 4, 2, 4, 4}, {4, 4, 1, 4, 8}, {5, 0, 2, 8, 2}, {6, 4, 1, 8, 8},
};
const int kNumVariants = (int)(sizeof(kVariants) / sizeof(kVariants[0]));

// One warp per (row, strip): class 1 when every cell of the strip row carries exactly SB2_F_DOMAIN in every
// geometry (interior cell, no boundary bit), class 2 when no cell is in any domain, else 0.
__global__ void __launch_bounds__(256) classify_kernel(const uint8_t* f0, const uint8_t* f1, const uint8_t* f2, const uint8_t* f3,
                                                       int G, int nx, int ny, int nz, int64_t P, int64_t PL, int64_t first, int sw, uint8_t* cls,
                                                       int nstrips) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= ny * nz * nstrips) return;
  const int row = warp / nstrips, strip = warp % nstrips;
  const int64_t row0 = first + (int64_t)(row / ny) * PL + (int64_t)(row % ny) * P;
  const uint8_t* fl[4] = {f0, f1, f2, f3};
  bool all_interior = true, any_domain = false;
  for (int i = lane; i < sw; i += 32) {
    const int x = strip * sw + i;
    for (int g = 0; g < G; ++g) {
      const uint8_t v = x < nx ? fl[g][row0 + x] : 0;
      all_interior = all_interior && v == SB2_F_DOMAIN;
      any_domain = any_domain || (v & SB2_F_DOMAIN);
    }
  }
  all_interior = __all_sync(0xffffffffu, all_interior);
  any_domain = __any_sync(0xffffffffu, any_domain);
  if (lane == 0) cls[warp] = all_interior ? 1 : any_domain ? 0 : 2;
}
}  // namespace

int sb2_rd_pick_variant(int dim, int nx, int nspecies, int depth, Sb2RdVariant* out, int64_t ncells) {
  if ((dim != 2 && dim != 3) || nspecies < 1 || nspecies > 4) return -1;
  int id;
  if (dim == 3) {
    // the z-marching kernel keeps three plane tiles of W + 2 rows in shared memory: 8 warps and 2 cells per thread leave
    // room for three CTAs per SM with two species
    id = nspecies <= 2 ? 2 : 6;
  } else {
    // four cells per thread on wide grids (measured at 8192^2: variant 1 beats 0 and 5).  Grids of up to 2^18 cells are bound by
    // the latency of a warp's instruction stream, not by throughput: two cells per thread there (301^2 Brusselator: 16.4 ->
    // 14.4 us per step)
    const bool wide = depth <= 4 && nx >= 192 && (ncells <= 0 || ncells > ((int64_t)1 << 18));
    if (nspecies <= 2) id = wide ? 1 : 2;
    else id = wide ? 3 : 4;
  }
  const char* force = getenv("SB2_RD_VARIANT");  // experiments: force a variant that fits the model
  if (force && atoi(force) >= 0 && atoi(force) < kNumVariants && kVariants[atoi(force)].ns >= nspecies) id = atoi(force);
  if (depth > kVariants[id].depth) return -1;
  if (out) *out = kVariants[id];
  return id;
}

cudaError_t sb2_rd_launch(const Sb2RdArgs& a, cudaStream_t stream) {
  switch (a.rd_variant) {
    case 1: return sb2_rd_launch_v1(a, stream);
    case 2: return sb2_rd_launch_v2(a, stream);
    case 3: return sb2_rd_launch_v3(a, stream);
    case 4: return sb2_rd_launch_v4(a, stream);
    case 6: return sb2_rd_launch_v6(a, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t sb2_rd_classify(const uint8_t* const* flags, int G, int nx, int ny, int nz, int64_t P, int64_t PL, int64_t first,
                            int sw, uint8_t* cls, int nstrips, cudaStream_t stream) {
  const int64_t warps = (int64_t)ny * nz * nstrips;
  const int blocks = (int)((warps * 32 + 255) / 256);
  classify_kernel<<<blocks, 256, 0, stream>>>(flags[0], G > 1 ? flags[1] : nullptr, G > 2 ? flags[2] : nullptr,
                                               G > 3 ? flags[3] : nullptr, G, nx, ny, nz, P, PL, first, sw, cls, nstrips);
  return cudaGetLastError();
}
