// Floor of a chain of dependent tiny kernels replayed from a CUDA graph on this GPU: per-node time for empty kernels, with 1 KB of
// parameters, with 100 KB of dynamic shared memory, with programmatic dependent launch, and for a kernel that does one global
// load -> store round trip per CTA.  nvcc -arch=sm_100a -o graph_floor graph_floor.cu
#include <cstdio>
#include <cuda_runtime.h>
struct Big { double x[120]; };
__global__ void k_empty() {}
__global__ void k_params(const __grid_constant__ Big b, double* out) { if (b.x[3] == 12345.0) out[0] = 1; }
__global__ void k_smem(double* out) { extern __shared__ double s[]; if (out == nullptr) s[threadIdx.x] = 1; }
__global__ void k_pdl(double* out) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (out == nullptr) out[0] = 1;
}
__global__ void k_rt(const double* in, double* out) { const int i = blockIdx.x * blockDim.x + threadIdx.x; out[i] = in[i] + 1.0; }
__global__ void k_rt_pdl(const double* in, double* out) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int i = blockIdx.x * blockDim.x + threadIdx.x; out[i] = in[i] + 1.0;
}
template <class F> float run(const char* name, F launch, cudaStream_t st, int nodes = 8, int reps = 2000) {
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
  for (int i = 0; i < nodes; ++i) launch(i);
  cudaStreamEndCapture(st, &g);
  cudaGraphInstantiate(&ge, g, 0);
  for (int i = 0; i < 50; ++i) cudaGraphLaunch(ge, st);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  for (int i = 0; i < reps; ++i) cudaGraphLaunch(ge, st);
  cudaEventRecord(e1, st); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-44s %.2f us per kernel node (%s)\n", name, 1e3f * ms / (reps * nodes), cudaGetErrorString(cudaGetLastError()));
  return ms;
}
int main() {
  cudaStream_t st; cudaStreamCreate(&st);
  double *a, *b; cudaMalloc(&a, 1 << 22); cudaMalloc(&b, 1 << 22); cudaMemset(a, 0, 1 << 22);
  Big big = {};
  cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  auto pdl = [&](auto kern, dim3 grid, dim3 block, size_t smem, auto... args) {
    cudaLaunchConfig_t cfg = {}; cudaLaunchAttribute at[1];
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, args...);
  };
  run("empty, 114 CTAs x 256", [&](int) { k_empty<<<114, 256, 0, st>>>(); }, st);
  run("1 KB of parameters", [&](int) { k_params<<<114, 256, 0, st>>>(big, b); }, st);
  run("100 KB dynamic shared memory", [&](int) { k_smem<<<114, 256, 100 * 1024, st>>>(b); }, st);
  run("alternating 56 / 100 KB shared memory", [&](int i) { k_smem<<<114, 256, (i & 1) ? 100 * 1024 : 56 * 1024, st>>>(b); }, st);
  run("empty with programmatic dependent launch", [&](int) { pdl(k_pdl, dim3(114), dim3(256), 0, b); }, st);
  run("load + store round trip", [&](int i) { k_rt<<<114, 256, 0, st>>>((i & 1) ? b : a, (i & 1) ? a : b); }, st);
  run("load + store round trip, PDL", [&](int i) { pdl(k_rt_pdl, dim3(114), dim3(256), 0, (const double*)((i & 1) ? b : a), (i & 1) ? a : b); }, st);
  return 0;
}
