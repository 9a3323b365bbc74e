// Micro-benchmark: cost of one dependent kernel boundary inside a CUDA graph on this GPU, plain stream order vs
// programmatic dependent launch (PDL), for a chain of tiny kernels (the shape of one MCTS simulation: 8 dependent
// launches of a few microseconds each).   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o launch_chain launch_chain.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <bool kPdl, bool kEarlyTrigger>
__global__ void k_tiny(int* p, int work) {
  extern __shared__ int sm[];
  if (kPdl && kEarlyTrigger) cudaTriggerProgrammaticLaunchCompletion();
  sm[threadIdx.x] = threadIdx.x;           // prologue that does not depend on the previous kernel
  __syncthreads();
  if (kPdl) cudaGridDependencySynchronize();
  if (kPdl && !kEarlyTrigger) cudaTriggerProgrammaticLaunchCompletion();
  int v = p[blockIdx.x * blockDim.x + threadIdx.x];
  for (int i = 0; i < work; ++i) v = v * 3 + sm[(threadIdx.x + i) & 127];
  p[blockIdx.x * blockDim.x + threadIdx.x] = v + 1;
}

template <bool kPdl, bool kEarly>
static float run(int* d, int grid, int smem, int chain, int work, bool pdl_attr) {
  cudaStream_t s;
  cudaStreamCreate(&s);
  cudaFuncSetAttribute(k_tiny<kPdl, kEarly>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < chain; ++i) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute a[1];
    a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    a[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = a;
    cfg.numAttrs = pdl_attr ? 1 : 0;
    cudaLaunchKernelEx(&cfg, k_tiny<kPdl, kEarly>, d, work);
  }
  cudaStreamEndCapture(s, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 5; ++i) cudaGraphLaunch(ge, s);
  cudaStreamSynchronize(s);
  const int reps = 50;
  cudaEventRecord(e0, s);
  for (int i = 0; i < reps; ++i) cudaGraphLaunch(ge, s);
  cudaEventRecord(e1, s);
  cudaStreamSynchronize(s);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(err));
  cudaGraphExecDestroy(ge);
  cudaGraphDestroy(g);
  cudaStreamDestroy(s);
  return ms * 1e3f / (reps * chain);
}

int main() {
  int* d;
  cudaMalloc(&d, 1 << 22);
  cudaMemset(d, 0, 1 << 22);
  const int chain = 56;
  printf("per-kernel time inside a CUDA graph of %d dependent kernels (us)\n", chain);
  printf("%6s %8s %6s | %8s %12s %12s\n", "grid", "smem", "work", "plain", "pdl(late)", "pdl(early)");
  for (int grid : {16, 64, 148, 592}) {
    for (int smem : {1024, 100 * 1024}) {
      for (int work : {0, 2000}) {
        const float a = run<false, false>(d, grid, smem, chain, work, false);
        const float b = run<true, false>(d, grid, smem, chain, work, true);
        const float c = run<true, true>(d, grid, smem, chain, work, true);
        printf("%6d %8d %6d | %8.2f %12.2f %12.2f\n", grid, smem, work, a, b, c);
      }
    }
  }
  return 0;
}
