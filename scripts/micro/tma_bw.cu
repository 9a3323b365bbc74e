// Micro-benchmark: per-SM TMA load rate for [128 rows x 128 B] boxes (row stride 1024 B, 128-byte swizzle),
// 8 boxes in flight per CTA, with and without a cluster launch.   nvcc -arch=sm_100a tma_bw.cu -o tma_bw -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap map, int iters, int nbox, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + 9 * 16384;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 9; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8 * i));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int it = 0; it < iters; ++it) {
      for (int j = 0; j < nbox; ++j) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 16384;" ::"r"(bars + 8 * j) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(base + j * 16384), "l"(&map), "r"(bars + 8 * j), "r"(j * 64), "r"((int)(blockIdx.x % 32) * 128), "r"(0) : "memory");
      }
      for (int j = 0; j < nbox; ++j) {
        uint32_t done = 0;
        while (!done) {
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(done) : "r"(bars + 8 * j), "r"(it & 1) : "memory");
        }
      }
    }
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    out[blockIdx.x] = t1 - t0;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 128;
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeTiledFn fn = (EncodeTiledFn)fnp;
  void* buf;
  cudaMalloc(&buf, 4096 * 576 * 2);
  cudaMemset(buf, 1, 4096 * 576 * 2);
  CUtensorMap map;
  cuuint64_t dims[3] = {576, 4096, 1}, strides[2] = {576 * 2, 576 * 2};
  cuuint32_t box[3] = {64, 128, 1}, es[3] = {1, 1, 1};
  CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  unsigned long long* out;
  cudaMalloc(&out, grid * 8);
  const int smem = 9 * 16384 + 1024 + 128;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int cluster = 1; cluster <= 4; cluster *= 4) {
    for (int nbox = 1; nbox <= 9; nbox += 4) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      const int iters = 20;
      for (int rep = 0; rep < 2; ++rep) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, k, map, iters, nbox, out);
        if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
        cudaDeviceSynchronize();
      }
      unsigned long long h[256];
      cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
      double avg = 0; unsigned long long mx = 0;
      for (int i = 0; i < grid; ++i) { avg += h[i]; if (h[i] > mx) mx = h[i]; }
      avg /= grid;
      printf("cluster=%d grid=%d boxes/iter=%d: per iteration avg %.0f ns, max %.0f ns -> %.1f GB/s per SM (avg)\n", cluster, grid, nbox,
             avg / iters, (double)mx / iters, nbox * 16384.0 / (avg / iters));
    }
  }
  return 0;
}
