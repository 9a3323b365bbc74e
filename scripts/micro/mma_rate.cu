// Micro-benchmark: issue rate of tcgen05.mma.cta_group::1.kind::f16 (M=128, N in {64,128,256}, K=16) from 128B-swizzled
// K-major shared-memory tiles, 32 MMAs per batch + one commit, timed on the issuing thread.  nvcc -arch=sm_100a
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_k128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) k(int N, int reps, unsigned long long* out, int commit_each) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 12 * 16384, slot = bar + 8;   // second barrier at bar + 16
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar + 16));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  for (uint32_t i = threadIdx.x; i < 12 * 16384 / 4; i += blockDim.x) asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + 4 * i), "r"(0x3c003c00u));
  asm volatile("fence.proxy.async;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    unsigned long long t0, t1, t2;
    uint32_t parity = 0;
    unsigned long long issue = 0, total = 0;
    for (int r = 0; r < reps; ++r) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      for (int kb = 0; kb < 8; ++kb) {
        const uint64_t da = desc_k128(base + kb * 16384), dw = desc_k128(base + 8 * 16384 + (kb & 1) * 32768);
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16) {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tmem), "l"(da + 2 * k16), "l"(dw + 2 * k16), "r"(idesc), "r"((uint32_t)((kb | k16) != 0)) : "memory");
        }
        if (commit_each) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar + 16) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      uint32_t done = 0;
      while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
      }
      parity ^= 1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t2));
      if (r > 0) { issue += t1 - t0; total += t2 - t0; }
    }
    out[2 * blockIdx.x] = issue / (reps - 1);
    out[2 * blockIdx.x + 1] = total / (reps - 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  unsigned long long* out;
  cudaMalloc(&out, 148 * 16);
  const int smem = 12 * 16384 + 1024 + 64;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int grid = 1; grid <= 128; grid *= 128) {
    for (int N = 64; N <= 256; N *= 2) {
     for (int ce = 0; ce < 2; ++ce) {
      k<<<grid, 128, smem>>>(N, 20, out, ce);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      unsigned long long h[256];
      cudaMemcpy(h, out, grid * 16, cudaMemcpyDeviceToHost);
      printf("grid=%d N=%d commit per K-block=%d: 32 MMAs (8 K-blocks of 64): issue %llu ns, issue+complete %llu ns\n", grid, N, ce, h[0], h[1]);
     }
    }
  }
  return 0;
}
