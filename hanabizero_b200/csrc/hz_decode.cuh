// Categorical-support decoding shared by the standalone decode kernel (hz_nn.cu) and the fused
// search-step kernel (hz_tree.cu): /root/reference/core/config.py:210-232 (inverse_scalar_transform).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "hz_common.cuh"

namespace hz {

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// Whole warp: softmax over x[0..width) -> expectation over support -> / delta -> inverse of
// h(v) = sign(v)(sqrt(|v|+1)-1) + 0.001 v -> * delta, NaN -> 0.  Every lane returns the result.
template <typename T>
__device__ __forceinline__ float warp_support_decode(const T* __restrict__ x, const float* __restrict__ support,
                                                     int width, float delta, int lane) {
  float m = -INFINITY;
  for (int i = lane; i < width; i += 32) m = fmaxf(m, to_f(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(HZ_FULL, m, o));
  float se = 0.0f, sw = 0.0f;
  for (int i = lane; i < width; i += 32) {
    const float e = expf(to_f(x[i]) - m);
    se += e;
    sw += e * support[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    se += __shfl_xor_sync(HZ_FULL, se, o);
    sw += __shfl_xor_sync(HZ_FULL, sw, o);
  }
  const float eps = 0.001f;
  const float v = (sw / se) / delta;
  float r = (sqrtf(1.0f + 4.0f * eps * (fabsf(v) + 1.0f + eps)) - 1.0f) / (2.0f * eps);
  r = r * r - 1.0f;
  r = (v < 0.0f ? -r : r) * delta;
  return (r != r) ? 0.0f : r;
}

}  // namespace hz
