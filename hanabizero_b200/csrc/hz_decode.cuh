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

// eight consecutive logits of a row starting at column c0 (16-byte vector loads); the caller
// guarantees c0 + 8 <= ld and 16-byte aligned rows
__device__ __forceinline__ void load8(const float* row, int c0, float out[8]) {
  const float4 a = *reinterpret_cast<const float4*>(row + c0), b = *reinterpret_cast<const float4*>(row + c0 + 4);
  out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w; out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
}
__device__ __forceinline__ void load8(const __half* row, int c0, float out[8]) {
  const uint4 v = *reinterpret_cast<const uint4*>(row + c0);
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = __half22float2(h[k]);
    out[2 * k] = f.x;
    out[2 * k + 1] = f.y;
  }
}

struct Logits8 {
  float v[8];
};

// Lane `lane` owns columns [8*lane, 8*lane+8).  Vector form needs width <= 256, ld % 8 == 0 and
// 16-byte aligned rows (vec == true); otherwise columns are read one by one.
template <typename T>
__device__ __forceinline__ Logits8 load_logits8(const T* __restrict__ x, int width, bool vec, int lane) {
  Logits8 r;
  const int c0 = lane * 8;
  if (vec) {
    if (c0 < width) load8(x, c0, r.v);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = (c0 + k < width) ? to_f(x[c0 + k]) : 0.0f;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (c0 + k >= width) r.v[k] = -INFINITY;   // padded / absent columns never contribute
  return r;
}

// Whole warp: softmax over the row -> expectation over support -> / delta -> inverse of
// h(v) = sign(v)(sqrt(|v|+1)-1) + 0.001 v -> * delta, NaN -> 0.  Every lane returns the result.
// (width <= 256; one fixed summation order shared by every caller, so results are reproducible
// between the standalone decode kernel and the fused search step.)
__device__ __forceinline__ float warp_decode8(const Logits8& x, const float* __restrict__ support, int width,
                                              float delta, int lane) {
  const int c0 = lane * 8;
  float m = x.v[0];
#pragma unroll
  for (int k = 1; k < 8; ++k) m = fmaxf(m, x.v[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(HZ_FULL, m, o));
  float se = 0.0f, sw = 0.0f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (c0 + k < width) {
      const float e = __expf(x.v[k] - m);   // decode tolerance is float-level, not bit-level (network output)
      se += e;
      sw += e * support[c0 + k];
    }
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

// Two rows at once (value and reward of one tree): the same arithmetic per row as warp_decode8, with the
// two rows' shuffle reductions issued side by side so their latencies overlap.
__device__ __forceinline__ void warp_decode8_pair(const Logits8& xa, const Logits8& xb, const float* __restrict__ support,
                                                  int width, float delta, int lane, float& out_a, float& out_b) {
  const int c0 = lane * 8;
  float ma = xa.v[0], mb = xb.v[0];
#pragma unroll
  for (int k = 1; k < 8; ++k) {
    ma = fmaxf(ma, xa.v[k]);
    mb = fmaxf(mb, xb.v[k]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ta = __shfl_xor_sync(HZ_FULL, ma, o), tb = __shfl_xor_sync(HZ_FULL, mb, o);
    ma = fmaxf(ma, ta);
    mb = fmaxf(mb, tb);
  }
  float sea = 0.0f, swa = 0.0f, seb = 0.0f, swb = 0.0f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (c0 + k < width) {
      const float sp = support[c0 + k];
      const float ea = __expf(xa.v[k] - ma), eb = __expf(xb.v[k] - mb);
      sea += ea;
      swa += ea * sp;
      seb += eb;
      swb += eb * sp;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t0 = __shfl_xor_sync(HZ_FULL, sea, o), t1 = __shfl_xor_sync(HZ_FULL, swa, o);
    const float t2 = __shfl_xor_sync(HZ_FULL, seb, o), t3 = __shfl_xor_sync(HZ_FULL, swb, o);
    sea += t0;
    swa += t1;
    seb += t2;
    swb += t3;
  }
  const float eps = 0.001f;
  const float va = (swa / sea) / delta, vb = (swb / seb) / delta;
  float ra = (sqrtf(1.0f + 4.0f * eps * (fabsf(va) + 1.0f + eps)) - 1.0f) / (2.0f * eps);
  float rb = (sqrtf(1.0f + 4.0f * eps * (fabsf(vb) + 1.0f + eps)) - 1.0f) / (2.0f * eps);
  ra = ra * ra - 1.0f;
  rb = rb * rb - 1.0f;
  ra = (va < 0.0f ? -ra : ra) * delta;
  rb = (vb < 0.0f ? -rb : rb) * delta;
  out_a = (ra != ra) ? 0.0f : ra;
  out_b = (rb != rb) ? 0.0f : rb;
}

template <typename T>
__device__ __forceinline__ bool decode_vec_ok(const T* x, int64_t ld) {
  return (ld % 8) == 0 && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
}

template <typename T>
__device__ __forceinline__ float warp_support_decode(const T* __restrict__ x, const float* __restrict__ support,
                                                     int width, float delta, int64_t ld, int lane) {
  return warp_decode8(load_logits8<T>(x, width, decode_vec_ok(x, ld), lane), support, width, delta, lane);
}

}  // namespace hz
