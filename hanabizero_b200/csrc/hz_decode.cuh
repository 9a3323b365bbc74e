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

// The decode is not part of the reference's tree engine (its callers do it in PyTorch, core/config.py:210-232), so its
// bar is float tolerance, not bits: fused multiply-adds and the fast exp2 / reciprocal units are used on purpose —
// this code runs once per tree per simulation inside an issue-bound kernel.  One arithmetic sequence is shared by the
// standalone kernel and the fused search step, so the two paths agree bit for bit.

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// this lane's eight support values (zero past the row's width)
__device__ __forceinline__ void load_support8(const float* __restrict__ support, int width, int c0, float sp[8]) {
  if (c0 + 8 <= width && ((reinterpret_cast<uintptr_t>(support + c0) & 15) == 0)) {
    const float4 a = *reinterpret_cast<const float4*>(support + c0), b = *reinterpret_cast<const float4*>(support + c0 + 4);
    sp[0] = a.x; sp[1] = a.y; sp[2] = a.z; sp[3] = a.w; sp[4] = b.x; sp[5] = b.y; sp[6] = b.z; sp[7] = b.w;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) sp[k] = (c0 + k < width) ? support[c0 + k] : 0.0f;
  }
}

// per-lane partial sums of exp(x - m) and exp(x - m) * support; absent columns hold -inf and contribute exp2(-inf) = 0
__device__ __forceinline__ void decode_partial(const Logits8& x, const float sp[8], float m, float& se, float& sw) {
  const float kLog2e = 1.4426950408889634f;
  const float nm = -m * kLog2e;
  se = 0.0f;
  sw = 0.0f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float e = fast_exp2(__fmaf_rn(x.v[k], kLog2e, nm));
    se += e;
    sw = __fmaf_rn(e, sp[k], sw);
  }
}

// expectation -> / delta -> inverse of h(v) = sign(v)(sqrt(|v|+1)-1) + 0.001 v -> * delta, NaN -> 0
__device__ __forceinline__ float decode_finish(float se, float sw, float delta) {
  const float eps = 0.001f;
  const float v = __fdividef(__fdividef(sw, se), delta);
  float r = (fast_sqrt(__fmaf_rn(4.0f * eps, fabsf(v) + 1.0f + eps, 1.0f)) - 1.0f) * (1.0f / (2.0f * eps));
  r = __fmaf_rn(r, r, -1.0f);
  r = (v < 0.0f ? -r : r) * delta;
  return (r != r) ? 0.0f : r;
}

// Whole warp: softmax over the row -> expectation over support -> inverse transform.  Every lane returns the result.
// (width <= 256; one fixed summation order shared by every caller.)
__device__ __forceinline__ float warp_decode8(const Logits8& x, const float* __restrict__ support, int width,
                                              float delta, int lane) {
  float sp[8];
  load_support8(support, width, lane * 8, sp);
  float m = x.v[0];
#pragma unroll
  for (int k = 1; k < 8; ++k) m = fmaxf(m, x.v[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(HZ_FULL, m, o));
  float se, sw;
  decode_partial(x, sp, m, se, sw);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    se += __shfl_xor_sync(HZ_FULL, se, o);
    sw += __shfl_xor_sync(HZ_FULL, sw, o);
  }
  return decode_finish(se, sw, delta);
}

// Two rows at once (value and reward of one tree): the same arithmetic per row as warp_decode8, with the
// two rows' shuffle reductions issued side by side so their latencies overlap.
__device__ __forceinline__ void warp_decode8_pair(const Logits8& xa, const Logits8& xb, const float* __restrict__ support,
                                                  int width, float delta, int lane, float& out_a, float& out_b) {
  float sp[8];
  load_support8(support, width, lane * 8, sp);
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
  float sea, swa, seb, swb;
  decode_partial(xa, sp, ma, sea, swa);
  decode_partial(xb, sp, mb, seb, swb);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t0 = __shfl_xor_sync(HZ_FULL, sea, o), t1 = __shfl_xor_sync(HZ_FULL, swa, o);
    const float t2 = __shfl_xor_sync(HZ_FULL, seb, o), t3 = __shfl_xor_sync(HZ_FULL, swb, o);
    sea += t0;
    swa += t1;
    seb += t2;
    swb += t3;
  }
  out_a = decode_finish(sea, swa, delta);
  out_b = decode_finish(seb, swb, delta);
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
