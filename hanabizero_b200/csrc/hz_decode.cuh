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

// Raw (unconverted) segment of a logit row: lane owns columns [8*lane, 8*lane + 8) (width <= 256).  Issuing the
// load and converting it are separate steps so that a kernel can put every independent load in flight before the
// first conversion stalls on one of them, and only the raw words stay live between the passes of the decode.
template <typename T> struct RawRow8;
template <> struct RawRow8<__half> { uint4 a; };
template <> struct RawRow8<float> { float4 a, b; };

template <typename T>
__device__ __forceinline__ bool decode_vec_ok(const T* x, int64_t ld) {
  return (ld % 8) == 0 && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
}

// vector form: rows 16-byte aligned with ld % 8 == 0 (decode_vec_ok)
__device__ __forceinline__ RawRow8<__half> load_raw8(const __half* __restrict__ row, int width, int lane) {
  RawRow8<__half> r;
  r.a = make_uint4(0, 0, 0, 0);
  if (lane * 8 < width) r.a = *reinterpret_cast<const uint4*>(row + lane * 8);
  return r;
}
__device__ __forceinline__ RawRow8<float> load_raw8(const float* __restrict__ row, int width, int lane) {
  RawRow8<float> r;
  r.a = r.b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane * 8 < width) {
    r.a = *reinterpret_cast<const float4*>(row + lane * 8);
    r.b = *reinterpret_cast<const float4*>(row + lane * 8 + 4);
  }
  return r;
}
// any alignment / leading dimension: element by element (cold path, kept out of line)
static __device__ __noinline__ RawRow8<__half> load_raw8_scalar(const __half* __restrict__ row, int width, int lane) {
  RawRow8<__half> r;
  __half h[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) h[k] = (lane * 8 + k < width) ? row[lane * 8 + k] : __float2half_rn(0.0f);
  r.a = *reinterpret_cast<const uint4*>(h);
  return r;
}
static __device__ __noinline__ RawRow8<float> load_raw8_scalar(const float* __restrict__ row, int width, int lane) {
  RawRow8<float> r;
  float f[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) f[k] = (lane * 8 + k < width) ? row[lane * 8 + k] : 0.0f;
  r.a = make_float4(f[0], f[1], f[2], f[3]);
  r.b = make_float4(f[4], f[5], f[6], f[7]);
  return r;
}

// the eight logits of this lane as floats; columns past the row's width become -inf (they never contribute)
__device__ __forceinline__ void raw_to_floats(const RawRow8<__half>& raw, int width, int lane, float v[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&raw.a);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = __half22float2(h[k]);
    v[2 * k] = f.x;
    v[2 * k + 1] = f.y;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (lane * 8 + k >= width) v[k] = -INFINITY;
}
__device__ __forceinline__ void raw_to_floats(const RawRow8<float>& raw, int width, int lane, float v[8]) {
  v[0] = raw.a.x; v[1] = raw.a.y; v[2] = raw.a.z; v[3] = raw.a.w;
  v[4] = raw.b.x; v[5] = raw.b.y; v[6] = raw.b.z; v[7] = raw.b.w;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (lane * 8 + k >= width) v[k] = -INFINITY;
}

// Arithmetic of the decode (core/config.py:210-232): float32 throughout.  exp(x - m) is one fused multiply-add into
// the hardware exp2 unit (MUFU.EX2, 2 ulp — the unit libdevice's own expf ends in; the argument x*log2(e) - m*log2(e)
// is a single rounding, so a softmax term within e^-20 of the maximum carries <= 1e-6 relative error and the terms
// that matter <= 3e-7); everything after the sums — the two divisions, the square root — is correctly rounded.
// The inverse transform
//     out = sign(v) * (((sqrt(1 + 4 eps (|v| + 1 + eps)) - 1) / (2 eps))^2 - 1) * delta
// is evaluated in the algebraically identical cancellation-free form: with s = sqrt(1 + 4 eps (|v| + 1 + eps)) and
// s0 = 1 + 2 eps (s0^2 = 1 + 4 eps (1 + eps)), (s - 1) / (2 eps) - 1 = (s - s0) / (2 eps) = 2 |v| / (s + s0) =: u and the
// result is u (u + 2).  The reference's own float32 evaluation of the textbook form loses ~1e-4 absolute near zero
// (sqrt(1.004) - 1 and r^2 - 1 both cancel); this one adds a few ulp to the expectation's own rounding
// (tests/test_nn_glue_gpu.py: <= 1e-5 relative against a float64 evaluation, closer to it than torch's float32).
// One arithmetic sequence is shared by the standalone kernel and the fused search step, so the two paths agree
// bit for bit.

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

template <typename T>
__device__ __forceinline__ float lane_max8(const RawRow8<T>& raw, int width, int lane) {
  float v[8];
  raw_to_floats(raw, width, lane, v);
  float m = v[0];
#pragma unroll
  for (int k = 1; k < 8; ++k) m = fmaxf(m, v[k]);
  return m;
}

// per-lane partial sums of exp(x - m) and exp(x - m) * support; absent columns hold -inf and contribute exp(-inf) = 0
template <typename T>
__device__ __forceinline__ void decode_partial(const RawRow8<T>& raw, int width, int lane, const float sp[8], float m,
                                               float& se, float& sw) {
  float v[8];
  raw_to_floats(raw, width, lane, v);
  se = 0.0f;
  sw = 0.0f;
  const float kLog2e = 1.4426950408889634f;
  const float nm = __fmul_rn(-m, kLog2e);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fmaf_rn(v[k], kLog2e, nm)));   // exp(v - m); -inf -> 0
    se = __fadd_rn(se, e);
    sw = __fmaf_rn(e, sp[k], sw);
  }
}

// expectation -> / delta -> inverse of h(v) = sign(v)(sqrt(|v|+1)-1) + 0.001 v -> * delta, NaN -> 0
static __device__ __noinline__ float decode_finish(float se, float sw, float delta) {
  const float eps = 0.001f;
  const float v = __fdiv_rn(__fdiv_rn(sw, se), delta);
  const float a = fabsf(v);
  const float y = __fmul_rn(4.0f * eps, __fadd_rn(__fadd_rn(a, 1.0f), eps));
  const float s = __fsqrt_rn(__fadd_rn(1.0f, y));
  const float u = __fdiv_rn(__fmul_rn(2.0f, a), __fadd_rn(s, 1.0f + 2.0f * eps));
  float r = __fmul_rn(u, __fadd_rn(u, 2.0f));
  r = __fmul_rn(v < 0.0f ? -r : r, delta);
  return (r != r) ? 0.0f : r;
}

// Whole warp: softmax over the row -> expectation over support -> inverse transform.  Every lane returns the result.
// (width <= 256; one fixed summation order shared by every caller.)
template <typename T>
__device__ __forceinline__ float warp_decode8(const RawRow8<T>& x, const float* __restrict__ support, int width,
                                              float delta, int lane) {
  float sp[8];
  load_support8(support, width, lane * 8, sp);
  float m = lane_max8(x, width, lane);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(HZ_FULL, m, o));
  float se, sw;
  decode_partial(x, width, lane, sp, m, se, sw);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t0 = __shfl_xor_sync(HZ_FULL, se, o), t1 = __shfl_xor_sync(HZ_FULL, sw, o);
    se = __fadd_rn(se, t0);
    sw = __fadd_rn(sw, t1);
  }
  return decode_finish(se, sw, delta);
}

// Two rows at once (value and reward of one tree): the same arithmetic per row as warp_decode8, with the two rows'
// shuffle reductions issued side by side so their latencies overlap.
template <typename T>
__device__ __forceinline__ void warp_decode8_pair(const RawRow8<T>& xa, const RawRow8<T>& xb,
                                                  const float* __restrict__ support, int width, float delta, int lane,
                                                  float& out_a, float& out_b) {
  float sp[8];
  load_support8(support, width, lane * 8, sp);
  float ma = lane_max8(xa, width, lane), mb = lane_max8(xb, width, lane);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ta = __shfl_xor_sync(HZ_FULL, ma, o), tb = __shfl_xor_sync(HZ_FULL, mb, o);
    ma = fmaxf(ma, ta);
    mb = fmaxf(mb, tb);
  }
  float sea, swa, seb, swb;
  decode_partial(xa, width, lane, sp, ma, sea, swa);
  decode_partial(xb, width, lane, sp, mb, seb, swb);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t0 = __shfl_xor_sync(HZ_FULL, sea, o), t1 = __shfl_xor_sync(HZ_FULL, swa, o);
    const float t2 = __shfl_xor_sync(HZ_FULL, seb, o), t3 = __shfl_xor_sync(HZ_FULL, swb, o);
    sea = __fadd_rn(sea, t0);
    swa = __fadd_rn(swa, t1);
    seb = __fadd_rn(seb, t2);
    swb = __fadd_rn(swb, t3);
  }
  // one pass through the (out-of-line) inverse transform: the lower half-warp finishes row a, the upper half row b
  const bool lo = lane < 16;
  const float r = decode_finish(lo ? sea : seb, lo ? swa : swb, delta);
  out_a = __shfl_sync(HZ_FULL, r, 0);
  out_b = __shfl_sync(HZ_FULL, r, 16);
}

template <typename T>
__device__ __forceinline__ float warp_support_decode(const T* __restrict__ x, const float* __restrict__ support,
                                                     int width, float delta, int64_t ld, int lane) {
  const RawRow8<T> raw = decode_vec_ok(x, ld) ? load_raw8(x, width, lane) : load_raw8_scalar(x, width, lane);
  return warp_decode8(raw, support, width, delta, lane);
}

}  // namespace hz
