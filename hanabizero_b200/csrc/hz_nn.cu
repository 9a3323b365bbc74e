// Small device-side glue around the PyTorch network so that a simulation stays on the GPU with few
// launches (the GEMMs themselves remain cuBLAS through torch):
//   hz_support_decode  softmax over the categorical value/reward support -> expectation ->
//                      inverse scalar transform, NaN -> 0  (/root/reference/core/config.py:210-232,
//                      which the reference runs as ~10 torch kernels and then copies to the host)
//   hz_bias_act        GEMM epilogue: out = relu?(x + bias + residual + table[idx[row]])
#include "hz_decode.cuh"

namespace hz {

// one warp per row of logits[rows][width]
template <typename T>
__global__ void __launch_bounds__(128) k_support_decode(const T* __restrict__ logits,
                                                        const float* __restrict__ support,
                                                        float* __restrict__ out, int rows, int width,
                                                        int64_t ld, float delta) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float r = warp_support_decode<T>(logits + (size_t)row * ld, support, width, delta, ld, lane);
  if (lane == 0) out[row] = r;
}

// out[row][c] = act(x[row][c] + bias[c] + residual[row][c] + table[idx[row]][c]); V columns per thread
template <typename T, int V>
struct alignas(sizeof(T) * V) Vec {
  T v[V];
};

template <typename T, int V>
__global__ void __launch_bounds__(256) k_bias_act(T* __restrict__ out, int64_t ld_out,
                                                  const T* __restrict__ x, int64_t ld_x,
                                                  const T* __restrict__ bias,
                                                  const T* __restrict__ residual, int64_t ld_res,
                                                  const T* __restrict__ table,
                                                  const int64_t* __restrict__ idx, int rows, int cols,
                                                  int relu) {
  using VT = Vec<T, V>;
  const int cv = cols / V;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * cv) return;
  const int row = (int)(i / cv), c = (int)(i - (int64_t)row * cv) * V;
  float acc[V];
  const VT xv = *reinterpret_cast<const VT*>(x + (size_t)row * ld_x + c);
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] = to_f(xv.v[k]);
  if (bias) {
    const VT b = *reinterpret_cast<const VT*>(bias + c);
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] += to_f(b.v[k]);
  }
  if (residual) {
    const VT r = *reinterpret_cast<const VT*>(residual + (size_t)row * ld_res + c);
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] += to_f(r.v[k]);
  }
  if (table) {
    const VT t = *reinterpret_cast<const VT*>(table + (size_t)idx[row] * cols + c);
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] += to_f(t.v[k]);
  }
  VT o;
#pragma unroll
  for (int k = 0; k < V; ++k) o.v[k] = from_f<T>(relu ? fmaxf(acc[k], 0.0f) : acc[k]);
  *reinterpret_cast<VT*>(out + (size_t)row * ld_out + c) = o;
}

template <typename T>
static int launch_bias_act(cudaStream_t s, void* out, int64_t ld_out, const void* x, int64_t ld_x, const void* bias,
                           const void* residual, int64_t ld_res, const void* table, const int64_t* idx,
                           int rows, int cols, int relu) {
  constexpr int V = 16 / sizeof(T);
  auto al = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  const bool vec = cols % V == 0 && ld_out % V == 0 && ld_x % V == 0 && (!residual || ld_res % V == 0) && al(out) &&
                   al(x) && al(bias) && al(residual) && al(table);
  if (vec) {
    const int64_t total = (int64_t)rows * (cols / V);
    k_bias_act<T, V><<<(unsigned)((total + 255) / 256), 256, 0, s>>>((T*)out, ld_out, (const T*)x, ld_x, (const T*)bias,
                                                                     (const T*)residual, ld_res, (const T*)table, idx, rows, cols, relu);
  } else {
    const int64_t total = (int64_t)rows * cols;
    k_bias_act<T, 1><<<(unsigned)((total + 255) / 256), 256, 0, s>>>((T*)out, ld_out, (const T*)x, ld_x, (const T*)bias,
                                                                     (const T*)residual, ld_res, (const T*)table, idx, rows, cols, relu);
  }
  HZ_LAUNCH_CHECK("k_bias_act");
  return HZ_OK;
}

}  // namespace hz

using namespace hz;

extern "C" {
#pragma GCC visibility push(default)

int hz_support_decode(void* stream, const void* logits, int elem_bytes, const float* support,
                      float* out, int rows, int width, int64_t ld, float delta) {
  if (!logits || !support || !out || rows <= 0 || width <= 0 || width > 256 || ld < width ||
      (elem_bytes != 2 && elem_bytes != 4)) {
    set_error("hz_support_decode: bad argument");
    return HZ_ERR_ARG;
  }
  dim3 grid((rows + 3) / 4), block(128);
  if (elem_bytes == 4) {
    k_support_decode<float><<<grid, block, 0, (cudaStream_t)stream>>>((const float*)logits, support, out, rows, width, ld, delta);
  } else {
    k_support_decode<__half><<<grid, block, 0, (cudaStream_t)stream>>>((const __half*)logits, support, out, rows, width, ld, delta);
  }
  HZ_LAUNCH_CHECK("k_support_decode");
  return HZ_OK;
}

int hz_bias_act(void* stream, void* out, int64_t ld_out, const void* x, int64_t ld_x, const void* bias,
                const void* residual, int64_t ld_res, const void* table, const int64_t* idx,
                int rows, int cols, int relu, int elem_bytes) {
  if (!out || !x || rows <= 0 || cols <= 0 || (table && !idx) || (elem_bytes != 2 && elem_bytes != 4)) {
    set_error("hz_bias_act: bad argument");
    return HZ_ERR_ARG;
  }
  if (elem_bytes == 4) {
    return launch_bias_act<float>((cudaStream_t)stream, out, ld_out, x, ld_x, bias, residual, ld_res, table, idx, rows, cols, relu);
  }
  return launch_bias_act<__half>((cudaStream_t)stream, out, ld_out, x, ld_x, bias, residual, ld_res, table, idx, rows, cols, relu);
}

#pragma GCC visibility pop
}  // extern "C"
