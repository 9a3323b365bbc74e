// Small device-side glue around the PyTorch network so that a simulation stays on the GPU with few
// launches (the GEMMs themselves remain cuBLAS through torch):
//   hz_support_decode  softmax over the categorical value/reward support -> expectation ->
//                      inverse scalar transform, NaN -> 0  (/root/reference/core/config.py:210-232,
//                      which the reference runs as ~10 torch kernels and then copies to the host)
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

#pragma GCC visibility pop
}  // extern "C"
