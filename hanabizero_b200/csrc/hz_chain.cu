// Persistent fused executor for a hz_gemm_plan: the whole chain of nn.Linear-shaped GEMMs of one
// recurrent_inference (7 steps for Hanabi-Full, 5 for Hanabi-Small; hanabizero_b200/plan.py) in ONE launch.
//
// Why: at 4096 rows every GEMM of the chain is a single wave of tiles whose cost is fill + drain, not
// math (~2 GFLOP, ~5 us per cuBLASLt launch, 35 us per chain against ~9 us of tensor time), and the chain
// is strictly sequential.  One resident CTA per SM walks all steps; between steps the CTAs meet at a
// grid barrier in global memory, and the next step's weight tiles (which do not depend on the barrier)
// are already in flight while the CTA waits.
//
// Per step:  D[m][n] = act( A[m][k] . W[n][k]^T + bias[n] + C[m][n] ), fp16 in / fp32 accumulate / fp16 out,
// row-major, strided batches.  Tiles of 128 rows x BN columns (BN a multiple of 16 chosen per step).
//   warp 0    TMA producer: A [128 x 64] and W [BN x 64] K-blocks, 128-byte swizzle, 4-stage ring
//   warp 1    tcgen05.mma issuer (one elected lane), accumulators in TMEM (2 x 256 columns)
//   warps 2-9 epilogue: tcgen05.ld -> + bias + residual -> ReLU -> fp16 -> global
// The fp32 plan (amp off) stays on cuBLASLt.
#include <cuda.h>
#include <cuda_fp16.h>

#include <stdlib.h>
#include <string.h>

#include "hz_chain.h"
#include "hz_common.cuh"

namespace hz {

constexpr int kBM = 128, kBK = 64, kStages = 4, kMaxBN = 256;
constexpr int kABytes = kBM * kBK * 2;          // 16 KB
constexpr int kWBytes = kMaxBN * kBK * 2;       // 32 KB
constexpr int kStageBytes = kABytes + kWBytes;  // 48 KB
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*alignment slack*/ + 2048 /*barriers, bias*/;
constexpr int kThreads = 320;            // producer warp, MMA warp, 8 epilogue warps
constexpr int kTmemCols = 512;
constexpr unsigned kSpinLimit = 1u << 28;       // a stuck wait traps instead of hanging the GPU

struct ChainStep {
  CUtensorMap tm_a, tm_w;
  const __half* bias; int64_t stride_bias;
  const __half* c; int64_t ldc, stride_c;
  __half* d; int64_t ldd, stride_d;
  int32_t m, n, k, batch, relu;
  int32_t bn, tiles_m, tiles_n, tiles, num_k;
  uint32_t idesc;
  uint32_t pad;
};

struct ChainParams {
  ChainStep step[kChainMaxSteps];
  int32_t n_steps;
  uint32_t* sync;   // [kChainMaxSteps] arrival counters, all zero between launches
  unsigned long long* trace;   // optional [gridDim][kChainMaxSteps][8] globaltimer stamps (HZ_CHAIN_TRACE=1)
};

__device__ __forceinline__ void stamp(const ChainParams& P, int step, int slot) {
  if (P.trace) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    P.trace[((size_t)blockIdx.x * kChainMaxSteps + step) * 8 + slot] = t;
  }
}

// ---- PTX wrappers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (unsigned spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile in shared memory, rows of 128 bytes, 128-byte swizzle (what TMA wrote):
// 8-row groups are 1024 bytes apart (SBO), descriptor version 1 (sm_100), layout type 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffffu) >> 4);        // start address, 16-byte units
  d |= (uint64_t)1 << 16;                          // leading byte offset (unused with swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset
  d |= (uint64_t)1 << 46;                          // version
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void tile_coords(const ChainStep& s, int t, int& b, int& tm, int& tn) {
  tn = t % s.tiles_n;
  const int r = t / s.tiles_n;
  tm = r % s.tiles_m;
  b = r / s.tiles_m;
}

// grid barrier: arrival on counter[step]; the last arriver clears the previous step's counter (every CTA has
// passed its wait on it by then) and, at the last step, its own, so all counters are zero again at exit
__device__ __forceinline__ void grid_arrive(uint32_t* sync, int step, int n_steps) {
  uint32_t old;   // release: cumulative over the other epilogue threads' fenced stores this thread met at the CTA barrier
  asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(sync + step) : "memory");
  if (old == gridDim.x - 1) {
    if (step > 0) atomicExch(&sync[step - 1], 0u);
    if (step == n_steps - 1) atomicExch(&sync[step], 0u);
  }
}
__device__ __forceinline__ void grid_wait(const uint32_t* sync, int step) {
  unsigned v = 0;
  for (unsigned spin = 0;; ++spin) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(sync + step) : "memory");
    if (v >= gridDim.x) break;
    if (spin > kSpinLimit) __trap();
  }
  asm volatile("fence.proxy.async;" ::: "memory");
}

__global__ void __launch_bounds__(kThreads, 1) k_gemm_chain(const __grid_constant__ ChainParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // swizzle-128B tiles need 1024-byte alignment
  const uint32_t bars = base + kStages * kStageBytes;            // full[4], empty[4], acc_full[2], acc_empty[2], tmem ptr
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  auto accf_bar = [&](int a) { return bars + 8u * (2 * kStages + a); };
  auto acce_bar = [&](int a) { return bars + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * kStages + 4);
  const uint32_t sbias = bars + 256u;                            // [2][kMaxBN] halfs: bias of the tile per accumulator
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x, cta = blockIdx.x;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(accf_bar(a), 1); mbar_init(acce_bar(a), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int s = 0; s < P.n_steps; ++s) { tma_prefetch_desc(&P.step[s].tm_a); tma_prefetch_desc(&P.step[s].tm_w); }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int si = 0; si < P.n_steps; ++si) {
        const ChainStep& s = P.step[si];
        const uint32_t bytes = kABytes + (uint32_t)s.bn * kBK * 2;
        bool passed = (si == 0);   // step 0 reads what the previous kernel wrote: stream order covers it
        for (int t = cta; t < s.tiles; t += G) {
          int b, tm, tn;
          tile_coords(s, t, b, tm, tn);
          int kb = 0;
          if (!passed) {
            // weights do not depend on the other CTAs: put the first K-blocks' W tiles in flight, then wait
            // for every CTA to have finished the previous step, then fetch the matching A tiles
            const int pre = s.num_k < kStages ? s.num_k : kStages;
            int st = stage; uint32_t ph = phase;
            for (int j = 0; j < pre; ++j) {
              mbar_wait(empty_bar(st), ph ^ 1u);
              mbar_expect_tx(full_bar(st), bytes);
              tma_load_3d(base + st * kStageBytes + kABytes, &s.tm_w, full_bar(st), j * kBK, tn * s.bn, b);
              if (++st == kStages) { st = 0; ph ^= 1u; }
            }
            stamp(P, si, 0);
            grid_wait(P.sync, si - 1);
            stamp(P, si, 1);
            passed = true;
            for (int j = 0; j < pre; ++j) {
              tma_load_3d(base + stage * kStageBytes, &s.tm_a, full_bar(stage), j * kBK, tm * kBM, b);
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
            kb = pre;
          }
          for (; kb < s.num_k; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_expect_tx(full_bar(stage), bytes);
            tma_load_3d(base + stage * kStageBytes + kABytes, &s.tm_w, full_bar(stage), kb * kBK, tn * s.bn, b);
            tma_load_3d(base + stage * kStageBytes, &s.tm_a, full_bar(stage), kb * kBK, tm * kBM, b);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int si = 0; si < P.n_steps; ++si) {
        const ChainStep& s = P.step[si];
        for (int t = cta; t < s.tiles; t += G) {
          mbar_wait(acce_bar(acc), acc_phase ^ 1u);   // epilogue has drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)acc * kMaxBN;
          for (int kb = 0; kb < s.num_k; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t a_addr = base + stage * kStageBytes, w_addr = a_addr + kABytes;
            const uint64_t da = umma_desc_k128(a_addr), dw = umma_desc_k128(w_addr);
#pragma unroll
            for (int k16 = 0; k16 < kBK / 16; ++k16) {
              // +32 bytes along K inside the 128-byte swizzle atom = +2 in the descriptor's 16-byte units
              tc_mma_f16(d_tmem, da + (uint64_t)(2 * k16), dw + (uint64_t)(2 * k16), s.idesc, (kb | k16) != 0);
            }
            tc_commit(empty_bar(stage));              // frees the stage once these MMAs have read it
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          tc_commit(accf_bar(acc));                   // accumulator complete -> epilogue
          stamp(P, si, 2);
          if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
      }
    }
  } else {
    // ===== epilogue (warps 2..9): TMEM lane quarter = warp % 4; the two warps of a quarter split the columns =====
    const int q = warp & 3, half_id = (warp - 2) >> 2, et = threadIdx.x - 64;   // et: 0..255 among epilogue threads
    int acc = 0; uint32_t acc_phase = 0;
    for (int si = 0; si < P.n_steps; ++si) {
      const ChainStep& s = P.step[si];
      const int n16 = s.bn >> 4, n16_lo = (n16 + 1) >> 1;
      const int c_begin = half_id == 0 ? 0 : n16_lo * 16, c_end = half_id == 0 ? n16_lo * 16 : s.bn;
      for (int t = cta; t < s.tiles; t += G) {
        int b, tm, tn;
        tile_coords(s, t, b, tm, tn);
        const int col0 = tn * s.bn;
        // bias of this tile -> shared memory (broadcast reads later), overlapped with the MMAs still running
        const uint32_t sb = sbias + (uint32_t)acc * (kMaxBN * 2);
        if (et < s.bn) {
          const __half bv = s.bias ? s.bias[(size_t)b * s.stride_bias + col0 + et] : __float2half_rn(0.0f);
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(sb + 2u * et), "h"(__half_as_ushort(bv)) : "memory");
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int row = tm * kBM + q * 32 + lane;
        const bool row_ok = row < s.m;
        __half* drow = s.d + (size_t)b * s.stride_d + (size_t)row * s.ldd + col0;
        const __half* crow = (s.c && row_ok) ? s.c + (size_t)b * s.stride_c + (size_t)row * s.ldc + col0 : nullptr;
        const uint32_t t_addr = tmem_base + (uint32_t)acc * kMaxBN + ((uint32_t)(q * 32) << 16);
        uint4 res0 = make_uint4(0, 0, 0, 0), res1 = res0;
        if (crow && c_begin < c_end) {   // residual of the first chunk: in flight while the accumulator completes
          res0 = *reinterpret_cast<const uint4*>(crow + c_begin);
          res1 = *reinterpret_cast<const uint4*>(crow + c_begin + 8);
        }
        mbar_wait(accf_bar(acc), acc_phase);
        tc_fence_after();
        if (warp == 2 && lane == 0) stamp(P, si, 3);
        uint32_t r[16], rn[16];
        if (c_begin < c_end) {
          tc_ld16(t_addr + (uint32_t)c_begin, r);
          tc_wait_ld();
        }
        for (int c = c_begin; c < c_end; c += 16) {
          const bool more = c + 16 < c_end;
          uint4 nres0 = make_uint4(0, 0, 0, 0), nres1 = nres0;
          if (more) {
            tc_ld16(t_addr + (uint32_t)(c + 16), rn);    // next chunk's accumulators and residual while this one is processed
            if (crow) {
              nres0 = *reinterpret_cast<const uint4*>(crow + c + 16);
              nres1 = *reinterpret_cast<const uint4*>(crow + c + 24);
            }
          }
          float v[16];
          uint4 b0, b1;
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(b0.x), "=r"(b0.y), "=r"(b0.z), "=r"(b0.w) : "r"(sb + 2u * c));
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(b1.x), "=r"(b1.y), "=r"(b1.z), "=r"(b1.w) : "r"(sb + 2u * c + 16u));
          {
            const __half2* h0 = reinterpret_cast<const __half2*>(&b0);
            const __half2* h1 = reinterpret_cast<const __half2*>(&b1);
            const __half2* g0 = reinterpret_cast<const __half2*>(&res0);
            const __half2* g1 = reinterpret_cast<const __half2*>(&res1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f0 = __half22float2(h0[j]), f1 = __half22float2(h1[j]);
              const float2 e0 = __half22float2(g0[j]), e1 = __half22float2(g1[j]);
              v[2 * j] = __uint_as_float(r[2 * j]) + f0.x + e0.x;
              v[2 * j + 1] = __uint_as_float(r[2 * j + 1]) + f0.y + e0.y;
              v[8 + 2 * j] = __uint_as_float(r[8 + 2 * j]) + f1.x + e1.x;
              v[8 + 2 * j + 1] = __uint_as_float(r[8 + 2 * j + 1]) + f1.y + e1.y;
            }
          }
          if (s.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.0f);
          }
          if (row_ok) {
            uint4 o0, o1;
            __half2* p0 = reinterpret_cast<__half2*>(&o0);
            __half2* p1 = reinterpret_cast<__half2*>(&o1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              p0[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
              p1[j] = __floats2half2_rn(v[8 + 2 * j], v[8 + 2 * j + 1]);
            }
            *reinterpret_cast<uint4*>(drow + c) = o0;
            *reinterpret_cast<uint4*>(drow + c + 8) = o1;
          }
          if (more) {
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = rn[j];
            res0 = nres0;
            res1 = nres1;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acce_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
      // this CTA's part of the step is on its way to global memory: every thread publishes its own stores (gpu scope,
      // and towards the async proxy that the other SMs' TMA loads read through), then one thread joins the grid barrier
      __threadfence();
      asm volatile("fence.proxy.async;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (warp == 2 && lane == 0) {
        stamp(P, si, 4);
        grid_arrive(P.sync, si, P.n_steps);
        stamp(P, si, 5);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      p = nullptr;
    }
    cudaGetLastError();
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// fp16 [batch][rows][k] operand with row stride ld and batch stride `stride` (elements); box = [1][box_rows][64]
static bool make_map(CUtensorMap* map, const void* ptr, int64_t k, int64_t rows, int64_t batch, int64_t ld,
                     int64_t stride, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(batch > 1 ? stride : ld) * 2};
  cuuint32_t box[3] = {(cuuint32_t)kBK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int pick_bn(int n, int tiles_m_batch, int G) {
  if (n < 64) return n;
  int best = 0;
  for (int bn = 64; bn <= kMaxBN; bn += 16) {
    if (n % bn) continue;
    if (best == 0 || (int64_t)tiles_m_batch * (n / bn) > G) best = bn;   // smallest BN whose tile count fits one wave,
    if ((int64_t)tiles_m_batch * (n / bn) <= G) { best = bn; break; }    // else the largest divisor
  }
  return best;
}

struct ChainExec {
  ChainParams params;
  int grid = 0;
  uint32_t* sync = nullptr;
  unsigned long long* trace = nullptr;
};

bool chain_supported(const hz_gemm_step* steps, int n_steps, int elem_bytes, const char** why) {
  static const char* reason = "";
  auto no = [&](const char* r) { reason = r; if (why) *why = reason; return false; };
  if (elem_bytes != 2) return no("fp16 plans only");
  if (n_steps > kChainMaxSteps) return no("too many steps");
  const char* env = getenv("HZ_FUSED_CHAIN");
  if (!env || env[0] != '1') return no("opt-in: set HZ_FUSED_CHAIN=1");
  if (!encode_fn()) return no("cuTensorMapEncodeTiled unavailable");
  for (int i = 0; i < n_steps; ++i) {
    const hz_gemm_step& s = steps[i];
    if (s.n % 16 || s.k % 8 || s.lda % 8 || s.ldw % 8 || s.ldd % 8 || (s.c && s.ldc % 8)) return no("shape not a multiple of 16 / 8");
    if (s.batch > 1 && (s.stride_a % 8 || s.stride_w % 8 || s.stride_d % 8 || (s.c && s.stride_c % 8) ||
                        (s.bias && s.stride_bias % 8))) return no("batch stride not a multiple of 8");
    if (((uintptr_t)s.a | (uintptr_t)s.w | (uintptr_t)s.d | (uintptr_t)s.c | (uintptr_t)s.bias) & 15) return no("unaligned pointer");
    if (s.n > 64 && pick_bn(s.n, 1, 1) == 0) return no("no column tile divides n");
  }
  return true;
}

int chain_create(ChainExec** out, int device, const hz_gemm_step* steps, int n_steps) {
  cudaDeviceProp prop;
  HZ_CUDA(cudaGetDeviceProperties(&prop, device));
  ChainExec* e = new ChainExec;
  memset(&e->params, 0, sizeof(e->params));
  const int G = prop.multiProcessorCount;
  int max_tiles = 1;
  for (int i = 0; i < n_steps; ++i) {
    const hz_gemm_step& s = steps[i];
    ChainStep& c = e->params.step[i];
    c.bias = (const __half*)s.bias; c.stride_bias = s.stride_bias;
    c.c = (const __half*)s.c; c.ldc = s.ldc; c.stride_c = s.stride_c;
    c.d = (__half*)s.d; c.ldd = s.ldd; c.stride_d = s.stride_d;
    c.m = s.m; c.n = s.n; c.k = s.k; c.batch = s.batch; c.relu = s.relu;
    c.tiles_m = (s.m + kBM - 1) / kBM;
    c.bn = pick_bn(s.n, c.tiles_m * s.batch, G);
    c.tiles_n = s.n / c.bn;
    c.tiles = c.tiles_m * c.tiles_n * s.batch;
    c.num_k = (s.k + kBK - 1) / kBK;
    c.idesc = (1u << 4) | ((uint32_t)(c.bn >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);   // f32 accumulate, f16 x f16, K-major
    if (!make_map(&c.tm_a, s.a, s.k, s.m, s.batch, s.lda, s.stride_a, kBM) ||
        !make_map(&c.tm_w, s.w, s.k, s.n, s.batch, s.ldw, s.stride_w, c.bn)) {
      delete e;
      set_error("fused chain: cuTensorMapEncodeTiled failed for step %d", i);
      return HZ_ERR_CUDA;
    }
    if (c.tiles > max_tiles) max_tiles = c.tiles;
  }
  e->params.n_steps = n_steps;
  e->grid = max_tiles < G ? max_tiles : G;
  cudaError_t err = cudaMalloc(&e->sync, kChainMaxSteps * sizeof(uint32_t));
  if (err == cudaSuccess) err = cudaMemset(e->sync, 0, kChainMaxSteps * sizeof(uint32_t));
  if (err == cudaSuccess) err = cudaFuncSetAttribute(k_gemm_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (err != cudaSuccess) {
    cudaFree(e->sync);
    delete e;
    return fail_cuda(err, "fused chain: setup");
  }
  e->params.sync = e->sync;
  if (const char* tr = getenv("HZ_CHAIN_TRACE")) {
    if (tr[0] == '1') {
      const size_t bytes = (size_t)e->grid * kChainMaxSteps * 8 * sizeof(unsigned long long);
      if (cudaMalloc(&e->trace, bytes) == cudaSuccess) cudaMemset(e->trace, 0, bytes);
      e->params.trace = e->trace;
    }
  }
  *out = e;
  return HZ_OK;
}

void chain_destroy(ChainExec* e) {
  if (!e) return;
  cudaFree(e->sync);
  cudaFree(e->trace);
  delete e;
}

int chain_run(ChainExec* e, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(e->grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs must be co-resident: they meet at grid barriers
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t err = cudaLaunchKernelEx(&cfg, k_gemm_chain, e->params);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (err != cudaSuccess) return fail_cuda(err, "k_gemm_chain");
  return HZ_OK;
}

int chain_grid(const ChainExec* e) { return e ? e->grid : 0; }

int chain_trace(const ChainExec* e, unsigned long long* host_out, size_t count) {
  if (!e || !e->trace) return 0;
  const size_t have = (size_t)e->grid * kChainMaxSteps * 8;
  cudaMemcpy(host_out, e->trace, (count < have ? count : have) * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  return (int)have;
}

}  // namespace hz
