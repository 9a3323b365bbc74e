// Persistent fused executor for a hz_gemm_plan: the whole chain of nn.Linear-shaped GEMMs of one
// recurrent_inference (7 steps for Hanabi-Full, 5 for Hanabi-Small; hanabizero_b200/plan.py) in ONE launch.
//
// Why: at 4096 rows every GEMM of the chain is a single wave of tiles whose cost is fill + drain, not
// math (~2 GFLOP, ~5 us per cuBLASLt launch, 35 us per chain against ~9 us of tensor time), and the chain
// is strictly sequential.  But every step is row-local: rows [128 r, 128 r + 128) of a step's output depend
// only on the same rows of its input.  So a CLUSTER of 4 CTAs owns one 128-row block for the whole chain:
// per step each CTA computes one column tile of the block, the tiles go to global memory (L2) and the four
// CTAs meet at a cluster-scope mbarrier — no grid-wide barrier, no cooperative launch.  The next step's
// weight tiles do not depend on that barrier and are fetched while the epilogue and the barrier run.
//
// Per step:  D[m][n] = act( A[m][k] . W[n][k]^T + bias[n] + C[m][n] ), fp16 in / fp32 accumulate / fp16 out,
// row-major, strided batches.  Tile = 128 rows x BN columns (BN = n/4 for plain steps, n for batched ones).
//   warp 0    TMA producer: W K-blocks [BN x 64] into an 80 KB slot ring; the A row block [128 x K] stays resident
//             (9 x 16 KB), its K-blocks fetched once per cluster and multicast to the four CTAs; 128-byte swizzle
//   warp 1    tcgen05.mma issuer (one lane), accumulators in TMEM (2 x 256 columns)
//   warps 2-17 epilogue: tcgen05.ld -> + bias + residual -> ReLU -> fp16 -> staging tile in shared memory (aliases the
//             idle A region) -> row-contiguous global stores -> fence -> cluster barrier arrival
// The fp32 plan (amp off) stays on cuBLASLt.
//
// STATUS: experimental, opt-in (HZ_FUSED_CHAIN=1).  Parity-tested (tests/test_chain_gpu.py) but at 4096 rows it runs
// the Hanabi-Full chain in 43.7 us against 33.8 us for the seven cuBLASLt launches (B200, in a CUDA graph).  Per step:
// A fetch + MMA 1.8 us, epilogue 1.6 us, and 3.2 us of exchange latency that no restructuring tried here removes —
// making the tile visible to the peers (fence after the stores, 1.7 us) and the barrier completing and waking the
// producers (1.5 us).  Measured and rejected on the way: a grid-wide barrier instead of clusters (54 us), TMA stores
// for the tiles (their completion wait costs 1.8 us), TMA multicast of the shared A block (no L2 saving at cluster
// size 4, +1 us latency), pushing tiles through distributed shared memory (17-21 B/clk per SM: slower than L2).
// profiles/README.md has the stamps; scripts/exp_fused_chain.py and scripts/micro/ reproduce them.
#include <cuda.h>
#include <cuda_fp16.h>

#include <stdlib.h>
#include <string.h>

#include "hz_chain.h"
#include "hz_common.cuh"

namespace hz {

constexpr int kBM = 128, kBK = 64, kMaxBN = 256, kCluster = 4;
constexpr int kASlots = 9, kABytes = kBM * kBK * 2;      // the whole A row block of a step: up to 9 K-blocks of 16 KB
constexpr int kWRegion = 80 * 1024;                       // W K-block slots of BN x 128 bytes
constexpr int kMaxWSlots = 10;
constexpr int kOffA = 0, kOffW = kOffA + kASlots * kABytes, kOffAux = kOffW + kWRegion;
constexpr int kSmemBytes = kOffAux + 2048 + 1024 /*alignment slack*/;   // 227 KB: the sm_100 maximum
constexpr int kEpiWarps = 16, kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 64 + kEpiThreads;                // producer warp, MMA warp, 16 epilogue warps
constexpr int kTmemCols = 512;
constexpr unsigned kSpinLimit = 1u << 26;                 // a stuck wait traps instead of hanging the GPU

struct ChainStep {
  CUtensorMap tm_a, tm_w;
  const __half* bias; int64_t stride_bias;
  const __half* c; int64_t ldc, stride_c;
  __half* d; int64_t ldd, stride_d;
  int32_t m, n, k, batch, relu;
  int32_t bn, tiles_n, tiles, num_k, wslots;   // tiles = column tiles x batches of ONE 128-row block
  uint32_t idesc, cpr_magic;
};

struct ChainParams {
  ChainStep step[kChainMaxSteps];
  int32_t n_steps, row_blocks;
  unsigned long long* trace;   // optional [gridDim][kChainMaxSteps][16] globaltimer stamps (HZ_CHAIN_TRACE=1)
};

__device__ __forceinline__ void stamp(const ChainParams& P, int step, int slot) {
  if (P.trace) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    P.trace[((size_t)blockIdx.x * kChainMaxSteps + step) * 16 + slot] = t;
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
// arrive on the barrier at the same shared-memory offset in CTA `rank` of this cluster (release, cluster scope)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
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
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (unsigned spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
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

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__global__ void __launch_bounds__(kThreads, 1) k_gemm_chain(const __grid_constant__ ChainParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // swizzle-128B tiles need 1024-byte alignment
  // the staging tile of the epilogue aliases the A region: both are idle whenever the other is in use
  const uint32_t a_base = base + kOffA, out_base = a_base, w_base = base + kOffW, aux = base + kOffAux;
  auto a_full = [&](int s) { return aux + 8u * s; };
  auto w_full = [&](int s) { return aux + 8u * (kASlots + s); };
  auto w_empty = [&](int s) { return aux + 8u * (kASlots + kMaxWSlots + s); };
  auto accf_bar = [&](int a) { return aux + 8u * (kASlots + 2 * kMaxWSlots + a); };
  auto acce_bar = [&](int a) { return aux + 8u * (kASlots + 2 * kMaxWSlots + 2 + a); };
  const uint32_t row_ready = aux + 8u * (kASlots + 2 * kMaxWSlots + 4);   // 4 arrivals: one per CTA of the cluster
  const uint32_t tmem_slot = row_ready + 8u;
  const uint32_t sbias = aux + 512u;                             // [2][kMaxBN] halfs: bias of the tile per accumulator
  // warp index through a shuffle: the compiler then knows the role branches are warp-uniform and keeps the
  // producer's and the MMA issuer's operands in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  uint32_t rank, cluster_id, n_clusters;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(cluster_id));
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(n_clusters));

  if (warp == 0 && lane == 0) {
    stamp(P, 0, 12);
    for (int s = 0; s < kASlots; ++s) mbar_init(a_full(s), 1);
    for (int s = 0; s < kMaxWSlots; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(accf_bar(a), 1); mbar_init(acce_bar(a), kEpiWarps); }
    mbar_init(row_ready, kCluster);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int s = 0; s < P.n_steps; ++s) { tma_prefetch_desc(&P.step[s].tm_a); tma_prefetch_desc(&P.step[s].tm_w); }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // every CTA's barriers are initialised before a peer may arrive on them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  // Every step boundary (also the one between two row blocks of this cluster) is a cluster barrier: once it has
  // completed, all four CTAs have finished their MMAs, drained their staging tile and published their part of the
  // step.  A slots are therefore free by construction and need no empty barriers.
  if (warp == 0) {
    // ===== TMA producer: the whole warp walks the loops, one elected lane issues =====
    const bool leader = elect_one();
    uint32_t w_par = 0xffffffffu;     // per W slot: parity to wait for on its empty barrier (fresh barrier: 1 passes)
    uint32_t tile_count = 0, rr_phase = 0;
    bool first = true;
    for (int rb = cluster_id; rb < P.row_blocks; rb += n_clusters) {
      for (int si = 0; si < P.n_steps; ++si) {
        const ChainStep& s = P.step[si];
        const int bn = s.bn, num_k = s.num_k, wslots = s.wslots, tiles_n = s.tiles_n;
        const uint32_t wbytes = (uint32_t)bn * kBK * 2;
        const bool mine = (int)rank < s.tiles;       // <= kCluster tiles: CTA `rank` owns tile `rank`
        int pre = 0, b = 0, tn = 0;
        if (mine) {
          b = (int)rank / tiles_n; tn = (int)rank - b * tiles_n;
          if (tile_count > 0) {   // the W region is re-partitioned per tile: all MMAs of the previous tile must be done
            const uint32_t prev = tile_count - 1;
            mbar_wait(accf_bar(prev & 1), (prev >> 1) & 1);
          }
          // weights do not depend on the other CTAs: put the first K-blocks' W tiles in flight before the barrier
          pre = num_k < wslots ? num_k : wslots;
          for (int j = 0; j < pre; ++j) {
            mbar_wait(w_empty(j), (w_par >> j) & 1u);
            w_par ^= 1u << j;
            if (leader) {
              mbar_expect_tx(w_full(j), wbytes);
              tma_load_3d(w_base + j * wbytes, &s.tm_w, w_full(j), j * kBK, tn * bn, b);
            }
          }
        }
        if (!first) {
          if (leader) stamp(P, si, 0);
          mbar_wait_cluster(row_ready, rr_phase);
          rr_phase ^= 1u;
          asm volatile("fence.proxy.async;" ::: "memory");
          if (leader) stamp(P, si, 1);
        }
        first = false;
        if (mine) {
          // (TMA multicast of the shared A row block was measured: no L2 saving at cluster size 4 and ~1 us more
          // latency than plain loads, so every CTA fetches its own copy)
          if (leader) {
            for (int kb = 0; kb < num_k; ++kb) {
              mbar_expect_tx(a_full(kb), kABytes);
              tma_load_3d(a_base + kb * kABytes, &s.tm_a, a_full(kb), kb * kBK, rb * kBM, b);
            }
            stamp(P, si, 7);
          }
          for (int kb = pre; kb < num_k; ++kb) {
            const int j = kb % wslots;
            mbar_wait(w_empty(j), (w_par >> j) & 1u);
            w_par ^= 1u << j;
            if (leader) {
              mbar_expect_tx(w_full(j), wbytes);
              tma_load_3d(w_base + j * wbytes, &s.tm_w, w_full(j), kb * kBK, tn * bn, b);
            }
          }
          ++tile_count;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp walks the loops, one elected lane issues =====
    const bool leader = elect_one();
    uint32_t a_par = 0, w_par = 0;    // per slot: parity of its full barrier's next completion
    int acc = 0; uint32_t acc_phase = 0;
    for (int rb = cluster_id; rb < P.row_blocks; rb += n_clusters) {
      for (int si = 0; si < P.n_steps; ++si) {
        const ChainStep& s = P.step[si];
        if ((int)rank >= s.tiles) continue;
        const int num_k = s.num_k, wslots = s.wslots;
        const uint32_t wbytes = (uint32_t)s.bn * kBK * 2, idesc = s.idesc;
        mbar_wait(acce_bar(acc), acc_phase ^ 1u);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * kMaxBN;
        uint64_t da = umma_desc_k128(a_base);
        const uint64_t dw0 = umma_desc_k128(w_base);
        int j = 0;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(a_full(kb), (a_par >> kb) & 1u);
          a_par ^= 1u << kb;
          mbar_wait(w_full(j), (w_par >> j) & 1u);
          w_par ^= 1u << j;
          tc_fence_after();
          if (leader) {
            const uint64_t dw = dw0 + (uint64_t)((j * wbytes) >> 4);
#pragma unroll
            for (int k16 = 0; k16 < kBK / 16; ++k16) {
              // +32 bytes along K inside the 128-byte swizzle atom = +2 in the descriptor's 16-byte units
              tc_mma_f16(d_tmem, da + (uint64_t)(2 * k16), dw + (uint64_t)(2 * k16), idesc, (kb | k16) != 0);
            }
            tc_commit(w_empty(j));                  // the W slot is free once these MMAs have read it
          }
          __syncwarp();
          da += (uint64_t)(kABytes >> 4);
          if (++j == wslots) j = 0;
        }
        if (leader) {
          tc_commit(accf_bar(acc));                 // accumulator complete -> epilogue (and the producer)
          stamp(P, si, 2);
        }
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ===== epilogue (warps 2..17): TMEM lane quarter = warp % 4; the four warps of a quarter split the columns =====
    const int q = warp & 3, part = (warp - 2) >> 2, et = threadIdx.x - 64;   // et: 0..511 among epilogue threads
    const int rloc = q * 32 + lane;                                          // row inside the 128-row block
    int acc = 0; uint32_t acc_phase = 0;
    for (int rb = cluster_id; rb < P.row_blocks; rb += n_clusters) {
      for (int si = 0; si < P.n_steps; ++si) {
        const ChainStep& s = P.step[si];
        if ((int)rank < s.tiles) {
          const int bn = s.bn, n16 = bn >> 4;
          const int c_begin = ((part * n16) >> 2) * 16, c_end = (((part + 1) * n16) >> 2) * 16;
          const int b = (int)rank / s.tiles_n, tn = (int)rank - b * s.tiles_n;
          const int col0 = tn * bn;
          // bias of this tile -> shared memory (broadcast reads later), overlapped with the MMAs still running
          const uint32_t sb = sbias + (uint32_t)acc * (kMaxBN * 2);
          if (et < bn) {
            const __half bv = s.bias ? s.bias[(size_t)b * s.stride_bias + col0 + et] : __float2half_rn(0.0f);
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(sb + 2u * et), "h"(__half_as_ushort(bv)) : "memory");
          }
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
          const int row = rb * kBM + rloc;
          const __half* crow = (s.c && row < s.m) ? s.c + (size_t)b * s.stride_c + (size_t)row * s.ldc + col0 : nullptr;
          const uint32_t t_addr = tmem_base + (uint32_t)acc * kMaxBN + ((uint32_t)(q * 32) << 16);
          uint4 res0 = make_uint4(0, 0, 0, 0), res1 = res0;
          if (crow && c_begin < c_end) {   // residual of the first chunk: in flight while the accumulator completes
            res0 = *reinterpret_cast<const uint4*>(crow + c_begin);
            res1 = *reinterpret_cast<const uint4*>(crow + c_begin + 8);
          }
          mbar_wait(accf_bar(acc), acc_phase);
          tc_fence_after();
          if (et == 0) stamp(P, si, 3);
          // staging tile: row-major, 16 bytes of padding per row so that a quarter-warp's 16-byte stores to eight
          // consecutive rows fall into distinct banks
          const uint32_t pitch = (uint32_t)bn * 2u + 16u;
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
            uint4 o0, o1;
            __half2* p0 = reinterpret_cast<__half2*>(&o0);
            __half2* p1 = reinterpret_cast<__half2*>(&o1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              p0[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
              p1[j] = __floats2half2_rn(v[8 + 2 * j], v[8 + 2 * j + 1]);
            }
            const uint32_t dst = out_base + (uint32_t)rloc * pitch + 2u * (uint32_t)c;
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o0.x), "r"(o0.y), "r"(o0.z), "r"(o0.w) : "memory");
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 16u), "r"(o1.x), "r"(o1.y), "r"(o1.z), "r"(o1.w) : "memory");
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
          if (et == 0) stamp(P, si, 6);
          // the whole tile is staged: copy it out with row-contiguous 16-byte stores (a warp covers whole rows)
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
          const int cpr = bn >> 3;                                     // 16-byte chunks per row
          const int rows_here = s.m - rb * kBM < kBM ? s.m - rb * kBM : kBM;
          __half* dtile = s.d + (size_t)b * s.stride_d + (size_t)(rb * kBM) * s.ldd + col0;
          const uint32_t magic = s.cpr_magic;                          // idx / cpr == (idx * magic) >> 20 for idx < 128 * cpr
          for (int idx = et; idx < rows_here * cpr; idx += kEpiThreads) {
            const int rr = (int)(((uint32_t)idx * magic) >> 20), ch = idx - rr * cpr;
            uint4 o;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w)
                         : "r"(out_base + (uint32_t)rr * pitch + 16u * (uint32_t)ch));
            *reinterpret_cast<uint4*>(dtile + (size_t)rr * s.ldd + 8 * ch) = o;
          }
          if (et == 0) stamp(P, si, 4);
        }
        // publish: every thread fences its own stores (gpu scope, and towards the async proxy the peers' TMA loads
        // read through), the CTA's epilogue threads meet, one of them tells the four CTAs of the cluster
        const bool last = si + 1 == P.n_steps && rb + (int)n_clusters >= P.row_blocks;
        if (!last) {
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
          asm volatile("fence.proxy.async;" ::: "memory");
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
          if (et == 0) {
            for (uint32_t r_ = 0; r_ < kCluster; ++r_) mbar_arrive_remote(row_ready, r_);
            stamp(P, si, 5);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // no CTA exits while a peer may still arrive on its barriers
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

// fp16 [batch][rows][cols] matrix with row stride ld and batch stride `stride` (elements); box = [1][box_rows][64]
static bool make_map(CUtensorMap* map, const void* ptr, int64_t cols, int64_t rows, int64_t batch, int64_t ld,
                     int64_t stride, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(batch > 1 ? stride : ld) * 2};
  cuuint32_t box[3] = {(cuuint32_t)kBK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// column tile of a step: one tile per batch for batched steps, otherwise n / 4 (one tile per CTA of the cluster) when
// that is a multiple of 64 (whole 64-column store blocks), else the widest multiple of 64 dividing n, else n itself
static int pick_bn(int n, int batch) {
  if (n <= kMaxBN && (batch > 1 || n <= 64)) return n;
  if (n % kCluster == 0 && (n / kCluster) % 64 == 0 && n / kCluster <= kMaxBN) return n / kCluster;
  for (int bn = kMaxBN; bn >= 64; bn -= 64)
    if (n % bn == 0) return bn;
  return n <= kMaxBN ? n : 0;
}

struct ChainExec {
  ChainParams params;
  int clusters = 0;
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
    if (s.m != steps[0].m) return no("steps with different row counts");
    if (s.n % 16 || s.k % 8 || s.lda % 8 || s.ldw % 8 || s.ldd % 8 || (s.c && s.ldc % 8)) return no("shape not a multiple of 16 / 8");
    if (s.batch > 1 && (s.stride_a % 8 || s.stride_w % 8 || s.stride_d % 8 || (s.c && s.stride_c % 8) ||
                        (s.bias && s.stride_bias % 8))) return no("batch stride not a multiple of 8");
    if (((uintptr_t)s.a | (uintptr_t)s.w | (uintptr_t)s.d | (uintptr_t)s.c | (uintptr_t)s.bias) & 15) return no("unaligned pointer");
    const int bn = pick_bn(s.n, s.batch);
    if (bn == 0) return no("no column tile for n");
    if ((s.n / bn) * s.batch > kCluster) return no("more than one tile per CTA of the cluster in a step");
    if ((s.k + kBK - 1) / kBK > kASlots) return no("K exceeds the resident A row block (576)");
  }
  return true;
}

int chain_create(ChainExec** out, int device, const hz_gemm_step* steps, int n_steps) {
  ChainExec* e = new ChainExec;
  memset(&e->params, 0, sizeof(e->params));
  for (int i = 0; i < n_steps; ++i) {
    const hz_gemm_step& s = steps[i];
    ChainStep& c = e->params.step[i];
    c.bias = (const __half*)s.bias; c.stride_bias = s.stride_bias;
    c.c = (const __half*)s.c; c.ldc = s.ldc; c.stride_c = s.stride_c;
    c.d = (__half*)s.d; c.ldd = s.ldd; c.stride_d = s.stride_d;
    c.m = s.m; c.n = s.n; c.k = s.k; c.batch = s.batch; c.relu = s.relu;
    c.bn = pick_bn(s.n, s.batch);
    c.tiles_n = s.n / c.bn;
    c.tiles = c.tiles_n * s.batch;
    c.num_k = (s.k + kBK - 1) / kBK;
    c.wslots = kWRegion / (c.bn * kBK * 2);
    if (c.wslots > kMaxWSlots) c.wslots = kMaxWSlots;
    c.cpr_magic = ((1u << 20) + (uint32_t)(c.bn >> 3) - 1) / (uint32_t)(c.bn >> 3);
    c.idesc = (1u << 4) | ((uint32_t)(c.bn >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);   // f32 accumulate, f16 x f16, K-major
    if (!make_map(&c.tm_a, s.a, s.k, s.m, s.batch, s.lda, s.stride_a, kBM) ||
        !make_map(&c.tm_w, s.w, s.k, s.n, s.batch, s.ldw, s.stride_w, c.bn)) {
      delete e;
      set_error("fused chain: cuTensorMapEncodeTiled failed for step %d", i);
      return HZ_ERR_CUDA;
    }
  }
  e->params.n_steps = n_steps;
  e->params.row_blocks = (steps[0].m + kBM - 1) / kBM;
  cudaError_t err = cudaFuncSetAttribute(k_gemm_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  int max_clusters = 0;
  if (err == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kCluster);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    err = cudaOccupancyMaxActiveClusters(&max_clusters, k_gemm_chain, &cfg);
  }
  if (err != cudaSuccess || max_clusters <= 0) {
    delete e;
    if (err == cudaSuccess) { set_error("fused chain: no cluster of %d CTAs fits on this device", kCluster); return HZ_ERR_CUDA; }
    return fail_cuda(err, "fused chain: setup");
  }
  e->clusters = e->params.row_blocks < max_clusters ? e->params.row_blocks : max_clusters;
  if (const char* tr = getenv("HZ_CHAIN_TRACE")) {
    if (tr[0] == '1') {
      const size_t bytes = (size_t)e->clusters * kCluster * kChainMaxSteps * 16 * sizeof(unsigned long long);
      if (cudaMalloc(&e->trace, bytes) == cudaSuccess) cudaMemset(e->trace, 0, bytes);
      e->params.trace = e->trace;
    }
  }
  *out = e;
  return HZ_OK;
}

void chain_destroy(ChainExec* e) {
  if (!e) return;
  cudaFree(e->trace);
  delete e;
}

int chain_run(ChainExec* e, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(e->clusters * kCluster);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;   // a cluster = the four column tiles of one 128-row block
  attr[0].val.clusterDim.x = kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t err = cudaLaunchKernelEx(&cfg, k_gemm_chain, e->params);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (err != cudaSuccess) return fail_cuda(err, "k_gemm_chain");
  return HZ_OK;
}

int chain_grid(const ChainExec* e) { return e ? e->clusters * kCluster : 0; }

int chain_trace(const ChainExec* e, unsigned long long* host_out, size_t count) {
  if (!e || !e->trace) return 0;
  const size_t have = (size_t)e->clusters * kCluster * kChainMaxSteps * 16;
  cudaMemcpy(host_out, e->trace, (count < have ? count : have) * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  return (int)have;
}

}  // namespace hz
