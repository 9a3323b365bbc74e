// GEMM-chain executor: the network's Linear layers as a fixed list of cuBLASLt matmuls with fused
// epilogues (bias, ReLU, residual add, strided batches), launched back to back on one stream.
//
// "The network stays in PyTorch" as far as weights, training and the module are concerned; at
// search time its eval-mode forward is this chain of plain library GEMMs (cuBLASLt owns the tensor
// cores).  What the chain buys is launch count: PyTorch needs ~70 kernels per recurrent_inference,
// torch-level fusion ~16, this chain 7 (Full) / 5 (Small) — and a simulation is latency-bound
// (SURVEY.md §7.4-9), so launches are the cost.
//
// Row-major convention of nn.Linear: D[M,N] = act(A[M,K] · W[N,K]^T + bias[N] + C[M,N]).  cuBLASLt is
// column-major: D^T[N,M] = W(col-major view [K,N], op T) · A^T(col-major view [K,M], op N).
#include <cublasLt.h>

#include <stdlib.h>

#include <vector>

#include "hz_common.cuh"

namespace hz {

struct LtStep {
  hz_gemm_step s;
  cublasLtMatmulDesc_t op = nullptr;
  cublasLtMatrixLayout_t la = nullptr, lb = nullptr, lc = nullptr, ld = nullptr;
  cublasLtMatmulAlgo_t algo;
  float beta = 0.f;
};

static size_t align_of(const void* p, int64_t ld_elems, int elem_bytes) {
  size_t a = 256;
  auto shrink = [&](uintptr_t v) {
    while (a > 1 && (v % a) != 0) a >>= 1;
  };
  shrink((uintptr_t)p);
  shrink((uintptr_t)(ld_elems * elem_bytes));
  return a;
}

}  // namespace hz

using namespace hz;

static std::atomic<int64_t> g_gemm_launches{0};

struct hz_gemm_plan {
  int device = 0, elem_bytes = 2;
  cublasLtHandle_t lt = nullptr;
  void* workspace = nullptr;
  size_t ws_bytes = 0;
  bool autotune = false;
  int sm_target = 0;   // 0 = the whole device
  std::vector<LtStep> steps;
};

#define HZ_LT(call)                                                         \
  do {                                                                      \
    cublasStatus_t _s = (call);                                             \
    if (_s != CUBLAS_STATUS_SUCCESS) {                                      \
      set_error("%s failed with cublas status %d", #call, (int)_s);         \
      return HZ_ERR_CUDA;                                                   \
    }                                                                       \
  } while (0)

// Heuristic choice of the library kernel for one step.  sm_target > 0 sets CUBLASLT_MATMUL_DESC_SM_COUNT_TARGET: the
// heuristic then sizes the kernel for that many SMs (fewer, fatter CTAs) because concurrent streams are expected to use
// the rest of the device — the setting for several searches in flight (SearchPipeline).
static int select_algo(hz_gemm_plan* p, LtStep& st) {
  const hz_gemm_step& s = st.s;
  int32_t target = p->sm_target > 0 ? p->sm_target : 0;
  HZ_LT(cublasLtMatmulDescSetAttribute(st.op, CUBLASLT_MATMUL_DESC_SM_COUNT_TARGET, &target, sizeof(target)));
  cublasLtMatmulPreference_t pref = nullptr;
  HZ_LT(cublasLtMatmulPreferenceCreate(&pref));
  HZ_LT(cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &p->ws_bytes, sizeof(p->ws_bytes)));
  const int eb = p->elem_bytes;
  uint32_t aa = (uint32_t)align_of(s.w, s.ldw, eb), ab = (uint32_t)align_of(s.a, s.lda, eb);
  uint32_t ac = (uint32_t)(s.c ? align_of(s.c, s.ldc, eb) : align_of(s.d, s.ldd, eb));
  uint32_t ad = (uint32_t)align_of(s.d, s.ldd, eb);
  HZ_LT(cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MIN_ALIGNMENT_A_BYTES, &aa, sizeof(aa)));
  HZ_LT(cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MIN_ALIGNMENT_B_BYTES, &ab, sizeof(ab)));
  HZ_LT(cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MIN_ALIGNMENT_C_BYTES, &ac, sizeof(ac)));
  HZ_LT(cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MIN_ALIGNMENT_D_BYTES, &ad, sizeof(ad)));
  constexpr int kCand = 12;
  cublasLtMatmulHeuristicResult_t res[kCand];
  int found = 0;
  cublasStatus_t hs = cublasLtMatmulAlgoGetHeuristic(p->lt, st.op, st.la, st.lb, st.lc, st.ld, pref, kCand, res, &found);
  cublasLtMatmulPreferenceDestroy(pref);
  if (hs != CUBLAS_STATUS_SUCCESS || found == 0) {
    set_error("cuBLASLt has no algorithm for GEMM m=%d n=%d k=%d batch=%d (status %d)", s.m, s.n, s.k, s.batch, (int)hs);
    return HZ_ERR_CUDA;
  }
  st.algo = res[0].algo;
  if (p->autotune && found > 1) {
    // the chain is latency-bound and the shapes are fixed: time the heuristic's candidates once and keep
    // the fastest (outputs are scratch at this point; every candidate computes the same D)
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const float alpha = 1.0f;
    float best = 1e30f;
    for (int c = 0; c < found; ++c) {
      if (res[c].state != CUBLAS_STATUS_SUCCESS || res[c].workspaceSize > p->ws_bytes) continue;
      bool ok = true;
      for (int rep = 0; rep < 3 && ok; ++rep) {
        ok = cublasLtMatmul(p->lt, st.op, &alpha, s.w, st.la, s.a, st.lb, &st.beta, s.c ? s.c : s.d, st.lc, s.d, st.ld,
                            &res[c].algo, p->workspace, p->ws_bytes, 0) == CUBLAS_STATUS_SUCCESS;
      }
      if (!ok) continue;
      cudaEventRecord(e0, 0);
      for (int rep = 0; rep < 20; ++rep) {
        cublasLtMatmul(p->lt, st.op, &alpha, s.w, st.la, s.a, st.lb, &st.beta, s.c ? s.c : s.d, st.lc, s.d, st.ld,
                       &res[c].algo, p->workspace, p->ws_bytes, 0);
      }
      cudaEventRecord(e1, 0);
      if (cudaEventSynchronize(e1) != cudaSuccess) { ok = false; }
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ok && ms < best) {
        best = ms;
        st.algo = res[c].algo;
      }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaGetLastError();
  }
  return HZ_OK;
}

static int build_step(hz_gemm_plan* p, LtStep& st) {
  const hz_gemm_step& s = st.s;
  const cudaDataType_t dt = p->elem_bytes == 2 ? CUDA_R_16F : CUDA_R_32F;
  HZ_LT(cublasLtMatmulDescCreate(&st.op, CUBLAS_COMPUTE_32F, CUDA_R_32F));
  const cublasOperation_t opT = CUBLAS_OP_T, opN = CUBLAS_OP_N;
  HZ_LT(cublasLtMatmulDescSetAttribute(st.op, CUBLASLT_MATMUL_DESC_TRANSA, &opT, sizeof(opT)));
  HZ_LT(cublasLtMatmulDescSetAttribute(st.op, CUBLASLT_MATMUL_DESC_TRANSB, &opN, sizeof(opN)));
  cublasLtEpilogue_t epi = s.bias ? (s.relu ? CUBLASLT_EPILOGUE_RELU_BIAS : CUBLASLT_EPILOGUE_BIAS)
                                  : (s.relu ? CUBLASLT_EPILOGUE_RELU : CUBLASLT_EPILOGUE_DEFAULT);
  HZ_LT(cublasLtMatmulDescSetAttribute(st.op, CUBLASLT_MATMUL_DESC_EPILOGUE, &epi, sizeof(epi)));
  if (s.bias) {
    HZ_LT(cublasLtMatmulDescSetAttribute(st.op, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &s.bias, sizeof(s.bias)));
    HZ_LT(cublasLtMatmulDescSetAttribute(st.op, CUBLASLT_MATMUL_DESC_BIAS_DATA_TYPE, &dt, sizeof(dt)));
    if (s.batch > 1) {
      int64_t bs = s.stride_bias;
      HZ_LT(cublasLtMatmulDescSetAttribute(st.op, CUBLASLT_MATMUL_DESC_BIAS_BATCH_STRIDE, &bs, sizeof(bs)));
    }
  }
  // column-major views: A_lt = W [k x n] ld=ldw, B_lt = A [k x m] ld=lda, C/D_lt [n x m]
  HZ_LT(cublasLtMatrixLayoutCreate(&st.la, dt, s.k, s.n, s.ldw));
  HZ_LT(cublasLtMatrixLayoutCreate(&st.lb, dt, s.k, s.m, s.lda));
  HZ_LT(cublasLtMatrixLayoutCreate(&st.lc, dt, s.n, s.m, s.c ? s.ldc : s.ldd));
  HZ_LT(cublasLtMatrixLayoutCreate(&st.ld, dt, s.n, s.m, s.ldd));
  if (s.batch > 1) {
    const int32_t bc = s.batch;
    struct { cublasLtMatrixLayout_t l; int64_t stride; } lay[4] = {
        {st.la, s.stride_w}, {st.lb, s.stride_a}, {st.lc, s.c ? s.stride_c : s.stride_d}, {st.ld, s.stride_d}};
    for (auto& x : lay) {
      HZ_LT(cublasLtMatrixLayoutSetAttribute(x.l, CUBLASLT_MATRIX_LAYOUT_BATCH_COUNT, &bc, sizeof(bc)));
      HZ_LT(cublasLtMatrixLayoutSetAttribute(x.l, CUBLASLT_MATRIX_LAYOUT_STRIDED_BATCH_OFFSET, &x.stride, sizeof(x.stride)));
    }
  }
  st.beta = s.c ? 1.0f : 0.0f;
  return select_algo(p, st);
}

static void free_step(LtStep& st) {
  if (st.la) cublasLtMatrixLayoutDestroy(st.la);
  if (st.lb) cublasLtMatrixLayoutDestroy(st.lb);
  if (st.lc) cublasLtMatrixLayoutDestroy(st.lc);
  if (st.ld) cublasLtMatrixLayoutDestroy(st.ld);
  if (st.op) cublasLtMatmulDescDestroy(st.op);
}

extern "C" {
#pragma GCC visibility push(default)

int hz_gemm_plan_create(hz_gemm_plan** out, int device, int elem_bytes, const hz_gemm_step* steps, int n_steps) {
  if (!out || !steps || n_steps <= 0 || (elem_bytes != 2 && elem_bytes != 4)) {
    set_error("hz_gemm_plan_create: bad argument");
    return HZ_ERR_ARG;
  }
  for (int i = 0; i < n_steps; ++i) {
    const hz_gemm_step& s = steps[i];
    if (!s.a || !s.w || !s.d || s.m <= 0 || s.n <= 0 || s.k <= 0 || s.batch <= 0 || s.lda < s.k || s.ldw < s.k ||
        s.ldd < s.n || (s.c && s.ldc < s.n)) {
      set_error("hz_gemm_plan_create: step %d is malformed", i);
      return HZ_ERR_ARG;
    }
  }
  DeviceGuard dg(device);
  if (!dg.ok) { set_error("hz_gemm_plan_create: cannot select device %d", device); return HZ_ERR_CUDA; }
  hz_gemm_plan* p = new hz_gemm_plan;
  p->device = device;
  p->elem_bytes = elem_bytes;
  p->ws_bytes = 32u << 20;
  p->autotune = false;   // timing the heuristic's candidates changed nothing measurable on B200 (34.9 vs 34.0 us per chain)
  if (cublasLtCreate(&p->lt) != CUBLAS_STATUS_SUCCESS) { delete p; set_error("cublasLtCreate failed"); return HZ_ERR_CUDA; }
  cudaError_t e = cudaMalloc(&p->workspace, p->ws_bytes);
  if (e != cudaSuccess) { cublasLtDestroy(p->lt); delete p; return fail_cuda(e, "hz_gemm_plan_create: workspace"); }
  p->steps.resize(n_steps);
  for (int i = 0; i < n_steps; ++i) {
    p->steps[i].s = steps[i];
    if (int rc = build_step(p, p->steps[i])) {
      hz_gemm_plan_destroy(p);
      return rc;
    }
  }
  *out = p;
  return HZ_OK;
}

int hz_gemm_plan_destroy(hz_gemm_plan* p) {
  if (!p) return HZ_OK;
  DeviceGuard dg(p->device);
  for (auto& st : p->steps) free_step(st);
  if (p->workspace) cudaFree(p->workspace);
  if (p->lt) cublasLtDestroy(p->lt);
  delete p;
  return HZ_OK;
}

int64_t hz_gemm_launch_count(void) { return g_gemm_launches.load(); }

int hz_gemm_plan_steps(const hz_gemm_plan* p) { return p ? (int)p->steps.size() : 0; }

int hz_gemm_plan_set_operand(hz_gemm_plan* p, int step, int which, void* ptr) {
  if (!p || step < 0 || step >= (int)p->steps.size() || which < 0 || which > 2 || !ptr) {
    set_error("hz_gemm_plan_set_operand: bad argument");
    return HZ_ERR_ARG;
  }
  if ((uintptr_t)ptr & 255) {   // the algorithms were chosen for 256-byte aligned operands
    set_error("hz_gemm_plan_set_operand: operand must be 256-byte aligned");
    return HZ_ERR_ARG;
  }
  hz_gemm_step& s = p->steps[step].s;
  if (which == 1 && !s.c) { set_error("hz_gemm_plan_set_operand: step %d has no residual operand", step); return HZ_ERR_ARG; }
  if (which == 0) s.a = ptr;
  else if (which == 1) s.c = ptr;
  else s.d = ptr;
  return HZ_OK;
}

int hz_gemm_plan_set_sm_target(hz_gemm_plan* p, int sm_count) {
  if (!p || sm_count < 0) { set_error("hz_gemm_plan_set_sm_target: bad argument"); return HZ_ERR_ARG; }
  DeviceGuard dg(p->device);
  p->sm_target = sm_count;
  for (auto& st : p->steps) {
    if (int rc = select_algo(p, st)) return rc;
  }
  return HZ_OK;
}

int hz_gemm_plan_run(hz_gemm_plan* p, void* stream, int first, int count) {
  if (!p || first < 0 || count < 0 || first + count > (int)p->steps.size()) {
    set_error("hz_gemm_plan_run: bad argument");
    return HZ_ERR_ARG;
  }
  DeviceGuard dg(p->device);
  const float alpha = 1.0f;
  for (int i = first; i < first + count; ++i) {
    LtStep& st = p->steps[i];
    const hz_gemm_step& s = st.s;
    HZ_LT(cublasLtMatmul(p->lt, st.op, &alpha, s.w, st.la, s.a, st.lb, &st.beta, s.c ? s.c : s.d, st.lc, s.d, st.ld,
                         &st.algo, p->workspace, p->ws_bytes, (cudaStream_t)stream));
    g_gemm_launches.fetch_add(1, std::memory_order_relaxed);
  }
  return HZ_OK;
}

#pragma GCC visibility pop
}  // extern "C"
