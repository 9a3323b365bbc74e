// Internal interface of the persistent fused GEMM-chain executor (hz_chain.cu) used by hz_gemm.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "../../include/hzb200.h"

namespace hz {

constexpr int kChainMaxSteps = 8;

struct ChainExec;

// true if the fused kernel can run this plan (fp16, shapes/strides it tiles, not disabled by HZ_FUSED_CHAIN=0)
bool chain_supported(const hz_gemm_step* steps, int n_steps, int elem_bytes, const char** why);
int chain_create(ChainExec** out, int device, const hz_gemm_step* steps, int n_steps);
void chain_destroy(ChainExec* e);
int chain_run(ChainExec* e, cudaStream_t stream);
int chain_grid(const ChainExec* e);
int chain_trace(const ChainExec* e, unsigned long long* host_out, size_t count);   // debug: HZ_CHAIN_TRACE=1

}  // namespace hz
