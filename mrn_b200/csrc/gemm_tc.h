// Internal descriptor of the tcgen05 / TMEM / TMA bf16 GEMM (gemm_tc.cu).
#pragma once
#include <cuda_runtime.h>

struct MrnbTcGemm {
  // out[g, m, n] = epi( sum_k A[g, m, k] * W[g, n, k] + bias[g, n] )     A, W: bf16, k-contiguous
  const void* A; long lda; long a_gstride;     // elements
  const void* W; long ldw; long w_gstride;
  const float* bias; long bias_gstride;
  void* out; long ldo; long o_gstride; int out_f32;    // fp32 or bf16 output
  const float* res;                                    // fp32 residual at the output address (out_f32 only)
  const float* rowscale; int rows_per_scale; long rowscale_gstride;   // DropPath: * rowscale[g*gs + m / rows_per_scale]
  int M, N, K, groups, gelu;
};

int mrnb_tc_gemm(const MrnbTcGemm& g, cudaStream_t st);
