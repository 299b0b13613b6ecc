// Internal descriptor of the fused MLP kernel (mlp_tc.cu).
#pragma once
#include <cuda_runtime.h>

struct MrnbMlp {
  const void* A;          // bf16 [groups][M][D]  = LN2(x)
  const void* W1;         // bf16 [groups][4D][D]   (mlp.fc1.weight)
  const void* W2;         // bf16 [groups][D][4D]   (mlp.fc2.weight)
  const float* b1;        // [groups][4D]
  const float* b2;        // [groups][D]
  float* x; long x_gstride;                       // fp32 residual stream [g][M][D], updated in place
  const float* rowscale; int rows_per_scale; long rowscale_gstride;    // DropPath multipliers (tile-uniform)
  void* ln_out; const float* ln_gamma; const float* ln_beta; float ln_eps;   // optional next LayerNorm (D <= 128), may alias A
  void* cast_out;         // optional bf16 copy of the updated x [groups][M][D] (feeds the SubSample implicit GEMM); not with ln_out
  int M, D, groups;
  void* trace;            // optional device buffer [10][64] u64: debug timeline of CTA 0 (tools/mlp_trace.py)
};

int mrnb_mlp_tc(const MrnbMlp& p, cudaStream_t st);
