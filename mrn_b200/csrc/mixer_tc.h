// Internal descriptor of the fused SVTR mixer-branch kernel (mixer_tc.cu).
#pragma once
#include <cuda_runtime.h>

// One launch = the whole first branch of a Block for every (expert, sample) unit (modules/svtr.py:200-203, :133-152):
//     x <- x + rs * ( proj( softmax(q k^T [+ local window]) v ) + b_proj ),   [ ln_out = LN2(x) ]
// with q, k, v = LN1(x) Wqkv^T + b (computed on chip from A = LN1(x)); q, k, v, the scores and the probabilities never
// reach HBM.  Token grid is H x 64 with H * 64 = 32768 / D tokens per unit; head_dim = 32.
struct MrnbMixer {
  const void* A;            // bf16 [units][N][D] = LN1(x), units = groups * units_per_group, N = 32768 / D
  const void* Wqkv;         // bf16 [groups][3D][D]   (mixer.qkv.weight)
  const float* bqkv;        // [groups][3D]
  const void* Wproj;        // bf16 [groups][D][D]    (mixer.proj.weight)
  const float* bproj;       // [groups][D]
  float* x; long x_gstride; // fp32 residual stream, unit (g, b) at x + g * x_gstride + b * N * D, updated in place
  const float* rowscale; long rowscale_gstride;   // optional DropPath multiplier per unit: rowscale[g * gs + b]
  void* ln_out; const float* ln_gamma; const float* ln_beta; float ln_eps;   // optional LN2 (D <= 128): bf16 [units][N][D], may alias A
  int D, groups, units_per_group, local;
};

int mrnb_mixer_tc(const MrnbMixer& p, cudaStream_t st);
