// clip_grad_norm_(max_norm) + Adam over one flat fp32 arena (parameters, gradients and both moments share the
// layout, which is also the NCCL all-reduce buffer).  Replaces il_modules/mrn.py:364-367 (reference: torch
// foreach kernels + a host-side norm).  Nothing synchronises with the host: the clip coefficient is computed
// on the device from the reduced norm.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, long n, double* __restrict__ partial) {
  __shared__ double sh[8];
  double s = 0.0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const double v = g[i];
    s += v * v;
  }
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += sh[k];
    partial[blockIdx.x] = t;
  }
}

// one warp: lane l sums partial[l], partial[l + 32], ... (all loads in flight), then a fixed-order butterfly -- the order
// of the additions never depends on timing, so the norm is deterministic (a single serial thread took 23 us here)
__global__ void norm_finish_kernel(const double* __restrict__ partial, int n, float max_norm, float* __restrict__ norm_out,
                                   float* __restrict__ coef_out) {
  const int lane = threadIdx.x;
  double t = 0.0;
#pragma unroll 4
  for (int k = lane; k < n; k += 32) t += partial[k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (lane != 0) return;
  const float norm = (float)sqrt(t);
  if (norm_out) norm_out[0] = norm;
  const float c = max_norm / (norm + 1e-6f);        // torch.nn.utils.clip_grad_norm_
  coef_out[0] = c < 1.f ? c : 1.f;
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long n,
            const float* __restrict__ coef, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gg = g[i] * coef[0];
  const float mm = b1 * m[i] + (1.f - b1) * gg;
  const float vv = b2 * v[i] + (1.f - b2) * gg * gg;
  m[i] = mm; v[i] = vv;
  const float denom = sqrtf(vv) / bc2_sqrt + eps;   // torch.optim.Adam (no amsgrad, no weight decay)
  p[i] -= (lr / bc1) * (mm / denom);
}

}  // namespace

extern "C" int mrnb_clip_adam(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long n, float lr,
                              float beta1, float beta2, float eps, float max_norm, int step, float* norm_out,
                              void* workspace, cudaStream_t stream) {
  MRNB_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && workspace && n > 0 && step >= 1, "clip_adam: bad argument");
  double* partial = (double*)workspace;          // [<= 448] doubles
  float* coef = (float*)((char*)workspace + 448 * sizeof(double));
  int blocks = cdiv(n, 256 * 8);
  if (blocks > 444) blocks = 444;                // 3 x 148 SMs
  sumsq_kernel<<<blocks, 256, 0, stream>>>(grads, n, partial);
  MRNB_CHECK_LAUNCH("sumsq_kernel");
  norm_finish_kernel<<<1, 32, 0, stream>>>(partial, blocks, max_norm, norm_out, coef);
  MRNB_CHECK_LAUNCH("norm_finish_kernel");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2s = sqrtf(1.f - powf(beta2, (float)step));
  adam_kernel<<<cdiv(n, 256), 256, 0, stream>>>(params, grads, exp_avg, exp_avg_sq, n, coef, lr, beta1, beta2, eps, bc1, bc2s);
  MRNB_CHECK_LAUNCH("adam_kernel");
  return MRNB_OK;
}
