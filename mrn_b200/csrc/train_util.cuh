// Building blocks shared by the expert-training translation units (svtr_train.cu, crnn_train.cu): column sums (bias
// gradients, BatchNorm statistics), BatchNorm finalisation, the forward / backward GEMM dispatch (CUDA-core fp32 engine or
// tcgen05 bf16 engines) and small casts.  Everything has internal linkage.
#pragma once
#include "common.cuh"
#include "gemm_f32.h"
#include "gemm_tc.h"
#include "gemm_tc2.h"
#include "expert_util.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// Column sums over rows (bias gradients; BatchNorm statistics with SQ).  block (32, 8), grid (cdiv(C,32), chunks).
// out must be zeroed; partial sums are combined with atomics.
// ------------------------------------------------------------------------------------------------
template <typename T, bool SQ>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, long ld, long rows, int C, float* __restrict__ out, double* __restrict__ out2) {
  __shared__ double sh[8][32][SQ ? 2 : 1];
  const int c = blockIdx.x * 32 + threadIdx.x, ty = threadIdx.y;
  const long per = (rows + gridDim.y - 1) / gridDim.y;
  const long r0 = (long)blockIdx.y * per, r1 = (r0 + per < rows) ? r0 + per : rows;
  double s1 = 0.0, s2 = 0.0;
  float f1 = 0.f;
  if (c < C) {
    for (long r = r0 + ty; r < r1; r += 8) {
      const float v = to_f32<T>(x[r * ld + c]);
      if (SQ) { s1 += v; s2 += (double)v * v; } else f1 += v;
    }
  }
  if (!SQ) s1 = f1;
  sh[ty][threadIdx.x][0] = s1;
  if (SQ) sh[ty][threadIdx.x][SQ ? 1 : 0] = s2;
  __syncthreads();
  if (ty == 0 && c < C) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { a += sh[k][threadIdx.x][0]; if (SQ) b += sh[k][threadIdx.x][SQ ? 1 : 0]; }
    if (SQ) { atomicAdd(out2 + c * 2, a); atomicAdd(out2 + c * 2 + 1, b); }
    else atomicAdd(out + c, (float)a);
  }
}

// Vector variant for C % 64 == 0 with aligned rows: a thread owns four adjacent columns (one 16- or 8-byte load per row),
// a block owns 64 columns x 16 rows in flight.  grid (C / 64, chunks), block 256.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_vec4_kernel(const T* __restrict__ x, long ld, long rows, float* __restrict__ out) {
  __shared__ float sh[16][64];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int c = blockIdx.x * 64 + tx * 4;
  const long per = (rows + gridDim.y - 1) / gridDim.y;
  const long r0 = (long)blockIdx.y * per, r1 = (r0 + per < rows) ? r0 + per : rows;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (long r = r0 + ty; r < r1; r += 16) {
    if constexpr (sizeof(T) == 4) {
      const float4 v = *reinterpret_cast<const float4*>(x + r * ld + c);
      a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
    } else {
      const uint2 v = *reinterpret_cast<const uint2*>(x + r * ld + c);
      const __nv_bfloat162 p = *reinterpret_cast<const __nv_bfloat162*>(&v.x), q = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
      a0 += __bfloat162float(p.x); a1 += __bfloat162float(p.y); a2 += __bfloat162float(q.x); a3 += __bfloat162float(q.y);
    }
  }
  sh[ty][tx * 4] = a0; sh[ty][tx * 4 + 1] = a1; sh[ty][tx * 4 + 2] = a2; sh[ty][tx * 4 + 3] = a3;
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) t += sh[k][threadIdx.x];
    atomicAdd(out + blockIdx.x * 64 + threadIdx.x, t);
  }
}

template <typename T>
int launch_colsum(const T* x, long ld, long rows, int C, float* out, cudaStream_t st) {
  int chunks = (int)(rows / 256); if (chunks < 1) chunks = 1; if (chunks > 64) chunks = 64;
  if (C % 64 == 0 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    int ch = (int)(rows / 512); if (ch < 1) ch = 1;
    const int want = (148 * 4 + C / 64 - 1) / (C / 64);            // ~4 blocks per SM over all column groups
    if (ch > want) ch = want;
    colsum_vec4_kernel<T><<<dim3(C / 64, ch), 256, 0, st>>>(x, ld, rows, out);
    MRNB_CHECK_LAUNCH("colsum_vec4_kernel");
    return MRNB_OK;
  }
  colsum_kernel<T, false><<<dim3(cdiv(C, 32), chunks), dim3(32, 8), 0, st>>>(x, ld, rows, C, out, nullptr);
  MRNB_CHECK_LAUNCH("colsum_kernel");
  return MRNB_OK;
}
int launch_colstats(const float* x, long rows, int C, double* stats, cudaStream_t st) {
  int chunks = (int)(rows / 256); if (chunks < 1) chunks = 1; if (chunks > 148) chunks = 148;
  colsum_kernel<float, true><<<dim3(cdiv(C, 32), chunks), dim3(32, 8), 0, st>>>(x, C, rows, C, nullptr, stats);
  MRNB_CHECK_LAUNCH("colsum_kernel");
  return MRNB_OK;
}

// BatchNorm (train): finalize keeps (scale, shift) and (mean, rstd)
__global__ void bn_finalize_train_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, float* __restrict__ run_mean,
                                         float* __restrict__ run_var, float* __restrict__ ss, float* __restrict__ mr, int C,
                                         double count, int use_batch, int update_running, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, var;
  if (use_batch) {
    const double m = stats[c * 2] / count;
    double v = stats[c * 2 + 1] / count - m * m;
    if (v < 0) v = 0;
    mean = (float)m; var = (float)v;
    if (update_running) {
      run_mean[c] = 0.9f * run_mean[c] + 0.1f * mean;
      run_var[c] = 0.9f * run_var[c] + 0.1f * (float)(v * count / (count - 1.0));
    }
  } else {
    mean = run_mean[c]; var = run_var[c];
  }
  const float rstd = rsqrtf(var + eps);
  const float sc = gamma[c] * rstd;
  ss[c * 2] = sc; ss[c * 2 + 1] = beta[c] - mean * sc;
  mr[c * 2] = mean; mr[c * 2 + 1] = rstd;
}

// ------------------------------------------------------------------------------------------------
// GEMM dispatch.  Forward Linear: linear<AT>() (CUDA cores fp32 / persistent tcgen05 bf16).  Backward:
//   dX[M,K] = dY[M,N] . W[N,K];   dW[N,K] (+)= dY[M,N]^T . X[M,K]  (split over the M rows, atomics into a zeroed dW).
// fp32: gemm_f32.cu.  bf16: the general tcgen05 engine (gemm_tc2.cu) reading every tensor where it lies -- dY K-major
// and W MN-major for dX; dY and X both MN-major for dW -- so no transposed copy is ever made.  A contraction length
// that is not a multiple of 64 (the ragged charset C) is rounded up: TMA zero-fills both operands beyond their extent.
// ------------------------------------------------------------------------------------------------
typedef __nv_bfloat16 bf16;

template <typename AT>
int lin(const AT* A, long lda, const float* W32, const void* W16, const float* bias, void* out, long ldo, bool out_f32,
        int M, int N, int K, const float* res, const float* rowscale, int rows_per_scale, cudaStream_t st) {
  LinearArgs a{};
  a.A = A; a.lda = lda; a.a_gstride = (long)M * lda;
  a.W32 = W32; a.W16 = W16; a.w_gstride = (long)N * K;
  a.bias = bias; a.bias_gstride = N;
  a.out = out; a.ldo = ldo; a.o_gstride = (long)M * ldo; a.out_is_f32 = out_f32 ? 1 : 0;
  a.res = res; a.rowscale = rowscale; a.rows_per_scale = rows_per_scale;
  a.M = M; a.N = N; a.K = K; a.groups = 1;
  if (sizeof(AT) == 2) MRNB_CHECK_ARG(W16, "svtr_train: bf16 mode needs the 16-bit weight shadow (pack->h / fc_w16)");
  return linear<AT>(a, st);
}

int gemm_dx_f32(const float* dY, long ldy, const float* W, float* dX, long ldx, int M, int N, int K, cudaStream_t st) {
  MrnbGemm g{};
  g.A = dY; g.am = mrnb_axis(ldy); g.ak = mrnb_axis(1); g.a_kfast = 1;
  g.B = W; g.bk = mrnb_axis(K); g.bn = mrnb_axis(1); g.b_kfast = 0;
  g.C = dX; g.cm = mrnb_axis(ldx); g.cn = mrnb_axis(1);
  g.M = M; g.N = K; g.K = N; g.batch = 1; g.splitk = 1; g.alpha = 1.f; g.rows_per_scale = 1;
  return mrnb_sgemm(g, st);
}
int gemm_dw_f32(const float* dY, long ldy, const float* X, long ldx, float* dW, int rows, int N, int K, cudaStream_t st) {
  MrnbGemm g{};
  g.A = dY; g.am = mrnb_axis(1); g.ak = mrnb_axis(ldy); g.a_kfast = 0;
  g.B = X; g.bk = mrnb_axis(ldx); g.bn = mrnb_axis(1); g.b_kfast = 0;
  g.C = dW; g.cm = mrnb_axis(K); g.cn = mrnb_axis(1);
  g.M = N; g.N = K; g.K = rows; g.batch = 1; g.alpha = 1.f; g.rows_per_scale = 1;
  const int tile = (N >= 96 && K >= 96) ? 128 : 64;
  const long tiles = (long)cdiv(N, tile) * cdiv(K, tile);
  long sk = (148L * 4 + tiles - 1) / tiles;
  const long maxsk = rows / 128 > 0 ? rows / 128 : 1;
  if (sk > maxsk) sk = maxsk;
  if (sk < 1) sk = 1;
  g.splitk = (int)sk;
  return mrnb_sgemm(g, st);
}
int gemm_dx_tc(const bf16* dY, long ldy, const void* W16, float* dX, bf16* dX16, long ldx, int M, int N, int K, cudaStream_t st) {
  MrnbTcGemm2 g{};
  g.a = mrnb_operand_k2d(dY, M, N, ldy, 128, 1);
  g.b = mrnb_operand_mn2d(W16, K, N, K, 1);
  g.out32 = dX; g.out16 = dX16; g.cm = mrnb_axis(ldx); g.cn = mrnb_axis(1);
  g.M = M; g.N = K; g.K = (N + 63) / 64 * 64; g.groups = 1; g.splitk = 1; g.alpha = 1.f;
  return mrnb_tc_gemm2(g, st);
}
int gemm_dw_tc(const bf16* dY, long ldy, const bf16* X, long ldx, float* dW, int rows, int N, int K, cudaStream_t st) {
  // a contraction length (rows) that is not a multiple of 64 is rounded up: both MN-major operands are zero-filled by TMA
  MrnbTcGemm2 g{};
  g.a = mrnb_operand_mn2d(dY, N, rows, ldy, 1);
  g.b = mrnb_operand_mn2d(X, K, rows, ldx, 1);
  g.out32 = dW; g.cm = mrnb_axis(K); g.cn = mrnb_axis(1);
  g.M = N; g.N = K; g.K = (rows + 63) / 64 * 64; g.groups = 1; g.alpha = 1.f;
  const long tiles = (long)cdiv(N, 128) * cdiv(K, K >= 128 ? 128 : 64);
  long sk = (148L * 3 + tiles - 1) / tiles;
  const long maxsk = rows / 256 > 0 ? rows / 256 : 1;
  if (sk > maxsk) sk = maxsk;
  if (sk < 1) sk = 1;
  g.splitk = (int)sk;
  return mrnb_tc_gemm2(g, st);
}
// one gradient tensor in both flavours: fp32 (elementwise math, bias sums) and, in bf16 mode, its 16-bit GEMM operand
struct Grad { const float* f; const bf16* h; long ld; };

template <typename AT>
int gemm_dx(const Grad& dY, const float* W32, const void* W16, float* dX, bf16* dX16, long ldx, int M, int N, int K, cudaStream_t st) {
  if constexpr (sizeof(AT) == 4) return gemm_dx_f32(dY.f, dY.ld, W32, dX, ldx, M, N, K, st);
  else return gemm_dx_tc(dY.h, dY.ld, W16, dX, dX16, ldx, M, N, K, st);
}
template <typename AT>
int gemm_dw(const Grad& dY, const AT* X, long ldx, float* dW, int rows, int N, int K, cudaStream_t st) {
  if constexpr (sizeof(AT) == 4) return gemm_dw_f32(dY.f, dY.ld, X, ldx, dW, rows, N, K, st);
  else return gemm_dw_tc(dY.h, dY.ld, X, ldx, dW, rows, N, K, st);
}

// fp32 [rows, C] (ld; vector loads when it is a multiple of 4) -> bf16 [rows, ld16] (a multiple of 8), columns C..ld16 zero; eight columns per thread
__global__ void cast_pad_rows_kernel(const float* __restrict__ x, long ld, int C, bf16* __restrict__ y, long ld16, long total) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;          // one group of 8 output columns
  const long per_row = ld16 / 8;
  if (i >= total / 8) return;
  const long r = i / per_row;
  const int c = (int)(i % per_row) * 8;
  float v[8];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int cc = c + 4 * h;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((ld & 3) == 0 && cc + 4 <= ld) {
      t = *reinterpret_cast<const float4*>(x + r * ld + cc);
    } else {                                               // unaligned rows: scalar loads
      const float* s = x + r * ld + cc;
      if (cc < C) t.x = s[0];
      if (cc + 1 < C) t.y = s[1];
      if (cc + 2 < C) t.z = s[2];
      if (cc + 3 < C) t.w = s[3];
    }
    v[4 * h] = cc < C ? t.x : 0.f; v[4 * h + 1] = cc + 1 < C ? t.y : 0.f;
    v[4 * h + 2] = cc + 2 < C ? t.z : 0.f; v[4 * h + 3] = cc + 3 < C ? t.w : 0.f;
  }
  uint32_t pk[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * u], v[2 * u + 1]);
    pk[u] = *reinterpret_cast<uint32_t*>(&hh);
  }
  *reinterpret_cast<uint4*>(y + r * ld16 + c) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
}
}  // namespace
