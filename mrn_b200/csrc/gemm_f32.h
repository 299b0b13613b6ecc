// Internal descriptor of the generic strided fp32 GEMM (gemm_f32.cu).
#pragma once
#include <cuda_runtime.h>

struct MrnbAxis {       // off(idx) = (idx / inner) * so + (idx % inner) * si
  int inner;
  long si, so;
};
static inline MrnbAxis mrnb_axis(long stride) { return MrnbAxis{1 << 30, stride, 0}; }
static inline MrnbAxis mrnb_axis2(int inner, long si, long so) { return MrnbAxis{inner, si, so}; }

struct MrnbGemm {
  const float* A; MrnbAxis am, ak; long sAb; int a_kfast;   // A(b,m,k)
  const float* B; MrnbAxis bn, bk; long sBb; int b_kfast;   // B(b,k,n)
  float* C; MrnbAxis cm, cn; long sCb;                      // C(b,m,n); mul / res share C's addressing
  int M, N, K, batch, splitk;
  float alpha;
  const float* bias_n; long bias_bstride;                   // + bias_n[b*bias_bstride + n]
  const float* bias_m;                                      // + bias_m[m]
  const float* mul;                                         // * mul[addr]
  const float* rowscale; long rowscale_bstride; int rows_per_scale;   // * rowscale[b*bstride + m / rows_per_scale]
  const float* res;                                         // + res[addr]
  int act;                                                  // 1 = exact GELU
  int accumulate;                                           // C += result
};

// out[M,N] = A[M,K] . W[N,K]^T   (both operands k-contiguous)
static inline MrnbGemm mrnb_gemm_nt(const float* A, long lda, const float* W, long ldw, float* C, long ldc, int M, int N, int K) {
  MrnbGemm g{};
  g.A = A; g.am = mrnb_axis(lda); g.ak = mrnb_axis(1); g.a_kfast = 1;
  g.B = W; g.bn = mrnb_axis(ldw); g.bk = mrnb_axis(1); g.b_kfast = 1;
  g.C = C; g.cm = mrnb_axis(ldc); g.cn = mrnb_axis(1);
  g.M = M; g.N = N; g.K = K; g.batch = 1; g.splitk = 1; g.alpha = 1.f; g.rows_per_scale = 1;
  return g;
}

int mrnb_sgemm(const MrnbGemm& p, cudaStream_t st);
