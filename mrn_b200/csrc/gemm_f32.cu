// Generic strided / batched fp32 GEMM on the CUDA cores with a fused epilogue.
//
// This is the fp32 ("parity") arithmetic mode of every dense contraction on the path and the v1 engine of
// the DM-Router backward; the bf16 tensor-core mode is gemm_tc.cu (tcgen05 + TMA).  C = epi(A . B) with
//   A(b,m,k), B(b,k,n), C(b,m,n) addressed through two-level strides per axis:
//        off(idx) = (idx / inner) * outer_stride + (idx % inner) * inner_stride
// so that the reference's permutes / rearranges (modules/dm_router.py:58-65, modules/model.py:402) become
// addressing instead of copies.
#include "common.cuh"
#include "gemm_f32.h"

namespace {

__device__ __forceinline__ long ax_off(const MrnbAxis& a, int idx) {
  return (long)(idx / a.inner) * a.so + (long)(idx % a.inner) * a.si;
}

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
sgemm_kernel(const MrnbGemm p) {
  constexpr int BK = 16;
  // each thread owns TM/4 x TN/4 chunks of 4x4 outputs, chunk stride CSM/CSN (bank-conflict-free LDS.128)
  constexpr int CSM = BM / (TM / 4), CSN = BN / (TN / 4);
  constexpr int NT = (BM / TM) * (BN / TN);
  __shared__ float As[2][BK][BM + 4];
  __shared__ float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  int batch = blockIdx.z, ks = 0, ke = p.K;
  if (p.splitk > 1) {
    batch = 0;
    const int per = ((p.K + p.splitk - 1) / p.splitk + BK - 1) / BK * BK;
    ks = blockIdx.z * per;
    ke = min(p.K, ks + per);
    if (ks >= ke) return;
  }
  const float* A = p.A + (long)batch * p.sAb;
  const float* B = p.B + (long)batch * p.sBb;

  // loader mappings: "k fastest" (operand contiguous along k) or "mn fastest"
  constexpr int A_ELEMS = BM * BK / NT, B_ELEMS = BN * BK / NT;
  static_assert(A_ELEMS * NT == BM * BK && B_ELEMS * NT == BN * BK, "tile/threads mismatch");
  float ra[A_ELEMS], rb[B_ELEMS];

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int e = 0; e < A_ELEMS; ++e) {
      const int lin = e * NT + tid;
      int mm, kk;
      if (p.a_kfast) { kk = lin % BK; mm = lin / BK; } else { mm = lin % BM; kk = lin / BM; }
      const int m = m0 + mm, k = k0 + kk;
      ra[e] = (m < p.M && k < ke) ? __ldg(A + ax_off(p.am, m) + ax_off(p.ak, k)) : 0.f;
    }
#pragma unroll
    for (int e = 0; e < B_ELEMS; ++e) {
      const int lin = e * NT + tid;
      int nn, kk;
      if (p.b_kfast) { kk = lin % BK; nn = lin / BK; } else { nn = lin % BN; kk = lin / BN; }
      const int n = n0 + nn, k = k0 + kk;
      rb[e] = (n < p.N && k < ke) ? __ldg(B + ax_off(p.bn, n) + ax_off(p.bk, k)) : 0.f;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int e = 0; e < A_ELEMS; ++e) {
      const int lin = e * NT + tid;
      int mm, kk;
      if (p.a_kfast) { kk = lin % BK; mm = lin / BK; } else { mm = lin % BM; kk = lin / BM; }
      As[buf][kk][mm] = ra[e];
    }
#pragma unroll
    for (int e = 0; e < B_ELEMS; ++e) {
      const int lin = e * NT + tid;
      int nn, kk;
      if (p.b_kfast) { kk = lin % BK; nn = lin / BK; } else { nn = lin % BN; kk = lin / BN; }
      Bs[buf][kk][nn] = rb[e];
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  load_tiles(ks);
  store_tiles(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = ks; k0 < ke; k0 += BK) {
    const bool more = k0 + BK < ke;
    if (more) load_tiles(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&As[buf][kk][(i / 4) * CSM + ty * 4]);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][(j / 4) * CSN + tx * 4]);
        b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      store_tiles(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

  // epilogue
  float* C = p.C + (long)batch * p.sCb;
  const float* mul = p.mul ? p.mul + (long)batch * p.sCb : nullptr;
  const float* res = p.res ? p.res + (long)batch * p.sCb : nullptr;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + (i / 4) * CSM + ty * 4 + (i % 4);
    if (m >= p.M) continue;
    const long om = ax_off(p.cm, m);
    const float bm = p.bias_m ? p.bias_m[m] : 0.f;
    const float rs = p.rowscale ? p.rowscale[(long)batch * p.rowscale_bstride + m / p.rows_per_scale] : 1.f;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + (j / 4) * CSN + tx * 4 + (j % 4);
      if (n >= p.N) continue;
      const long o = om + ax_off(p.cn, n);
      float v = acc[i][j] * p.alpha;
      if (p.splitk > 1) { atomicAdd(C + o, v); continue; }
      v += bm;
      if (p.bias_n) v += p.bias_n[(long)batch * p.bias_bstride + n];
      if (p.act == 1) v = gelu_erf(v);
      else if (p.act == 2) v = fmaxf(v, 0.f);
      if (mul) v *= mul[o];
      v *= rs;
      if (res) v += res[o];
      if (p.accumulate) v += C[o];
      C[o] = v;
    }
  }
}

}  // namespace

int mrnb_sgemm(const MrnbGemm& p, cudaStream_t st) {
  MRNB_CHECK_ARG(p.A && p.B && p.C && p.M > 0 && p.N > 0 && p.K > 0 && p.batch > 0, "sgemm: bad argument");
  MRNB_CHECK_ARG(p.splitk <= 1 || p.batch == 1, "sgemm: split-K only for batch == 1");
  const int z = p.splitk > 1 ? p.splitk : p.batch;
  MrnbProfScope prof(MRNB_PROF_SGEMM, st, 2.0 * p.M * p.N * p.K * p.batch,
                     4.0 * p.batch * ((double)p.M * p.K + (double)p.N * p.K + (double)p.M * p.N));
  // 128 x 128 tiles only when they fill the machine; small problems (the LSTM recurrences) take 64 x 64 tiles
  if (p.M >= 96 && p.N >= 96 && (long)cdiv(p.N, 128) * cdiv(p.M, 128) * z >= 148) {
    dim3 grid(cdiv(p.N, 128), cdiv(p.M, 128), z);
    sgemm_kernel<128, 128, 8, 8><<<grid, 256, 0, st>>>(p);
  } else {
    dim3 grid(cdiv(p.N, 64), cdiv(p.M, 64), z);
    sgemm_kernel<64, 64, 4, 4><<<grid, 256, 0, st>>>(p);
  }
  MRNB_CHECK_LAUNCH("sgemm_kernel");
  return MRNB_OK;
}

// C-ABI: plain row-major Linear, out[M,N] = act(A[M,K] . W[N,K]^T + bias) (+ residual)   (tests, fp32 mode)
extern "C" int mrnb_linear_f32(const float* A, const float* W, const float* bias, const float* residual, float* out,
                               int M, int N, int K, int act_gelu, cudaStream_t stream) {
  MrnbGemm g = mrnb_gemm_nt(A, K, W, K, out, N, M, N, K);
  g.bias_n = bias;
  g.res = residual;
  g.act = act_gelu ? 1 : 0;
  return mrnb_sgemm(g, stream);
}
