// Grouped bf16 GEMM on the 5th-generation tensor cores: TMA-fed shared-memory tiles (128B swizzle), tcgen05.mma
// issued by one elected thread, fp32 accumulator in TMEM, epilogue warps read it back with tcgen05.ld and fuse
// bias / exact GELU / DropPath row scale / fp32 residual before the store.
//
//   out[g, m, n] = epi( sum_k A[g, m, k] * W[g, n, k] + bias[g, n] )      g = expert (blockIdx.z)
//
// This is the tensor-core mode of every nn.Linear / conv-as-GEMM on the expert path (reference call sites:
// modules/svtr.py:56-58,106-108,277; modules/model.py:75-78,164,181), one launch for all experts.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue
// (TMEM lane quarter = warp_id % 4).  Pipelines: smem full/empty mbarriers (TMA <-> MMA) and one TMEM-full
// mbarrier (MMA -> epilogue).  sm_100a only.
#include "common.cuh"
#include "gemm_tc.h"
#include <cuda.h>

namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int MAX_STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KiB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
    if (spin > (1u << 28)) __trap();
  }
}

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row swizzle atoms 1024 B apart (SBO), LBO unused.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);        // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                              // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                              // layout type: SWIZZLE_128B
  return d;
}

// kind::f16 instruction descriptor: D = fp32, A = B = bf16, both K-major, M = 128, N = BN
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct TcEpi {
  const float* bias; long bias_gs;
  void* out; long ldo; long o_gs;
  const float* res;
  const float* rowscale; int rows_per_scale; long rowscale_gs;
  int M, N, KB, stages, gelu;
};

template <int BN, bool OUT_F32>
__global__ void __launch_bounds__(192)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const TcEpi ep) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int B_STAGE_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_sh;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM, g = blockIdx.z;
  const int stages = ep.stages, KB = ep.KB;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)), "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_sh;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % stages;
        const uint32_t ph = (uint32_t)(kb / stages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        uint8_t* sa = smem + (size_t)s * STAGE_BYTES;
        tma_load_3d(sa, &tmA, &full_bar[s], kb * BK, m0, g);
        tma_load_3d(sa + A_STAGE_BYTES, &tmW, &full_bar[s], kb * BK, n0, g);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % stages;
        const uint32_t ph = (uint32_t)(kb / stages) & 1u;
        mbar_wait(&full_bar[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + A_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in the (>>4) address field
          umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);          // frees the smem slot once these MMAs have read it
      }
      umma_commit(&tmem_full_bar);           // accumulator complete
    }
  } else {
    // ---- epilogue: warps 2..5, TMEM lane quarter q = warp % 4, one output row per thread
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    mbar_wait(&tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const float* bias = ep.bias ? ep.bias + (long)g * ep.bias_gs : nullptr;
    float rs = 1.f;
    if (ep.rowscale && row < ep.M) rs = ep.rowscale[(long)g * ep.rowscale_gs + row / ep.rows_per_scale];
    const long obase = (long)g * ep.o_gs + (long)row * ep.ldo;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);   // warp-collective: no early exit above
      const int col0 = n0 + c0;
      if (row >= ep.M || col0 >= ep.N) continue;
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float x = __uint_as_float(r[j]);
        if (bias && col0 + j < ep.N) x += __ldg(bias + col0 + j);
        if (ep.gelu) x = gelu_erf(x);
        v[j] = x * rs;
      }
      const bool full = (col0 + 16 <= ep.N);
      if (OUT_F32) {
        float* o = reinterpret_cast<float*>(ep.out) + obase + col0;
        const float* rp = ep.res ? ep.res + obase + col0 : nullptr;
        if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 t = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            if (rp) { const float4 rr = *reinterpret_cast<const float4*>(rp + j); t.x += rr.x; t.y += rr.y; t.z += rr.z; t.w += rr.w; }
            *reinterpret_cast<float4*>(o + j) = t;
          }
        } else {
          for (int j = 0; j < 16 && col0 + j < ep.N; ++j) o[j] = v[j] + (rp ? rp[j] : 0.f);
        }
      } else {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ep.out) + obase + col0;
        if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            pk[j] = *reinterpret_cast<uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(o + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        } else {
          for (int j = 0; j < 16 && col0 + j < ep.N; ++j) o[j] = __float2bfloat16_rn(v[j]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 3-D map over a [groups][rows][K] bf16 operand (K contiguous): box = 64 x box_rows x 1, 128B swizzle, zero OOB fill
int make_map(CUtensorMap* map, const void* ptr, long K, long rows, long groups, long ld, long gstride, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { mrnb_set_error("tc_gemm: cuTensorMapEncodeTiled is not available from the driver"); return MRNB_ERR_UNSUPPORTED; }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld % 8) || (groups > 1 && gstride % 8)) {
    mrnb_set_error("tc_gemm: operand must be 16-byte aligned with ld %% 8 == 0 (ptr=%p ld=%ld gstride=%ld)", ptr, ld, gstride);
    return MRNB_ERR_ARG;
  }
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)groups};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(groups > 1 ? gstride : rows * ld) * 2};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { mrnb_set_error("tc_gemm: cuTensorMapEncodeTiled failed (%d)", (int)r); return MRNB_ERR_ARG; }
  return MRNB_OK;
}

template <int BN, bool OUT_F32>
int launch_tc(const MrnbTcGemm& p, cudaStream_t st) {
  CUtensorMap tmA, tmW;
  MRNB_TRY(make_map(&tmA, p.A, p.K, p.M, p.groups, p.lda, p.a_gstride, BM));
  MRNB_TRY(make_map(&tmW, p.W, p.K, p.N, p.groups, p.ldw, p.w_gstride, BN));
  TcEpi ep;
  ep.bias = p.bias; ep.bias_gs = p.bias_gstride;
  ep.out = p.out; ep.ldo = p.ldo; ep.o_gs = p.o_gstride;
  ep.res = p.res; ep.rowscale = p.rowscale; ep.rows_per_scale = p.rows_per_scale > 0 ? p.rows_per_scale : 1;
  ep.rowscale_gs = p.rowscale_gstride;
  ep.M = p.M; ep.N = p.N; ep.KB = p.K / BK; ep.gelu = p.gelu;
  // two stages keep the tile at <= 64 KiB of smem: 3 CTAs per SM interleave their (short) K loops and epilogues
  ep.stages = ep.KB < 2 ? ep.KB : 2;
  const size_t smem = 1024 + (size_t)ep.stages * (A_STAGE_BYTES + BN * BK * 2);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_gemm_kernel<BN, OUT_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         1024 + MAX_STAGES * (A_STAGE_BYTES + BN * BK * 2));
    attr_set = true;
  }
  dim3 grid(cdiv(p.N, BN), cdiv(p.M, BM), p.groups);
  tc_gemm_kernel<BN, OUT_F32><<<grid, 192, smem, st>>>(tmA, tmW, ep);
  MRNB_CHECK_LAUNCH("tc_gemm_kernel");
  return MRNB_OK;
}

}  // namespace

int mrnb_tc_gemm(const MrnbTcGemm& p, cudaStream_t st) {
  MRNB_CHECK_ARG(p.A && p.W && p.out && p.M > 0 && p.N > 0 && p.K > 0 && p.groups > 0, "tc_gemm: bad argument");
  MRNB_CHECK_ARG(p.K % BK == 0, "tc_gemm: K=%d must be a multiple of %d", p.K, BK);
  MRNB_CHECK_ARG(!p.res || p.out_f32, "tc_gemm: residual needs an fp32 output");
  MrnbProfScope prof(MRNB_PROF_TCGEMM, st, 2.0 * p.M * p.N * p.K * p.groups,
                     (double)p.groups * (2.0 * p.M * p.K + 2.0 * p.N * p.K + (double)p.M * p.N * (p.out_f32 ? 4 : 2) +
                                         (p.res ? 4.0 * p.M * p.N : 0.0)));
  const bool wide = p.N >= 128 && (p.N % 128 == 0 || p.N > 256);
  if (wide) return p.out_f32 ? launch_tc<128, true>(p, st) : launch_tc<128, false>(p, st);
  return p.out_f32 ? launch_tc<64, true>(p, st) : launch_tc<64, false>(p, st);
}

extern "C" int mrnb_linear_bf16(const void* A, const void* W, const float* bias, const float* residual, void* out,
                                int out_is_f32, int M, int N, int K, int act_gelu, cudaStream_t stream) {
  MrnbTcGemm g{};
  g.A = A; g.lda = K; g.W = W; g.ldw = K; g.bias = bias; g.out = out; g.ldo = N; g.out_f32 = out_is_f32;
  g.res = residual; g.M = M; g.N = N; g.K = K; g.groups = 1; g.gelu = act_gelu; g.rows_per_scale = 1;
  return mrnb_tc_gemm(g, stream);
}
