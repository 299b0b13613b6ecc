// Image preparation on the device: ragged RGBA uint8 crops -> [B, 4, outH, outW] fp32 in [-1, 1], bit-identical to the
// reference's host pipeline.
//
// Reference (paths relative to /root/reference): data/dataset.py:235-246 ResizeNormalize (PIL Image.resize((imgW, imgH),
// BICUBIC) -> torchvision ToTensor -> sub_(0.5).div_(0.5)), applied per image by AlignCollate (data/dataset.py:169-197).
// The arithmetic is Pillow's (third-party, restated from its published algorithm; oracle/resize_oracle.py is pinned
// bit-exactly against Pillow itself):
//   Image.resize on RGBA: premultiply alpha (Convert.c rgbA2rgba, MULDIV255) -> resample -> un-premultiply (rgba2rgbA);
//     an image that already has the target size is returned unchanged;
//   Resample.c: horizontal pass then vertical pass, 8-bit intermediate; per output index the taps are
//     precompute_coeffs() in double (bicubic a = -0.5, support 2 x max(scale, 1): antialiased when shrinking) normalised
//     and converted to 22-bit fixed point; pixel = clip8((2^21 + sum tap * value) >> 22).
// The taps are recomputed on the device in IEEE double with explicitly unfused operations (__dmul_rn / __dadd_rn), so the
// fixed-point taps -- and with them every output byte -- equal Pillow's.
//
// Layout: `pixels` is one packed byte buffer, image b = rows of RGBA at pixels + offsets[b], widths[b] x heights[b].
// HBM-bound byte work: pass H writes an [h, outW] RGBA intermediate per image, pass V reads it once.
#include "common.cuh"
#include "svtr.h"

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;
constexpr int KMAX_H = 64;      // taps per output column:  ceil(2 * w / outW) * 2 + 1 <= 64
constexpr int KMAX_V = 128;     // taps per output row

__device__ __forceinline__ double bicubic_filter(double x) {
  // ((a + 2) x - (a + 3)) x x + 1   |   (((x - 5) x + 8) x - 4) a      with a = -0.5, evaluated left to right, no FMA
  if (x < 0.0) x = -x;
  if (x < 1.0) return __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(1.5, x), 2.5), x), x), 1.0);
  if (x < 2.0) return __dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0), x), 4.0), -0.5);
  return 0.0;
}

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for one output index; taps written to k[0 .. stride*(count-1)].
__device__ void taps_for(int in_size, int out_size, int xx, int* k, int stride, int kmax, int& xmin_out, int& count_out) {
  const double scale = __ddiv_rn((double)in_size, (double)out_size);
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = __dmul_rn(2.0, filterscale);
  const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
  const double ss = __ddiv_rn(1.0, filterscale);
  int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
  if (xmax > in_size) xmax = in_size;
  int n = xmax - xmin;
  if (n > kmax) n = kmax;                       // guarded by the host-side size check
  double ww = 0.0;
  for (int x = 0; x < n; ++x)
    ww = __dadd_rn(ww, bicubic_filter(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss)));
  for (int x = 0; x < n; ++x) {
    double w = bicubic_filter(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss));
    if (ww != 0.0) w = __ddiv_rn(w, ww);
    const double f = __dmul_rn(w, (double)(1 << PRECISION_BITS));
    k[x * stride] = w < 0.0 ? (int)__dadd_rn(-0.5, f) : (int)__dadd_rn(0.5, f);
  }
  xmin_out = xmin; count_out = n;
}

__device__ __forceinline__ int muldiv255(int a, int b) { const int t = a * b + 128; return ((t >> 8) + t) >> 8; }
__device__ __forceinline__ int clip8(int v) { v >>= PRECISION_BITS; return v < 0 ? 0 : (v > 255 ? 255 : v); }

// Pass H.  grid (B, row chunks), block outW threads: thread = output column; taps in shared memory [tap][column].
__global__ void __launch_bounds__(256)
resize_h_kernel(const unsigned char* __restrict__ pixels, const long long* __restrict__ offsets, const int* __restrict__ widths,
                const int* __restrict__ heights, int out_h, int out_w, int max_h, unsigned char* __restrict__ tmp, int rows_per_cta) {
  extern __shared__ int sk[];                    // [KMAX_H][out_w]
  const int b = blockIdx.x, xx = threadIdx.x;
  const int w = widths[b], h = heights[b];
  const int y0 = blockIdx.y * rows_per_cta;
  if (y0 >= h || (w == out_w && h == out_h)) return;          // identical size: Image.resize returns a copy
  int xmin, n;
  taps_for(w, out_w, xx, sk + xx, out_w, KMAX_H, xmin, n);
  const uchar4* src = reinterpret_cast<const uchar4*>(pixels + offsets[b]);
  uchar4* dst = reinterpret_cast<uchar4*>(tmp) + (size_t)b * max_h * out_w;
  const int y1 = min(h, y0 + rows_per_cta);
  for (int y = y0; y < y1; ++y) {
    const uchar4* row = src + (size_t)y * w + xmin;
    int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0, s3 = s0;
    for (int x = 0; x < n; ++x) {
      const uchar4 p = row[x];
      const int kx = sk[x * out_w + xx];
      const int a = p.w;
      s0 += muldiv255(p.x, a) * kx; s1 += muldiv255(p.y, a) * kx; s2 += muldiv255(p.z, a) * kx; s3 += a * kx;
    }
    dst[(size_t)y * out_w + xx] = make_uchar4((unsigned char)clip8(s0), (unsigned char)clip8(s1), (unsigned char)clip8(s2),
                                              (unsigned char)clip8(s3));
  }
}

// Pass V + un-premultiply + ToTensor + normalise.  grid (B), block outW threads.
__global__ void __launch_bounds__(256)
resize_v_kernel(const unsigned char* __restrict__ pixels, const long long* __restrict__ offsets, const int* __restrict__ widths,
                const int* __restrict__ heights, int out_h, int out_w, int max_h, const unsigned char* __restrict__ tmp,
                float* __restrict__ out) {
  extern __shared__ int sk[];                    // [out_h][KMAX_V] taps, then [out_h][2] bounds
  int* sb = sk + out_h * KMAX_V;
  const int b = blockIdx.x, xx = threadIdx.x;
  const int w = widths[b], h = heights[b];
  const bool same = (w == out_w && h == out_h);
  if (!same && xx < out_h) {
    int ymin, n;
    taps_for(h, out_h, xx, sk + xx * KMAX_V, 1, KMAX_V, ymin, n);
    sb[xx * 2] = ymin; sb[xx * 2 + 1] = n;
  }
  __syncthreads();
  const uchar4* src = same ? reinterpret_cast<const uchar4*>(pixels + offsets[b])
                           : reinterpret_cast<const uchar4*>(tmp) + (size_t)b * max_h * out_w;
  float* o = out + (size_t)b * 4 * out_h * out_w;
  for (int yy = 0; yy < out_h; ++yy) {
    int r, g, bl, a;
    if (same) {
      const uchar4 p = src[(size_t)yy * out_w + xx];
      r = p.x; g = p.y; bl = p.z; a = p.w;
    } else {
      const int ymin = sb[yy * 2], n = sb[yy * 2 + 1];
      const int* k = sk + yy * KMAX_V;
      int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0, s3 = s0;
      for (int y = 0; y < n; ++y) {
        const uchar4 p = src[(size_t)(ymin + y) * out_w + xx];
        const int ky = k[y];
        s0 += p.x * ky; s1 += p.y * ky; s2 += p.z * ky; s3 += p.w * ky;
      }
      r = clip8(s0); g = clip8(s1); bl = clip8(s2); a = clip8(s3);
      if (a != 255 && a != 0) {                    // Convert.c rgba2rgbA
        r = min(255, 255 * r / a); g = min(255, 255 * g / a); bl = min(255, 255 * bl / a);
      }
    }
    // ToTensor: uint8 -> float / 255 ; then (x - 0.5) / 0.5, each step rounded to fp32 like the torch ops
    const size_t plane = (size_t)out_h * out_w, at = (size_t)yy * out_w + xx;
    o[at] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)r, 255.0f), 0.5f), 0.5f);
    o[plane + at] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)g, 255.0f), 0.5f), 0.5f);
    o[2 * plane + at] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)bl, 255.0f), 0.5f), 0.5f);
    o[3 * plane + at] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)a, 255.0f), 0.5f), 0.5f);
  }
}

}  // namespace

extern "C" size_t mrnb_resize_workspace_bytes(int B, int max_h, int out_w) {
  return (size_t)B * (size_t)(max_h > 0 ? max_h : 1) * out_w * 4 + 256;
}

extern "C" int mrnb_resize_normalize_rgba(const unsigned char* pixels, const long long* offsets, const int* widths,
                                          const int* heights, int B, int max_w, int max_h, int out_h, int out_w, float* out,
                                          void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MRNB_CHECK_ARG(pixels && offsets && widths && heights && out && workspace && B > 0, "resize_normalize: null/empty argument");
  MRNB_CHECK_ARG(out_w >= 32 && out_w <= 256 && out_w % 32 == 0 && out_h >= 1 && out_h <= out_w,
                 "resize_normalize: output size %dx%d unsupported (width a multiple of 32 up to 256, height <= width)", out_w, out_h);
  MRNB_CHECK_ARG(max_w >= 1 && max_h >= 1, "resize_normalize: empty image");
  const double sx = (double)max_w / out_w, sy = (double)max_h / out_h;
  const int kx = (int)ceil(2.0 * (sx < 1.0 ? 1.0 : sx)) * 2 + 1, ky = (int)ceil(2.0 * (sy < 1.0 ? 1.0 : sy)) * 2 + 1;
  MRNB_CHECK_ARG(kx <= KMAX_H && ky <= KMAX_V, "resize_normalize: image %dx%d shrinks by more than the supported factor "
                 "(taps %d/%d > %d/%d)", max_w, max_h, kx, ky, KMAX_H, KMAX_V);
  MRNB_CHECK_ARG(workspace_bytes >= mrnb_resize_workspace_bytes(B, max_h, out_w), "resize_normalize: workspace too small");
  const size_t smem_h = (size_t)KMAX_H * out_w * sizeof(int), smem_v = (size_t)out_h * (KMAX_V + 2) * sizeof(int);
  static bool set = false;
  if (!set) {
    cudaFuncSetAttribute(resize_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KMAX_H * 256 * 4);
    cudaFuncSetAttribute(resize_v_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * (KMAX_V + 2) * 4);
    set = true;
  }
  const int rows_per_cta = 16;
  MrnbProfScope prof(MRNB_PROF_MISC, stream, 0.0, 0.0);
  resize_h_kernel<<<dim3(B, cdiv(max_h, rows_per_cta)), out_w, smem_h, stream>>>(pixels, offsets, widths, heights, out_h, out_w,
                                                                                  max_h, (unsigned char*)workspace, rows_per_cta);
  MRNB_CHECK_LAUNCH("resize_h_kernel");
  resize_v_kernel<<<B, out_w, smem_v, stream>>>(pixels, offsets, widths, heights, out_h, out_w, max_h,
                                                (const unsigned char*)workspace, out);
  MRNB_CHECK_LAUNCH("resize_v_kernel");
  return MRNB_OK;
}
