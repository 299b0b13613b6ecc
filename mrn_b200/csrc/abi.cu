// Library-wide pieces of the C ABI: version, thread-local error text, launch counter, dtype cast.
#include "common.cuh"
#include <stdarg.h>
#include <atomic>

static thread_local char g_err[512] = "";
static std::atomic<long> g_launches{0};

void mrnb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" void mrnb_count_launch(int n) { g_launches += n; }
extern "C" long mrnb_launch_count(void) { return g_launches.load(); }
extern "C" void mrnb_reset_launch_count(void) { g_launches = 0; }
extern "C" const char* mrnb_last_error(void) { return g_err; }
extern "C" int mrnb_version(void) { return 100; }

namespace {
__global__ void cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __float2bfloat16_rn(x[i]);
}
}  // namespace

extern "C" int mrnb_cast_f32_to_bf16(const float* x, void* y, long n, cudaStream_t stream) {
  MRNB_CHECK_ARG(x && y && n > 0, "cast: bad argument");
  cast_bf16_kernel<<<cdiv(n, 256), 256, 0, stream>>>(x, (__nv_bfloat16*)y, n);
  MRNB_CHECK_LAUNCH("cast_bf16_kernel");
  return MRNB_OK;
}
