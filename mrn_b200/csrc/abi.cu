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

// ---- profiling ---------------------------------------------------------------------------------------
#include <vector>
namespace {
struct ProfRec { int family; cudaEvent_t a, b; };
bool g_prof_on = false;
std::vector<ProfRec> g_recs;
std::vector<cudaEvent_t> g_pool;
double g_flops[MRNB_PROF_COUNT], g_bytes[MRNB_PROF_COUNT];
long g_calls[MRNB_PROF_COUNT];
cudaEvent_t g_open[MRNB_PROF_COUNT];
cudaEvent_t take_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
}  // namespace

void mrnb_prof_begin(int family, cudaStream_t st, double flops, double bytes) {
  if (!g_prof_on) return;
  cudaEvent_t e = take_event();
  cudaEventRecord(e, st);
  g_open[family] = e;
  g_flops[family] += flops; g_bytes[family] += bytes; g_calls[family] += 1;
}
void mrnb_prof_end(int family, cudaStream_t st) {
  if (!g_prof_on) return;
  cudaEvent_t e = take_event();
  cudaEventRecord(e, st);
  g_recs.push_back(ProfRec{family, g_open[family], e});
}
extern "C" void mrnb_profile_enable(int on) {
  g_prof_on = on != 0;
}
extern "C" void mrnb_profile_reset(void) {
  for (auto& r : g_recs) { g_pool.push_back(r.a); g_pool.push_back(r.b); }
  g_recs.clear();
  for (int k = 0; k < MRNB_PROF_COUNT; ++k) { g_flops[k] = 0; g_bytes[k] = 0; g_calls[k] = 0; }
}
// Sums the recorded intervals of one family (synchronises on the recorded events).
extern "C" int mrnb_profile_read(int family, double* ms, long* calls, double* flops, double* bytes) {
  if (family < 0 || family >= MRNB_PROF_COUNT) return MRNB_ERR_ARG;
  double t = 0.0;
  for (auto& r : g_recs) {
    if (r.family != family) continue;
    cudaEventSynchronize(r.b);
    float f = 0.f;
    if (cudaEventElapsedTime(&f, r.a, r.b) == cudaSuccess) t += f;
  }
  if (ms) *ms = t;
  if (calls) *calls = g_calls[family];
  if (flops) *flops = g_flops[family];
  if (bytes) *bytes = g_bytes[family];
  return MRNB_OK;
}

namespace {
__global__ void cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __float2bfloat16_rn(x[i]);
}
}  // namespace

namespace {
__global__ void cast_f16_kernel(const float* __restrict__ x, __half* __restrict__ y, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __float2half_rn(x[i]);
}
}  // namespace

extern "C" int mrnb_cast_f32_to_f16(const float* x, void* y, long n, cudaStream_t stream) {
  MRNB_CHECK_ARG(x && y && n > 0, "cast: bad argument");
  cast_f16_kernel<<<cdiv(n, 256), 256, 0, stream>>>(x, (__half*)y, n);
  MRNB_CHECK_LAUNCH("cast_f16_kernel");
  return MRNB_OK;
}

extern "C" int mrnb_cast_f32_to_bf16(const float* x, void* y, long n, cudaStream_t stream) {
  MRNB_CHECK_ARG(x && y && n > 0, "cast: bad argument");
  cast_bf16_kernel<<<cdiv(n, 256), 256, 0, stream>>>(x, (__nv_bfloat16*)y, n);
  MRNB_CHECK_LAUNCH("cast_bf16_kernel");
  return MRNB_OK;
}
