// Library-level C-ABI pieces: error string, version, device info, workspace
// queries.
#include <stdarg.h>
#include <string.h>

#include "rr_common.cuh"

namespace rr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
unsigned long long launches() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

size_t slm_workspace_bytes(int op, int64_t N, const rr_plan* pl, int engine);
size_t glm_workspace_bytes(int op, int64_t M, const rr_plan* pl, int S);

}  // namespace rr

extern "C" int rr_version(void) { return 100; }

extern "C" uint64_t rr_launch_count(void) { return (uint64_t)rr::launches(); }

extern "C" const char* rr_last_error(void) { return rr::g_err; }

extern "C" int rr_device_info(int32_t* sms, int32_t* maj, int32_t* min) {
  int dev = 0;
  RR_CUDA_CHECK(cudaGetDevice(&dev));
  int a = 0, b = 0, c = 0;
  RR_CUDA_CHECK(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev));
  RR_CUDA_CHECK(cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev));
  RR_CUDA_CHECK(cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev));
  if (sms) *sms = a;
  if (maj) *maj = b;
  if (min) *min = c;
  return RR_OK;
}

extern "C" size_t rr_workspace_bytes(int32_t op, int64_t N, int32_t d,
                                     int32_t ktot, int32_t D, int32_t aux0,
                                     int32_t aux1, int32_t engine) {
  rr_plan pl;
  memset(&pl, 0, sizeof(pl));
  pl.d = d;
  pl.ktot = ktot;
  pl.D = D;
  pl.next = D - 2 * ktot;
  switch (op) {
    case RR_OP_SUFFSTATS:
    case RR_OP_GRADPASS:
    case RR_OP_PREDICT:
    case RR_OP_RESIDUAL:
      return rr::slm_workspace_bytes(op, N, &pl, engine);
    case RR_OP_GLM_STEP:
      return rr::glm_workspace_bytes(op, N, &pl, aux0 * aux1);
    case RR_OP_GLM_PREDICT:
      return rr::glm_workspace_bytes(op, N, &pl, aux0);
    default:
      return 0;
  }
}

extern "C" int rr_tcgen05_supported(int32_t d, int32_t ktot, int32_t next,
                                    int32_t D) {
  rr_plan pl;
  memset(&pl, 0, sizeof(pl));
  pl.d = d;
  pl.ktot = ktot;
  pl.next = next;
  pl.D = D;
  return rr::tc_suffstats_supported(&pl);
}
