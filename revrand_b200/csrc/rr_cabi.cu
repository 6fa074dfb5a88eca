// Library-level C-ABI pieces: error string, version, device info, workspace
// queries.
#include <stdarg.h>
#include <new>
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

// RR_ENGINE_AUTO runs the tensor-core engine from this many rows on; below it the
// job is launch-latency sized and the chunked SIMT engine is as fast.
int64_t tc_auto_min_rows() { return 16384; }

size_t slm_workspace_bytes(int op, int64_t N, const rr_plan* pl, int engine);
size_t glm_workspace_bytes(int op, int64_t M, const rr_plan* pl, int S);

}  // namespace rr

// Caller-owned helper stream + events (header: Conventions).
struct rr_context {
  cudaStream_t aux;
  cudaEvent_t fork, a[2], b[2];
};

namespace rr {
int ctx_aux(rr_context* ctx, cudaStream_t* stream, cudaEvent_t* fork, cudaEvent_t a[2],
            cudaEvent_t b[2]) {
  if (ctx == nullptr) return RR_OK;
  *stream = ctx->aux;
  *fork = ctx->fork;
  for (int i = 0; i < 2; ++i) {
    a[i] = ctx->a[i];
    b[i] = ctx->b[i];
  }
  return RR_OK;
}
}  // namespace rr

extern "C" int rr_context_create(rr_context** out) {
  RR_REQUIRE(out != nullptr, "null pointer");
  rr_context* c = new (std::nothrow) rr_context();
  RR_REQUIRE(c != nullptr, "out of host memory");
  memset(c, 0, sizeof(*c));
  bool ok = cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&c->fork, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; i < 2; ++i) {
    ok = ok && cudaEventCreateWithFlags(&c->a[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->b[i], cudaEventDisableTiming) == cudaSuccess;
  }
  if (!ok) {
    rr::set_error("rr_context_create: %s", cudaGetErrorString(cudaGetLastError()));
    rr_context_destroy(c);
    return RR_ERR_CUDA;
  }
  *out = c;
  return RR_OK;
}

extern "C" int rr_context_destroy(rr_context* c) {
  if (c == nullptr) return RR_OK;
  if (c->fork) cudaEventDestroy(c->fork);
  for (int i = 0; i < 2; ++i) {
    if (c->a[i]) cudaEventDestroy(c->a[i]);
    if (c->b[i]) cudaEventDestroy(c->b[i]);
  }
  if (c->aux) cudaStreamDestroy(c->aux);
  delete c;
  return RR_OK;
}

extern "C" int64_t rr_engine_auto_min_rows(void) { return rr::tc_auto_min_rows(); }

extern "C" int rr_version(void) { return 200; }

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
    case RR_OP_GRADPASS_KEPT:
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
  return rr::tc3_suffstats_supported(&pl);
}
