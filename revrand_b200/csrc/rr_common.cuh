// Shared helpers for the revrand_b200 CUDA sources (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/revrand_b200.h"

namespace rr {

void set_error(const char* fmt, ...);
void count_launch();

#define RR_CUDA_CHECK(expr)                                                  \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) {                                                 \
      rr::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,     \
                    cudaGetErrorString(_e));                                 \
      return RR_ERR_CUDA;                                                    \
    }                                                                        \
  } while (0)

#define RR_LAUNCH_CHECK(name)                                                \
  do {                                                                       \
    rr::count_launch();                                                      \
    cudaError_t _e = cudaGetLastError();                                     \
    if (_e != cudaSuccess) {                                                 \
      rr::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));\
      return RR_ERR_CUDA;                                                    \
    }                                                                        \
  } while (0)

#define RR_REQUIRE(cond, msg)                                                \
  do {                                                                       \
    if (!(cond)) {                                                           \
      rr::set_error("invalid argument: %s (%s)", msg, #cond);                \
      return RR_ERR_INVALID;                                                 \
    }                                                                        \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over the caller-provided workspace.
struct Workspace {
  char* base;
  size_t size;
  size_t off;
  Workspace(void* p, size_t n) : base((char*)p), size(n), off(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t o = align_up(off, 256);
    size_t need = o + count * sizeof(T);
    if (need > size || base == nullptr) return nullptr;
    off = need;
    return (T*)(base + o);
  }
};

int sm_count();
int64_t tc_auto_min_rows();

// Helper stream + events of an rr_context (NULL ctx: *stream is left alone).
int ctx_aux(rr_context* ctx, cudaStream_t* stream, cudaEvent_t* fork, cudaEvent_t a[2],
            cudaEvent_t b[2]);

// ---- device helpers --------------------------------------------------------

// sin/cos of 2*pi*u for u in turns, with exact range reduction in fp32:
// u - rint(u) is exact for |u| < 2^23, leaving r in [-0.5, 0.5].
__device__ __forceinline__ void sincos_turns(float u, float* s, float* c) {
  float r = u - rintf(u);
  sincospif(2.0f * r, s, c);
}

// x^p for a small non-negative integer p (polynomial columns of a plan).
__device__ __forceinline__ float ipowf(float x, int p) {
  float r = 1.0f;
  for (int i = 0; i < p; ++i) r *= x;
  return r;
}
// value of extra (affine / polynomial / constant) column j of a plan for one row
__device__ __forceinline__ float ext_value(const rr_plan& plan, int j, const float* xrow,
                                           int stride) {
  const int src = plan.ext_src[j];
  if (src < 0) return plan.ext_val[j];
  const float x = xrow[src * stride];
  return plan.ext_pow ? ipowf(x, plan.ext_pow[j]) : x;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- launchers implemented across the .cu files ----------------------------

// Generic strided SGEMM on CUDA cores: C (+)= alpha * A' B', A'(m,k) =
// A[m*sAm + k*sAk], B'(k,n) = B[k*sBk + n*sBn]; C row-major with ldc.  Output
// either float (Cf) or float64 (Cd); accumulate != 0 adds to C.
int sgemm(int M, int N, int K, float alpha, const float* A, int64_t sAm,
          int64_t sAk, const float* B, int64_t sBk, int64_t sBn, float* Cf,
          double* Cd, int64_t ldc, int accumulate, cudaStream_t st);

// The same product (fp32 output) on the tcgen05 tensor cores, operands split into two
// tf32 parts, three products (rr_tc_gemm3.cu).  imgA / imgB: scratch of
// gemm3_image_bytes(M, K) / gemm3_image_bytes(N, K) bytes.
size_t gemm3_image_bytes(int64_t R, int64_t K);
bool gemm3_worthwhile(int M, int N, int K);
int gemm3(int M, int N, int K, float alpha, const float* A, int64_t sAm, int64_t sAk,
          const float* B, int64_t sBk, int64_t sBn, float* C, int64_t ldc, int accumulate,
          uint8_t* imgA, uint8_t* imgB, cudaStream_t st);

// R (d x kt, float64) += X^T Q for row-major fp32 X (rows x d), Q (rows x kt).
int xtq(const float* X, const float* Q, int rows, int d, int kt, double* R, cudaStream_t st);

int launch_features(const rr_plan* plan, const float* X, int64_t N, float* Phi,
                    int64_t ldphi, cudaStream_t st);

// tcgen05 engines (rr_tc_*.cu)
int tc_suffstats_supported(const rr_plan* plan);
size_t tc_suffstats_workspace(const rr_plan* plan, int64_t N);
int tc_suffstats(const rr_plan* plan, const float* X, const float* y, int64_t N,
                 double* G, double* p, void* ws, size_t ws_bytes, int grid_bits,
                 cudaStream_t st);
int tc3_suffstats_supported(const rr_plan* plan);
size_t tc3_suffstats_workspace(const rr_plan* plan, int64_t N);
// kept != NULL: also leave the fp16 feature image of the gradient pass behind
int tc3_suffstats(const rr_plan* plan, const float* X, const float* y, int64_t N,
                  double* G, double* p, void* ws, size_t ws_bytes, rr_context* ctx,
                  cudaStream_t st, void* kept = nullptr);
// Kept features (rr_slm_suffstats_keep -> rr_slm_gradpass_kept): padded column count of
// the image (a multiple of 64) and its size in bytes for N rows.
int64_t kept_features_cols(const rr_plan* plan);
size_t kept_features_bytes(const rr_plan* plan, int64_t N);
size_t tc_gradpass_kept_workspace(const rr_plan* plan, int64_t N, bool split_c);
int tc_gradpass_kept(const rr_plan* plan, const float* X, const float* y, int64_t N,
                     const float* m, const float* C, double* R, double* sqerr,
                     const void* kept, void* ws, size_t ws_bytes, bool split_c,
                     cudaStream_t st);
// split_c: a second GEMM over the fp16 rounding residual of C (RR_GRAD_SPLIT_C)
size_t tc_gradpass_workspace(const rr_plan* plan, int64_t N, bool split_c);
int tc_gradpass_supported(const rr_plan* plan);
int tc_gradpass(const rr_plan* plan, const float* X, const float* y, int64_t N,
                const float* m, const float* C, double* R, double* sqerr, void* ws,
                size_t ws_bytes, rr_context* ctx, bool split_c, cudaStream_t st);
int phi_residual(const rr_plan* plan, const float* X, const float* y, int64_t N,
                 const float* m, float* err, double* sqerr, float* fbuf,
                 cudaStream_t st);

}  // namespace rr
