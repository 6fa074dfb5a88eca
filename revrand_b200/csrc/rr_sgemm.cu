// Generic strided fp32 GEMM on CUDA cores (FFMA), used by the chunked SIMT
// engine (any feature plan) and by the GLM step.  128x128x16 block tile, 256
// threads, 8x8 register tile per thread, fp32 accumulation, optional float64
// read-modify-write epilogue so long row sums are carried in double.  (An
// mma.sync tf32 three-product variant for the fp32 outputs was measured at
// 1.65 ms against 0.99 ms for this kernel on the GLM step's GEMMs -- 182
// registers, one block per SM -- and dropped.)
#include <type_traits>

#include "rr_common.cuh"

namespace rr {

constexpr int GB_M = 128, GB_N = 128, GB_K = 16, G_THREADS = 256;

template <bool OUT_DOUBLE>
__global__ void __launch_bounds__(G_THREADS)
sgemm_kernel(int M, int N, int K, float alpha, const float* __restrict__ A,
             int64_t sAm, int64_t sAk, const float* __restrict__ B, int64_t sBk,
             int64_t sBn, float* __restrict__ Cf, double* __restrict__ Cd,
             int64_t ldc, int accumulate) {
  __shared__ float As[GB_K][GB_M + 4];
  __shared__ float Bs[GB_K][GB_N + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GB_M, n0 = blockIdx.x * GB_N;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 thread grid

  // float64 outputs carry float64 accumulators (exact products of the fp32
  // inputs): posterior solves amplify Gram errors by the conditioning of the
  // problem, so fp32 accumulation is not enough for parity with the reference.
  using acc_t = typename std::conditional<OUT_DOUBLE, double, float>::type;
  acc_t acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = (acc_t)0;

  // Load mapping: pick the orientation whose fastest index is contiguous.
  const bool a_m_fast = (sAm == 1);
  const bool b_n_fast = (sBn == 1);

  for (int k0 = 0; k0 < K; k0 += GB_K) {
#pragma unroll
    for (int it = 0; it < (GB_M * GB_K) / G_THREADS; ++it) {
      int e = tid + it * G_THREADS;
      int mm, kk;
      if (a_m_fast) { mm = e % GB_M; kk = e / GB_M; }
      else          { kk = e % GB_K; mm = e / GB_K; }
      int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < K) ? A[gm * sAm + gk * sAk] : 0.0f;
    }
#pragma unroll
    for (int it = 0; it < (GB_N * GB_K) / G_THREADS; ++it) {
      int e = tid + it * G_THREADS;
      int nn, kk;
      if (b_n_fast) { nn = e % GB_N; kk = e / GB_N; }
      else          { kk = e % GB_K; nn = e / GB_K; }
      int gn = n0 + nn, gk = k0 + kk;
      Bs[kk][nn] = (gn < N && gk < K) ? B[gk * sBk + gn * sBn] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GB_K; ++kk) {
      float a[8], b[8];
      // rows ty*4..+3 and 64+ty*4..+3; cols tx*4..+3 and 64+tx*4..+3
      float4 a0 = *(const float4*)&As[kk][ty * 4];
      float4 a1 = *(const float4*)&As[kk][64 + ty * 4];
      float4 b0 = *(const float4*)&Bs[kk][tx * 4];
      float4 b1 = *(const float4*)&Bs[kk][64 + tx * 4];
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (OUT_DOUBLE) acc[i][j] = fma((double)a[i], (double)b[j], (double)acc[i][j]);
          else acc[i][j] = fmaf(a[i], b[j], (float)acc[i][j]);
        }
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (gn >= N) continue;
      if (OUT_DOUBLE) {
        const double v = (double)alpha * (double)acc[i][j];
        double* c = Cd + (int64_t)gm * ldc + gn;
        *c = accumulate ? (*c + v) : v;
      } else {
        const float v = alpha * (float)acc[i][j];
        float* c = Cf + (int64_t)gm * ldc + gn;
        *c = accumulate ? (*c + v) : v;
      }
    }
  }
}

// R[i][k] += sum_r X[r][i] * Q[r][k]  (d x kt float64, rows x d and rows x kt
// fp32 inputs): the X^T Q reduction of the hyper-gradient.  The generic kernel
// above would run this short-and-wide product on kt / 128 blocks; here the rows
// are split over the grid as well and the partial sums meet in float64 atomics.
constexpr int XTQ_ROWS = 256, XTQ_DMAX = 64;
__global__ void __launch_bounds__(128)
xtq_kernel(const float* __restrict__ X, const float* __restrict__ Q, int rows, int d, int kt,
           double* __restrict__ R) {
  extern __shared__ float xs[];   // XTQ_ROWS x d
  const int k = blockIdx.x * 128 + threadIdx.x;
  const int r0 = blockIdx.y * XTQ_ROWS;
  const int nr = min(XTQ_ROWS, rows - r0);
  for (int e = threadIdx.x; e < nr * d; e += 128) xs[e] = X[(int64_t)r0 * d + e];
  __syncthreads();
  if (k >= kt) return;
  for (int i0 = 0; i0 < d; i0 += 16) {      // 16 input dimensions per sweep of the rows
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.0f;
    for (int r = 0; r < nr; ++r) {
      const float q = Q[(int64_t)(r0 + r) * kt + k];
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i0 + i < d) acc[i] = fmaf(xs[r * d + i0 + i], q, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i0 + i < d && acc[i] != 0.0f) atomicAdd(R + (int64_t)(i0 + i) * kt + k, (double)acc[i]);
  }
}

int xtq(const float* X, const float* Q, int rows, int d, int kt, double* R, cudaStream_t st) {
  if (rows <= 0 || kt <= 0 || d <= 0) return RR_OK;
  dim3 grid((kt + 127) / 128, (rows + XTQ_ROWS - 1) / XTQ_ROWS);
  const size_t smem = (size_t)XTQ_ROWS * d * sizeof(float);
  if (smem > 48 * 1024) {      // very wide inputs: the generic kernel
    return sgemm(d, kt, rows, 1.0f, X, 1, d, Q, kt, 1, nullptr, R, kt, 1, st);
  }
  (void)XTQ_DMAX;
  xtq_kernel<<<grid, 128, smem, st>>>(X, Q, rows, d, kt, R);
  RR_LAUNCH_CHECK("xtq_kernel");
  return RR_OK;
}

int sgemm(int M, int N, int K, float alpha, const float* A, int64_t sAm,
          int64_t sAk, const float* B, int64_t sBk, int64_t sBn, float* Cf,
          double* Cd, int64_t ldc, int accumulate, cudaStream_t st) {
  if (M <= 0 || N <= 0) return RR_OK;
  dim3 grid((N + GB_N - 1) / GB_N, (M + GB_M - 1) / GB_M);
  if (Cd)
    sgemm_kernel<true><<<grid, G_THREADS, 0, st>>>(M, N, K, alpha, A, sAm, sAk, B,
                                                  sBk, sBn, nullptr, Cd, ldc,
                                                  accumulate);
  else
    sgemm_kernel<false><<<grid, G_THREADS, 0, st>>>(M, N, K, alpha, A, sAm, sAk,
                                                   B, sBk, sBn, Cf, nullptr, ldc,
                                                   accumulate);
  RR_LAUNCH_CHECK("sgemm_kernel");
  return RR_OK;
}

}  // namespace rr
