// Generic strided fp32 GEMM, used by the chunked SIMT engine (any feature plan)
// and by the GLM step.  128x128x16 block tile, 256 threads.  float64 outputs
// (long row sums feeding the posterior solve) accumulate exact fp32 products in
// float64 on the CUDA cores, 8x8 register tile per thread; fp32 outputs run on
// the tensor cores (mma.sync tf32, three-product split).
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

// ---------------------------------------------------------------------------
// fp32-output variant on the tensor cores: mma.sync m16n8k8 tf32 with the
// three-product split  hi*hi + lo*hi + hi*lo  of tf32-rounded operands (about 21
// bits per product: fp32 grade for the fp32 outputs it serves).  Same block tile
// and global -> shared staging as above; the operands are split once when they
// are staged.  8 warps as 2 x 4, warp tile 64 x 32 = 4 x 4 mma tiles.
// ---------------------------------------------------------------------------
constexpr int GT_PAD = 8;    // row pitch 136 floats: fragment loads touch 32 distinct banks

__device__ __forceinline__ float tf32_round(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, "
      "{%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(G_THREADS)
sgemm_tf32x3_kernel(int M, int N, int K, float alpha, const float* __restrict__ A,
                    int64_t sAm, int64_t sAk, const float* __restrict__ B, int64_t sBk,
                    int64_t sBn, float* __restrict__ Cf, int64_t ldc, int accumulate) {
  __shared__ float Ah[GB_K][GB_M + GT_PAD], Al[GB_K][GB_M + GT_PAD];
  __shared__ float Bh[GB_K][GB_N + GT_PAD], Bl[GB_K][GB_N + GT_PAD];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.y * GB_M, n0 = blockIdx.x * GB_N;
  const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;   // warp tile origin in the block tile
  const int g = lane >> 2, q = lane & 3;
  float acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.0f;
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
      const float v = (gm < M && gk < K) ? A[gm * sAm + gk * sAk] : 0.0f;
      const float h = tf32_round(v);
      Ah[kk][mm] = h;
      Al[kk][mm] = tf32_round(v - h);
    }
#pragma unroll
    for (int it = 0; it < (GB_N * GB_K) / G_THREADS; ++it) {
      int e = tid + it * G_THREADS;
      int nn, kk;
      if (b_n_fast) { nn = e % GB_N; kk = e / GB_N; }
      else          { kk = e % GB_K; nn = e / GB_K; }
      int gn = n0 + nn, gk = k0 + kk;
      const float v = (gn < N && gk < K) ? B[gk * sBk + gn * sBn] : 0.0f;
      const float h = tf32_round(v);
      Bh[kk][nn] = h;
      Bl[kk][nn] = tf32_round(v - h);
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < GB_K; ks += 8) {
      // B fragments (k x n, "col"): b0 = (k = q, n = g), b1 = (k = q + 4, n = g)
      uint32_t bh[4][2], bl[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = wn + 8 * j + g;
        bh[j][0] = __float_as_uint(Bh[ks + q][n]);
        bh[j][1] = __float_as_uint(Bh[ks + q + 4][n]);
        bl[j][0] = __float_as_uint(Bl[ks + q][n]);
        bl[j][1] = __float_as_uint(Bl[ks + q + 4][n]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // A fragment (m x k, "row"): a0 = (g, q), a1 = (g + 8, q), a2 = (g, q + 4), a3 = (g + 8, q + 4)
        const int m = wm + 16 * i + g;
        uint32_t ah[4], al[4];
        ah[0] = __float_as_uint(Ah[ks + q][m]);
        ah[1] = __float_as_uint(Ah[ks + q][m + 8]);
        ah[2] = __float_as_uint(Ah[ks + q + 4][m]);
        ah[3] = __float_as_uint(Ah[ks + q + 4][m + 8]);
        al[0] = __float_as_uint(Al[ks + q][m]);
        al[1] = __float_as_uint(Al[ks + q][m + 8]);
        al[2] = __float_as_uint(Al[ks + q + 4][m]);
        al[3] = __float_as_uint(Al[ks + q + 4][m + 8]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          mma_tf32(acc[i][j], al, bh[j][0], bh[j][1]);
          mma_tf32(acc[i][j], ah, bl[j][0], bl[j][1]);
          mma_tf32(acc[i][j], ah, bh[j][0], bh[j][1]);
        }
      }
    }
    __syncthreads();
  }
  // accumulator fragment: c0 = (g, 2q), c1 = (g, 2q + 1), c2 = (g + 8, 2q), c3 = (g + 8, 2q + 1)
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int gm = m0 + wm + 16 * i + g + ((e & 2) ? 8 : 0);
        const int gn = n0 + wn + 8 * j + 2 * q + (e & 1);
        if (gm < M && gn < N) {
          float* c = Cf + (int64_t)gm * ldc + gn;
          const float v = alpha * acc[i][j][e];
          *c = accumulate ? (*c + v) : v;
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
    sgemm_tf32x3_kernel<<<grid, G_THREADS, 0, st>>>(M, N, K, alpha, A, sAm, sAk, B, sBk,
                                                   sBn, Cf, ldc, accumulate);
  RR_LAUNCH_CHECK("sgemm_kernel");
  return RR_OK;
}

}  // namespace rr
