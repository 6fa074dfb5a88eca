// HBM-bound feature-map kernels: the public transform/grad API of the random
// kernel bases.  These MUST write Phi (the caller asked for it); the fused SLM
// passes in rr_tc_*.cu never do.
//
// Reference call sites: revrand/basis_functions.py:859-864 (transform),
// :888-901 (grad), :1356-1371 + mathfun/linalg.py:182-220 (FastFood).
#include "rr_common.cuh"

namespace rr {

// Block = FEAT_ROWS rows x all columns (looped).  Each thread owns a frequency
// column k and FEAT_ROWS accumulators, so W is read once per row tile and the
// stores of cos/sin are coalesced across k.
constexpr int FEAT_ROWS = 16;
constexpr int FEAT_THREADS = 256;

__global__ void __launch_bounds__(FEAT_THREADS)
features_kernel(rr_plan plan, const float* __restrict__ X, int64_t N,
                float* __restrict__ Phi, int64_t ldphi) {
  extern __shared__ float xs[];  // FEAT_ROWS x d
  const int d = plan.d;
  const int64_t n0 = (int64_t)blockIdx.x * FEAT_ROWS;
  const int rows = (int)min((int64_t)FEAT_ROWS, N - n0);
  for (int t = threadIdx.x; t < FEAT_ROWS * d; t += blockDim.x) {
    int r = t / d;
    xs[t] = (r < rows) ? X[(n0 + r) * d + (t - r * d)] : 0.0f;
  }
  __syncthreads();

  for (int k = threadIdx.x; k < plan.ktot; k += blockDim.x) {
    float u[FEAT_ROWS];
#pragma unroll
    for (int r = 0; r < FEAT_ROWS; ++r) u[r] = 0.0f;
    for (int i = 0; i < d; ++i) {
      float w = __ldg(plan.Wt + (int64_t)i * plan.ktot + k);
#pragma unroll
      for (int r = 0; r < FEAT_ROWS; ++r) u[r] = fmaf(xs[r * d + i], w, u[r]);
    }
    const float a = plan.amp[k];
    const int cc = plan.col_cos[k], cs = plan.col_sin[k];
#pragma unroll
    for (int r = 0; r < FEAT_ROWS; ++r) {
      if (r < rows) {
        float s, c;
        sincos_turns(u[r], &s, &c);
        Phi[(n0 + r) * ldphi + cc] = a * c;
        Phi[(n0 + r) * ldphi + cs] = a * s;
      }
    }
  }
  for (int t = threadIdx.x; t < rows * plan.next; t += blockDim.x) {
    int r = t / plan.next, j = t - r * plan.next;
    int src = plan.ext_src[j];
    float v = src >= 0 ? xs[r * d + src] : plan.ext_val[j];
    Phi[(n0 + r) * ldphi + plan.ext_col[j]] = v;
  }
}

int launch_features(const rr_plan* plan, const float* X, int64_t N, float* Phi,
                    int64_t ldphi, cudaStream_t st) {
  if (N == 0) return RR_OK;
  RR_REQUIRE(plan->d > 0 && plan->d <= 2048, "input dimension out of range");
  int64_t blocks = (N + FEAT_ROWS - 1) / FEAT_ROWS;
  size_t smem = (size_t)FEAT_ROWS * plan->d * sizeof(float);
  if (smem > 48 * 1024)
    RR_CUDA_CHECK(cudaFuncSetAttribute(
        features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  features_kernel<<<(unsigned)blocks, FEAT_THREADS, smem, st>>>(*plan, X, N, Phi,
                                                              ldphi);
  RR_LAUNCH_CHECK("features_kernel");
  return RR_OK;
}

// d Phi / d lenscale for one trig block, reference layout (N, 2K[, P]).
__global__ void __launch_bounds__(256)
trig_grad_kernel(const float* __restrict__ X, int64_t N, int d,
                 const float* __restrict__ W, int K,
                 const float* __restrict__ ls, int P, int compat,
                 float* __restrict__ out) {
  extern __shared__ float sh[];  // x row (d), inverse lenscales (d)
  float* xr = sh;
  float* il = sh + d;
  const int64_t n = blockIdx.x;
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    xr[i] = X[n * d + i];
    il[i] = 1.0f / ls[P == 1 ? 0 : i];
  }
  __syncthreads();
  const float amp = rsqrtf((float)K);
  const float inv2pi = 0.15915494309189535f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float u = 0.0f;
    for (int i = 0; i < d; ++i)
      u = fmaf(xr[i], W[(int64_t)i * K + k] * il[i] * inv2pi, u);
    float s, c;
    sincos_turns(u, &s, &c);
    if (P == 1) {
      // dWX = x_i * (-W_ik / l^2); the reference sums over i = 0 only.
      float g;
      if (compat) {
        g = -xr[0] * W[k] * il[0] * il[0];
      } else {
        float th = 0.0f;
        for (int i = 0; i < d; ++i) th = fmaf(xr[i], W[(int64_t)i * K + k], th);
        g = -th * il[0] * il[0];
      }
      out[n * 2 * K + k] = g * (-s) * amp;
      out[n * 2 * K + K + k] = g * c * amp;
    } else {
      float* oc = out + (n * 2 * K + k) * P;
      float* os = out + (n * 2 * K + K + k) * P;
      for (int i = 0; i < P; ++i) {
        float g = -xr[i] * W[(int64_t)i * K + k] * il[i] * il[i];
        oc[i] = g * (-s) * amp;
        os[i] = g * c * amp;
      }
    }
  }
}

// FastFood: one warp per (row, block) when d2 == 32 would be ideal; to keep
// any power of two d2 <= 1024 working, one thread block handles one row and
// loops over the k blocks with the butterfly done in shared memory.
__global__ void __launch_bounds__(256)
fastfood_kernel(const float* __restrict__ Xs, int64_t N, int d, int d2, int kb,
                const float* __restrict__ B, const float* __restrict__ G,
                const int* __restrict__ PI, const float* __restrict__ S,
                float* __restrict__ Phi, float* __restrict__ VX) {
  extern __shared__ float sh[];  // 2 * d2
  float* a = sh;
  float* b = sh + d2;
  const int64_t n = blockIdx.x;
  const int nfreq = kb * d2;
  const float amp = rsqrtf((float)nfreq);
  const float sq = sqrtf((float)d2);
  const float inv2pi = 0.15915494309189535f;
  for (int blk = 0; blk < kb; ++blk) {
    for (int j = threadIdx.x; j < d2; j += blockDim.x)
      a[j] = (j < d ? Xs[n * d + j] : 0.0f) * B[blk * d2 + j];
    __syncthreads();
    // H/d2 as log2(d2) halving butterflies (mathfun/linalg.py:212-216).
    for (int h = 1; h < d2; h <<= 1) {
      for (int j = threadIdx.x; j < d2; j += blockDim.x) {
        int lo = j & ~h, hi = j | h;
        float v = (j & h) ? (a[lo] - a[hi]) : (a[lo] + a[hi]);
        b[j] = 0.5f * v;
      }
      __syncthreads();
      float* t = a; a = b; b = t;
    }
    for (int j = threadIdx.x; j < d2; j += blockDim.x)
      b[j] = a[PI[blk * d2 + j]] * G[blk * d2 + j];
    __syncthreads();
    { float* t = a; a = b; b = t; }
    for (int h = 1; h < d2; h <<= 1) {
      for (int j = threadIdx.x; j < d2; j += blockDim.x) {
        int lo = j & ~h, hi = j | h;
        float v = (j & h) ? (a[lo] - a[hi]) : (a[lo] + a[hi]);
        b[j] = 0.5f * v;
      }
      __syncthreads();
      float* t = a; a = b; b = t;
    }
    for (int j = threadIdx.x; j < d2; j += blockDim.x) {
      float v = a[j] * S[blk * d2 + j] * sq;
      int col = blk * d2 + j;
      if (VX) VX[n * nfreq + col] = v;
      if (Phi) {
        float s, c;
        sincos_turns(v * inv2pi, &s, &c);
        Phi[n * 2 * nfreq + col] = amp * c;
        Phi[n * 2 * nfreq + nfreq + col] = amp * s;
      }
    }
    __syncthreads();
  }
}

}  // namespace rr

extern "C" int rr_features(const rr_plan* plan, const float* X, int64_t N,
                           float* Phi, int64_t ldphi, void* stream) {
  RR_REQUIRE(plan && X && Phi, "null pointer");
  RR_REQUIRE(ldphi >= plan->D, "ldphi < D");
  return rr::launch_features(plan, X, N, Phi, ldphi, (cudaStream_t)stream);
}

extern "C" int rr_trig_grad(const float* X, int64_t N, int32_t d, const float* W,
                            int32_t K, const float* lenscale, int32_t n_ls,
                            int32_t compat, float* dPhi, void* stream) {
  RR_REQUIRE(X && W && lenscale && dPhi, "null pointer");
  RR_REQUIRE(n_ls == 1 || n_ls == d, "lenscale must have 1 or d entries");
  if (N == 0) return RR_OK;
  size_t smem = 2 * (size_t)d * sizeof(float);
  rr::trig_grad_kernel<<<(unsigned)N, 256, smem, (cudaStream_t)stream>>>(
      X, N, d, W, K, lenscale, n_ls, compat, dPhi);
  RR_LAUNCH_CHECK("trig_grad_kernel");
  return RR_OK;
}

extern "C" int rr_fastfood_features(const float* Xs, int64_t N, int32_t d,
                                    int32_t d2, int32_t k, const float* B,
                                    const float* G, const int32_t* PI,
                                    const float* S, float* Phi, float* VX_out,
                                    void* stream) {
  RR_REQUIRE(Xs && B && G && PI && S, "null pointer");
  RR_REQUIRE(d2 >= 1 && (d2 & (d2 - 1)) == 0 && d <= d2, "d2 must be a power of two >= d");
  if (N == 0) return RR_OK;
  size_t smem = 2 * (size_t)d2 * sizeof(float);
  rr::fastfood_kernel<<<(unsigned)N, 256, smem, (cudaStream_t)stream>>>(
      Xs, N, d, d2, k, B, G, PI, S, Phi, VX_out);
  RR_LAUNCH_CHECK("fastfood_kernel");
  return RR_OK;
}
