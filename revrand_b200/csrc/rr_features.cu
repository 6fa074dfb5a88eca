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
    Phi[(n0 + r) * ldphi + plan.ext_col[j]] = ext_value(plan, j, xs + r * d, 1);
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

// Radial / sigmoidal bases: thread per (row, centre).
__global__ void __launch_bounds__(256)
centre_features_kernel(const float* __restrict__ X, int64_t N, int d,
                       const float* __restrict__ C, int M, const float* __restrict__ ls,
                       int P, int kind, float* __restrict__ Phi, float* __restrict__ dPhi) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * M) return;
  const int64_t n = idx / M;
  const int c = (int)(idx - n * M);
  const float* x = X + n * d;
  const float* cc = C + (int64_t)c * d;
  float acc = 0.0f;
  for (int i = 0; i < d; ++i) {
    const float l = ls[P == 1 ? 0 : i];
    // radial: the reference divides by 2 l^2 BEFORE squaring (:686-688)
    const float t = (x[i] - cc[i]) / (kind == 0 ? 2.0f * l * l : l);
    acc = fmaf(t, t, acc);
  }
  const float phi = kind == 0 ? expf(-acc) : 1.0f / (1.0f + expf(-sqrtf(acc)));
  Phi[idx] = phi;
  if (dPhi) {
    for (int i = 0; i < P; ++i) {      // P == 1: input dimension 0 only, as the reference
      const float l = ls[i];
      const float df = x[i] - cc[i];
      float g;
      if (kind == 0) {
        const float l3 = l * l * l;
        g = phi * (df / l3) * (df / l3);
      } else {
        g = -fabsf(df) / (l * l) * phi * (1.0f - phi);
      }
      dPhi[idx * P + i] = g;
    }
  }
}

// FastFoodGM.grad from the dense projection image: thread per (row, frequency).
__global__ void __launch_bounds__(256)
gm_grad_kernel(const float* __restrict__ X, int64_t N, int d, const float* __restrict__ V,
               int n, const float* __restrict__ mean, const float* __restrict__ ls,
               float* __restrict__ dmean, float* __restrict__ dlen) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * n) return;
  const int64_t r = idx / n;
  const int k = (int)(idx - r * n);
  const float* x = X + r * d;
  const float inv2pi = 0.15915494309189535f;
  float p = 0.0f, q = 0.0f;          // phases in turns
  for (int i = 0; i < d; ++i) {
    p = fmaf(x[i], V[(int64_t)i * n + k] / ls[i] * inv2pi, p);
    q = fmaf(x[i], mean[i] * inv2pi, q);
  }
  float sp, cp, sm, cm;
  sincos_turns(p + q, &sp, &cp);
  sincos_turns(p - q, &sm, &cm);
  const float amp = rsqrtf(2.0f * (float)n);
  const int64_t F = 4 * (int64_t)n;
  for (int i = 0; i < d; ++i) {
    const float dmx = x[i];
    const float dvx = -x[i] * V[(int64_t)i * n + k] / (ls[i] * ls[i]);
    // blocks [cos(p+q) | sin(p+q) | cos(p-q) | sin(p-q)]; d/dphase cos = -sin, sin = cos
    float* om = dmean + (r * F) * d + i;
    float* ol = dlen + (r * F) * d + i;
    om[(int64_t)(k) * d] = amp * dmx * (-sp);
    om[(int64_t)(n + k) * d] = amp * dmx * cp;
    om[(int64_t)(2 * n + k) * d] = amp * (-dmx) * (-sm);
    om[(int64_t)(3 * n + k) * d] = amp * (-dmx) * cm;
    ol[(int64_t)(k) * d] = amp * dvx * (-sp);
    ol[(int64_t)(n + k) * d] = amp * dvx * cp;
    ol[(int64_t)(2 * n + k) * d] = amp * dvx * (-sm);
    ol[(int64_t)(3 * n + k) * d] = amp * dvx * cm;
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

extern "C" int rr_centre_features(const float* X, int64_t N, int32_t d, const float* C,
                                  int32_t M, const float* lenscale, int32_t n_ls,
                                  int32_t kind, float* Phi, float* dPhi, void* stream) {
  using namespace rr;
  RR_REQUIRE(X && C && lenscale && Phi, "null pointer");
  RR_REQUIRE(d > 0 && M > 0 && (n_ls == 1 || n_ls == d), "bad shape");
  RR_REQUIRE(kind == 0 || kind == 1, "kind must be 0 (radial) or 1 (sigmoidal)");
  if (N == 0) return RR_OK;
  const int64_t total = N * M;
  centre_features_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      X, N, d, C, M, lenscale, n_ls, kind, Phi, dPhi);
  RR_LAUNCH_CHECK("centre_features_kernel");
  return RR_OK;
}

extern "C" int rr_gm_grad(const float* X, int64_t N, int32_t d, const float* V, int32_t n,
                          const float* mean, const float* lenscale, float* dmean,
                          float* dlen, void* stream) {
  using namespace rr;
  RR_REQUIRE(X && V && mean && lenscale && dmean && dlen, "null pointer");
  RR_REQUIRE(d > 0 && n > 0, "bad shape");
  if (N == 0) return RR_OK;
  const int64_t total = N * n;
  gm_grad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      X, N, d, V, n, mean, lenscale, dmean, dlen);
  RR_LAUNCH_CHECK("gm_grad_kernel");
  return RR_OK;
}
