// Data-dependent part of one GLM SVI step (all mixture components and
// reparameterised draws batched) and the GLM predictive sampler.
//
// Reference: revrand/glm.py:296-322 (_reparam_k), :274-275 (EdPhi contraction),
// :404-418 / :572-620 (predict_moments / _sample_func);
// revrand/likelihoods.py (loglike / df / dp / Ey per likelihood).
#include "rr_common.cuh"

namespace rr {

constexpr int64_t GLM_CHUNK = 8192;

__device__ __forceinline__ float softplusf(float f) {
  // log(1 + exp(f)), stable on both tails (mathfun/special.py:91-124).
  return fmaxf(f, 0.0f) + log1pf(expf(-fabsf(f)));
}
__device__ __forceinline__ float expitf(float f) { return 1.0f / (1.0f + expf(-f)); }

// likelihoods.py: Gaussian :298-396, Bernoulli :46-104, Binomial :171-233,
// Poisson :456-521.
__device__ __forceinline__ void lik_eval(int lik, float y, float f, float par,
                                         float arg, bool want_ll, float* df,
                                         float* dp, float* ll) {
  *dp = 0.0f;
  *ll = 0.0f;
  switch (lik) {
    case RR_LIK_GAUSSIAN: {
      float iv = 1.0f / par, r = y - f;
      *df = r * iv;
      *dp = 0.5f * (r * iv * r * iv - iv);
      if (want_ll) *ll = -0.5f * (logf(6.283185307179586f * par) + r * r * iv);
      break;
    }
    case RR_LIK_BERNOULLI:
      *df = y - expitf(f);
      if (want_ll) *ll = y * f - softplusf(f);
      break;
    case RR_LIK_BINOMIAL:
      *df = y - expitf(f) * arg;
      if (want_ll)
        *ll = lgammaf(arg + 1.0f) - lgammaf(y + 1.0f) - lgammaf(arg - y + 1.0f) +
              y * f - arg * softplusf(f);
      break;
    case RR_LIK_POISSON_EXP: {
      float g = expf(f);
      *df = y - g;
      if (want_ll) *ll = y * f - g - lgammaf(y + 1.0f);
      break;
    }
    default: {  // RR_LIK_POISSON_SOFTPLUS
      float g = fmaxf(softplusf(f), 1e-37f);
      *df = expitf(f) * (y / g - 1.0f);
      if (want_ll) *ll = y * logf(g) - g - lgammaf(y + 1.0f);
      break;
    }
  }
}

__device__ __forceinline__ float lik_Ey(int lik, float f, float arg) {
  switch (lik) {
    case RR_LIK_GAUSSIAN: return f;
    case RR_LIK_BERNOULLI: return expitf(f);
    case RR_LIK_BINOMIAL: return expitf(f) * arg;
    case RR_LIK_POISSON_EXP: return expf(f);
    default: return softplusf(f);
  }
}

// ws[(k,l), j] = m[j,k] + sqrt(C[j,k]) * eps[k,l,j]      (glm.py:300-302)
__global__ void __launch_bounds__(256)
draw_weights_kernel(const float* __restrict__ mq, const float* __restrict__ Cq,
                    const float* __restrict__ eps, int D, int Kmix, int L,
                    float* __restrict__ Ws) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)Kmix * L * D;
  if (idx >= total) return;
  int j = (int)(idx % D);
  int k = (int)(idx / ((int64_t)L * D));
  Ws[idx] = mq[(int64_t)j * Kmix + k] + sqrtf(Cq[(int64_t)j * Kmix + k]) * eps[idx];
}

// In place F -> dF = df(y, F); per-mixture sums of loglike and dp.
// grid = (Kmix, row blocks); block reduces then one float64 atomic per output.
constexpr int LIK_ROWS = 32;
__global__ void __launch_bounds__(256)
lik_kernel(float* __restrict__ F, int rows, int S, int L,
           const float* __restrict__ y, const float* __restrict__ larg, int lik,
           float par, double* __restrict__ Ell, double* __restrict__ dlpar) {
  const int k = blockIdx.x;
  const int r0 = blockIdx.y * LIK_ROWS;
  const int nr = min(LIK_ROWS, rows - r0);
  const bool want_ll = Ell != nullptr;
  float sll = 0.0f, sdp = 0.0f;
  for (int e = threadIdx.x; e < nr * L; e += blockDim.x) {
    int r = e / L, l = e - r * L;
    int64_t off = (int64_t)(r0 + r) * S + (int64_t)k * L + l;
    float df, dp, ll;
    lik_eval(lik, y[r0 + r], F[off], par, larg ? larg[r0 + r] : 0.0f, want_ll,
             &df, &dp, &ll);
    F[off] = df;
    sll += ll;
    sdp += dp;
  }
  __shared__ float red[2][8];
  sll = warp_sum(sll);
  sdp = warp_sum(sdp);
  int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[0][w] = sll; red[1][w] = sdp; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += red[0][i]; b += red[1][i]; }
    if (want_ll) atomicAdd(Ell + k, a / (double)L);
    if (dlpar && lik == RR_LIK_GAUSSIAN) atomicAdd(dlpar, b / (double)L);
  }
}

// Edm[j,k] = mean_l Edws[(k,l),j];  EdC[j,k] = mean_l Edws*eps / sqrt(C[j,k])
// (glm.py:307-309).
__global__ void __launch_bounds__(256)
reduce_draws_kernel(const float* __restrict__ Edws, const float* __restrict__ eps,
                    const float* __restrict__ Cq, int D, int Kmix, int L,
                    float* __restrict__ Edm, float* __restrict__ EdC) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int k = blockIdx.y;
  if (j >= D) return;
  float a = 0.0f, b = 0.0f;
  for (int l = 0; l < L; ++l) {
    int64_t off = ((int64_t)k * L + l) * D + j;
    float v = Edws[off];
    a += v;
    b = fmaf(v, eps[off], b);
  }
  float invL = 1.0f / (float)L;
  Edm[(int64_t)j * Kmix + k] = a * invL;
  EdC[(int64_t)j * Kmix + k] = b * invL * rsqrtf(Cq[(int64_t)j * Kmix + k]);
}

// Q[r,k] = -Phi_sin * T[:,col_cos] + Phi_cos * T[:,col_sin] with T := EdPhi.
__global__ void __launch_bounds__(256)
q_plain_kernel(rr_plan plan, const float* __restrict__ Phi,
               const float* __restrict__ T, int64_t ld, int rows,
               float* __restrict__ Q) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  int r = blockIdx.y;
  if (k >= plan.ktot || r >= rows) return;
  int cc = plan.col_cos[k], cs = plan.col_sin[k];
  float tc = T[(int64_t)r * ld + cc], ts = T[(int64_t)r * ld + cs];
  float pc = Phi[(int64_t)r * ld + cc], ps = Phi[(int64_t)r * ld + cs];
  Q[(int64_t)r * plan.ktot + k] = -ps * tc + pc * ts;
}

// Warp per row: mean over draws of Ey(f) and Ey(f)^2.
__global__ void __launch_bounds__(256)
predict_reduce_kernel(const float* __restrict__ F, int rows, int S, int lik,
                      const float* __restrict__ larg, float* __restrict__ Ey,
                      float* __restrict__ Ey2) {
  int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float arg = larg ? larg[r] : 0.0f;
  float a = 0.0f, b = 0.0f;
  for (int s = lane; s < S; s += 32) {
    float e = lik_Ey(lik, F[(int64_t)r * S + s], arg);
    a += e;
    b = fmaf(e, e, b);
  }
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) {
    Ey[r] = a / (float)S;
    if (Ey2) Ey2[r] = b / (float)S;
  }
}

// ---- predictive CDF and quantiles (glm.py:468-570, 669-694) ---------------------
// Discrete CDFs sum the pmf outward from the side of floor(q) that holds the
// smaller tail, in float64, starting from a log-space pmf value: O(sqrt(mean))
// terms, no under/overflow for any rate.

// P(Y <= q), Y ~ Poisson(mu)                               (scipy.stats.poisson.cdf)
__device__ double poisson_cdf(double q, double mu) {
  if (q < 0.0) return 0.0;
  if (!(mu > 0.0)) return 1.0;
  const double kq = floor(q);
  const double span = 12.0 * sqrt(mu) + 40.0;
  if (kq >= mu + span) return 1.0;
  if (kq < mu - span) return 0.0;
  if (kq >= floor(mu)) {       // upper tail is the small side: 1 - sum_{k > kq} pmf
    double k = kq + 1.0;
    double p = exp(k * log(mu) - mu - lgamma(k + 1.0));
    double tail = 0.0;
    for (int it = 0; it < 100000 && p > 1e-18 * (tail + 1e-300); ++it) {
      tail += p;
      k += 1.0;
      p *= mu / k;
    }
    return fmax(0.0, 1.0 - tail);
  }
  double k = kq;
  double p = exp(k * log(mu) - mu - lgamma(k + 1.0));
  double sum = 0.0;
  for (int it = 0; it < 100000 && k >= 0.0 && p > 1e-18 * (sum + 1e-300); ++it) {
    sum += p;
    p *= k / mu;
    k -= 1.0;
  }
  return fmin(1.0, sum);
}

// P(Y <= q), Y ~ Binomial(n, pr)                             (scipy.stats.binom.cdf)
__device__ double binom_cdf(double q, double n, double pr) {
  if (q < 0.0) return 0.0;
  const double kq = floor(q);
  if (kq >= n) return 1.0;
  if (pr <= 0.0) return 1.0;
  if (pr >= 1.0) return 0.0;
  const double lp = log(pr), lq = log1p(-pr), odds = pr / (1.0 - pr);
  const double mode = floor((n + 1.0) * pr);
  auto lpmf = [&](double k) {
    return lgamma(n + 1.0) - lgamma(k + 1.0) - lgamma(n - k + 1.0) + k * lp + (n - k) * lq;
  };
  if (kq >= mode) {
    double k = kq + 1.0;
    double p = exp(lpmf(k));
    double tail = 0.0;
    for (int it = 0; it < 100000 && k <= n && p > 1e-18 * (tail + 1e-300); ++it) {
      tail += p;
      p *= (n - k) / (k + 1.0) * odds;
      k += 1.0;
    }
    return fmax(0.0, 1.0 - tail);
  }
  double k = kq;
  double p = exp(lpmf(k));
  double sum = 0.0;
  for (int it = 0; it < 100000 && k >= 0.0 && p > 1e-18 * (sum + 1e-300); ++it) {
    sum += p;
    p *= k / ((n - k + 1.0) * odds);
    k -= 1.0;
  }
  return fmin(1.0, sum);
}

// likelihoods.py: cdf of Bernoulli :129-146, Binomial :235-254, Gaussian :398-419,
// Poisson :523-541.
__device__ double lik_cdf(int lik, double q, float f, float par, float arg) {
  switch (lik) {
    case RR_LIK_GAUSSIAN:
      return normcdf((q - (double)f) / sqrt((double)par));
    case RR_LIK_BERNOULLI: {
      const double pr = 1.0 / (1.0 + exp(-(double)f));
      return q < 0.0 ? 0.0 : (q < 1.0 ? 1.0 - pr : 1.0);
    }
    case RR_LIK_BINOMIAL:
      return binom_cdf(q, (double)arg, 1.0 / (1.0 + exp(-(double)f)));
    case RR_LIK_POISSON_EXP:
      return poisson_cdf(q, exp((double)f));
    default: {
      const double ff = (double)f;
      return poisson_cdf(q, fmax(ff, 0.0) + log1p(exp(-fabs(ff))));
    }
  }
}

// Warp per row: mean / min / max over the draws of cdf(quantile | f).
__global__ void __launch_bounds__(256)
glm_cdf_kernel(const float* __restrict__ F, int rows, int S, int lik, float par,
               const float* __restrict__ larg, double quantile, float* __restrict__ pm,
               float* __restrict__ pmin, float* __restrict__ pmax) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float arg = larg ? larg[r] : 0.0f;
  double a = 0.0, lo = 2.0, hi = -1.0;
  for (int s = lane; s < S; s += 32) {
    const double c = lik_cdf(lik, quantile, F[(int64_t)r * S + s], par, arg);
    a += c;
    lo = fmin(lo, c);
    hi = fmax(hi, c);
  }
  a = warp_sum(a);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) {
    pm[r] = (float)(a / S);
    if (pmin) pmin[r] = (float)lo;
    if (pmax) pmax[r] = (float)hi;
  }
}

// Warp per row: both ends of the central `percentile` interval of the Monte-Carlo
// predictive distribution, i.e. the roots of  mean_s cdf(q | f_s) - p  for
// p = (1 -+ percentile) / 2 inside the reference's bracket +-1000 max(E[y], 1)
// (glm.py:681-683), by bisection (the function is monotone in q; for discrete
// likelihoods it is a step function and the root is the jump that brentq also
// converges to).  NaN where the bracket holds no sign change, as the reference.
__global__ void __launch_bounds__(256)
glm_quantile_kernel(const float* __restrict__ F, int rows, int S, int lik, float par,
                    const float* __restrict__ larg, double lo_p, double hi_p,
                    double* __restrict__ ql, double* __restrict__ qu) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float arg = larg ? larg[r] : 0.0f;
  const float* fr = F + (int64_t)r * S;
  double ey = 0.0;
  for (int s = lane; s < S; s += 32) ey += (double)lik_Ey(lik, fr[s], arg);
  ey = warp_sum(ey) / S;
  const double bound = 1000.0 * fmax(ey, 1.0);
  auto gap = [&](double q, double pct) {
    double a = 0.0;
    for (int s = lane; s < S; s += 32) a += lik_cdf(lik, q, fr[s], par, arg);
    return warp_sum(a) / S - pct;
  };
  for (int side = 0; side < 2; ++side) {
    const double pct = side ? hi_p : lo_p;
    double a = -bound, b = bound;
    const double ga = gap(a, pct), gb = gap(b, pct);
    double root = nan("");
    if (ga == 0.0) root = a;
    else if (gb == 0.0) root = b;
    else if ((ga < 0.0) != (gb < 0.0)) {
      for (int it = 0; it < 200 && (b - a) > 4e-12 + 8.9e-16 * fabs(b); ++it) {   // brentq's xtol, 4 rtol
        const double mid = 0.5 * (a + b);
        const double gm = gap(mid, pct);
        if ((gm < 0.0) == (ga < 0.0)) a = mid;
        else b = mid;
      }
      root = 0.5 * (a + b);
    }
    if (lane == 0) (side ? qu : ql)[r] = root;
  }
}

// scratch for the tile-major tf32 operand images of the tensor-core GEMMs
// (rr_tc_gemm3.cu): the largest A-side and B-side operand of the pass
static size_t glm_img_a(int op, int64_t R, int D, int S) {
  size_t a = gemm3_image_bytes(R, D);                       // F = Phi Ws^T
  if (op == RR_OP_GLM_STEP) {
    const size_t b = gemm3_image_bytes(S, R), c = gemm3_image_bytes(R, S);
    a = a > b ? a : b;                                      // Edws = dF^T Phi
    a = a > c ? a : c;                                      // EdPhi = dF Ws
  }
  return align_up(a, 1024);
}
static size_t glm_img_b(int op, int64_t R, int D, int S) {
  size_t a = gemm3_image_bytes(S, D);
  if (op == RR_OP_GLM_STEP) {
    const size_t b = gemm3_image_bytes(D, R), c = gemm3_image_bytes(D, S);
    a = a > b ? a : b;
    a = a > c ? a : c;
  }
  return align_up(a, 1024);
}
// R (d x kt, float64) += C (d x kt, float32)
__global__ void add_f32_to_f64_kernel(const float* __restrict__ Cf, double* __restrict__ R,
                                      int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) R[i] += (double)Cf[i];
}

// R += X^T Q (the lengthscale-gradient contraction): on the tensor cores as a
// (d x kt x rows) product, d padded to one 256-row tile, when the scratch is there
static int glm_xtq(const float* X, const float* Q, int rows, int d, int kt, double* R,
                   float* Ctmp, uint8_t* imgA, uint8_t* imgB, cudaStream_t st) {
  if (Ctmp && imgA && imgB && gemm3_worthwhile(256, kt, rows)) {
    int rc = gemm3(d, kt, rows, 1.0f, X, 1, d, Q, kt, 1, Ctmp, kt, 0, imgA, imgB, st);
    if (rc) return rc;
    const int64_t n = (int64_t)d * kt;
    add_f32_to_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Ctmp, R, n);
    RR_LAUNCH_CHECK("add_f32_to_f64_kernel");
    return RR_OK;
  }
  return xtq(X, Q, rows, d, kt, R, st);
}

// fp32-output product: tensor cores when the product is large enough to pay for
// packing its operands, CUDA cores otherwise
static int glm_gemm(int M, int N, int K, float alpha, const float* A, int64_t sAm, int64_t sAk,
                    const float* B, int64_t sBk, int64_t sBn, float* C, int64_t ldc,
                    int accumulate, uint8_t* imgA, uint8_t* imgB, cudaStream_t st) {
  if (imgA && imgB && gemm3_worthwhile(M, N, K))
    return gemm3(M, N, K, alpha, A, sAm, sAk, B, sBk, sBn, C, ldc, accumulate, imgA, imgB, st);
  return sgemm(M, N, K, alpha, A, sAm, sAk, B, sBk, sBn, C, nullptr, ldc, accumulate, st);
}

size_t glm_workspace_bytes(int op, int64_t M, const rr_plan* pl, int S) {
  int64_t R = M < GLM_CHUNK ? M : GLM_CHUNK;
  if (R < 1) R = 1;
  size_t phi = align_up((size_t)R * pl->D * 4, 256);
  size_t f = align_up((size_t)R * S * 4, 256);
  size_t ws = align_up((size_t)S * pl->D * 4, 256);
  size_t q = align_up((size_t)R * (pl->ktot > 0 ? pl->ktot : 1) * 4, 256);
  size_t img = 0;
  if (gemm3_worthwhile((int)R, S, pl->D))
    img = glm_img_a(op, R, pl->D, S) + glm_img_b(op, R, pl->D, S) + 2048 +
          align_up((size_t)pl->d * (pl->ktot > 0 ? pl->ktot : 1) * 4, 256);
  if (op == RR_OP_GLM_STEP) return 2 * phi + f + 2 * ws + q + img + 2048;
  return phi + f + img + 1024;
}

}  // namespace rr

using namespace rr;

extern "C" int rr_glm_step(const rr_plan* plan, const float* X, const float* y,
                           const float* larg, int64_t M, const float* mq,
                           const float* Cq, int32_t Kmix, const float* eps,
                           int32_t L, int32_t lik, float lik_param, float* Edm,
                           float* EdC, double* Rout, double* Ell, double* dlpar,
                           void* workspace, size_t workspace_bytes,
                           void* stream) {
  RR_REQUIRE(plan && X && y && mq && Cq && eps && Edm && EdC, "null pointer");
  RR_REQUIRE(Kmix > 0 && L > 0 && M > 0, "empty problem");
  RR_REQUIRE(lik >= 0 && lik <= RR_LIK_POISSON_SOFTPLUS, "unknown likelihood");
  RR_REQUIRE(lik != RR_LIK_BINOMIAL || larg, "Binomial needs its n argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int D = plan->D, d = plan->d, kt = plan->ktot, S = Kmix * L;
  const int64_t R = M < GLM_CHUNK ? M : GLM_CHUNK;
  Workspace W(workspace, workspace_bytes);
  float* Phi = W.take<float>((size_t)R * D);
  float* T = W.take<float>((size_t)R * D);
  float* F = W.take<float>((size_t)R * S);
  float* Ws = W.take<float>((size_t)S * D);
  float* Edws = W.take<float>((size_t)S * D);
  float* Q = W.take<float>((size_t)R * (kt > 0 ? kt : 1));
  if (!Phi || !T || !F || !Ws || !Edws || !Q) {
    set_error("glm_step workspace too small");
    return RR_ERR_WORKSPACE;
  }
  // operand images of the tensor-core GEMMs (absent in a caller's smaller, older
  // workspace: the CUDA-core kernel then does the products)
  uint8_t* imgA = nullptr;
  uint8_t* imgB = nullptr;
  if (gemm3_worthwhile((int)R, S, D)) {
    imgA = W.take<uint8_t>(glm_img_a(RR_OP_GLM_STEP, R, D, S));
    imgB = W.take<uint8_t>(glm_img_b(RR_OP_GLM_STEP, R, D, S));
  }
  // (d x kt) float staging of X^T Q; its operand images (256 x rows, kt x rows) fit in
  // the ones above (kt < D, 256 <= S or D)
  float* Ctmp = (imgA && imgB && kt > 0 && S >= 256) ? W.take<float>((size_t)d * kt) : nullptr;
  {
    int64_t total = (int64_t)S * D;
    draw_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        mq, Cq, eps, D, Kmix, L, Ws);
    RR_LAUNCH_CHECK("draw_weights_kernel");
  }
  for (int64_t s = 0; s < M; s += R) {
    int rows = (int)((M - s) < R ? (M - s) : R);
    int rc = launch_features(plan, X + s * d, rows, Phi, D, st);
    if (rc) return rc;
    // F = Phi Ws^T  (glm.py:303)
    rc = glm_gemm(rows, S, D, 1.0f, Phi, D, 1, Ws, 1, D, F, S, 0, imgA, imgB, st);
    if (rc) return rc;
    dim3 lg(Kmix, (rows + LIK_ROWS - 1) / LIK_ROWS);
    lik_kernel<<<lg, 256, 0, st>>>(F, rows, S, L, y + s, larg ? larg + s : nullptr,
                                   lik, lik_param, Ell, dlpar);
    RR_LAUNCH_CHECK("lik_kernel");
    // Edws (+)= dF^T Phi  (glm.py:307)
    rc = glm_gemm(S, D, rows, 1.0f, F, 1, S, Phi, D, 1, Edws, D, s > 0, imgA, imgB, st);
    if (rc) return rc;
    if (Rout && kt > 0) {
      // EdPhi = dF Ws / (L Kmix)  (glm.py:310, :246)
      rc = glm_gemm(rows, D, S, 1.0f / (float)(L * Kmix), F, S, 1, Ws, D, 1, T, D, 0, imgA,
                    imgB, st);
      if (rc) return rc;
      dim3 qg((kt + 255) / 256, rows);
      q_plain_kernel<<<qg, 256, 0, st>>>(*plan, Phi, T, D, rows, Q);
      RR_LAUNCH_CHECK("q_plain_kernel");
      rc = glm_xtq(X + s * d, Q, rows, d, kt, Rout, Ctmp, imgA, imgB, st);
      if (rc) return rc;
    }
  }
  dim3 rg((D + 255) / 256, Kmix);
  reduce_draws_kernel<<<rg, 256, 0, st>>>(Edws, eps, Cq, D, Kmix, L, Edm, EdC);
  RR_LAUNCH_CHECK("reduce_draws_kernel");
  return RR_OK;
}

extern "C" int rr_glm_predict(const rr_plan* plan, const float* X, int64_t N,
                              const float* ws, int32_t S, int32_t lik,
                              float lik_param, const float* larg, float* Ey,
                              float* Ey2, void* workspace,
                              size_t workspace_bytes, void* stream) {
  RR_REQUIRE(plan && X && ws && Ey, "null pointer");
  RR_REQUIRE(S > 0, "no draws");
  (void)lik_param;
  if (N == 0) return RR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int D = plan->D;
  const int64_t R = N < GLM_CHUNK ? N : GLM_CHUNK;
  Workspace W(workspace, workspace_bytes);
  float* Phi = W.take<float>((size_t)R * D);
  float* F = W.take<float>((size_t)R * S);
  if (!Phi || !F) { set_error("glm_predict workspace too small"); return RR_ERR_WORKSPACE; }
  uint8_t* imgA = nullptr;
  uint8_t* imgB = nullptr;
  if (gemm3_worthwhile((int)R, S, D)) {
    imgA = W.take<uint8_t>(glm_img_a(RR_OP_GLM_PREDICT, R, D, S));
    imgB = W.take<uint8_t>(glm_img_b(RR_OP_GLM_PREDICT, R, D, S));
  }
  for (int64_t s = 0; s < N; s += R) {
    int rows = (int)((N - s) < R ? (N - s) : R);
    int rc = launch_features(plan, X + s * plan->d, rows, Phi, D, st);
    if (rc) return rc;
    rc = glm_gemm(rows, S, D, 1.0f, Phi, D, 1, ws, 1, D, F, S, 0, imgA, imgB, st);
    if (rc) return rc;
    predict_reduce_kernel<<<(rows + 7) / 8, 256, 0, st>>>(
        F, rows, S, lik, larg ? larg + s : nullptr, Ey + s, Ey2 ? Ey2 + s : nullptr);
    RR_LAUNCH_CHECK("predict_reduce_kernel");
  }
  return RR_OK;
}

extern "C" int rr_glm_cdf(const float* F, int64_t N, int32_t S, int32_t lik, float lik_param,
                          const float* larg, double quantile, float* p_mean, float* p_min,
                          float* p_max, void* stream) {
  RR_REQUIRE(F && p_mean, "null pointer");
  RR_REQUIRE(S > 0, "no draws");
  RR_REQUIRE(lik >= 0 && lik <= RR_LIK_POISSON_SOFTPLUS, "unknown likelihood");
  RR_REQUIRE(lik != RR_LIK_BINOMIAL || larg, "Binomial needs its n argument");
  if (N == 0) return RR_OK;
  glm_cdf_kernel<<<(unsigned)((N + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      F, (int)N, S, lik, lik_param, larg, quantile, p_mean, p_min, p_max);
  RR_LAUNCH_CHECK("glm_cdf_kernel");
  return RR_OK;
}

extern "C" int rr_glm_quantiles(const float* F, int64_t N, int32_t S, int32_t lik,
                                float lik_param, const float* larg, double lo_p, double hi_p,
                                double* ql, double* qu, void* stream) {
  RR_REQUIRE(F && ql && qu, "null pointer");
  RR_REQUIRE(S > 0, "no draws");
  RR_REQUIRE(lik >= 0 && lik <= RR_LIK_POISSON_SOFTPLUS, "unknown likelihood");
  RR_REQUIRE(lik != RR_LIK_BINOMIAL || larg, "Binomial needs its n argument");
  if (N == 0) return RR_OK;
  glm_quantile_kernel<<<(unsigned)((N + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      F, (int)N, S, lik, lik_param, larg, lo_p, hi_p, ql, qu);
  RR_LAUNCH_CHECK("glm_quantile_kernel");
  return RR_OK;
}
