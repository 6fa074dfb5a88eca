/*
 * revrand_b200 C-ABI: the drop-in boundary for the random-feature
 * log-marginal-likelihood hot path.
 *
 * The reference (NICTA/revrand @ 4c1881b) has no FFI; its boundary for this
 * path is the Python protocol between slm.py / glm.py and
 * basis_functions.py.  Every entry point below replaces one NumPy call site
 * of that protocol (cited per function, paths relative to the reference
 * checkout) and is what a binding on the reference side would load with
 * ctypes.
 *
 * Conventions
 *  - All data pointers are DEVICE pointers owned by the caller; the library
 *    allocates nothing persistent and keeps no global state besides a
 *    thread-local error string and a launch counter.  Matrices are row-major.
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *    Calls are asynchronous with respect to the host.
 *  - Two passes can overlap their feature generation with their tensor-core
 *    work on a helper stream.  That stream and its events belong to an
 *    `rr_context` the CALLER creates and destroys (one per host thread and
 *    device); passing ctx == NULL runs everything on `stream`.  Every call
 *    joins the helper stream back into `stream` before it returns, on error
 *    paths too, so a context is capturable in a CUDA graph with the call.
 *  - Return value: 0 on success, negative rr_status otherwise;
 *    rr_last_error() returns a thread-local message.  No C++ exceptions
 *    cross this boundary.
 *  - Accumulating outputs (G, p, R, ...) are float64 and are ADDED to, so
 *    row-sharded callers (one rank per GPU) can allreduce them directly.
 */
#ifndef REVRAND_B200_H
#define REVRAND_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum rr_status {
  RR_OK = 0,
  RR_ERR_INVALID = -1,   /* bad argument / shape */
  RR_ERR_WORKSPACE = -2, /* workspace too small */
  RR_ERR_CUDA = -3,      /* CUDA runtime error (see rr_last_error) */
  RR_ERR_UNSUPPORTED = -4
} rr_status;

/*
 * Feature plan: a concatenation of random trigonometric blocks and "extra"
 * affine columns, i.e. what BasisCat.transform (basis_functions.py:1599-1627)
 * evaluates for RandomRBF/Laplace/Cauchy/Matern32/Matern52/OrthogonalRBF/
 * FastFoodRBF blocks (basis_functions.py:838-864, 1263-1289) concatenated
 * with LinearBasis / BiasBasis columns (:468-485, :415-432).
 *
 * Frequency k of the plan produces two feature columns
 *   Phi[n, col_cos[k]] = amp[k] * cos(2*pi*u),  Phi[n, col_sin[k]] = amp[k] * sin(2*pi*u),
 *   u = sum_i X[n,i] * Wt[i*ktot + k]            (Wt = W / lenscale / (2*pi): "turns")
 * and extra column j is  X[n, ext_src[j]]  (ext_src[j] >= 0)  or the constant
 * ext_val[j] (ext_src[j] < 0), written to column ext_col[j].
 */
typedef struct rr_plan {
  int32_t d;            /* input dimension of X                              */
  int32_t ktot;         /* number of random frequencies over all trig blocks */
  int32_t next;         /* number of extra (non-trigonometric) columns       */
  int32_t D;            /* total number of feature columns                   */
  const float* Wt;      /* (d, ktot) projection in turns                     */
  const float* amp;     /* (ktot) amplitude, 1/sqrt(K_block)                 */
  const int32_t* col_cos; /* (ktot)                                          */
  const int32_t* col_sin; /* (ktot)                                          */
  const int32_t* ext_src; /* (next)                                          */
  const float* ext_val;   /* (next)                                          */
  const int32_t* ext_col; /* (next)                                          */
  /* Optional (NULL = all trigonometric): per-frequency slot kind for the fused
   * tcgen05 value pass, which can carry affine columns as pseudo-frequencies:
   *   0: cos/sin pair of 2*pi*u (as above)
   *   1: linear slot -- the single feature amp[k] * u, written to col_cos[k]
   *      (u = sum_i X[n,i] Wt[i,k]: Wt selects and scales one input column so
   *      that |u| <= 1; amp undoes the scale); col_sin[k] must be -1
   *   2: constant slot -- the single feature amp[k] in col_cos[k]; col_sin[k] = -1
   * A plan with kind != NULL must have next == 0 (the affine columns ARE slots). */
  const uint8_t* kind;  /* (ktot) or NULL                                    */
  /* Optional (NULL = all 1): integer power of extra column j, i.e. the column is
   * X[n, ext_src[j]] ^ ext_pow[j] -- PolynomialBasis (basis_functions.py:537-566). */
  const int32_t* ext_pow; /* (next) or NULL                                   */
  /* Optional (NULL = taken from the rows of the call): the fixed-point scales of the
   * int8 value pass, col_scale[i] >= max |X[:, i]| (i < d) and col_scale[d] >= max |y|
   * over ALL rows of the job.  A row-sharded caller passes the job-wide maxima so
   * that every rank quantises alike (the ranks' statistics then add up to the
   * one-process result bit for bit); it also saves the pass over X that finds them. */
  const float* col_scale; /* (d + 1) or NULL                                  */
} rr_plan;

/* Likelihood ids for rr_glm_step; revrand/likelihoods.py:18-545. */
typedef enum rr_likelihood {
  RR_LIK_GAUSSIAN = 0,
  RR_LIK_BERNOULLI = 1,
  RR_LIK_BINOMIAL = 2,
  RR_LIK_POISSON_EXP = 3,
  RR_LIK_POISSON_SOFTPLUS = 4
} rr_likelihood;

/* Engine selection flags for the SLM passes. */
#define RR_ENGINE_AUTO 0   /* tensor-core path when the plan allows it and the
                              job has at least rr_engine_auto_min_rows() rows */
#define RR_ENGINE_SIMT 1   /* chunked CUDA-core path (any plan)          */
#define RR_ENGINE_TCGEN05 2 /* tensor-core path, error if unsupported: value pass in
                               exact 24-bit fixed point on tcgen05 kind::i8, gradient
                               pass on kind::f16 */
#define RR_ENGINE_TCGEN05_FINE 3    /* round-1 fused kind::f16 value pass, 2^-7 grid */
#define RR_ENGINE_TCGEN05_FUSED16 4 /* round-1 fused kind::f16 value pass, 2^-5 grid
                                       (kept for A/B measurements) */

#define RR_ENGINE_MASK 0xff
/* Flag for the gradient pass, OR-ed into `engine` (rr_slm_gradpass, rr_workspace_bytes)
 * or passed as `flags` (rr_slm_gradpass_kept).  The tensor-core gradient pass multiplies
 * Phi by an fp16 image of C; its rounding (2^-12 per entry) is harmless where the
 * quadratic form Phi C dPhi does not cancel (config 2: 1.6e-5 on the gradients), but on
 * a strongly correlated feature set (64 frequencies on 3-D inputs, cond(C) 3e5) it cost
 * 1.5e-2.  With this flag a second GEMM over the rounding residual restores C to ~22
 * bits, at twice the tensor-core work of the pass. */
#define RR_GRAD_SPLIT_C 0x100

/* Helper stream + events for the passes that overlap generation with tensor-core
 * work (see Conventions).  Created on the current device. */
typedef struct rr_context rr_context;
int rr_context_create(rr_context** out);
int rr_context_destroy(rr_context* ctx);

/* Row count from which RR_ENGINE_AUTO uses the tensor-core engine. */
int64_t rr_engine_auto_min_rows(void);

int rr_version(void);
const char* rr_last_error(void);
/* Number of kernels this library has launched in the process so far. */
uint64_t rr_launch_count(void);
/* SM count and compute capability of the current device. */
int rr_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/*
 * Phi = transform(X): replaces np.dot/np.cos/np.sin/np.hstack at
 * basis_functions.py:862-864 (and BasisCat.transform :1622-1627).
 * Phi is (N, ldphi) with ldphi >= plan->D.
 */
int rr_features(const rr_plan* plan, const float* X, int64_t N, float* Phi,
                int64_t ldphi, void* stream);

/*
 * dPhi = grad(X) of ONE trig block wrt its lengthscale(s): replaces
 * basis_functions.py:888-901.  W is the raw (d,K) frequency matrix,
 * lenscale has n_ls entries (1 or d).  Output layout is the reference's:
 * (N, 2K) for n_ls == 1, else (N, 2K, d) with the lengthscale index fastest.
 * compat != 0 reproduces the reference's scalar-lengthscale behaviour (only
 * input dimension 0 contributes, :896-899); compat == 0 gives the full
 * derivative for a shared scalar lengthscale.
 */
int rr_trig_grad(const float* X, int64_t N, int32_t d, const float* W,
                 int32_t K, const float* lenscale, int32_t n_ls,
                 int32_t compat, float* dPhi, void* stream);

/*
 * FastFood projection  VX = hstack_b( H(PI_b(H(X~ * B_b)) * G_b) * S_b * sqrt(d2) )
 * followed by the trig map: replaces FastFoodRBF._makeVX + transform,
 * basis_functions.py:1356-1371, 1285-1289, with mathfun/linalg.py:182-220
 * (hadamard, ordering=False) done as an in-register butterfly.
 * Xs is X / lenscale, (N, d) with d <= d2; B,G,S are (k, d2) float, PI (k, d2)
 * int32.  Phi is (N, 2*k*d2): [cos | sin] / sqrt(k*d2).  If VX_out != NULL
 * the raw projection (N, k*d2) is also written.
 */
int rr_fastfood_features(const float* Xs, int64_t N, int32_t d, int32_t d2,
                         int32_t k, const float* B, const float* G,
                         const int32_t* PI, const float* S, float* Phi,
                         float* VX_out, void* stream);

/*
 * Centre-based bases (basis_functions.py:616-815).  C is (M, d) centres, lenscale
 * has n_ls entries (1 or d).
 *   kind 0, RadialBasis   :665-688   Phi = exp(-sum_i ((x_i - c_i) / (2 l_i^2))^2)
 *                         :690-722   dPhi_i = Phi (x_i - c_i)^2 / l_i^6
 *   kind 1, SigmoidalBasis:770-790   Phi = expit(sqrt(sum_i ((x_i - c_i) / l_i)^2))
 *                         :792-815   dPhi_i = -|x_i - c_i| / l_i^2 Phi (1 - Phi)
 * Phi is (N, M); dPhi (optional) is (N, M) for n_ls == 1 -- where, as in the
 * reference, only input dimension 0 contributes -- else (N, M, d).
 */
int rr_centre_features(const float* X, int64_t N, int32_t d, const float* C,
                       int32_t M, const float* lenscale, int32_t n_ls,
                       int32_t kind, float* Phi, float* dPhi, void* stream);

/*
 * Gradients of one Gaussian spectral-mixture component (FastFoodGM.grad,
 * basis_functions.py:1474-1527) from the dense image V (d, n) of its FastFood
 * projection: phases p = x (V / l), q = x . mean; outputs (N, 4n, d) (or (N, 4n)
 * for d == 1) wrt the means and wrt the lengthscales, block order
 * [cos(p+q) | sin(p+q) | cos(p-q) | sin(p-q)] / sqrt(2n).
 */
int rr_gm_grad(const float* X, int64_t N, int32_t d, const float* V, int32_t n,
               const float* mean, const float* lenscale, float* dmean,
               float* dlen, void* stream);

/*
 * Value pass of StandardLinearModel._elbo: G += Phi^T Phi, p += Phi^T y,
 * yy += y^T y over this rank's rows; replaces slm.py:145-146 and the
 * Phi.T.dot(y) of :157.  G is (D,D) float64, p (D) float64, yy (1) float64.
 * The tensor-core engine rounds every feature value once to 24-bit fixed point
 * and forms G and p from those integers EXACTLY (int8 digits, int32
 * accumulators): the result is independent of tiling and row partitioning.
 */
int rr_slm_suffstats(const rr_plan* plan, const float* X, const float* y,
                     int64_t N, double* G, double* p, double* yy,
                     void* workspace, size_t workspace_bytes, int32_t engine,
                     rr_context* ctx, void* stream);

/*
 * Residual pass: sqerr += sum_n (y_n - phi_n^T m)^2 (slm.py:161-162);
 * if err != NULL also writes the N residuals (and no workspace is needed).
 */
int rr_slm_residual(const rr_plan* plan, const float* X, const float* y,
                    int64_t N, const float* m, float* err, double* sqerr,
                    void* workspace, size_t workspace_bytes, void* stream);

/*
 * Residual + gradient pass of StandardLinearModel._elbo wrt the basis
 * hyper-parameters (slm.py:161-162 and :193-197 with basis_functions.py:888-901
 * and apply_grad :109-152), restated so that dPhi is never formed:
 *   Err = y - Phi m,  sqerr += sum Err^2,
 *   T = Err (x) m - Phi C,  Q[n,k] = -Phi_sin[n,k] T[n,col_cos k] + Phi_cos[n,k] T[n,col_sin k],
 *   R += X^T Q     (d, ktot) float64.
 * The caller turns R into d(-ELBO)/d lenscale_i = sum_k W[i,k] R[i,k] / (var * l_i^2).
 * m (D) and C (D,D) are float32; the residuals are computed from the same
 * trigonometric values that feed the Phi C product (one pass over the rows).
 */
int rr_slm_gradpass(const rr_plan* plan, const float* X, const float* y,
                    int64_t N, const float* m, const float* C, double* R,
                    double* sqerr, void* workspace, size_t workspace_bytes,
                    int32_t engine, rr_context* ctx, void* stream);

/*
 * The two passes of one evaluation, sharing the feature map.  slm.py:145 forms
 * Phi = basis.transform(X) ONCE per _elbo call and uses it for Phi^T Phi (:146), the
 * residuals (:161) and the gradients (:193-197).  rr_slm_suffstats_keep is
 * rr_slm_suffstats on the tensor-core engine that also leaves Phi behind as an fp16
 * image in the caller's buffer `kept` (rr_slm_kept_features_bytes(plan, N) bytes,
 * 1024-byte aligned; 0 = this plan / row count cannot keep its features);
 * rr_slm_gradpass_kept is rr_slm_gradpass reading that image instead of evaluating
 * the feature map a second time.  The image is only meaningful for the SAME plan
 * contents (Wt), X and N as the call that wrote it.  Workspace of the second call:
 * rr_workspace_bytes(RR_OP_GRADPASS_KEPT, ...).  sqerr here comes from fp16 feature
 * values (relative accuracy ~1e-5): callers that need more take
 * sum Err^2 = y'y - 2 p'm + m'G m from the float64 statistics, or rr_slm_residual.
 */
size_t rr_slm_kept_features_bytes(const rr_plan* plan, int64_t N);
int rr_slm_suffstats_keep(const rr_plan* plan, const float* X, const float* y,
                          int64_t N, double* G, double* p, double* yy, void* kept,
                          size_t kept_bytes, void* workspace, size_t workspace_bytes,
                          rr_context* ctx, void* stream);
int rr_slm_gradpass_kept(const rr_plan* plan, const float* X, const float* y,
                         int64_t N, const float* m, const float* C, double* R,
                         double* sqerr, const void* kept, size_t kept_bytes,
                         void* workspace, size_t workspace_bytes, int32_t flags,
                         void* stream);

/*
 * Predictive moments (slm.py:239-242): Ey = Phi m, Vf = rowsum((Phi C) * Phi).
 */
int rr_slm_predict(const rr_plan* plan, const float* X, int64_t N,
                   const float* m, const float* C, float* Ey, float* Vf,
                   void* workspace, size_t workspace_bytes, void* stream);

/*
 * Data-dependent part of one GLM SVI step, all K_mix mixture components and
 * L reparameterised draws at once: replaces GeneralizedLinearModel._reparam_k
 * (glm.py:296-322) for k = 0..Kmix-1 and the EdPhi contraction of
 * glm.py:274-275.
 *   inputs : minibatch X (M,d), y (M), optional per-row likelihood argument
 *            larg (M) (Binomial n) or NULL; variational mean/var mq, Cq
 *            (D,Kmix) row-major; eps (Kmix,L,D) noise; lik / lik_param.
 *   outputs: Edm (D,Kmix), EdC (D,Kmix) as glm.py:306-309 (NOT yet scaled by
 *            B); R (d,ktot) float64 += X^T Q with T := EdPhi (mean over k);
 *            Ell (Kmix) float64 expected log-likelihood sums (glm.py:320);
 *            dlpar (1) float64 += sum over k of mean_l sum_n dp (glm.py:313-316).
 */
int rr_glm_step(const rr_plan* plan, const float* X, const float* y,
                const float* larg, int64_t M, const float* mq, const float* Cq,
                int32_t Kmix, const float* eps, int32_t L, int32_t lik,
                float lik_param, float* Edm, float* EdC, double* R,
                double* Ell, double* dlpar, void* workspace,
                size_t workspace_bytes, void* stream);

/*
 * GLM predictive sampling (glm.py:404-418, 572-620): f = Phi(Xs) ws^T for S
 * posterior weight draws ws (S,D); accumulates Ey (N) = mean_s Ey(f) and
 * optionally its second moment for the variance.
 */
int rr_glm_predict(const rr_plan* plan, const float* X, int64_t N,
                   const float* ws, int32_t S, int32_t lik, float lik_param,
                   const float* larg, float* Ey, float* Ey2, void* workspace,
                   size_t workspace_bytes, void* stream);

/*
 * Monte-Carlo predictive CDF (glm.py:468-516): for latent draws F (N, S) --
 * f = Phi(X*) w_s, as rr_glm_predict forms them -- mean / min / max over the draws
 * of likelihood.cdf(quantile, f) (likelihoods.py:129-146, 235-254, 398-419,
 * 523-541).  p_min / p_max may be NULL.
 */
int rr_glm_cdf(const float* F, int64_t N, int32_t S, int32_t lik, float lik_param,
               const float* larg, double quantile, float* p_mean, float* p_min,
               float* p_max, void* stream);

/*
 * Predictive interval (glm.py:518-570 with _rootfinding :669-694): per row the two
 * roots of  mean_s cdf(q | f_s) = lo_p, hi_p  inside +-1000 max(mean_s Ey(f_s), 1),
 * all rows bisected concurrently on the device instead of brentq in a process
 * pool.  NaN where the bracket holds no sign change (the reference's ValueError
 * branch).  ql, qu: (N) float64.
 */
int rr_glm_quantiles(const float* F, int64_t N, int32_t S, int32_t lik, float lik_param,
                     const float* larg, double lo_p, double hi_p, double* ql, double* qu,
                     void* stream);

/* Workspace query: op is one of the RR_OP_* codes. */
#define RR_OP_SUFFSTATS 1
#define RR_OP_GRADPASS 2
#define RR_OP_PREDICT 3
#define RR_OP_GLM_STEP 4
#define RR_OP_GLM_PREDICT 5
#define RR_OP_RESIDUAL 6
#define RR_OP_GRADPASS_KEPT 7
size_t rr_workspace_bytes(int32_t op, int64_t N, int32_t d, int32_t ktot,
                          int32_t D, int32_t aux0, int32_t aux1,
                          int32_t engine);

/* 1 if the fused tcgen05 engine can run this plan shape, else 0. */
int rr_tcgen05_supported(int32_t d, int32_t ktot, int32_t next, int32_t D);

/*
 * Self-test of the tcgen05 / TMEM building blocks (descriptor encodings,
 * swizzled shared-memory layout, TMEM load mapping) against a CUDA-core
 * reference on random data; returns 0 when bit-level layout checks pass.
 * max_abs_err (host pointer) receives the largest deviation.
 */
int rr_tcgen05_selftest(double* max_abs_err);

/*
 * Self-test of the kind::i8 path: one CTA pair (cta_group::2, M = 256,
 * N = 160), `kblocks` K blocks of 64 through the 64-byte-swizzled layout and
 * the six digit products of the value pass, compared BIT-EXACTLY with a host
 * integer reference.  mismatches (host pointer) receives the number of wrong
 * accumulator entries; returns 0 when it is zero.
 */
int rr_tcgen05_i8_selftest(int32_t kblocks, int64_t* mismatches);

/*
 * The fp32-grade tensor-core GEMM behind the GLM step and predictive sampler,
 * callable on its own (tests, diagnostics): C (+)= alpha * A B^T for row-major fp32
 * DEVICE matrices A (M x K, leading dimension lda), B (N x K, ldb), C (M x N, ldc);
 * transa / transb != 0: the operand is stored K x M / K x N instead.  Operands are
 * split into two tf32 parts, three tcgen05.mma.kind::tf32 products per k-step
 * (every product to ~2^-21), fp32 accumulation in TMEM.  Workspace: at least
 * 65536 * (ceil(M/256) + ceil(N/256)) * ceil(K/32) + 4096 bytes.
 */
int rr_tcgen05_gemm3(int32_t M, int32_t N, int32_t K, float alpha, const float* A,
                     int64_t lda, int32_t transa, const float* B, int64_t ldb,
                     int32_t transb, float* C, int64_t ldc, int32_t accumulate,
                     void* workspace, size_t workspace_bytes, void* stream);

/*
 * Diagnostic: repeat D += A B^T over one 128 x 256 x 64 fp16 tile `reps`
 * times in TMEM and return the fp32 accumulator (HOST pointers; A* are
 * (128,64), B* (256,64) fp16 bit patterns, D and Aux (128,256) float).
 * mode 0: hi*hi only; 1: hi*hi + lo*hi + hi*lo into one accumulator;
 * 2: the two cross terms into a second accumulator (Aux).  Used to calibrate
 * how many rows may be accumulated before a flush to float64.
 */
int rr_tcgen05_accum_probe(const uint16_t* Ahi, const uint16_t* Alo,
                           const uint16_t* Bhi, const uint16_t* Blo,
                           int32_t reps, int32_t mode, float* D, float* Aux);

#ifdef __cplusplus
}
#endif
#endif /* REVRAND_B200_H */
