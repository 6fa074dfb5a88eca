// Standard linear model passes: C-ABI entry points, engine dispatch and the
// chunked SIMT engine (works for any feature plan; Phi is materialised one
// row chunk at a time in the caller's workspace).  The fused tcgen05 engine
// lives in rr_tc3_syrk.cu (value pass) / rr_tc_gradpass.cu.
//
// Reference: revrand/slm.py:142-199 (_elbo), :219-244 (predict_moments).
#include "rr_common.cuh"

namespace rr {

constexpr int64_t SIMT_CHUNK = 8192;  // rows of Phi held in the workspace

// Value pass: 1 = int8 fixed-point tensor-core engine (rr_tc3_syrk.cu), 2 = the
// round-1 fused kind::f16 kernel (rr_tc2_suffstats.cu, kept for A/B runs),
// 0 = SIMT, -1 = a tensor-core engine was demanded but cannot run this plan.
// Gradient pass: 1 = tcgen05 (rr_tc_gradpass.cu), 0 = SIMT, -1 as above.
static int pick_engine(int engine, const rr_plan* plan, int64_t N, bool grad = false) {
  engine &= RR_ENGINE_MASK;
  if (engine == RR_ENGINE_SIMT) return 0;
  if (grad) {
    const bool ok = tc_gradpass_supported(plan) != 0;
    if (engine != RR_ENGINE_AUTO) return ok ? 1 : -1;
    return (ok && N >= tc_auto_min_rows()) ? 1 : 0;
  }
  if (engine == RR_ENGINE_TCGEN05_FINE || engine == RR_ENGINE_TCGEN05_FUSED16)
    return tc_suffstats_supported(plan) ? 2 : -1;
  const bool ok = tc3_suffstats_supported(plan) != 0;
  if (engine == RR_ENGINE_TCGEN05) return ok ? 1 : -1;
  return (ok && N >= tc_auto_min_rows()) ? 1 : 0;
}

// p[j] += sum_r Phi[r,j] * y[r].  Thread per column, grid.y splits rows;
// float64 accumulation (the posterior mean inherits every bit lost here,
// amplified by the conditioning of the problem).
__global__ void __launch_bounds__(256)
colsum_weighted_kernel(const float* __restrict__ Phi, int64_t ld, int rows,
                       int D, const float* __restrict__ y,
                       double* __restrict__ p) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int per = (rows + gridDim.y - 1) / gridDim.y;
  int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  if (j >= D) return;
  double acc = 0.0;
  for (int r = r0; r < r1; ++r) acc = fma((double)Phi[(int64_t)r * ld + j], (double)y[r], acc);
  atomicAdd(p + j, acc);
}

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ y, int64_t n, double* __restrict__ out) {
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    double v = y[i];
    acc += v * v;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// T <- Err (x) m - T   restricted to what Q needs, then
// Q[r,k] = -Phi_sin[r,k] * T[r,col_cos k] + Phi_cos[r,k] * T[r,col_sin k].
// (amp and 1/sqrt(K) are already inside Phi.)
__global__ void __launch_bounds__(256)
q_kernel(rr_plan plan, const float* __restrict__ Phi, const float* __restrict__ T,
         int64_t ld, int rows, const float* __restrict__ err,
         const float* __restrict__ m, float* __restrict__ Q) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  int r = blockIdx.y;
  if (k >= plan.ktot || r >= rows) return;
  int cc = plan.col_cos[k], cs = plan.col_sin[k];
  float e = err ? err[r] : 0.0f;
  float tc = e * m[cc] - T[(int64_t)r * ld + cc];
  float ts = e * m[cs] - T[(int64_t)r * ld + cs];
  float pc = Phi[(int64_t)r * ld + cc], ps = Phi[(int64_t)r * ld + cs];
  Q[(int64_t)r * plan.ktot + k] = -ps * tc + pc * ts;
}

// Ey[r] = Phi[r,:] . m ; Vf[r] = sum_j T[r,j] * Phi[r,j].  Warp per row.
__global__ void __launch_bounds__(256)
rowdot_kernel(const float* __restrict__ Phi, const float* __restrict__ T,
              int64_t ld, int rows, int D, const float* __restrict__ m,
              float* __restrict__ Ey, float* __restrict__ Vf) {
  int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float a = 0.0f, b = 0.0f;
  for (int j = lane; j < D; j += 32) {
    float ph = Phi[(int64_t)r * ld + j];
    a = fmaf(ph, m[j], a);
    if (T) b = fmaf(ph, T[(int64_t)r * ld + j], b);
  }
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) {
    if (Ey) Ey[r] = a;
    if (Vf) Vf[r] = b;
  }
}

static size_t simt_ws(int op, int64_t N, const rr_plan* pl) {
  int64_t R = N < SIMT_CHUNK ? N : SIMT_CHUNK;
  if (R < 1) R = 1;
  size_t phi = align_up((size_t)R * pl->D * 4, 256);
  size_t q = align_up((size_t)R * (pl->ktot > 0 ? pl->ktot : 1) * 4, 256);
  switch (op) {
    case RR_OP_SUFFSTATS: return phi + 256;
    case RR_OP_GRADPASS: return 2 * phi + q + align_up((size_t)(N > 0 ? N : 1) * 4, 256) + 1024;
    case RR_OP_PREDICT: {
      // + the tf32 operand images of Phi (chunk) and C for the tensor-core product
      size_t img = 0;
      if (gemm3_worthwhile((int)R, pl->D, pl->D))
        img = align_up(gemm3_image_bytes(R, pl->D), 1024) +
              align_up(gemm3_image_bytes(pl->D, pl->D), 1024) + 2048;
      return 2 * phi + img + 1024;
    }
    case RR_OP_RESIDUAL: return align_up((size_t)(N > 0 ? N : 1) * 4, 256) + 512;
    default: return 0;
  }
}

static int simt_suffstats(const rr_plan* pl, const float* X, const float* y,
                          int64_t N, double* G, double* p, void* ws,
                          size_t wsb, cudaStream_t st) {
  Workspace W(ws, wsb);
  int64_t R = N < SIMT_CHUNK ? N : SIMT_CHUNK;
  float* Phi = W.take<float>((size_t)R * pl->D);
  if (!Phi) { set_error("suffstats workspace too small"); return RR_ERR_WORKSPACE; }
  const int D = pl->D;
  for (int64_t s = 0; s < N; s += R) {
    int rows = (int)((N - s) < R ? (N - s) : R);
    int rc = launch_features(pl, X + s * pl->d, rows, Phi, D, st);
    if (rc) return rc;
    // G += Phi^T Phi : A'(m,k) = Phi[k*D + m], B'(k,n) = Phi[k*D + n]
    rc = sgemm(D, D, rows, 1.0f, Phi, 1, D, Phi, D, 1, nullptr, G, D, 1, st);
    if (rc) return rc;
    if (p) {
      dim3 grid((D + 255) / 256, 32);
      colsum_weighted_kernel<<<grid, 256, 0, st>>>(Phi, D, rows, D, y + s, p);
      RR_LAUNCH_CHECK("colsum_weighted_kernel");
    }
  }
  return RR_OK;
}

static int simt_gradpass(const rr_plan* pl, const float* X, const float* y,
                         int64_t N, const float* m, const float* C, double* Rout,
                         double* sqerr, void* ws, size_t wsb, cudaStream_t st) {
  Workspace W(ws, wsb);
  int64_t R = N < SIMT_CHUNK ? N : SIMT_CHUNK;
  const int D = pl->D, d = pl->d, kt = pl->ktot;
  float* Phi = W.take<float>((size_t)R * D);
  float* T = W.take<float>((size_t)R * D);
  float* Q = W.take<float>((size_t)R * kt);
  float* err = W.take<float>((size_t)N);
  if (!Phi || !T || !Q || !err) { set_error("gradpass workspace too small"); return RR_ERR_WORKSPACE; }
  {
    int rc = phi_residual(pl, X, y, N, m, err, sqerr, err, st);
    if (rc) return rc;
  }
  for (int64_t s = 0; s < N; s += R) {
    int rows = (int)((N - s) < R ? (N - s) : R);
    int rc = launch_features(pl, X + s * d, rows, Phi, D, st);
    if (rc) return rc;
    rc = sgemm(rows, D, D, 1.0f, Phi, D, 1, C, D, 1, T, nullptr, D, 0, st);
    if (rc) return rc;
    dim3 grid((kt + 255) / 256, rows);
    q_kernel<<<grid, 256, 0, st>>>(*pl, Phi, T, D, rows, err + s, m, Q);
    RR_LAUNCH_CHECK("q_kernel");
    // R += X^T Q : A'(i,r) = X[r*d + i], B'(r,k) = Q[r*kt + k]
    rc = xtq(X + s * d, Q, rows, d, kt, Rout, st);
    if (rc) return rc;
  }
  return RR_OK;
}

}  // namespace rr

using namespace rr;

extern "C" int rr_slm_suffstats(const rr_plan* plan, const float* X,
                                const float* y, int64_t N, double* G, double* p,
                                double* yy, void* workspace,
                                size_t workspace_bytes, int32_t engine,
                                rr_context* ctx, void* stream) {
  RR_REQUIRE(plan && X && G, "null pointer");
  RR_REQUIRE(N >= 0, "negative N");
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0) return RR_OK;
  if (yy && y) {
    sumsq_kernel<<<sm_count() * 4, 256, 0, st>>>(y, N, yy);
    RR_LAUNCH_CHECK("sumsq_kernel");
  }
  const int use_tc = pick_engine(engine, plan, N);
  if (use_tc < 0) {
    set_error("tcgen05 engine does not support this plan");
    return RR_ERR_UNSUPPORTED;
  }
  if (use_tc == 1)
    return tc3_suffstats(plan, X, y, N, G, y ? p : nullptr, workspace, workspace_bytes, ctx, st);
  if (use_tc == 2)
    return tc_suffstats(plan, X, y, N, G, p, workspace, workspace_bytes,
                        engine == RR_ENGINE_TCGEN05_FINE ? 7 : 5, st);
  if (plan->kind != nullptr) {
    set_error("a plan with pseudo-frequency slots (kind != NULL) is only understood by "
              "the fused kind::f16 engine; pass the plain plan to the other engines");
    return RR_ERR_UNSUPPORTED;
  }
  return simt_suffstats(plan, X, y, N, G, y ? p : nullptr, workspace,
                        workspace_bytes, st);
}

extern "C" size_t rr_slm_kept_features_bytes(const rr_plan* plan, int64_t N) {
  if (plan == nullptr || N < tc_auto_min_rows()) return 0;
  return kept_features_bytes(plan, N);
}

extern "C" int rr_slm_suffstats_keep(const rr_plan* plan, const float* X, const float* y,
                                     int64_t N, double* G, double* p, double* yy, void* kept,
                                     size_t kept_bytes, void* workspace,
                                     size_t workspace_bytes, rr_context* ctx, void* stream) {
  RR_REQUIRE(plan && X && G && kept, "null pointer");
  RR_REQUIRE(N > 0, "no rows");
  const size_t need = kept_features_bytes(plan, N);
  if (need == 0 || plan->kind != nullptr) {
    set_error("kept features need a plan both tensor-core passes support");
    return RR_ERR_UNSUPPORTED;
  }
  RR_REQUIRE(kept_bytes >= need, "kept feature buffer too small (rr_slm_kept_features_bytes)");
  cudaStream_t st = (cudaStream_t)stream;
  if (yy && y) {
    sumsq_kernel<<<sm_count() * 4, 256, 0, st>>>(y, N, yy);
    RR_LAUNCH_CHECK("sumsq_kernel");
  }
  return tc3_suffstats(plan, X, y, N, G, y ? p : nullptr, workspace, workspace_bytes, ctx, st, kept);
}

extern "C" int rr_slm_gradpass_kept(const rr_plan* plan, const float* X, const float* y,
                                    int64_t N, const float* m, const float* C, double* R,
                                    double* sqerr, const void* kept, size_t kept_bytes,
                                    void* workspace, size_t workspace_bytes, int32_t flags,
                                    void* stream) {
  RR_REQUIRE(plan && X && y && m && C && R && sqerr && kept, "null pointer");
  RR_REQUIRE(N > 0, "no rows");
  const size_t need = kept_features_bytes(plan, N);
  if (need == 0 || plan->ktot == 0) {
    set_error("kept features need a plan both tensor-core passes support");
    return RR_ERR_UNSUPPORTED;
  }
  RR_REQUIRE(kept_bytes >= need, "kept feature buffer too small (rr_slm_kept_features_bytes)");
  return tc_gradpass_kept(plan, X, y, N, m, C, R, sqerr, kept, workspace, workspace_bytes,
                          (flags & RR_GRAD_SPLIT_C) != 0, (cudaStream_t)stream);
}

extern "C" int rr_slm_residual(const rr_plan* plan, const float* X,
                               const float* y, int64_t N, const float* m,
                               float* err, double* sqerr, void* workspace,
                               size_t workspace_bytes, void* stream) {
  RR_REQUIRE(plan && X && y && m && sqerr, "null pointer");
  if (N == 0) return RR_OK;
  Workspace W(workspace, workspace_bytes);
  float* fbuf = err ? err : W.take<float>((size_t)N);
  if (!fbuf) { set_error("residual workspace too small"); return RR_ERR_WORKSPACE; }
  return phi_residual(plan, X, y, N, m, err, sqerr, fbuf, (cudaStream_t)stream);
}

extern "C" int rr_slm_gradpass(const rr_plan* plan, const float* X,
                               const float* y, int64_t N, const float* m,
                               const float* C, double* R, double* sqerr,
                               void* workspace, size_t workspace_bytes,
                               int32_t engine, rr_context* ctx, void* stream) {
  RR_REQUIRE(plan && X && y && m && C && R && sqerr, "null pointer");
  if (N == 0) return RR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (plan->ktot == 0) {
    Workspace W(workspace, workspace_bytes);
    float* fbuf = W.take<float>((size_t)N);
    if (!fbuf) { set_error("gradpass workspace too small"); return RR_ERR_WORKSPACE; }
    return phi_residual(plan, X, y, N, m, nullptr, sqerr, fbuf, st);
  }
  const int use_tc = pick_engine(engine, plan, N, true);
  if (use_tc < 0) {
    set_error("tcgen05 engine does not support this plan");
    return RR_ERR_UNSUPPORTED;
  }
  if (use_tc)
    return tc_gradpass(plan, X, y, N, m, C, R, sqerr, workspace, workspace_bytes, ctx,
                       (engine & RR_GRAD_SPLIT_C) != 0, st);
  return simt_gradpass(plan, X, y, N, m, C, R, sqerr, workspace, workspace_bytes, st);
}

extern "C" int rr_slm_predict(const rr_plan* plan, const float* X, int64_t N,
                              const float* m, const float* C, float* Ey,
                              float* Vf, void* workspace, size_t workspace_bytes,
                              void* stream) {
  RR_REQUIRE(plan && X && m, "null pointer");
  RR_REQUIRE(!Vf || C, "Vf requested without C");
  if (N == 0) return RR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  Workspace W(workspace, workspace_bytes);
  int64_t R = N < SIMT_CHUNK ? N : SIMT_CHUNK;
  const int D = plan->D;
  float* Phi = W.take<float>((size_t)R * D);
  float* T = Vf ? W.take<float>((size_t)R * D) : nullptr;
  if (!Phi || (Vf && !T)) { set_error("predict workspace too small"); return RR_ERR_WORKSPACE; }
  // Phi C on the tcgen05 tensor cores (tf32, three-product split: fp32 grade) when the
  // product is large enough to pay for packing its operands
  uint8_t* imgA = nullptr;
  uint8_t* imgB = nullptr;
  if (Vf && gemm3_worthwhile((int)R, D, D)) {
    imgA = W.take<uint8_t>(align_up(gemm3_image_bytes(R, D), 1024));
    imgB = W.take<uint8_t>(align_up(gemm3_image_bytes(D, D), 1024));
  }
  for (int64_t s = 0; s < N; s += R) {
    int rows = (int)((N - s) < R ? (N - s) : R);
    int rc = launch_features(plan, X + s * plan->d, rows, Phi, D, st);
    if (rc) return rc;
    if (Vf) {
      if (imgA && imgB)
        rc = gemm3(rows, D, D, 1.0f, Phi, D, 1, C, D, 1, T, D, 0, imgA, imgB, st);
      else
        rc = sgemm(rows, D, D, 1.0f, Phi, D, 1, C, D, 1, T, nullptr, D, 0, st);
      if (rc) return rc;
    }
    rowdot_kernel<<<(rows + 7) / 8, 256, 0, st>>>(Phi, T, D, rows, D, m,
                                                   Ey ? Ey + s : nullptr,
                                                   Vf ? Vf + s : nullptr);
    RR_LAUNCH_CHECK("rowdot_kernel");
  }
  return RR_OK;
}

namespace rr {
size_t slm_workspace_bytes(int op, int64_t N, const rr_plan* pl, int engine) {
  const bool split_c = (engine & RR_GRAD_SPLIT_C) != 0;
  if (op == RR_OP_GRADPASS_KEPT) return tc_gradpass_kept_workspace(pl, N, split_c);
  size_t s = simt_ws(op, N, pl);
  if (op != RR_OP_PREDICT && op != RR_OP_RESIDUAL) {
    const int e = pick_engine(engine, pl, N, op == RR_OP_GRADPASS);
    if (e > 0) {
      size_t t = op == RR_OP_GRADPASS ? tc_gradpass_workspace(pl, N, split_c)
                 : (e == 1 ? tc3_suffstats_workspace(pl, N) : tc_suffstats_workspace(pl, N));
      return t > 256 ? t : 256;
    }
  }
  return s;
}
}  // namespace rr
