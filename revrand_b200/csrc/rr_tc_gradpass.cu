// Gradient pass of the SLM log marginal likelihood on tcgen05.
//
//   T = Err (x) m - Phi C,   Q[n,k] = -Phi_sin[n,k] T[n,cos k] + Phi_cos[n,k] T[n,sin k],
//   R += X^T Q                                            (d x ktot, float64)
//
// dPhi (N x 2K x d in the reference, basis_functions.py:888-901) is never
// formed, and Phi only ever exists as an fp16 row chunk in a scratch buffer
// sized to stay L2-resident together with the fp16 image of C:
//   1. prep   : Bt[fo][fj] = s * amp_fj * C[col fo][col fj]  (fp16, internal
//               feature order = blocks of [64 cos | 64 sin]), s = 1/max|C|.
//   2. per row chunk: phi kernel writes trig values (fp16) for the chunk;
//      the GEMM kernel computes T' = Phi_chunk Bt^T on tcgen05 (cp.async ->
//      128B-swizzled smem -> UMMA, fp32 accumulators in TMEM, 256 rows x 256
//      output features per CTA) and its epilogue forms Q in registers, stages
//      it through shared memory and contracts it with X, adding X^T Q to R.
//
// Replaces: revrand/slm.py:193-197 + basis_functions.py:109-152, 888-901.
#include "rr_common.cuh"
#include "rr_tc.cuh"

namespace rr {

using namespace tc;

constexpr int GP_RM = 256;        // rows per CTA tile (two 128-lane accumulators)
constexpr int GP_RN = 256;        // output features per CTA tile
constexpr int GP_KT = 64;         // K extent per stage (one 128-byte line)
constexpr int GP_STAGES = 3;
constexpr int GP_TILE_BYTES = 256 * 128;
constexpr int GP_STAGE_BYTES = 2 * GP_TILE_BYTES;
constexpr int GP_THREADS = 32 + 128;
constexpr int64_t GP_SCRATCH_BYTES = 32ll << 20;  // Phi chunk budget (L2 resident)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- prep kernels -----------------------------------------------------------
__global__ void absmax_kernel(const float* __restrict__ C, int64_t n,
                              unsigned int* __restrict__ out_bits) {
  float mx = 0.0f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    mx = fmaxf(mx, fabsf(C[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(mx));
}

// internal feature f -> (frequency, is_sin)
__device__ __forceinline__ int feat_theta(int f) { return 64 * (f >> 7) + (f & 63); }
__device__ __forceinline__ int feat_is_sin(int f) { return (f >> 6) & 1; }

__global__ void __launch_bounds__(256)
prep_c_kernel(rr_plan plan, const float* __restrict__ C, int Dp,
              const unsigned int* __restrict__ cmax_bits, __half* __restrict__ Bt) {
  const int fj = blockIdx.x * blockDim.x + threadIdx.x;
  const int fo = blockIdx.y;
  if (fj >= Dp) return;
  const float cmax = __uint_as_float(*cmax_bits);
  const float s = cmax > 0.0f ? 1.0f / cmax : 0.0f;
  const int tj = feat_theta(fj), to = feat_theta(fo);
  float v = 0.0f;
  if (tj < plan.ktot && to < plan.ktot) {
    const int cj = feat_is_sin(fj) ? plan.col_sin[tj] : plan.col_cos[tj];
    const int co = feat_is_sin(fo) ? plan.col_sin[to] : plan.col_cos[to];
    v = s * plan.amp[tj] * C[(int64_t)co * plan.D + cj];
  }
  Bt[(int64_t)fo * Dp + fj] = __float2half_rn(v);
}

// Phi chunk (raw trig, no amplitude) in internal order, fp16, rows padded with
// zeros up to rows_pad.
constexpr int PH_ROWS = 16;
__global__ void __launch_bounds__(256)
phi_half_kernel(rr_plan plan, const float* __restrict__ X, int rows, int rows_pad,
                int Dp, __half* __restrict__ Ph) {
  extern __shared__ float xs[];
  const int d = plan.d;
  const int n0 = blockIdx.x * PH_ROWS;
  for (int t = threadIdx.x; t < PH_ROWS * d; t += blockDim.x) {
    int r = t / d;
    xs[t] = (n0 + r < rows) ? X[(int64_t)(n0 + r) * d + (t - r * d)] : 0.0f;
  }
  __syncthreads();
  const int nth = Dp / 2;  // padded frequency count
  for (int th = threadIdx.x; th < nth; th += blockDim.x) {
    float u[PH_ROWS];
#pragma unroll
    for (int r = 0; r < PH_ROWS; ++r) u[r] = 0.0f;
    const bool valid = th < plan.ktot;
    if (valid) {
      for (int i = 0; i < d; ++i) {
        float w = __ldg(plan.Wt + (int64_t)i * plan.ktot + th);
#pragma unroll
        for (int r = 0; r < PH_ROWS; ++r) u[r] = fmaf(xs[r * d + i], w, u[r]);
      }
    }
    const int fc = 128 * (th >> 6) + (th & 63);
#pragma unroll
    for (int r = 0; r < PH_ROWS; ++r) {
      if (n0 + r >= rows_pad) break;
      float s = 0.0f, c = 0.0f;
      if (valid && n0 + r < rows) {
        float fr = (u[r] - rintf(u[r])) * 6.283185307179586f;
        s = __sinf(fr);
        c = __cosf(fr);
      }
      Ph[(int64_t)(n0 + r) * Dp + fc] = __float2half_rn(c);
      Ph[(int64_t)(n0 + r) * Dp + fc + 64] = __float2half_rn(s);
    }
  }
}

// ---- GEMM + epilogue ---------------------------------------------------------
template <int DP>
__global__ void __launch_bounds__(GP_THREADS, 1)
tc_gradpass_kernel(rr_plan plan, const float* __restrict__ X,
                   const float* __restrict__ err, int rows,
                   const __half* __restrict__ Ph, const __half* __restrict__ Bt,
                   int Dp, const float* __restrict__ m,
                   const unsigned int* __restrict__ cmax_bits,
                   double* __restrict__ R) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[GP_STAGES], empty[GP_STAGES], acc_full;
  __shared__ uint32_t tmem_slot;
  __shared__ float m_loc[GP_RN];
  __shared__ float amp_loc[GP_RN / 2];
  __shared__ float err_loc[GP_RM];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ob = blockIdx.x;            // output feature block
  const int rb = blockIdx.y;            // row block
  const int row0 = rb * GP_RM;
  const int nk = Dp / GP_KT;
  const int d = plan.d, ktot = plan.ktot;

  if (tid == 0) {
    for (int s = 0; s < GP_STAGES; ++s) {
      mbar_init(&full[s], 128);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    // ---------------- MMA issuer ------------------------------------------------
    const uint32_t idesc = make_idesc_f16(128, GP_RN);
    for (int kt = 0; kt < nk; ++kt) {
      const int s = kt % GP_STAGES;
      mbar_wait(&full[s], (kt / GP_STAGES) & 1);
      tc_fence_after_sync();
      if (lane == 0) {
        const uint32_t a0 = smem_u32(tiles + s * GP_STAGE_BYTES);
        const uint32_t b0 = a0 + GP_TILE_BYTES;
        const uint64_t da0 = make_desc_sw128(a0);
        const uint64_t da1 = make_desc_sw128(a0 + 128 * 128);
        const uint64_t db = make_desc_sw128(b0);
#pragma unroll
        for (int k = 0; k < GP_KT / 16; ++k) {
          const uint64_t adv = (uint64_t)(2 * k);
          umma_f16_ss(tmem, da0 + adv, db + adv, idesc, (kt | k) != 0);
          umma_f16_ss(tmem + GP_RN, da1 + adv, db + adv, idesc, (kt | k) != 0);
        }
        umma_commit(&empty[s]);
        if (kt == nk - 1) umma_commit(&acc_full);
      }
      __syncwarp();
    }
  } else {
    // ---------------- loaders, then epilogue -----------------------------------
    const int lt = tid - 32;  // 0..127
    const __half* Abase = Ph + (int64_t)row0 * Dp;
    const __half* Bbase = Bt + (int64_t)ob * GP_RN * Dp;
    // tables for the epilogue
    for (int j = lt; j < GP_RN; j += 128) {
      int f = ob * GP_RN + j;
      int th = feat_theta(f);
      float mv = 0.0f;
      if (th < ktot) mv = m[feat_is_sin(f) ? plan.col_sin[th] : plan.col_cos[th]];
      m_loc[j] = mv;
    }
    {
      int th = ob * (GP_RN / 2) + lt;
      amp_loc[lt] = th < ktot ? plan.amp[th] : 0.0f;
    }
    for (int j = lt; j < GP_RM; j += 128)
      err_loc[j] = (row0 + j < rows) ? err[row0 + j] : 0.0f;

    auto issue_stage = [&](int kt) {
      const int s = kt % GP_STAGES;
      const uint32_t a0 = smem_u32(tiles + s * GP_STAGE_BYTES);
      const uint32_t b0 = a0 + GP_TILE_BYTES;
      const int k0 = kt * GP_KT;
#pragma unroll 4
      for (int j = 0; j < 16; ++j) {
        const int idx = lt + 128 * j;       // 0..2047
        const int row = idx >> 3, ch = idx & 7;
        cp_async16(a0 + sw128_off(row, ch), Abase + (int64_t)row * Dp + k0 + ch * 8);
        cp_async16(b0 + sw128_off(row, ch), Bbase + (int64_t)row * Dp + k0 + ch * 8);
      }
      cp_async_commit();
    };

    // software pipeline: keep GP_STAGES-1 stages of loads in flight
    for (int kt = 0; kt < nk + GP_STAGES - 1; ++kt) {
      if (kt < nk) {
        const int s = kt % GP_STAGES;
        mbar_wait(&empty[s], ((kt / GP_STAGES) & 1) ^ 1);
        issue_stage(kt);
      } else {
        cp_async_commit();  // empty group keeps the wait arithmetic uniform
      }
      const int done = kt - (GP_STAGES - 1);
      if (done >= 0) {
        cp_async_wait<GP_STAGES - 1>();
        fence_proxy_async_smem();
        mbar_arrive(&full[done % GP_STAGES]);
      }
    }

    // ---------------- epilogue ---------------------------------------------------
    mbar_wait(&acc_full, 0);
    tc_fence_after_sync();
    // all MMAs have completed: stage memory is free for Q and X staging
    float* Qs = reinterpret_cast<float*>(tiles);                       // [256][33]
    float* xs = reinterpret_cast<float*>(tiles + 2 * GP_STAGE_BYTES);  // [256][DP]
    for (int e = lt; e < GP_RM * DP; e += 128) {
      int r = e / DP, i = e - r * DP;
      xs[e] = (i < d && row0 + r < rows) ? X[(int64_t)(row0 + r) * d + i] : 0.0f;
    }
    const float cmax = __uint_as_float(*cmax_bits);
    const int q = warp & 3;
    constexpr int IG = DP / 4;  // input dims per reducing warp
    const int ew = warp - 1;    // 0..3 : which group of input dims this warp reduces
    for (int bb = 0; bb < 2; ++bb) {
      for (int c = 0; c < 2; ++c) {
        const int fcol = 128 * bb + 32 * c;  // first cos column of this chunk (tile-local)
        for (int h = 0; h < 2; ++h) {
          const int rl = 128 * h + 32 * q + lane;
          float tcv[32], tsv[32];
          const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(h * GP_RN + fcol);
          tmem_ld32(taddr, tcv);
          tmem_ld32(taddr + 64, tsv);
          const __half* prow = Ph + (int64_t)(row0 + rl) * Dp + ob * GP_RN + fcol;
          const float e_n = err_loc[rl];
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            uint4 pc = *reinterpret_cast<const uint4*>(prow + 8 * v4);
            uint4 ps = *reinterpret_cast<const uint4*>(prow + 64 + 8 * v4);
            const __half* hc = reinterpret_cast<const __half*>(&pc);
            const __half* hs = reinterpret_cast<const __half*>(&ps);
#pragma unroll
            for (int r8 = 0; r8 < 8; ++r8) {
              const int r = 8 * v4 + r8;
              const float Tc = e_n * m_loc[fcol + r] - tcv[r] * cmax;
              const float Ts = e_n * m_loc[fcol + 64 + r] - tsv[r] * cmax;
              const float a = amp_loc[64 * bb + 32 * c + r];
              Qs[rl * 33 + r] = a * (-__half2float(hs[r8]) * Tc + __half2float(hc[r8]) * Ts);
            }
          }
        }
        named_bar_sync(1, 128);
        // R[i, theta] += sum_rows x[row, i] * Q[row, theta]; lane = theta, warp = dim group
        float acc[IG];
#pragma unroll
        for (int j = 0; j < IG; ++j) acc[j] = 0.0f;
        for (int row = 0; row < GP_RM; ++row) {
          const float qv = Qs[row * 33 + lane];
#pragma unroll
          for (int j = 0; j < IG; ++j) acc[j] = fmaf(xs[row * DP + ew * IG + j], qv, acc[j]);
        }
        const int th = ob * (GP_RN / 2) + 64 * bb + 32 * c + lane;
        if (th < ktot) {
#pragma unroll
          for (int j = 0; j < IG; ++j) {
            const int i = ew * IG + j;
            if (i < d) atomicAdd(R + (int64_t)i * ktot + th, (double)acc[j]);
          }
        }
        named_bar_sync(1, 128);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- host ---------------------------------------------------------------------
// padded feature count: whole 256-feature output tiles (= 128 frequencies)
static int gp_dp(const rr_plan* pl) { return ((pl->ktot + 127) / 128) * 256; }

static int64_t gp_chunk_rows(const rr_plan* pl, int64_t N) {
  const int Dp = gp_dp(pl);
  int64_t rc = GP_SCRATCH_BYTES / ((int64_t)Dp * 2);
  rc = rc / GP_RM * GP_RM;
  if (rc < GP_RM) rc = GP_RM;
  int64_t npad = (N + GP_RM - 1) / GP_RM * GP_RM;
  return rc < npad ? rc : npad;
}

size_t tc_gradpass_workspace(const rr_plan* pl, int64_t N) {
  const int64_t Dp = gp_dp(pl);
  return align_up((size_t)gp_chunk_rows(pl, N) * Dp * 2, 256) +
         align_up((size_t)Dp * Dp * 2, 256) + 1024;
}

template <int DP>
static int launch_gp(const rr_plan* pl, const float* X, const float* err, int rows,
                     const __half* Ph, const __half* Bt, int Dp, const float* m,
                     const unsigned int* cmax, double* R, cudaStream_t st) {
  size_t smem = GP_STAGES * GP_STAGE_BYTES + 1024;
  RR_CUDA_CHECK(cudaFuncSetAttribute(tc_gradpass_kernel<DP>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
  dim3 grid(Dp / GP_RN, (rows + GP_RM - 1) / GP_RM);
  tc_gradpass_kernel<DP><<<grid, GP_THREADS, smem, st>>>(*pl, X, err, rows, Ph, Bt,
                                                        Dp, m, cmax, R);
  RR_LAUNCH_CHECK("tc_gradpass_kernel");
  return RR_OK;
}

int tc_gradpass(const rr_plan* pl, const float* X, const float* err, int64_t N,
                const float* m, const float* C, double* R, void* ws, size_t wsb,
                cudaStream_t st) {
  const int Dp = gp_dp(pl);
  const int64_t RC = gp_chunk_rows(pl, N);
  Workspace W(ws, wsb);
  __half* Ph = W.take<__half>((size_t)RC * Dp);
  __half* Bt = W.take<__half>((size_t)Dp * Dp);
  unsigned int* cmax = W.take<unsigned int>(1);
  if (!Ph || !Bt || !cmax) {
    set_error("tcgen05 gradpass workspace too small");
    return RR_ERR_WORKSPACE;
  }
  RR_CUDA_CHECK(cudaMemsetAsync(cmax, 0, sizeof(unsigned int), st));
  absmax_kernel<<<sm_count() * 4, 256, 0, st>>>(C, (int64_t)pl->D * pl->D, cmax);
  RR_LAUNCH_CHECK("absmax_kernel");
  {
    dim3 grid((Dp + 255) / 256, Dp);
    prep_c_kernel<<<grid, 256, 0, st>>>(*pl, C, Dp, cmax, Bt);
    RR_LAUNCH_CHECK("prep_c_kernel");
  }
  const int d = pl->d;
  for (int64_t s = 0; s < N; s += RC) {
    const int rows = (int)((N - s) < RC ? (N - s) : RC);
    const int rows_pad = (rows + GP_RM - 1) / GP_RM * GP_RM;
    phi_half_kernel<<<(rows_pad + PH_ROWS - 1) / PH_ROWS, 256,
                      PH_ROWS * d * sizeof(float), st>>>(*pl, X + s * d, rows,
                                                         rows_pad, Dp, Ph);
    RR_LAUNCH_CHECK("phi_half_kernel");
    int rc;
    const float* Xc = X + s * d;
    const float* ec = err + s;
    if (d <= 4) rc = launch_gp<4>(pl, Xc, ec, rows, Ph, Bt, Dp, m, cmax, R, st);
    else if (d <= 8) rc = launch_gp<8>(pl, Xc, ec, rows, Ph, Bt, Dp, m, cmax, R, st);
    else if (d <= 16) rc = launch_gp<16>(pl, Xc, ec, rows, Ph, Bt, Dp, m, cmax, R, st);
    else if (d <= 24) rc = launch_gp<24>(pl, Xc, ec, rows, Ph, Bt, Dp, m, cmax, R, st);
    else rc = launch_gp<32>(pl, Xc, ec, rows, Ph, Bt, Dp, m, cmax, R, st);
    if (rc) return rc;
  }
  return RR_OK;
}

}  // namespace rr
