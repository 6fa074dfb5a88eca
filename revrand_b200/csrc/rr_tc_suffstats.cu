// Fused value pass of the SLM log marginal likelihood on tcgen05:
//   G += Phi^T Phi,  p += Phi^T y   with  Phi = amp * [cos | sin](2 pi X Wt)
// Phi never exists in HBM: generator warps project a 64-row slab of X through
// the CTA's frequency columns, apply sin/cos in registers, split each value
// into fp16 hi + lo and store it straight into 128B-swizzled K-major operand
// tiles in shared memory; one thread issues the tcgen05 MMAs
//   D += A_hi B_hi^T + A_lo B_hi^T + A_hi B_lo^T        (fp32 in TMEM)
// which reproduces fp32-quality products from fp16 tensor-core inputs.
//
// Output-stationary tiling over the *internal* feature order (blocks of 64
// frequencies = [64 cos | 64 sin] features): a work item is a (128 x 256)
// tile of the upper triangle of G times a row range; items are distributed
// round-robin over persistent CTAs (one per SM).  At the end of an item the
// accumulator is scaled by amp_i amp_j and added (float64 atomics) to G at
// the caller's column positions, mirrored across the diagonal.
//
// Replaces: revrand/slm.py:145-146, :157 and basis_functions.py:859-864.
#include "rr_common.cuh"
#include "rr_tc.cuh"

namespace rr {

using namespace tc;

constexpr int TC_ROWS = 64;            // rows of X per pipeline stage
constexpr int TC_STAGES = 2;
constexpr int TC_TI = 128;             // features per I block (64 freqs)
constexpr int TC_TJ = 256;             // features per J block (128 freqs)
constexpr int TC_GEN = 192;            // generator threads (64 A + 128 B freqs)
constexpr int TC_THREADS = 32 + TC_GEN;
constexpr int TC_A_BYTES = TC_TI * 128;   // one fp16 tile, 64 rows deep
constexpr int TC_B_BYTES = TC_TJ * 128;
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;
constexpr float TWO_PI = 6.283185307179586f;

struct TcMaps {       // per item parity: output column / amplitude tables
  int icol[TC_TI];
  float iamp[TC_TI];
  int jcol[TC_TJ];
  float jamp[TC_TJ];
};

template <int DP>
struct TcSmem {
  // operand tiles first (1024-byte aligned)
  uint8_t tiles[TC_STAGES][TC_STAGE_BYTES];
  float xs[TC_STAGES][TC_ROWS * DP];
  float ys[TC_STAGES][TC_ROWS];
  TcMaps maps[2];
  uint64_t full[TC_STAGES];
  uint64_t empty[TC_STAGES];
  uint64_t acc_full;
  uint64_t acc_empty;
  uint32_t tmem_base;
};

__device__ __forceinline__ void decode_tile(int t, int NI, int* ib, int* jb) {
  int j = 0;
  for (;;) {
    int cnt = min(2 * j + 2, NI);
    if (t < cnt) break;
    t -= cnt;
    ++j;
  }
  *ib = t;
  *jb = j;
}

template <int DP>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_suffstats_kernel(rr_plan plan, const float* __restrict__ X,
                    const float* __restrict__ y, int64_t N, double* __restrict__ G,
                    double* __restrict__ p, int NI, int ntiles, int nsplit,
                    int64_t rows_per_item) {
  extern __shared__ uint8_t smem_raw[];
  TcSmem<DP>& sm = *reinterpret_cast<TcSmem<DP>*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int d = plan.d, ktot = plan.ktot, D = plan.D;
  const int total_items = ntiles * nsplit;

  // ---- one-time setup ------------------------------------------------------
  for (int i = tid; i < TC_STAGES * TC_ROWS * DP; i += TC_THREADS)
    (&sm.xs[0][0])[i] = 0.0f;
  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&sm.full[s], TC_GEN);
      mbar_init(&sm.empty[s], 1);
    }
    mbar_init(&sm.acc_full, 1);
    mbar_init(&sm.acc_empty, 128);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&sm.tmem_base, TC_TJ);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0) {
    // ======================= MMA issuer ====================================
    const uint32_t idesc = make_idesc_f16(TC_TI, TC_TJ);
    uint32_t gs = 0, it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      const int tile = item % ntiles, split = item / ntiles;
      int ib, jb;
      decode_tile(tile, NI, &ib, &jb);
      const bool diag = (ib >> 1) == jb;
      const int64_t r0 = (int64_t)split * rows_per_item;
      const int64_t r1 = min(N, r0 + rows_per_item);
      const int nst = (int)((r1 - r0 + TC_ROWS - 1) / TC_ROWS);
      mbar_wait(&sm.acc_empty, (it & 1) ^ 1);
      tc_fence_after_sync();
      for (int t = 0; t < nst; ++t, ++gs) {
        const int s = gs % TC_STAGES;
        mbar_wait(&sm.full[s], (gs / TC_STAGES) & 1);
        tc_fence_after_sync();
        if (lane == 0) {
          const uint32_t base = smem_u32(&sm.tiles[s][0]);
          const uint32_t b_hi = base + 2 * TC_A_BYTES;
          const uint32_t b_lo = b_hi + TC_B_BYTES;
          // diagonal items take A from inside the B tile
          const uint32_t a_hi = diag ? b_hi + (ib & 1) * TC_A_BYTES : base;
          const uint32_t a_lo = diag ? b_lo + (ib & 1) * TC_A_BYTES : base + TC_A_BYTES;
          const uint64_t dah = make_desc_sw128(a_hi), dal = make_desc_sw128(a_lo);
          const uint64_t dbh = make_desc_sw128(b_hi), dbl = make_desc_sw128(b_lo);
#pragma unroll
          for (int k = 0; k < TC_ROWS / 16; ++k) {
            const uint64_t adv = (uint64_t)(k * 2);  // 32 bytes >> 4
            umma_f16_ss(tmem, dah + adv, dbh + adv, idesc, (t | k) != 0);
            umma_f16_ss(tmem, dal + adv, dbh + adv, idesc, 1);
            umma_f16_ss(tmem, dah + adv, dbl + adv, idesc, 1);
          }
          umma_commit(&sm.empty[s]);
          if (t == nst - 1) umma_commit(&sm.acc_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ======================= generators + epilogue ==========================
    const int g = tid - 32;  // 0..191
    const bool is_a = g < 64;
    uint32_t gs = 0, it = 0;
    float w[DP];
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      const int tile = item % ntiles, split = item / ntiles;
      int ib, jb;
      decode_tile(tile, NI, &ib, &jb);
      const bool diag = (ib >> 1) == jb;
      const int64_t r0 = (int64_t)split * rows_per_item;
      const int64_t r1 = min(N, r0 + rows_per_item);
      const int nst = (int)((r1 - r0 + TC_ROWS - 1) / TC_ROWS);

      // frequency owned by this thread for this item
      const int b = g - 64;
      const int theta = is_a ? 64 * ib + g : 128 * jb + b;
      const bool active = is_a ? !diag : true;
      const bool valid = theta < ktot;
      // rows of the operand tile this thread fills
      const uint32_t row_cos = is_a ? g : (uint32_t)(128 * (b >> 6) + (b & 63));
      const uint32_t row_sin = row_cos + 64;
#pragma unroll
      for (int i = 0; i < DP; ++i)
        w[i] = (valid && i < d) ? __ldg(plan.Wt + (int64_t)i * ktot + theta) : 0.0f;
      // Phi^T y is produced once per frequency block: by the B generators of
      // the diagonal item whose I block contains this frequency.
      const bool do_p = (p != nullptr) && diag && !is_a && valid && ((b >> 6) == (ib & 1));
      double pc_d = 0.0, ps_d = 0.0;

      // column/amp tables for the epilogue of this item
      TcMaps& mp = sm.maps[it & 1];
      if (g < TC_TI) {
        int th = 64 * ib + (g & 63);
        bool ok = th < ktot;
        mp.icol[g] = ok ? ((g >> 6) ? plan.col_sin[th] : plan.col_cos[th]) : -1;
        mp.iamp[g] = ok ? plan.amp[th] : 0.0f;
      }
      for (int j = g; j < TC_TJ; j += TC_GEN) {
        int th = 128 * jb + 64 * (j >> 7) + (j & 63);
        bool ok = th < ktot;
        mp.jcol[j] = ok ? (((j >> 6) & 1) ? plan.col_sin[th] : plan.col_cos[th]) : -1;
        mp.jamp[j] = ok ? plan.amp[th] : 0.0f;
      }

      for (int t = 0; t < nst; ++t, ++gs) {
        const int s = gs % TC_STAGES;
        const int64_t row0 = r0 + (int64_t)t * TC_ROWS;
        const int vrows = (int)min((int64_t)TC_ROWS, r1 - row0);
        mbar_wait(&sm.empty[s], ((gs / TC_STAGES) & 1) ^ 1);
        // stage the X slab (coalesced) and y
        {
          const float* src = X + row0 * d;
          const int cnt = vrows * d;
          for (int e = g; e < TC_ROWS * d; e += TC_GEN) {
            int r = e / d, i = e - r * d;
            sm.xs[s][r * DP + i] = (e < cnt) ? __ldg(src + e) : 0.0f;
          }
          if (g < TC_ROWS) sm.ys[s][g] = (y != nullptr && g < vrows) ? __ldg(y + row0 + g) : 0.0f;
        }
        named_bar_sync(1, TC_GEN);
        if (active) {
          const uint32_t base = smem_u32(&sm.tiles[s][0]);
          const uint32_t t_hi = is_a ? base : base + 2 * TC_A_BYTES;
          const uint32_t t_lo = is_a ? base + TC_A_BYTES : base + 2 * TC_A_BYTES + TC_B_BYTES;
          float pc = 0.0f, ps = 0.0f;
#pragma unroll 1
          for (int rg = 0; rg < TC_ROWS / 8; ++rg) {
            float c8[8], s8[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              const int row = rg * 8 + r;
              const float4* xr = reinterpret_cast<const float4*>(&sm.xs[s][row * DP]);
              float u = 0.0f;
#pragma unroll
              for (int q = 0; q < DP / 4; ++q) {
                float4 xv = xr[q];
                u = fmaf(xv.x, w[4 * q + 0], u);
                u = fmaf(xv.y, w[4 * q + 1], u);
                u = fmaf(xv.z, w[4 * q + 2], u);
                u = fmaf(xv.w, w[4 * q + 3], u);
              }
              const float fr = (u - rintf(u)) * TWO_PI;
              const float live = row < vrows ? 1.0f : 0.0f;
              c8[r] = __cosf(fr) * live;
              s8[r] = __sinf(fr) * live;
              if (do_p) {
                const float yv = sm.ys[s][row];
                pc = fmaf(c8[r], yv, pc);
                ps = fmaf(s8[r], yv, ps);
              }
            }
            uint4 hi, lo;
            split8(c8, &hi, &lo);
            st_shared_v4(t_hi + sw128_off(row_cos, rg), hi);
            st_shared_v4(t_lo + sw128_off(row_cos, rg), lo);
            split8(s8, &hi, &lo);
            st_shared_v4(t_hi + sw128_off(row_sin, rg), hi);
            st_shared_v4(t_lo + sw128_off(row_sin, rg), lo);
          }
          pc_d += (double)pc;
          ps_d += (double)ps;
        }
        fence_proxy_async_smem();
        mbar_arrive(&sm.full[s]);
      }

      if (do_p) {
        const float a = plan.amp[theta];
        atomicAdd(p + plan.col_cos[theta], (double)a * pc_d);
        atomicAdd(p + plan.col_sin[theta], (double)a * ps_d);
      }

      // ---- epilogue: warps 1..4 drain the accumulator ----------------------
      if (warp <= 4) {
        mbar_wait(&sm.acc_full, it & 1);
        tc_fence_after_sync();
        const int q = warp & 3;            // TMEM lane quarter of this warp
        const int il = 32 * q + lane;      // row of the tile
        const int fi = TC_TI * ib + il;    // internal feature index
        const int ci = mp.icol[il];
        const float ai = mp.iamp[il];
        for (int c0 = 0; c0 < TC_TJ; c0 += 32) {
          float v[32];
          tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)c0, v);
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            const int jl = c0 + r;
            const int fj = TC_TJ * jb + jl;
            const int cj = mp.jcol[jl];
            if (ci >= 0 && cj >= 0 && fi <= fj) {
              const double val = (double)(v[r] * ai * mp.jamp[jl]);
              atomicAdd(G + (int64_t)cj * D + ci, val);
              if (fi != fj) atomicAdd(G + (int64_t)ci * D + cj, val);
            }
          }
        }
        tc_fence_before_sync();
        mbar_arrive(&sm.acc_empty);
      }
    }
  }

  // ---- teardown --------------------------------------------------------------
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TC_TJ);
}

// ---------------------------------------------------------------------------
// Self-test: D = A B^T for a 128 x 256 x 64 fp16 problem written into the
// swizzled layout by threads exactly as the generators do, twice (second pass
// accumulates), read back through tcgen05.ld.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
tc_selftest_kernel(const __half* __restrict__ A, const __half* __restrict__ B,
                   float* __restrict__ Dout) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = base;                 // 128 rows x 128 B
  uint8_t* sb = base + TC_A_BYTES;    // 256 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  // fill tiles: thread handles rows tid (A) and tid, tid+128 (B)
  for (int c = 0; c < 8; ++c) {
    uint4 va = *reinterpret_cast<const uint4*>(A + tid * 64 + c * 8);
    st_shared_v4(smem_u32(sa) + sw128_off(tid, c), va);
    for (int h = 0; h < 2; ++h) {
      int row = tid + 128 * h;
      uint4 vb = *reinterpret_cast<const uint4*>(B + row * 64 + c * 8);
      st_shared_v4(smem_u32(sb) + sw128_off(row, c), vb);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, 256);
    const uint64_t da = make_desc_sw128(smem_u32(sa));
    const uint64_t db = make_desc_sw128(smem_u32(sb));
    for (int pass = 0; pass < 2; ++pass)
      for (int k = 0; k < 4; ++k)
        umma_f16_ss(tmem, da + 2 * k, db + 2 * k, idesc, (pass | k) != 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  for (int c0 = 0; c0 < 256; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, v);
    for (int r = 0; r < 32; ++r) Dout[(32 * warp + lane) * 256 + c0 + r] = v[r];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------
int tc_suffstats_supported(const rr_plan* pl) {
  return (pl->d >= 1 && pl->d <= 32 && pl->ktot >= 1 && pl->next == 0 &&
          pl->D == 2 * pl->ktot) ? 1 : 0;
}

size_t tc_suffstats_workspace(const rr_plan*, int64_t) { return 256; }

template <int DP>
static int launch_tc_suffstats(const rr_plan* pl, const float* X, const float* y,
                               int64_t N, double* G, double* p, cudaStream_t st) {
  const int NI = (pl->ktot + 63) / 64;
  const int NJ = (NI + 1) / 2;
  int ntiles = 0;
  for (int j = 0; j < NJ; ++j) ntiles += (2 * j + 2 < NI) ? 2 * j + 2 : NI;
  const int sms = sm_count();
  // Row ranges: enough items to balance the persistent CTAs, short enough to
  // bound the fp32 accumulation length, long enough to amortise the epilogue.
  int64_t want_items = (int64_t)sms * 6;
  int64_t nsplit = (want_items + ntiles - 1) / ntiles;
  int64_t rpi = (N + nsplit - 1) / nsplit;
  if (rpi < 2048) rpi = 2048;
  if (rpi > 32768) rpi = 32768;
  rpi = (rpi + TC_ROWS - 1) / TC_ROWS * TC_ROWS;
  nsplit = (N + rpi - 1) / rpi;
  int64_t items = nsplit * ntiles;
  int grid = (int)(items < sms ? items : sms);
  size_t smem = sizeof(TcSmem<DP>) + 1024;
  RR_CUDA_CHECK(cudaFuncSetAttribute(tc_suffstats_kernel<DP>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
  tc_suffstats_kernel<DP><<<grid, TC_THREADS, smem, st>>>(*pl, X, y, N, G, p, NI,
                                                        ntiles, (int)nsplit, rpi);
  RR_LAUNCH_CHECK("tc_suffstats_kernel");
  return RR_OK;
}

int tc_suffstats(const rr_plan* pl, const float* X, const float* y, int64_t N,
                 double* G, double* p, void*, size_t, cudaStream_t st) {
  const int d = pl->d;
  if (d <= 4) return launch_tc_suffstats<4>(pl, X, y, N, G, p, st);
  if (d <= 8) return launch_tc_suffstats<8>(pl, X, y, N, G, p, st);
  if (d <= 16) return launch_tc_suffstats<16>(pl, X, y, N, G, p, st);
  if (d <= 24) return launch_tc_suffstats<24>(pl, X, y, N, G, p, st);
  return launch_tc_suffstats<32>(pl, X, y, N, G, p, st);
}

}  // namespace rr

extern "C" int rr_tcgen05_selftest(double* max_abs_err) {
  using namespace rr;
  const int M = 128, Nn = 256, K = 64;
  __half *hA = (__half*)malloc(M * K * 2), *hB = (__half*)malloc(Nn * K * 2);
  float* hD = (float*)malloc(M * Nn * 4);
  uint32_t seed = 12345u;
  auto rnd = [&]() {
    seed = seed * 1664525u + 1013904223u;
    return ((seed >> 8) & 0xFFFF) / 65536.0f - 0.5f;
  };
  for (int i = 0; i < M * K; ++i) hA[i] = __float2half(rnd());
  for (int i = 0; i < Nn * K; ++i) hB[i] = __float2half(rnd());
  __half *dA = nullptr, *dB = nullptr;
  float* dD = nullptr;
  int rc = RR_OK;
  cudaError_t e;
#define ST_CHECK(x) do { e = (x); if (e != cudaSuccess) { set_error("selftest: %s: %s", #x, cudaGetErrorString(e)); rc = RR_ERR_CUDA; goto done; } } while (0)
  ST_CHECK(cudaMalloc(&dA, M * K * 2));
  ST_CHECK(cudaMalloc(&dB, Nn * K * 2));
  ST_CHECK(cudaMalloc(&dD, M * Nn * 4));
  ST_CHECK(cudaMemcpy(dA, hA, M * K * 2, cudaMemcpyHostToDevice));
  ST_CHECK(cudaMemcpy(dB, hB, Nn * K * 2, cudaMemcpyHostToDevice));
  ST_CHECK(cudaMemset(dD, 0, M * Nn * 4));
  {
    size_t smem = TC_A_BYTES + TC_B_BYTES + 1024;
    ST_CHECK(cudaFuncSetAttribute(tc_selftest_kernel,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    tc_selftest_kernel<<<1, 128, smem>>>(dA, dB, dD);
    ST_CHECK(cudaGetLastError());
    ST_CHECK(cudaDeviceSynchronize());
  }
  ST_CHECK(cudaMemcpy(hD, dD, M * Nn * 4, cudaMemcpyDeviceToHost));
  {
    double worst = 0.0;
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < Nn; ++j) {
        double ref = 0.0;
        for (int k = 0; k < K; ++k)
          ref += (double)__half2float(hA[i * K + k]) * (double)__half2float(hB[j * K + k]);
        ref *= 2.0;  // two accumulating passes
        double err = fabs(ref - (double)hD[i * Nn + j]);
        if (err > worst) worst = err;
      }
    if (max_abs_err) *max_abs_err = worst;
    if (worst > 1e-3) {
      set_error("tcgen05 selftest mismatch: max abs err %g", worst);
      rc = RR_ERR_CUDA;
    }
  }
done:
#undef ST_CHECK
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dD);
  free(hA);
  free(hB);
  free(hD);
  return rc;
}
