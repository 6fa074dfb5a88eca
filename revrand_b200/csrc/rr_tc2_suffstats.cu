// Fused value pass of the SLM log marginal likelihood on tcgen05, CTA-pair
// edition:  G += Phi^T Phi,  p += Phi^T y,  Phi = amp * [cos | sin](2 pi X Wt).
// Phi never exists in HBM.
//
// One cluster of two CTAs (cta_group::2) owns a 256 x 224 tile of G in the
// "internal" feature order and streams row slabs of X through two chained
// tensor-core products:
//
//   MMA#1 (kind::tf32, 3-way split => fp32-grade):  U = Wt_tile^T X_slab^T
//         256 frequencies (128 per CTA) x 64 rows, accumulators in TMEM.
//   generator warps: tcgen05.ld U, exact range reduction in turns, sin/cos,
//         and split every trigonometric value c into
//             h1 = c rounded to the 2^-5 grid   (6-bit fixed point)
//             r  = fp16(c - h1)                  (|r| <= 2^-6)
//             cf = fp16(c)
//         written straight into 128B-swizzled K-major fp16 operand tiles.
//   MMA#2 (kind::f16):   MAIN += H1_a H1_b^T                      (exact)
//                        AUX  += R_a CF_b^T + H1_a R_b^T          (tiny)
//
// Why this split: the tensor core adds fp16 products into its fp32 accumulator
// with truncation (measured with rr_tcgen05_accum_probe: aligned to the
// accumulator exponent with two guard bits, then rounded toward zero), which
// biases long all-positive sums such as the diagonal of G by ~1e-5 after a few
// thousand rows.  Products of 2^-5-grid values are multiples of 2^-10 and their
// running sum stays below 2^14 for 16384 rows, so every partial sum of MAIN is
// exactly representable and nothing is ever truncated; AUX only holds terms
// <= 2^-6 whose truncation is far below fp32 resolution of the result.  Every
// 16384 rows the two accumulators are drained, combined and added to
// a float64 scratch image of G.
//
// The tile set covers every unordered feature pair exactly once (rows = 128
// frequencies of block ib, columns = 112 frequencies of block jb, kept when
// theta_row < theta_col or equal with type_row <= type_col); a finalize kernel
// mirrors the scratch image into the caller's G.
//
// Replaces: revrand/slm.py:145-146, :157 and basis_functions.py:859-864.
#include "rr_common.cuh"
#include "rr_tc.cuh"

namespace rr {

using namespace tc;

constexpr int T2_SLAB = 64;          // rows of X per pipeline step
constexpr int T2_NA = 64;            // row-block frequencies per CTA
constexpr int T2_NB = 56;            // column-block frequencies per CTA
constexpr int T2_IB = 2 * T2_NA;     // frequencies per row block (pair)
constexpr int T2_JB = 2 * T2_NB;     // frequencies per column block (pair)
constexpr int T2_NCOL = 4 * T2_NB;   // 224 accumulator columns
constexpr int T2_TILE_GROUP = 74;    // tiles whose scratch images are hot together
constexpr int T2_XSTAGES = 3;
constexpr int T2_PSTAGES = 2;
// The fixed-point grid of h1 is a run-time choice: grid 2^-g, magic 1.5 * 2^(23-g)
// ((c + M) - M rounds c to the grid), exact chains of 2^(24-2g) rows.  g = 5 is the
// default (16384-row chains); g = 7 quarters the rounding error of the remainder r
// (|r| <= 2^-8) at the price of a drain every 1024 rows.
constexpr int T2_SUPER = 32768;      // rows per work item
constexpr float T2_RINT_MAGIC = 12582912.0f;  // 1.5 * 2^23
constexpr float T2_TWO_PI = 6.283185307179586f;

constexpr int T2_A_BYTES = 128 * 128;            // one fp16 image of the A tile
constexpr int T2_B_BYTES = 2 * T2_NB * 128;      // one fp16 image of the B half tile
// image order [A.h1 | B.h1 | A.r | B.r | B.cf]: the h1 -> r distance is the same
// for both tiles, so one generator code path addresses either with immediates
constexpr int T2_OFF_AH = 0;
constexpr int T2_OFF_BH = T2_A_BYTES;
constexpr int T2_OFF_AR = T2_OFF_BH + T2_B_BYTES;
constexpr int T2_OFF_BR = T2_OFF_AR + T2_A_BYTES;
constexpr int T2_OFF_BC = T2_OFF_BR + T2_B_BYTES;
constexpr int T2_PHI_BYTES = T2_OFF_BC + T2_B_BYTES;   // 75776
constexpr int T2_H2R = T2_OFF_AR - T2_OFF_AH;          // == T2_OFF_BR - T2_OFF_BH
constexpr int T2_H2C = T2_OFF_BC - T2_OFF_BH;
static_assert(T2_OFF_BR - T2_OFF_BH == T2_H2R, "image layout");
constexpr int T2_W_BYTES = 128 * 128;            // one tf32 image of the W tile
constexpr int T2_X_BYTES = 32 * 128;             // one tf32 image of the X half slab
constexpr int T2_SMEM_PHI = 0;
constexpr int T2_SMEM_W = T2_PSTAGES * T2_PHI_BYTES;
constexpr int T2_SMEM_X = T2_SMEM_W + 2 * T2_W_BYTES;
constexpr int T2_SMEM_RAW = T2_SMEM_X + T2_XSTAGES * 2 * T2_X_BYTES;
constexpr int T2_LOAD_WARPS = 2;                 // loader warps; warp li takes slabs = li (mod 2)
constexpr int T2_RAW_AHEAD = 1;                  // own slabs of X in flight per loader warp
constexpr int T2_RAW_STAGES = T2_LOAD_WARPS * (T2_RAW_AHEAD + 1);
constexpr int T2_RAW_BYTES = 32 * 32 * 4;        // 32 rows x d <= 32 floats, as in HBM
constexpr int T2_SMEM_BYTES = T2_SMEM_RAW + T2_RAW_STAGES * T2_RAW_BYTES;

constexpr int T2_TMEM_MAIN = 0;
constexpr int T2_TMEM_AUX = T2_NCOL;
constexpr int T2_TMEM_U = 2 * T2_NCOL;           // 448 .. 511

constexpr int T2_EPI_WARPS = 4;                  // warps 0..3
constexpr int T2_GROWS = 16;                     // slab rows per generator thread
constexpr int T2_GSPLIT = T2_SLAB / T2_GROWS;    // generator warps per TMEM lane quadrant
constexpr int T2_GEN_WARPS = 4 * T2_GSPLIT;      // warps 4..19: four per SM sub-partition, so
                                                 // MUFU and FMA-pipe phases of different warps overlap
constexpr int T2_WARP_MMA = T2_EPI_WARPS + T2_GEN_WARPS;
constexpr int T2_WARP_LOAD = T2_WARP_MMA + 1;
constexpr int T2_THREADS = (T2_WARP_LOAD + T2_LOAD_WARPS) * 32;
constexpr int T2_GPAIRS = T2_GROWS / 2;          // packed row pairs per generator thread
constexpr int T2_GCHUNKS = T2_GROWS / 8;         // 16-byte chunks per image row per thread

static_assert(T2_PHI_BYTES % 1024 == 0, "operand tiles must stay 1024-byte aligned");
static_assert(T2_TMEM_U + 64 == 512, "TMEM budget");

// ---- optional event trace (debug builds only: -DRR_T2_TRACE) ---------------------
#ifdef RR_T2_TRACE
constexpr int T2_TR_SLAB0 = 200, T2_TR_NSLAB = 48, T2_TR_SLOTS = 32;
__device__ long long g_t2_trace[T2_TR_NSLAB * T2_TR_SLOTS];
#define T2_TRACE(cond, slab, slot)                                                     \
  do {                                                                                 \
    if ((cond) && blockIdx.x == 0 && (slab) >= T2_TR_SLAB0 &&                          \
        (slab) < T2_TR_SLAB0 + T2_TR_NSLAB)                                            \
      g_t2_trace[((slab) - T2_TR_SLAB0) * T2_TR_SLOTS + (slot)] = clock64();           \
  } while (0)
#else
#define T2_TRACE(cond, slab, slot) do { } while (0)
#endif

struct T2Bars {
  uint64_t x_full[T2_XSTAGES];    // leader waits; count 2 (one loader per CTA)
  uint64_t x_empty[T2_XSTAGES];   // multicast commit
  uint64_t w_full;                // leader waits; count 8 (4 writer warps per CTA)
  uint64_t u_full;                // multicast commit
  uint64_t u_empty;               // leader waits; count 16 (generator warps)
  // one barrier per (stage, 16-row k-step): generator group h writes exactly the
  // K = 16 columns that MMA k-step h reads, so the tensor pipe can start on a slab
  // as soon as its first group is done and no group waits for another's slot
  uint64_t phi_full[T2_PSTAGES][T2_GSPLIT];   // LOCAL to each CTA; count 4 (warps of the group)
  uint64_t peer_full[T2_PSTAGES];             // leader only: the peer CTA's slab is done; count 1
  uint64_t phi_empty[T2_PSTAGES][T2_GSPLIT];  // multicast commit
  uint64_t acc_full;              // multicast commit
  uint64_t acc_empty;             // leader waits; count 8 (epilogue warps)
  uint32_t tmem_base;
  // epilogue column tables for the current item
  int2 etab[T2_NCOL];   // {column * D (or -1), ordering key}
};

__host__ __device__ __forceinline__ int t2_jmin(int ib) {
  int v = T2_IB * ib - (T2_JB - 1);
  return v <= 0 ? 0 : (v + T2_JB - 1) / T2_JB;
}

struct T2Item {
  int ib, jb;
  int64_t r0, r1;
  bool designated;
};

__device__ __forceinline__ T2Item t2_decode(int item, int ntiles, int NIB, int NJB,
                                            int64_t N, int64_t rpi) {
  // Item order: tiles are taken in groups of T2_TILE_GROUP; inside a group the
  // row super-chunks are the outer loop.  At any time the CTA pairs then share ONE
  // super-chunk of X (2.75 MB) and add into one group of tiles of the float64
  // scratch image (34 MB), both L2-resident.  (Super-chunk-major over all 173 tiles
  // wrote the 134 MB image back to DRAM on every round: 1.1 GB per launch;
  // tile-major re-read X from DRAM for every tile: 3.2 GB.)
  T2Item it;
  const int nsuper = (int)((N + rpi - 1) / rpi);
  const int grp = item / (T2_TILE_GROUP * nsuper);
  const int rem = item - grp * (T2_TILE_GROUP * nsuper);
  const int left = ntiles - grp * T2_TILE_GROUP;
  const int ntg = left < T2_TILE_GROUP ? left : T2_TILE_GROUP;
  const int sc = rem / ntg;
  int t = grp * T2_TILE_GROUP + (rem - sc * ntg);
  int ib = 0;
  for (; ib < NIB; ++ib) {
    const int cnt = NJB - t2_jmin(ib);
    if (t < cnt) break;
    t -= cnt;
  }
  it.ib = ib;
  it.jb = t2_jmin(ib) + t;
  it.designated = (t == 0);
  it.r0 = (int64_t)sc * rpi;
  it.r1 = it.r0 + rpi < N ? it.r0 + rpi : N;
  return it;
}

// arrive on the leader CTA's copy of a barrier (local or remote)
__device__ __forceinline__ void t2_arrive_leader(uint64_t* bar, uint32_t my_rank) {
  const uint32_t a = smem_u32(bar);
  mbar_arrive_cluster(mapa_u32(a, 0));
  (void)my_rank;
}

struct T2GenCtx {
  uint32_t off_cos[T2_GCHUNKS], off_sin[T2_GCHUNKS];   // h1-image chunk offsets inside a Phi stage (tile base included)
  uint32_t base;                     // shared address of the current Phi stage
  uint64_t* empty_bar;               // stage-free barrier, waited before the first store
  uint32_t empty_parity;
  float grid_magic;
#ifdef RR_T2_TRACE
  bool trw;
  int trs, trb;
#endif
};

// ---- packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2): one issue slot per two
//      values; the generators are issue-bound, not FMA-pipe-bound -------------
__device__ __forceinline__ uint64_t f2_pack(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_sub(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint32_t f2_to_h2(uint64_t v) {
  float a, b;
  f2_unpack(v, a, b);
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// One generator thread, one slab: T2_GROWS projections (in turns) of its frequency
// -> cos/sin -> fixed-point head / fp16 remainder images in shared memory.
// A-tile threads write the images (h1, r), B-tile threads (h1, r, cf); tile
// kind and store predicate are runtime flags so that ONE copy of this code
// serves every generator warp (a six-way template expansion was 100 KB of SASS
// and spent half of its issue slots waiting on instruction fetch).
//   MASKED: some of the rows are dead (tail slab, padded frequency) -- rare.
//   DO_P:   also accumulate (sum cos*y, sum sin*y) -- designated A tiles only.
// The source order is a software pipeline over row pairs: the MUFU work of
// pair i+2 is written before the FMA-pipe split of pair i so that the two
// pipes overlap inside one warp (each SM sub-partition hosts only two
// generator warps, and they run in lock step).
template <bool MASKED, bool DO_P>
__device__ __forceinline__ void t2_gen_slab(const T2GenCtx& cx, const float* u,
                                            uint32_t live, bool is_b, bool store,
                                            const float* yrow, uint64_t& pc2,
                                            uint64_t& ps2, int kind = 0) {
  const uint64_t RM2 = f2_pack(T2_RINT_MAGIC, T2_RINT_MAGIC);
  const uint64_t GM2 = f2_pack(cx.grid_magic, cx.grid_magic);
  const uint64_t TP2 = f2_pack(T2_TWO_PI, T2_TWO_PI);
  uint64_t c2[T2_GPAIRS], s2[T2_GPAIRS];
  uint4 hc, rc, hs, rs, cf, sf;
  uint32_t* hcp = reinterpret_cast<uint32_t*>(&hc);
  uint32_t* rcp = reinterpret_cast<uint32_t*>(&rc);
  uint32_t* hsp = reinterpret_cast<uint32_t*>(&hs);
  uint32_t* rsp = reinterpret_cast<uint32_t*>(&rs);
  uint32_t* cfp = reinterpret_cast<uint32_t*>(&cf);
  uint32_t* sfp = reinterpret_cast<uint32_t*>(&sf);

  auto trig = [&](int i) {   // rows 2i, 2i+1
    const uint64_t uu = f2_pack(u[2 * i], u[2 * i + 1]);
    const uint64_t kk = f2_sub(f2_add(uu, RM2), RM2);
    const uint64_t ang = f2_mul(f2_sub(uu, kk), TP2);
    float a0, a1;
    f2_unpack(ang, a0, a1);
    float c0 = __cosf(a0), s0 = __sinf(a0), c1 = __cosf(a1), s1 = __sinf(a1);
    if (MASKED) {
      // pseudo-frequency slots (affine columns): "cos" = u or 1, "sin" = 0
      if (kind == 1) {
        c0 = u[2 * i];
        c1 = u[2 * i + 1];
        s0 = s1 = 0.0f;
      } else if (kind == 2) {
        c0 = c1 = 1.0f;
        s0 = s1 = 0.0f;
      }
      const bool on0 = (live >> (2 * i)) & 1u, on1 = (live >> (2 * i + 1)) & 1u;
      c0 = on0 ? c0 : 0.0f;
      s0 = on0 ? s0 : 0.0f;
      c1 = on1 ? c1 : 0.0f;
      s1 = on1 ? s1 : 0.0f;
    }
    c2[i] = f2_pack(c0, c1);
    s2[i] = f2_pack(s0, s1);
    if (DO_P) {
      float y0 = 0.0f, y1 = 0.0f;
      if (!MASKED || ((live >> (2 * i)) & 1u)) y0 = __ldg(yrow + 2 * i);
      if (!MASKED || ((live >> (2 * i + 1)) & 1u)) y1 = __ldg(yrow + 2 * i + 1);
      const uint64_t y2 = f2_pack(y0, y1);
      pc2 = f2_fma(c2[i], y2, pc2);
      ps2 = f2_fma(s2[i], y2, ps2);
    }
  };
  auto split = [&](int i) {
    const int j = i & 3;
    const uint64_t h_c = f2_sub(f2_add(c2[i], GM2), GM2);
    const uint64_t h_s = f2_sub(f2_add(s2[i], GM2), GM2);
    hcp[j] = f2_to_h2(h_c);
    hsp[j] = f2_to_h2(h_s);
    rcp[j] = f2_to_h2(f2_sub(c2[i], h_c));
    rsp[j] = f2_to_h2(f2_sub(s2[i], h_s));
    cfp[j] = f2_to_h2(c2[i]);   // dead code for A tiles unless stored below
    sfp[j] = f2_to_h2(s2[i]);
    if (j == 3) {               // 8 rows complete: one 16-byte chunk per image
      const int cg = i >> 2;
      if (cg == 0) {
        T2_TRACE(cx.trw, cx.trs, cx.trb + 5);
        mbar_wait_cl(cx.empty_bar, cx.empty_parity);
        T2_TRACE(cx.trw, cx.trs, cx.trb + 6);
      }
#ifdef RR_T2_EXP_NOSTS     // experiment: generator cost without shared-memory stores
      if (store && hcp[0] == 0x12345678u) {
#else
      if (store) {
#endif
        const uint32_t oc = cx.base + cx.off_cos[cg], os = cx.base + cx.off_sin[cg];
        st_shared_v4(oc, hc);
        st_shared_v4(os, hs);
        st_shared_v4(oc + T2_H2R, rc);
        st_shared_v4(os + T2_H2R, rs);
        if (is_b) {
          st_shared_v4(oc + T2_H2C, cf);
          st_shared_v4(os + T2_H2C, sf);
        }
      }
    }
  };
#ifdef RR_T2_EXP_HALFGEN   // experiment: half of the generator work per slab
  trig(0);
  trig(1);
#pragma unroll
  for (int i = 0; i < T2_GPAIRS / 2; ++i) {
    if (i + 2 < T2_GPAIRS / 2) trig(i + 2);
    split(i);
  }
  c2[T2_GPAIRS - 1] = c2[0];
#else
  trig(0);
  trig(1);
#pragma unroll
  for (int i = 0; i < T2_GPAIRS; ++i) {
    if (i + 2 < T2_GPAIRS) trig(i + 2);
    split(i);
  }
#endif
}

__global__ void __launch_bounds__(T2_THREADS, 1)
tc2_suffstats_kernel(rr_plan plan, const float* __restrict__ X,
                     const float* __restrict__ y, int64_t N, double* __restrict__ T,
                     double* __restrict__ p, int NIB, int NJB, int ntiles,
                     int nitems, int64_t rpi, float grid_magic, int chain_rows) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ T2Bars sb;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const uint32_t crank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int d = plan.d, ktot = plan.ktot, D = plan.D;
  const int nk1 = (d + 7) >> 3;   // tf32 k-steps of the projection

  // ---- one-time setup ----------------------------------------------------------
  for (int i = tid; i < T2_SMEM_BYTES / 16; i += T2_THREADS)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int s = 0; s < T2_XSTAGES; ++s) {
      mbar_init(&sb.x_full[s], 2);
      mbar_init(&sb.x_empty[s], 1);
    }
    mbar_init(&sb.w_full, 8);
    mbar_init(&sb.u_full, 1);
    mbar_init(&sb.u_empty, 2 * T2_GEN_WARPS);
    for (int s = 0; s < T2_PSTAGES; ++s) {
      for (int h = 0; h < T2_GSPLIT; ++h) {
        mbar_init(&sb.phi_full[s][h], 4);
        mbar_init(&sb.phi_empty[s][h], 1);
      }
      mbar_init(&sb.peer_full[s], 1);
    }
    mbar_init(&sb.acc_full, 1);
    mbar_init(&sb.acc_empty, 2 * T2_EPI_WARPS);
    mbar_fence_init();
  }
  if (warp == T2_WARP_MMA) tmem_alloc_2cta(&sb.tmem_base, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem = sb.tmem_base;

  if (warp == T2_WARP_MMA) {
    // ============================ MMA issuer (leader CTA) =========================
    if (crank == 0) {
      const uint32_t idesc1 = make_idesc(2, 256, T2_SLAB);     // tf32, N = 64 rows
      const uint32_t idesc2 = make_idesc(0, 256, T2_NCOL);     // f16, N = 224
      const uint32_t w_hi = smem_u32(smem + T2_SMEM_W), w_lo = w_hi + T2_W_BYTES;
      uint32_t gs = 0, gc = 0, itc = 0;

      auto issue_mma1 = [&](uint32_t g) {
        const uint32_t xs = g % T2_XSTAGES;
        mbar_wait_cl(&sb.x_full[xs], (g / T2_XSTAGES) & 1);
        mbar_wait_cl(&sb.u_empty, (g & 1) ^ 1);
        tc_fence_after_sync();
        if (elect_one()) {
          const uint32_t x_hi = smem_u32(smem + T2_SMEM_X + xs * 2 * T2_X_BYTES);
          const uint32_t x_lo = x_hi + T2_X_BYTES;
          const uint64_t dwh = make_desc_sw128(w_hi), dwl = make_desc_sw128(w_lo);
          const uint64_t dxh = make_desc_sw128(x_hi), dxl = make_desc_sw128(x_lo);
#ifdef RR_T2_EXP_NOPROJ   // experiment: one projection MMA instead of 3 * nk1
          for (int k = 0; k < 1; ++k) {
            const uint64_t adv = (uint64_t)(2 * k);
            umma2_tf32_ss(tmem + T2_TMEM_U, dwh + adv, dxh + adv, idesc1, k != 0);
          }
#else
          for (int k = 0; k < nk1; ++k) {
            const uint64_t adv = (uint64_t)(2 * k);
            umma2_tf32_ss(tmem + T2_TMEM_U, dwh + adv, dxh + adv, idesc1, k != 0);
            umma2_tf32_ss(tmem + T2_TMEM_U, dwl + adv, dxh + adv, idesc1, 1);
            umma2_tf32_ss(tmem + T2_TMEM_U, dwh + adv, dxl + adv, idesc1, 1);
          }
#endif
          umma2_commit_mc(&sb.u_full);
          umma2_commit_mc(&sb.x_empty[xs]);
        }
        __syncwarp();
      };

      for (int item = pair; item < nitems; item += npairs, ++itc) {
        const T2Item it = t2_decode(item, ntiles, NIB, NJB, N, rpi);
        const int nsl = (int)((it.r1 - it.r0 + T2_SLAB - 1) / T2_SLAB);
        const int SPC = chain_rows / T2_SLAB;
        mbar_wait_cl(&sb.w_full, itc & 1);
        // Tensor-pipe order  P(0) P(1) | P(2) G(0) | P(3) G(1) | ...  (P = projection,
        // G = Gram update).  U is single-buffered, so P(t+2) can only be issued once
        // the generators have pulled U(t+1) into registers -- which they do right
        // after publishing Phi(t).  Issuing it BEFORE G(t) keeps the projection of
        // the next slab off the critical path: with the order P(t+1) G(t-1) ... the
        // generators could not start slab t+1 before G(t-1) had drained through the
        // pipe, and the slab period became (generate + Gram + projection)/2.
        issue_mma1(gs);
        if (nsl > 1) issue_mma1(gs + 1);
        for (int t = 0; t < nsl; ++t, ++gs) {
          const bool first = (t % SPC) == 0;
          const bool last = (t % SPC) == SPC - 1 || t == nsl - 1;
          if (first) mbar_wait_cl(&sb.acc_empty, (gc & 1) ^ 1);
          const uint32_t ps = gs % T2_PSTAGES;
          const uint32_t base = smem_u32(smem + T2_SMEM_PHI + ps * T2_PHI_BYTES);
          const uint64_t dah = make_desc_sw128(base + T2_OFF_AH);
          const uint64_t dar = make_desc_sw128(base + T2_OFF_AR);
          const uint64_t dbh = make_desc_sw128(base + T2_OFF_BH);
          const uint64_t dbr = make_desc_sw128(base + T2_OFF_BR);
          const uint64_t dbc = make_desc_sw128(base + T2_OFF_BC);
          T2_TRACE(lane == 0, gs, 8);
#ifdef RR_T2_EXP_NOMMA2    // experiment: one Gram k-step instead of four
          constexpr int KSTEPS = 1;
#else
          constexpr int KSTEPS = T2_SLAB / 16;
#endif
          static_assert(T2_GROWS == 16, "one generator group per MMA k-step");
#pragma unroll
          for (int k = 0; k < T2_GSPLIT; ++k)
            mbar_wait_cl(&sb.phi_full[ps][k], (gs / T2_PSTAGES) & 1);
          mbar_wait_cl(&sb.peer_full[ps], (gs / T2_PSTAGES) & 1);
          T2_TRACE(lane == 0, gs, 9);
          // The generators publish their st.shared writes with a plain (release)
          // mbarrier arrive; the generic -> async proxy fence is executed HERE, once
          // per slab by the consumer (and by the relay warp in the peer CTA for its
          // shared memory), instead of once per generator warp: in lock step the 16
          // per-warp fences cost ~280 cycles of every slab.
          fence_proxy_async_smem();
          tc_fence_after_sync();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < T2_SLAB / 16; ++k) {
              const uint64_t adv = (uint64_t)(2 * k);
              const uint32_t acc = (first && k == 0) ? 0u : 1u;
              if (k < KSTEPS) {
                umma2_f16_ss(tmem + T2_TMEM_MAIN, dah + adv, dbh + adv, idesc2, acc);
                umma2_f16_ss(tmem + T2_TMEM_AUX, dar + adv, dbc + adv, idesc2, acc);
                umma2_f16_ss(tmem + T2_TMEM_AUX, dah + adv, dbr + adv, idesc2, 1);
              }
              umma2_commit_mc(&sb.phi_empty[ps][k]);
            }
            if (last) umma2_commit_mc(&sb.acc_full);
          }
          __syncwarp();
          T2_TRACE(lane == 0, gs, 10);
          if (t + 2 < nsl) issue_mma1(gs + 2);
          __syncwarp();
          T2_TRACE(lane == 0, gs, 11);
          if (last) ++gc;
        }
      }
    }
    else {
      // ===================== relay (peer CTA's otherwise idle MMA warp) =============
      // waits for this CTA's generator groups, makes their writes visible to the
      // async proxy of THIS SM and tells the leader.
      uint32_t gs = 0;
      for (int item = pair; item < nitems; item += npairs) {
        const T2Item it = t2_decode(item, ntiles, NIB, NJB, N, rpi);
        const int nsl = (int)((it.r1 - it.r0 + T2_SLAB - 1) / T2_SLAB);
        for (int t = 0; t < nsl; ++t, ++gs) {
          const uint32_t ps = gs % T2_PSTAGES;
#pragma unroll
          for (int k = 0; k < T2_GSPLIT; ++k)
            mbar_wait_cl(&sb.phi_full[ps][k], (gs / T2_PSTAGES) & 1);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&sb.peer_full[ps]), 0));
        }
      }
    }
  } else if (warp >= T2_WARP_LOAD) {
    // ============================ X slab loaders ==================================
    // Two warps, each taking every other slab (one warp needed ~2300 cycles per slab
    // and would cap the kernel once the generators speed up).
    // Two steps per slab, both by one warp.  (1) cp.async: the CTA's 32 rows
    // are one contiguous block of X; lanes copy it word by word (coalesced, any
    // 4-byte alignment, zero fill past the last row) into a raw staging ring,
    // T2_RAW_AHEAD slabs ahead, so HBM latency never sits on the slab period.
    // (2) lane r reads row r back (stride d words), splits it into tf32 hi/lo
    // and writes both K-major swizzled tiles with 16-byte stores.
    struct SlabIter {
      int item, t, nsl;
      int64_t r0, r1;
    };
    auto load_item = [&](SlabIter& si) {
      if (si.item < nitems) {
        const T2Item it = t2_decode(si.item, ntiles, NIB, NJB, N, rpi);
        si.r0 = it.r0;
        si.r1 = it.r1;
        si.nsl = (int)((it.r1 - it.r0 + T2_SLAB - 1) / T2_SLAB);
        si.t = 0;
      }
    };
    auto advance = [&](SlabIter& si) {
      if (si.item < nitems && ++si.t >= si.nsl) {
        si.item += npairs;
        load_item(si);
      }
    };
    const int li = warp - T2_WARP_LOAD;
    const uint32_t raw0 = smem_u32(smem + T2_SMEM_RAW) + (uint32_t)li * (T2_RAW_AHEAD + 1) * T2_RAW_BYTES;
    auto prefetch = [&](const SlabIter& si, uint32_t stage) {
      if (si.item < nitems) {
        const int64_t row0 = si.r0 + (int64_t)si.t * T2_SLAB + 32 * (int64_t)crank;
        int vr = (int)(si.r1 - row0);
        vr = vr < 0 ? 0 : (vr > 32 ? 32 : vr);
        const int cnt = vr * d;
        const float* src = X + row0 * d;
        const uint32_t dst = raw0 + stage * T2_RAW_BYTES + 4u * (uint32_t)lane;
        for (int j = 0; j < d; ++j) {
          const int e = lane + 32 * j;
          const bool ok = e < cnt;
          const float* g = ok ? src + e : X;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + 128u * (uint32_t)j),
                       "l"(g), "r"(ok ? 4 : 0)
                       : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    SlabIter pf, cs;
    pf.item = cs.item = pair;
    load_item(pf);
    load_item(cs);
    for (int i = 0; i < li; ++i) {      // first own slab
      advance(pf);
      advance(cs);
    }
    for (int i = 0; i < T2_RAW_AHEAD; ++i) {
      prefetch(pf, (uint32_t)i);
      for (int j = 0; j < T2_LOAD_WARPS; ++j) advance(pf);
    }
    const int nch = 2 * nk1;   // 16-byte chunks of a tile row that MMA#1 reads
    uint32_t gs = (uint32_t)li, own = 0;
    while (cs.item < nitems) {
      prefetch(pf, (own + T2_RAW_AHEAD) % (T2_RAW_AHEAD + 1));
      for (int j = 0; j < T2_LOAD_WARPS; ++j) advance(pf);
      asm volatile("cp.async.wait_group %0;" ::"n"(T2_RAW_AHEAD) : "memory");
      __syncwarp();
      const uint32_t rrow = raw0 + (own % (T2_RAW_AHEAD + 1)) * T2_RAW_BYTES + 4u * (uint32_t)(lane * d);
      const uint32_t xs = gs % T2_XSTAGES;
      T2_TRACE(lane == 0, gs, 12);
      mbar_wait_cl(&sb.x_empty[xs], ((gs / T2_XSTAGES) & 1) ^ 1);
      T2_TRACE(lane == 0, gs, 13);
      const uint32_t x_hi = smem_u32(smem + T2_SMEM_X + xs * 2 * T2_X_BYTES);
      const uint32_t x_lo = x_hi + T2_X_BYTES;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c < nch) {
          uint4 hi, lo;
          uint32_t* hp = reinterpret_cast<uint32_t*>(&hi);
          uint32_t* lp = reinterpret_cast<uint32_t*>(&lo);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float v = 0.0f;
            if (4 * c + k < d) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(rrow + 4u * (4 * c + k)));
            const uint32_t hb = __float_as_uint(v) & 0xFFFFE000u;
            hp[k] = hb;
            lp[k] = __float_as_uint(v - __uint_as_float(hb));
          }
          const uint32_t off = sw128_off((uint32_t)lane, (uint32_t)c);
          st_shared_v4(x_hi + off, hi);
          st_shared_v4(x_lo + off, lo);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) t2_arrive_leader(&sb.x_full[xs], crank);
      T2_TRACE(lane == 0, gs, 14);
      for (int j = 0; j < T2_LOAD_WARPS; ++j) advance(cs);
      gs += T2_LOAD_WARPS;
      ++own;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else if (warp >= T2_EPI_WARPS) {
    // ============================ generators ======================================
    const int gw = warp - T2_EPI_WARPS;       // 0..T2_GEN_WARPS-1
    const int q = warp & 3;                   // TMEM lane quadrant of this warp
    const int h = gw >> 2;                    // which T2_GROWS-row group of the slab
    const int fl = 32 * q + lane;             // local frequency row 0..127
    const bool is_a = q < 2;                  // warp-uniform: lanes 0..63 feed the A tile
    const bool has_row = fl < T2_NA + T2_NB;  // lanes 120..127 own no feature rows
    const uint32_t w_hi = smem_u32(smem + T2_SMEM_W), w_lo = w_hi + T2_W_BYTES;
    // byte offsets of this thread's 16-byte chunks inside one operand image
    // (fixed for the whole kernel: row and chunk index never change)
    T2GenCtx cx;
    cx.grid_magic = grid_magic;
    {
      const uint32_t row_cos = is_a ? (uint32_t)fl : (uint32_t)(fl - T2_NA);
      const uint32_t row_sin = row_cos + (is_a ? (uint32_t)T2_NA : (uint32_t)T2_NB);
#pragma unroll
      for (int cg = 0; cg < T2_GCHUNKS; ++cg) {
        const uint32_t tile = is_a ? (uint32_t)T2_OFF_AH : (uint32_t)T2_OFF_BH;
        cx.off_cos[cg] = tile + sw128_off(row_cos, (uint32_t)(T2_GCHUNKS * h + cg));
        cx.off_sin[cg] = tile + sw128_off(row_sin, (uint32_t)(T2_GCHUNKS * h + cg));
      }
    }
    // shared::cluster address of the leader CTA's u_empty barrier (mapa reads a
    // special register; keep it out of the per-slab path)
    const uint32_t u_empty_leader = mapa_u32(smem_u32(&sb.u_empty), 0);
    uint32_t gs = 0;
    for (int item = pair; item < nitems; item += npairs) {
      const T2Item it = t2_decode(item, ntiles, NIB, NJB, N, rpi);
      const int nsl = (int)((it.r1 - it.r0 + T2_SLAB - 1) / T2_SLAB);
      const int tail_rows = (int)(it.r1 - it.r0) - (nsl - 1) * T2_SLAB;   // rows of the last slab
      const int theta = is_a ? T2_IB * it.ib + T2_NA * (int)crank + fl
                             : T2_JB * it.jb + T2_NB * (int)crank + (fl - T2_NA);
      const bool valid = has_row && theta < ktot;
      const int kind = (valid && plan.kind != nullptr) ? (int)plan.kind[theta] : 0;
      // lanes without a feature row store nothing: they may run the unmasked code;
      // warps that hold pseudo-frequency slots take the masked (general) variant
      const bool all_valid = __all_sync(0xffffffffu, (valid && kind == 0) || !has_row);
      // ---- W tile of this item (previous item's projections have all completed:
      //      this thread has consumed their U) -------------------------------------
      if (h == 0) {   // one warp per lane quadrant writes the W tile
        for (int i = 0; i < 32; ++i) {
          const float w = (valid && i < d) ? __ldg(plan.Wt + (int64_t)i * ktot + theta) : 0.0f;
          const float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
          const float lo = w - hi;
          const uint32_t off = sw128_off((uint32_t)fl, (uint32_t)i >> 2) + ((uint32_t)i & 3u) * 4u;
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(w_hi + off), "f"(hi) : "memory");
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(w_lo + off), "f"(lo) : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) t2_arrive_leader(&sb.w_full, crank);
      }
      const bool want_p = (p != nullptr) && (y != nullptr) && is_a && it.designated;  // warp-uniform
      double pc_d = 0.0, ps_d = 0.0;

      for (int t = 0; t < nsl; ++t, ++gs) {
        const int vrows = (t == nsl - 1) ? tail_rows : T2_SLAB;
        float u[T2_GROWS];
        const bool trw = lane == 0 && (gw == 0 || gw == T2_GEN_WARPS - 1);
        const int trb = gw == 0 ? 0 : 16;
        T2_TRACE(trw, gs, trb + 0);
        mbar_wait_cl(&sb.u_full, gs & 1);
        T2_TRACE(trw, gs, trb + 1);
        tc_fence_after_sync();
        tmem_ld16_nowait(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(T2_TMEM_U + T2_GROWS * h), u);
        tmem_ld_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(u_empty_leader);
        T2_TRACE(trw, gs, trb + 2);

        const uint32_t ps = gs % T2_PSTAGES;
        cx.base = smem_u32(smem + T2_SMEM_PHI + ps * T2_PHI_BYTES);
        cx.empty_bar = &sb.phi_empty[ps][h];
        cx.empty_parity = ((gs / T2_PSTAGES) & 1) ^ 1;
#ifdef RR_T2_TRACE
        cx.trw = trw;
        cx.trs = (int)gs;
        cx.trb = trb;
#endif
        // rows of this thread's half that are live (tail slab / padded frequency)
        const int lim = vrows - T2_GROWS * h;
        const uint32_t live = !valid ? 0u : (lim >= T2_GROWS ? 0xffffffffu
                                             : (lim <= 0 ? 0u : ((1u << lim) - 1u)));
        const bool masked = !(all_valid && vrows == T2_SLAB);   // warp-uniform
        if (want_p) {   // warp-uniform, designated A tiles only
          uint64_t pc2 = 0ull, ps2 = 0ull;   // (+0.0f, +0.0f)
          const float* yrow = y + it.r0 + (int64_t)t * T2_SLAB + T2_GROWS * h;
          if (masked) t2_gen_slab<true, true>(cx, u, live, false, true, yrow, pc2, ps2, kind);
          else t2_gen_slab<false, true>(cx, u, live, false, true, yrow, pc2, ps2);
          float a0, a1, b0, b1;
          f2_unpack(pc2, a0, a1);
          f2_unpack(ps2, b0, b1);
          pc_d += (double)a0 + (double)a1;
          ps_d += (double)b0 + (double)b1;
        } else {
          uint64_t dummy0 = 0ull, dummy1 = 0ull;
          if (masked) t2_gen_slab<true, false>(cx, u, has_row ? live : 0u, !is_a, has_row, nullptr, dummy0, dummy1, kind);
          else t2_gen_slab<false, false>(cx, u, live, !is_a, has_row, nullptr, dummy0, dummy1);
        }
        T2_TRACE(trw, gs, trb + 3);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sb.phi_full[ps][h]);
        T2_TRACE(trw, gs, trb + 4);
      }
      if (want_p && valid) {
        const double a = (double)plan.amp[theta];
        const int cc = plan.col_cos[theta], cs = plan.col_sin[theta];
        if (cc >= 0) atomicAdd(p + cc, a * pc_d);
        if (cs >= 0) atomicAdd(p + cs, a * ps_d);
      }
    }
  } else {
    // ============================ epilogue (warps 0..3) ===========================
    const int q = warp;
    const int L = 32 * q + lane;              // accumulator lane = local feature row
    const int type_a = L >> 6;
    uint32_t gc = 0;
    for (int item = pair; item < nitems; item += npairs) {
      const T2Item it = t2_decode(item, ntiles, NIB, NJB, N, rpi);
      const int64_t rows = it.r1 - it.r0;
      const int nch = (int)((rows + chain_rows - 1) / chain_rows);
      // column tables (safe to rewrite: every epilogue warp has finished the
      // previous item before any of them passes the barrier below)
      asm volatile("bar.sync 2, 128;" ::: "memory");
      for (int j = tid; j < T2_NCOL; j += 32 * T2_EPI_WARPS) {
        const int cc = j / T2_JB, within = j % T2_JB;
        const int ty = within / T2_NB;
        const int th = T2_JB * it.jb + T2_NB * cc + within % T2_NB;
        const bool ok = th < ktot;
        const int col = ok ? (ty ? plan.col_sin[th] : plan.col_cos[th]) : -1;
        // invalid columns get a key below every row key, so one compare decides
        sb.etab[j] = make_int2(col >= 0 ? col * D : 0, col >= 0 ? 2 * th + ty : -1);
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");
      const int theta_a = T2_IB * it.ib + T2_NA * (int)crank + (L & 63);
      const bool valid_a = theta_a < ktot;
      const int ca_raw = valid_a ? (type_a ? plan.col_sin[theta_a] : plan.col_cos[theta_a]) : -1;
      const int ca = ca_raw >= 0 ? ca_raw : 0;
      // rows that own no feature never pass the key test
      const int key_a = ca_raw >= 0 ? 2 * theta_a + type_a : 0x7fffffff;
      double* Tca = T + ca;
      for (int ch = 0; ch < nch; ++ch, ++gc) {
        mbar_wait_cl(&sb.acc_full, gc & 1);
        tc_fence_after_sync();
        // Drain MAIN + AUX: the two are added in fp32 (MAIN is an exact multiple of
        // 2^-10 below 2^14, so the rounding is <= 2^-11 absolute per 16384-row chain,
        // unbiased), widened to float64 with integer ops (F2F would queue on the XU
        // pipe behind the generators' MUFU work) and added to the scratch image with
        // RED.F64.  Amplitudes are applied by the finalize kernel.
        float vm[2][8], vx[2][8];
        const uint32_t ta = tmem + ((uint32_t)(32 * q) << 16);
        tmem_ld8_nowait(ta + T2_TMEM_MAIN, vm[0]);
        tmem_ld8_nowait(ta + T2_TMEM_AUX, vx[0]);
#pragma unroll 2
        for (int c0 = 0; c0 < T2_NCOL; c0 += 8) {
          const int cur = (c0 >> 3) & 1;
          tmem_ld_wait();
          if (c0 + 8 < T2_NCOL) {
            tmem_ld8_nowait(ta + (uint32_t)(T2_TMEM_MAIN + c0 + 8), vm[cur ^ 1]);
            tmem_ld8_nowait(ta + (uint32_t)(T2_TMEM_AUX + c0 + 8), vx[cur ^ 1]);
          }
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int2 tb = sb.etab[c0 + r];
#ifdef RR_T2_EXP_NODRAIN
            if (key_a <= tb.y && vm[cur][r] == 123.456f) {
#else
            if (key_a <= tb.y) {
#endif
              const uint32_t b = __float_as_uint(vm[cur][r] + vx[cur][r]);
              // fp32 -> fp64 bit pattern (normal numbers; +-0 becomes +-2^-127)
              const uint32_t hi = (b & 0x80000000u) | (((b & 0x7fffffffu) >> 3) + 0x38000000u);
              const double val = __hiloint2double((int)hi, (int)(b << 29));
              atomicAdd(Tca + tb.x, val);
            }
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) t2_arrive_leader(&sb.acc_empty, crank);
      }
    }
  }

  // ---- teardown ------------------------------------------------------------------
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == T2_WARP_MMA) tmem_dealloc_2cta(tmem, 512);
}

// camp[c] = amplitude of feature column c (both columns of a frequency share it).
__global__ void t2_colamp_kernel(rr_plan plan, float* __restrict__ camp) {
  const int th = blockIdx.x * blockDim.x + threadIdx.x;
  if (th < plan.ktot) {
    const float a = plan.amp[th];
    if (plan.col_cos[th] >= 0) camp[plan.col_cos[th]] = a;
    if (plan.col_sin[th] >= 0) camp[plan.col_sin[th]] = a;
  }
}

// G[i][j] += a_i a_j (T[i][j] + T[j][i]) (i != j), G[i][i] += a_i^2 T[i][i]: every
// unordered feature pair was accumulated exactly once, at either position.
__global__ void __launch_bounds__(256)
t2_finalize_kernel(const double* __restrict__ T, const float* __restrict__ camp,
                   double* __restrict__ G, int D) {
  __shared__ double tile[32][33];
  const int bx = blockIdx.x, by = blockIdx.y;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  // transposed block T[bx-block rows][by-block cols] -> tile
  for (int r = ty; r < 32; r += 8) {
    const int i = bx * 32 + r, j = by * 32 + tx;
    tile[r][tx] = (i < D && j < D) ? T[(int64_t)i * D + j] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = by * 32 + r, j = bx * 32 + tx;
    if (i < D && j < D) {
      const double a = T[(int64_t)i * D + j];
      const double b = tile[tx][r];            // T[j][i]
      const double w = (double)camp[i] * (double)camp[j];
      G[(int64_t)i * D + j] += w * ((i == j) ? a : a + b);
    }
  }
}

// ---------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------
int tc_suffstats_supported(const rr_plan* pl) {
  // pure trigonometric plans, or plans whose affine columns ride along as
  // pseudo-frequency slots (kind != NULL; then D <= 2 * ktot)
  if (pl->d < 1 || pl->d > 32 || pl->ktot < 1 || pl->next != 0 || pl->ext_pow != nullptr) return 0;
  return (pl->kind != nullptr ? pl->D <= 2 * pl->ktot : pl->D == 2 * pl->ktot) ? 1 : 0;
}

size_t tc_suffstats_workspace(const rr_plan* pl, int64_t) {
  return align_up((size_t)pl->D * pl->D * sizeof(double), 256) +
         align_up((size_t)pl->D * sizeof(float), 256) + 256;
}

int tc_suffstats(const rr_plan* pl, const float* X, const float* y, int64_t N,
                 double* G, double* p, void* ws, size_t ws_bytes, int grid_bits,
                 cudaStream_t st) {
  if (grid_bits < 5 || grid_bits > 8) {
    set_error("tcgen05 suffstats: grid_bits must be 5..8");
    return RR_ERR_INVALID;
  }
  const float grid_magic = 1.5f * (float)(1 << (23 - grid_bits));
  const int chain_rows = 1 << (24 - 2 * grid_bits);
  const int D = pl->D;
  Workspace W(ws, ws_bytes);
  double* T = W.take<double>((size_t)D * D);
  float* camp = W.take<float>((size_t)D);
  if (!T || !camp) {
    set_error("tcgen05 suffstats workspace too small (need %zu bytes)",
              tc_suffstats_workspace(pl, N));
    return RR_ERR_WORKSPACE;
  }
  RR_CUDA_CHECK(cudaMemsetAsync(T, 0, (size_t)D * D * sizeof(double), st));
  t2_colamp_kernel<<<(pl->ktot + 255) / 256, 256, 0, st>>>(*pl, camp);
  RR_LAUNCH_CHECK("t2_colamp_kernel");
  const int NIB = (pl->ktot + T2_IB - 1) / T2_IB;
  const int NJB = (pl->ktot + T2_JB - 1) / T2_JB;
  int ntiles = 0;
  for (int ib = 0; ib < NIB; ++ib) ntiles += NJB - t2_jmin(ib);
  int64_t rpi = T2_SUPER;
  const int64_t nsuper = (N + rpi - 1) / rpi;
  const int64_t nitems = nsuper * ntiles;
  int npairs = sm_count() / 2;
  if (nitems < npairs) npairs = (int)nitems;
  const size_t smem = T2_SMEM_BYTES + 1024;
  RR_CUDA_CHECK(cudaFuncSetAttribute(tc2_suffstats_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * npairs);
  cfg.blockDim = dim3(T2_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  RR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, tc2_suffstats_kernel, *pl, X, y, N, T, p, NIB,
                                   NJB, ntiles, (int)nitems, rpi, grid_magic, chain_rows));
  RR_LAUNCH_CHECK("tc2_suffstats_kernel");
  dim3 fg((D + 31) / 32, (D + 31) / 32);
  t2_finalize_kernel<<<fg, 256, 0, st>>>(T, camp, G, D);
  RR_LAUNCH_CHECK("t2_finalize_kernel");
  return RR_OK;
}

}  // namespace rr

#ifdef RR_T2_TRACE
extern "C" int rr_debug_t2_trace(long long* out, int n) {
  const int m = rr::T2_TR_NSLAB * rr::T2_TR_SLOTS;
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out, rr::g_t2_trace, sizeof(long long) * (n < m ? n : m)) == cudaSuccess ? m : -1;
}
#endif
