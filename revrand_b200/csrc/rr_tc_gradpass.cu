// Residual + gradient pass of the SLM log marginal likelihood on tcgen05.
//
//   Err = y - Phi m,   sqerr = sum Err^2                        (slm.py:161-162)
//   T = Err (x) m - Phi C,
//   Q[n,k] = amp_k (-sin_nk T[n,cos k] + cos_nk T[n,sin k]),
//   R += X^T Q                                                  (d x ktot, float64)
//
// dPhi (N x 2K x d in the reference, basis_functions.py:888-901) is never
// formed.  Phi reaches the GEMM as a TILE-MAJOR fp16 image: every (256 rows x 64
// features) block is one contiguous 32 KB image already in the 128-byte-swizzled
// K-major layout tcgen05 wants, so an operand stage is a single cp.async.bulk.
//
//   A. tc_gradpass_kept (rr_slm_gradpass_kept; up to 6144 feature columns): the
//      value pass of the same evaluation left the whole image behind
//      (t3_digits_kernel<true>, rr_tc3_syrk.cu) -- the feature map is evaluated once
//      per evaluation, as at slm.py:145.  mint_kernel + fit16_kernel stream it once
//      for the residuals (89 % of DRAM peak), then ONE persistent gp2_kernel launch
//      covers all rows.
//   B. tc_gradpass (rr_slm_gradpass; large K, or no room for the image): per row
//      chunk (up to 320 MB, two ping-pong buffers) phi_fit_kernel regenerates the
//      chunk on the CUDA cores -- fp32 projection, exact range reduction, MUFU
//      sin/cos -- and accumulates f = Phi m from the same trig values on the
//      context's helper stream while gp2_kernel works on the previous chunk.
//
//   gp2_kernel: persistent CTA pairs (cta_group::2, M = 256 rows, N = 256 output
//   features, K = all features) stream the images through a 5-stage bulk-copy ring;
//   accumulators are double-buffered in TMEM (2 x 256 columns) so that the epilogue
//   of one tile (Q in registers, X^T Q on the CUDA cores, float64 atomics into R)
//   overlaps the tensor work of the next.  Tiles are visited in supertiles
//   (g2_tile) sized so that the operand panels in flight stay in L2.
//   RR_GRAD_SPLIT_C adds a second launch over the fp16 rounding residual of C.
//
// Replaces: revrand/slm.py:161-162, 193-197 + basis_functions.py:109-152, 888-901.
#include <stdlib.h>

#include "rr_common.cuh"
#include "rr_tc.cuh"

namespace rr {

using namespace tc;

constexpr int G2_TM = 256;              // rows per tile (128 per CTA of the pair)
constexpr int G2_TN = 256;              // output features per tile
constexpr int G2_KT = 64;               // reduction extent per stage (one 128-byte line)
constexpr int G2_IMG = G2_TM * 128;     // one tile-major operand image: 32 KB
constexpr int G2_HALF = G2_IMG / 2;     // what one CTA of the pair loads of it
constexpr int G2_STAGES = 5;
constexpr int G2_STAGE_BYTES = 2 * G2_HALF;   // A half + B half
constexpr int G2_THREADS = 6 * 32;      // producer, MMA / relay, 4 epilogue warps
constexpr int G2_RED = 4 * 32 * 32;     // floats of one cross-warp reduction buffer
// Phi chunk budget.  The chunk does not have to fit in L2: tiles are visited in
// supertiles (g2_tile), so the CTA pairs running at any time share a handful of
// row-block panels of Phi and of panels of the C image; each panel is fetched from
// HBM once per supertile and then hit in L2.  Large chunks amortise the per-launch
// pipeline fill and the exposed last epilogue.
constexpr int64_t G2_SCRATCH_BYTES = 320ll << 20;

// internal feature f -> (frequency, is_sin): blocks of [64 cos | 64 sin]
__device__ __forceinline__ int feat_theta(int f) { return 64 * (f >> 7) + (f & 63); }
__device__ __forceinline__ int feat_is_sin(int f) { return (f >> 6) & 1; }

// byte offset of element (row, col) inside the tile-major image array whose
// tiles are [row block of 256][k block of 64]
__device__ __forceinline__ int64_t tile_off(int row, int col, int nkb) {
  const int rb = row >> 8, r = row & 255, kb = col >> 6, c = col & 63;
  return ((int64_t)rb * nkb + kb) * G2_IMG + sw128_off((uint32_t)r, (uint32_t)(c >> 3)) + (c & 7) * 2;
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

// Bt[fo][fj] = s * amp_fj * C[col fo][col fj] as fp16, tile-major, s = 1/max|C|.
// Rows fo: the Dp (padded) trigonometric output features; columns fj: all Dk
// reduction features = the trigonometric ones followed by the affine columns
// (Linear / Bias bases, amplitude 1) padded to a multiple of 64.
// part 1 (RR_GRAD_SPLIT_C): what fp16 rounding removed from the same entry, scaled up
// by G2_LO_SCALE into fp16's normal range -- a second GEMM over it restores C to ~22
// bits (the lengthscale gradients are linear in C).
constexpr float G2_LO_SCALE = 2048.0f;
__global__ void __launch_bounds__(256)
prep_c_kernel(rr_plan plan, const float* __restrict__ C, int Dp, int Dk,
              const unsigned int* __restrict__ cmax_bits, uint8_t* __restrict__ BtT, int part) {
  const int fj = blockIdx.x * blockDim.x + threadIdx.x;
  const int fo = blockIdx.y;
  if (fj >= Dk) return;
  const float cmax = __uint_as_float(*cmax_bits);
  const float s = cmax > 0.0f ? 1.0f / cmax : 0.0f;
  const int to = feat_theta(fo);
  float v = 0.0f;
  if (to < plan.ktot) {
    const int co = feat_is_sin(fo) ? plan.col_sin[to] : plan.col_cos[to];
    if (fj < Dp) {
      const int tj = feat_theta(fj);
      if (tj < plan.ktot) {
        const int cj = feat_is_sin(fj) ? plan.col_sin[tj] : plan.col_cos[tj];
        v = s * plan.amp[tj] * C[(int64_t)co * plan.D + cj];
      }
    } else if (fj - Dp < plan.next) {
      v = s * C[(int64_t)co * plan.D + plan.ext_col[fj - Dp]];
    }
  }
  __half hv = __float2half_rn(v);
  if (part == 1) hv = __float2half_rn((v - __half2float(hv)) * G2_LO_SCALE);
  *reinterpret_cast<__half*>(BtT + tile_off(fo, fj, Dk / G2_KT)) = hv;
}

// ---- Phi chunk + fitted values --------------------------------------------------
// Block = PE_ROWS rows x PE_PAIRS frequency pairs; thread = one pair (2p, 2p+1)
// for all rows of the block: every X value read from shared memory feeds two
// FMAs, and the fp16 results leave as half2 so that a warp writes one whole
// 128-byte line of the tile-major image per instruction.  The grid's y dimension
// splits the frequencies, which gives a 4608-row chunk ~2300 blocks (the first
// version ran one 8-warp block per SM and was latency bound); partial fitted
// values f = Phi m are combined with float atomics in fbuf.
constexpr int PE_ROWS = 16;
constexpr int PE_PAIRS = 128;
template <bool STORE>
__global__ void __launch_bounds__(PE_PAIRS)
phi_fit_kernel(rr_plan plan, const float* __restrict__ X, int rows, int Dp, int Dk,
               int gy_trig, const float* __restrict__ m, uint8_t* __restrict__ PhT,
               float* __restrict__ fbuf) {
  extern __shared__ float xs[];            // PE_ROWS x d
  __shared__ float part[PE_PAIRS / 32][PE_ROWS];
  const int d = plan.d, ktot = plan.ktot;
  const int n0 = blockIdx.x * PE_ROWS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int t = tid; t < PE_ROWS * d; t += PE_PAIRS) {
    const int r = t / d;
    xs[t] = (n0 + r < rows) ? X[(int64_t)n0 * d + t] : 0.0f;
  }
  __syncthreads();
  if (STORE && (int)blockIdx.y >= gy_trig) {
    // affine columns (Linear / Bias bases) of the chunk: internal features
    // Dp + 2e, Dp + 2e + 1 as one half2 per row; zero padding up to Dk
    const int e0 = 2 * (((int)blockIdx.y - gy_trig) * PE_PAIRS + tid);
    if (Dp + e0 < Dk) {
      const int fc = Dp + e0;
      const int nkb = Dk / G2_KT;
      uint8_t* img = PhT + ((int64_t)(n0 >> 8) * nkb + (fc >> 6)) * G2_IMG + (int64_t)(n0 & 255) * 128;
      const uint32_t c4 = (uint32_t)((fc & 63) >> 3) << 4, inner = (uint32_t)(fc & 7) * 2;
      const bool in0 = e0 < plan.next, in1 = e0 + 1 < plan.next;
      const int s0 = in0 ? plan.ext_src[e0] : -1;           // < 0: constant column
      const int s1 = in1 ? plan.ext_src[e0 + 1] : -1;
      const float c0 = (in0 && s0 < 0) ? plan.ext_val[e0] : 0.0f;
      const float c1 = (in1 && s1 < 0) ? plan.ext_val[e0 + 1] : 0.0f;
#pragma unroll
      for (int r = 0; r < PE_ROWS; ++r) {
        const bool live = n0 + r < rows;
        const float v0 = !live ? 0.0f : (in0 ? ext_value(plan, e0, xs + r * d, 1) : c0);
        const float v1 = !live ? 0.0f : (in1 ? ext_value(plan, e0 + 1, xs + r * d, 1) : c1);
        *reinterpret_cast<__half2*>(img + r * 128 + (c4 ^ ((uint32_t)(r & 7) << 4)) + inner) =
            __floats2half2_rn(v0, v1);
      }
    }
    return;
  }
  const int th0 = 2 * (blockIdx.y * PE_PAIRS + tid);   // even frequency of this pair
  const bool v0 = th0 < ktot, v1 = th0 + 1 < ktot;
  float u0[PE_ROWS], u1[PE_ROWS];
#pragma unroll
  for (int r = 0; r < PE_ROWS; ++r) u0[r] = u1[r] = 0.0f;
  float mc0 = 0.0f, ms0 = 0.0f, mc1 = 0.0f, ms1 = 0.0f;
  if (v0) {
    for (int i = 0; i < d; ++i) {
      const float w0 = __ldg(plan.Wt + (int64_t)i * ktot + th0);
      const float w1 = v1 ? __ldg(plan.Wt + (int64_t)i * ktot + th0 + 1) : 0.0f;
#pragma unroll
      for (int r = 0; r < PE_ROWS; ++r) {
        const float x = xs[r * d + i];
        u0[r] = fmaf(x, w0, u0[r]);
        u1[r] = fmaf(x, w1, u1[r]);
      }
    }
    const float a0 = plan.amp[th0];
    mc0 = a0 * m[plan.col_cos[th0]];
    ms0 = a0 * m[plan.col_sin[th0]];
    if (v1) {
      const float a1 = plan.amp[th0 + 1];
      mc1 = a1 * m[plan.col_cos[th0 + 1]];
      ms1 = a1 * m[plan.col_sin[th0 + 1]];
    }
  }
  // tile-major addresses: rows n0.. lie in one 256-row image (n0 % 16 == 0), the
  // pair occupies 4 bytes of chunk (c >> 3) of the cos image and of the sin image
  // (the next k block); the 16-byte chunk index is XORed with (row & 7).
  uint8_t* img = nullptr;
  uint32_t inner = 0, c4 = 0;
  // a block covers 2 * PE_PAIRS frequencies; when the padded frequency count is
  // only half of that (ktot <= 64 mod 128) the upper threads own no column
  const bool st_ok = STORE && th0 < Dp / 2;
  if (st_ok) {
    const int fc = 128 * (th0 >> 6) + (th0 & 63);     // internal cos column of th0
    const int nkb = Dk / G2_KT;
    img = PhT + ((int64_t)(n0 >> 8) * nkb + (fc >> 6)) * G2_IMG + (int64_t)(n0 & 255) * 128;
    c4 = (uint32_t)((fc & 63) >> 3) << 4;
    inner = (uint32_t)(fc & 7) * 2;
  }
  float f[PE_ROWS];
#pragma unroll
  for (int r = 0; r < PE_ROWS; ++r) {
    float c0 = 0.0f, s0 = 0.0f, c1 = 0.0f, s1 = 0.0f;
    const bool live = n0 + r < rows;
    if (v0 && live) {
      const float fr = (u0[r] - rintf(u0[r])) * 6.283185307179586f;
      s0 = __sinf(fr);
      c0 = __cosf(fr);
    }
    if (v1 && live) {
      const float fr = (u1[r] - rintf(u1[r])) * 6.283185307179586f;
      s1 = __sinf(fr);
      c1 = __cosf(fr);
    }
    f[r] = fmaf(c0, mc0, fmaf(s0, ms0, fmaf(c1, mc1, s1 * ms1)));
    if (st_ok) {
      // (n0 + r) & 7 == r & 7 because n0 is a multiple of 16
      uint8_t* p = img + r * 128 + (c4 ^ ((uint32_t)(r & 7) << 4)) + inner;
      *reinterpret_cast<__half2*>(p) = __floats2half2_rn(c0, c1);
      *reinterpret_cast<__half2*>(p + G2_IMG) = __floats2half2_rn(s0, s1);
    }
  }
  // extra (non-trigonometric) columns: Linear / Bias bases (first frequency slice only)
  if (blockIdx.y == 0) {
    for (int j = tid; j < plan.next; j += PE_PAIRS) {
      const float mj = m[plan.ext_col[j]];
#pragma unroll
      for (int r = 0; r < PE_ROWS; ++r)
        f[r] = fmaf(ext_value(plan, j, xs + r * d, 1), mj, f[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < PE_ROWS; ++r) {
    float v = f[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) part[warp][r] = v;
  }
  __syncthreads();
  if (tid < PE_ROWS && n0 + tid < rows) {
    float v = 0.0f;
#pragma unroll
    for (int w = 0; w < PE_PAIRS / 32; ++w) v += part[w][tid];
    atomicAdd(fbuf + n0 + tid, v);
  }
}

// err = y - f (optional), sqerr += sum (y - f)^2.
__global__ void __launch_bounds__(256)
resid_finish_kernel(const float* __restrict__ y, const float* __restrict__ fbuf, int64_t n,
                    float* __restrict__ err, double* __restrict__ sqerr) {
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float e = y[i] - fbuf[i];
    if (err) err[i] = e;
    acc += (double)e * (double)e;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0 && acc != 0.0) atomicAdd(sqerr, acc);
}

static int launch_phi_fit(const rr_plan* pl, const float* X, int rows, int rows_pad, int Dp,
                          int Dk, const float* m, uint8_t* PhT, float* fbuf, cudaStream_t st) {
  const int nth = PhT ? Dp / 2 : pl->ktot;             // padded frequency count when storing
  int gy = (nth + 2 * PE_PAIRS - 1) / (2 * PE_PAIRS);
  if (gy < 1) gy = 1;                                  // plans with no trig block: extras only
  const int gy_ext = PhT ? ((Dk - Dp) / 2 + PE_PAIRS - 1) / PE_PAIRS : 0;   // affine column blocks
  dim3 grid((PhT ? rows_pad : rows + PE_ROWS - 1) / PE_ROWS, gy + gy_ext);
  const size_t smem = PE_ROWS * pl->d * sizeof(float);
  if (PhT) phi_fit_kernel<true><<<grid, PE_PAIRS, smem, st>>>(*pl, X, rows, Dp, Dk, gy, m, PhT, fbuf);
  else phi_fit_kernel<false><<<grid, PE_PAIRS, smem, st>>>(*pl, X, rows, Dp, Dk, gy, m, nullptr, fbuf);
  RR_LAUNCH_CHECK("phi_fit_kernel");
  return RR_OK;
}

// ---- fitted values from a kept fp16 feature image ---------------------------------
// mi[f] = weight of internal feature f in f = Phi m: amp * m[column] for the
// trigonometric features, m[column] for the affine ones, 0 in the padding.
__global__ void __launch_bounds__(256)
mint_kernel(rr_plan plan, const float* __restrict__ m, int Dp, int Dk, float* __restrict__ mi) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= Dk) return;
  float v = 0.0f;
  if (f < Dp) {
    const int th = feat_theta(f);
    if (th < plan.ktot) {
      const int c = feat_is_sin(f) ? plan.col_sin[th] : plan.col_cos[th];
      if (c >= 0) v = plan.amp[th] * m[c];
    }
  } else if (f - Dp < plan.next) {
    v = m[plan.ext_col[f - Dp]];
  }
  mi[f] = v;
}

// err = y - Phi m, sqerr += sum err^2 from the tile-major fp16 image.  Block = one
// 256-row image column (all k blocks), warp = 32 rows; a warp instruction reads four
// whole 128-byte lines (lane = row sub-index x 16-byte chunk), so the pass streams
// the image once at HBM speed.  MI_SMEM: the weights fit in shared memory.
template <bool MI_SMEM>
__global__ void __launch_bounds__(256)
fit16_kernel(const uint8_t* __restrict__ PhT, int nkb, const float* __restrict__ mi,
             const float* __restrict__ y, int64_t rows, float* __restrict__ err,
             double* __restrict__ sqerr) {
  extern __shared__ float mis[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (MI_SMEM) {
    for (int e = tid; e < nkb * 64; e += 256) mis[e] = mi[e];
    __syncthreads();
  }
  const float* mw = MI_SMEM ? mis : mi;
  const int rsub = lane >> 3, p = lane & 7;
  const uint8_t* base = PhT + (int64_t)blockIdx.x * nkb * G2_IMG + (32 * warp + rsub) * 128 + p * 16;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
  for (int kb = 0; kb < nkb; ++kb) {
    const uint8_t* img = base + (int64_t)kb * G2_IMG;
    uint4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldg(reinterpret_cast<const uint4*>(img + i * 512));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // row 32 warp + 4 i + rsub: its chunk at position p holds features 8 (p ^ (row & 7)) ..
      const int c = p ^ ((4 * i + rsub) & 7);
      const float4 w0 = *reinterpret_cast<const float4*>(mw + kb * 64 + 8 * c);
      const float4 w1 = *reinterpret_cast<const float4*>(mw + kb * 64 + 8 * c + 4);
      const __half2* h = reinterpret_cast<const __half2*>(&v[i]);
      const float2 a = __half22float2(h[0]), b = __half22float2(h[1]);
      const float2 c2 = __half22float2(h[2]), d2 = __half22float2(h[3]);
      float t = acc[i];
      t = fmaf(a.x, w0.x, t);
      t = fmaf(a.y, w0.y, t);
      t = fmaf(b.x, w0.z, t);
      t = fmaf(b.y, w0.w, t);
      t = fmaf(c2.x, w1.x, t);
      t = fmaf(c2.y, w1.y, t);
      t = fmaf(d2.x, w1.z, t);
      t = fmaf(d2.y, w1.w, t);
      acc[i] = t;
    }
  }
  double sq = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float t = acc[i];
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 4);
    const int64_t n = (int64_t)blockIdx.x * 256 + 32 * warp + 4 * i + rsub;
    if (p == 0 && n < rows) {
      const float e = y[n] - t;
      err[n] = e;
      sq += (double)e * (double)e;
    }
  }
  sq = warp_sum(sq);
  if (lane == 0 && sq != 0.0) atomicAdd(sqerr, sq);
}

// ---- GEMM + fused epilogue ------------------------------------------------------
struct G2Bars {
  uint64_t full[G2_STAGES];        // own bulk copies landed (complete_tx)
  uint64_t peer_full[G2_STAGES];   // leader only: the peer CTA's copies landed
  uint64_t empty[G2_STAGES];       // multicast commit: stage consumed by the MMAs
  uint64_t acc_full[2];            // multicast commit: accumulator buffer complete
  uint64_t acc_empty[2];           // leader waits; count 8 (epilogue warps of both CTAs)
  uint32_t tmem_base;
};

// Tile order.  Tile t of the launch -> (row block rb, output-feature block fb), visited
// in supertiles of `sup` row blocks x all FB feature blocks with the row block fastest.
// sup = 1 (feature block fastest) while the fp16 image of C fits in L2 next to a few
// row panels of Phi (32 MB at K = 2048: every panel of Phi is then read from HBM exactly
// once, 8.5 GB per pass under ncu; with sup = 8 it was 19 GB and 5 % slower).  For
// K >= 4096 the C image (135 MB) does not fit, and with the feature block fastest every
// row block streamed all of it from HBM: sup = 8 lets the ~74 tiles in flight share
// ~8 row panels and ~9 panels of C (gradient pass at K = 4096, 1.25e6 rows: 140 -> 130 ms).
constexpr int G2_SUPER_BIG = 8;
constexpr size_t G2_C_IMAGE_L2_BYTES = (size_t)48 << 20;
__device__ __forceinline__ void g2_tile(int t, int RB, int FB, int sup, int& rb, int& fb) {
  const int per = sup * FB;
  const int s = t / per, tl = t - s * per;
  const int left = RB - s * sup;
  const int rs = left < sup ? left : sup;
  fb = tl / rs;
  rb = s * sup + (tl - fb * rs);
}

template <int IG>   // input dimensions per reducing warp: d <= 4 * IG
__global__ void __launch_bounds__(G2_THREADS, 1)
gp2_kernel(rr_plan plan, const float* __restrict__ X, const float* __restrict__ err,
           int rows, int RB, int FB, int sup, int nkb, const uint8_t* __restrict__ PhT,
           const uint8_t* __restrict__ BtT, const float* __restrict__ m,
           const unsigned int* __restrict__ cmax_bits, float cscale, double* __restrict__ R) {
  constexpr int DPAD = 4 * IG;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ G2Bars sb;
  __shared__ float m_loc[G2_TN];
  __shared__ float amp_loc[G2_TN / 2];
  __shared__ float err_loc[G2_TM / 2];
  float* red = reinterpret_cast<float*>(smem + G2_STAGES * G2_STAGE_BYTES);  // [2][4][DPAD][32]
  float* xs = red + 2 * G2_RED;                                               // [128][DPAD]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t crank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  // nkb: k blocks over all reduction features (trigonometric + affine columns)
  const int ntiles = RB * FB;
  const int d = plan.d, ktot = plan.ktot;

  if (tid == 0) {
    for (int s = 0; s < G2_STAGES; ++s) {
      mbar_init(&sb.full[s], 1);
      mbar_init(&sb.peer_full[s], 1);
      mbar_init(&sb.empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sb.acc_full[b], 1);
      mbar_init(&sb.acc_empty[b], 8);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc_2cta(&sb.tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem = sb.tmem_base;

  if (warp == 0) {
    // ============================ producer (both CTAs) ============================
    if (elect_one()) {
      uint32_t g = 0;
      for (int t = pair; t < ntiles; t += npairs) {
        int rb, fb;
        g2_tile(t, RB, FB, sup, rb, fb);
        const uint8_t* a_src = PhT + ((int64_t)rb * nkb) * G2_IMG + (int64_t)crank * G2_HALF;
        const uint8_t* b_src = BtT + ((int64_t)fb * nkb) * G2_IMG + (int64_t)crank * G2_HALF;
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const uint32_t s = g % G2_STAGES;
          mbar_wait_cl(&sb.empty[s], ((g / G2_STAGES) & 1) ^ 1);
          const uint32_t dst = smem_u32(smem + s * G2_STAGE_BYTES);
          mbar_expect_tx(&sb.full[s], G2_STAGE_BYTES);
          bulk_g2s(dst, a_src + (int64_t)kb * G2_IMG, G2_HALF, &sb.full[s]);
          bulk_g2s(dst + G2_HALF, b_src + (int64_t)kb * G2_IMG, G2_HALF, &sb.full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (crank == 0) {
      // ========================== MMA issuer (leader CTA) ==========================
      const uint32_t idesc = make_idesc(0, G2_TM, G2_TN);
      uint32_t g = 0, it = 0;
      for (int t = pair; t < ntiles; t += npairs, ++it) {
        const uint32_t buf = it & 1;
        mbar_wait_cl(&sb.acc_empty[buf], ((it >> 1) & 1) ^ 1);
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const uint32_t s = g % G2_STAGES;
          const uint32_t ph = (g / G2_STAGES) & 1;
          mbar_wait_cl(&sb.full[s], ph);
          mbar_wait_cl(&sb.peer_full[s], ph);
          tc_fence_after_sync();
          if (elect_one()) {
            const uint32_t a0 = smem_u32(smem + s * G2_STAGE_BYTES);
            const uint64_t da = make_desc_sw128(a0);
            const uint64_t db = make_desc_sw128(a0 + G2_HALF);
#pragma unroll
            for (int k = 0; k < G2_KT / 16; ++k) {
              const uint64_t adv = (uint64_t)(2 * k);
              umma2_f16_ss(tmem + buf * G2_TN, da + adv, db + adv, idesc, (kb | k) != 0);
            }
            umma2_commit_mc(&sb.empty[s]);
            if (kb == nkb - 1) umma2_commit_mc(&sb.acc_full[buf]);
          }
          __syncwarp();
        }
      }
    } else {
      // ===================== relay (peer CTA): my stage landed =====================
      uint32_t g = 0;
      for (int t = pair; t < ntiles; t += npairs) {
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const uint32_t s = g % G2_STAGES;
          mbar_wait_cl(&sb.full[s], (g / G2_STAGES) & 1);
          if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&sb.peer_full[s]), 0));
          __syncwarp();
        }
      }
    }
  } else {
    // ============================ epilogue (warps 2..5) ============================
    const int et = tid - 64;              // 0..127
    const int q = warp & 3;               // TMEM lane quadrant this warp may read
    const int rl = 32 * q + lane;         // accumulator lane = row inside this CTA's half
    const int ew = warp - 2;              // which group of input dims this warp reduces
    // (err == NULL, cscale = 1 / G2_LO_SCALE: the correction pass over the low part of C)
    const float cmax = __uint_as_float(*cmax_bits) * cscale;
    uint32_t it = 0;
    for (int t = pair; t < ntiles; t += npairs, ++it) {
      int rb, fb;
      g2_tile(t, RB, FB, sup, rb, fb);
      const uint32_t buf = it & 1;
      const int row0 = rb * G2_TM + 128 * (int)crank;       // first row of this CTA's half
      // tables for this tile (the previous tile's readers are past their last barrier)
      for (int j = et; j < G2_TN; j += 128) {
        const int fo = fb * G2_TN + j;
        const int th = feat_theta(fo);
        float mv = 0.0f;
        if (th < ktot) mv = m[feat_is_sin(fo) ? plan.col_sin[th] : plan.col_cos[th]];
        m_loc[j] = mv;
      }
      {
        const int th = fb * (G2_TN / 2) + et;
        amp_loc[et] = th < ktot ? plan.amp[th] : 0.0f;
        err_loc[et] = (err != nullptr && row0 + et < rows) ? err[row0 + et] : 0.0f;
      }
      for (int e = et; e < 128 * DPAD; e += 128) {
        const int r = e / DPAD, i = e - r * DPAD;
        xs[e] = (i < d && row0 + r < rows) ? X[(int64_t)(row0 + r) * d + i] : 0.0f;
      }
      mbar_wait_cl(&sb.acc_full[buf], (it >> 1) & 1);
      tc_fence_after_sync();
      named_bar_sync(1, 128);
      const float e_n = err_loc[rl];
      const uint32_t tacc = tmem + ((uint32_t)(32 * q) << 16) + buf * G2_TN;
      // Phi values of this row: tile-major images of k blocks 4 fb .. 4 fb + 3
      const uint8_t* prow = PhT + ((int64_t)rb * nkb + 4 * fb) * G2_IMG;
      const uint32_t prl = (uint32_t)(128 * (int)crank + rl);   // row inside the 256-row image
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {      // (bb, c): 32 frequencies per pass
        const int bb = ch >> 1, c = ch & 1;
        const int fcol = 128 * bb + 32 * c;            // first cos column (tile-local)
        float tcv[32], tsv[32];
        tmem_ld32_nowait(tacc + fcol, tcv);
        tmem_ld32_nowait(tacc + fcol + 64, tsv);
        tmem_ld_wait();
        if (ch == 3) {                      // all TMEM reads of this tile are done
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&sb.acc_empty[buf]), 0));
        }
        // Q for this thread's row, 32 frequencies, in registers (reusing tcv)
        const uint8_t* pc_img = prow + (int64_t)(2 * bb) * G2_IMG;       // cos k block
        const uint8_t* ps_img = pc_img + G2_IMG;                         // sin k block
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
          const uint32_t off = sw128_off(prl, (uint32_t)(4 * c + v4));
          const uint4 pc = *reinterpret_cast<const uint4*>(pc_img + off);
          const uint4 ps = *reinterpret_cast<const uint4*>(ps_img + off);
          const __half* hc = reinterpret_cast<const __half*>(&pc);
          const __half* hs = reinterpret_cast<const __half*>(&ps);
#pragma unroll
          for (int r8 = 0; r8 < 8; ++r8) {
            const int r = 8 * v4 + r8;
            const float Tc = e_n * m_loc[fcol + r] - tcv[r] * cmax;
            const float Ts = e_n * m_loc[fcol + 64 + r] - tsv[r] * cmax;
            const float a = amp_loc[64 * bb + 32 * c + r];
            tcv[r] = a * (-__half2float(hs[r8]) * Tc + __half2float(hc[r8]) * Ts);
          }
        }
        // 32 x 32 transpose inside the warp (5 butterfly stages of shuffles): lane =
        // row, register = frequency  ->  lane = frequency, register = row.  The
        // contraction over rows then needs only broadcast reads of X from shared
        // memory; staging Q through shared memory cost more bandwidth than the
        // tensor core's own operand reads.
#pragma unroll
        for (int b = 16; b >= 1; b >>= 1) {
          const bool up = (lane & b) != 0;
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            if (!(k & b)) {
              const float send = up ? tcv[k] : tcv[k | b];
              const float recv = __shfl_xor_sync(0xffffffffu, send, b);
              if (up) tcv[k] = recv;
              else tcv[k | b] = recv;
            }
          }
        }
        // acc[i] = sum over this warp's 32 rows of x[row, i] * Q[row, theta = lane]
        float acc[DPAD];
#pragma unroll
        for (int i = 0; i < DPAD; ++i) acc[i] = 0.0f;
        const float4* xq = reinterpret_cast<const float4*>(xs + (32 * q) * DPAD);
#pragma unroll
        for (int n = 0; n < 32; ++n) {
#pragma unroll
          for (int g4 = 0; g4 < IG; ++g4) {
            const float4 xv = xq[n * IG + g4];
            acc[4 * g4 + 0] = fmaf(xv.x, tcv[n], acc[4 * g4 + 0]);
            acc[4 * g4 + 1] = fmaf(xv.y, tcv[n], acc[4 * g4 + 1]);
            acc[4 * g4 + 2] = fmaf(xv.z, tcv[n], acc[4 * g4 + 2]);
            acc[4 * g4 + 3] = fmaf(xv.w, tcv[n], acc[4 * g4 + 3]);
          }
        }
        // cross-warp (row quadrant) reduction through shared memory, then one
        // float64 atomic per (dimension, frequency) of the pass
        float* rb_ = red + (ch & 1) * G2_RED + ew * (32 * 32);
#pragma unroll
        for (int i = 0; i < DPAD; ++i) rb_[i * 32 + lane] = acc[i];
        named_bar_sync(1, 128);
        const int th = fb * (G2_TN / 2) + 64 * bb + 32 * c + lane;
        if (th < ktot) {
          const float* r0_ = red + (ch & 1) * G2_RED;
#pragma unroll
          for (int j = 0; j < IG; ++j) {
            const int i = ew * IG + j;
            if (i < d) {
              const float v = r0_[i * 32 + lane] + r0_[1024 + i * 32 + lane] +
                              r0_[2048 + i * 32 + lane] + r0_[3072 + i * 32 + lane];
              atomicAdd(R + (int64_t)i * ktot + th, (double)v);
            }
          }
        }
        // The two reduction buffers alternate, so the barrier of the next pass
        // already orders its writes after these reads.
      }
      named_bar_sync(1, 128);   // tables and xs are rewritten by the next tile
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2cta(tmem, 512);
}

// ---- host ---------------------------------------------------------------------
// padded feature count: whole 256-feature output tiles (= 128 frequencies)
static int gp_dp(const rr_plan* pl) { return ((pl->ktot + 127) / 128) * 256; }
// reduction extent: trigonometric features + affine columns padded to a k block
static int gp_dk(const rr_plan* pl) { return gp_dp(pl) + ((pl->next + G2_KT - 1) / G2_KT) * G2_KT; }

int tc_gradpass_supported(const rr_plan* pl) {
  return (pl->d >= 1 && pl->d <= 32 && pl->ktot >= 1 && pl->next >= 0 &&
          pl->D == 2 * pl->ktot + pl->next) ? 1 : 0;
}

// Row blocks (of 256 rows) per chunk: as many as the scratch budget allows, then
// trimmed so that the tile count fills whole rounds of the persistent CTA pairs.
static int gp_chunk_blocks(const rr_plan* pl, int64_t N) {
  const int Dp = gp_dp(pl), FB = Dp / G2_TN;
  int rbmax = (int)(G2_SCRATCH_BYTES / ((int64_t)G2_TM * gp_dk(pl) * 2));
  if (rbmax < 1) rbmax = 1;
  const int64_t need = (N + G2_TM - 1) / G2_TM;
  if (need <= rbmax) return (int)need;
  const int npairs = sm_count() / 2;
  int best = rbmax;
  double beste = 0.0;
  for (int rb = rbmax; rb >= (rbmax + 1) / 2; --rb) {
    const int tiles = rb * FB;
    const double eff = (double)tiles / ((double)((tiles + npairs - 1) / npairs) * npairs);
    if (eff > beste + 1e-9) {
      beste = eff;
      best = rb;
    }
  }
  return best;
}

size_t tc_gradpass_workspace(const rr_plan* pl, int64_t N, bool split_c) {
  const int64_t Dp = gp_dp(pl), Dk = gp_dk(pl);
  return 2 * (align_up((size_t)gp_chunk_blocks(pl, N) * G2_TM * Dk * 2, 1024) + 1024) +
         (split_c ? 2 : 1) * (align_up((size_t)Dp * Dk * 2, 1024) + 1024 + 256) +
         align_up((size_t)N * 4, 256) + 8192;
}

template <int IG>
static int launch_gp2(const rr_plan* pl, const float* X, const float* err, int rows,
                      int RB, int FB, int nkb, const uint8_t* PhT, const uint8_t* BtT,
                      const float* m, const unsigned int* cmax, float cscale, double* R,
                      cudaStream_t st) {
  const size_t smem = (size_t)G2_STAGES * G2_STAGE_BYTES + 2 * G2_RED * 4 +
                      128 * 4 * IG * 4 + 1024;
  RR_CUDA_CHECK(cudaFuncSetAttribute(gp2_kernel<IG>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int npairs = sm_count() / 2;
  if (RB * FB < npairs) npairs = RB * FB;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * npairs);
  cfg.blockDim = dim3(G2_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const int sup = ((size_t)FB * G2_TN * nkb * G2_KT * 2 > G2_C_IMAGE_L2_BYTES) ? G2_SUPER_BIG : 1;
  RR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gp2_kernel<IG>, *pl, X, err, rows, RB, FB, sup, nkb, PhT,
                                   BtT, m, cmax, cscale, R));
  RR_LAUNCH_CHECK("gp2_kernel");
  return RR_OK;
}

static int launch_gp2_d(const rr_plan* pl, const float* X, const float* err, int rows, int RB,
                        int FB, int nkb, const uint8_t* PhT, const uint8_t* BtT, const float* m,
                        const unsigned int* cmax, float cscale, double* R, cudaStream_t st) {
  const int d = pl->d;
  if (d <= 4) return launch_gp2<1>(pl, X, err, rows, RB, FB, nkb, PhT, BtT, m, cmax, cscale, R, st);
  if (d <= 8) return launch_gp2<2>(pl, X, err, rows, RB, FB, nkb, PhT, BtT, m, cmax, cscale, R, st);
  if (d <= 16) return launch_gp2<4>(pl, X, err, rows, RB, FB, nkb, PhT, BtT, m, cmax, cscale, R, st);
  if (d <= 24) return launch_gp2<6>(pl, X, err, rows, RB, FB, nkb, PhT, BtT, m, cmax, cscale, R, st);
  return launch_gp2<8>(pl, X, err, rows, RB, FB, nkb, PhT, BtT, m, cmax, cscale, R, st);
}

// ---- gradient pass from a kept feature image ----------------------------------------
// Keeping pays while ONE persistent launch over all rows keeps its operand panels in
// L2: the CTA pairs of a long launch drift apart, and then the ~17 panels (Dk x 512
// bytes each) their tiles touch must fit as a whole.  Measured at 1.25e6 rows
// (gpurun_out/final_r02z.log): K = 512 kept 14.1 ms / regenerated 15.2 ms per
// evaluation, K = 2048 100 / 100-108, K = 4096 425 / 405, K = 8192 1823 / 1663 -- past
// ~6000 columns the chunked pass (which restarts its pairs in step every 320 MB of
// Phi) is the faster one, and the generator's extra output no longer pays.
constexpr int G2_KEEP_MAX_COLS = 6144;
int64_t kept_features_cols(const rr_plan* pl) { return gp_dk(pl); }
size_t kept_features_bytes(const rr_plan* pl, int64_t N) {
  if (!tc_gradpass_supported(pl) || !tc3_suffstats_supported(pl) || N <= 0) return 0;
  if (gp_dk(pl) > G2_KEEP_MAX_COLS) return 0;
  return (size_t)((N + G2_TM - 1) / G2_TM) * (size_t)gp_dk(pl) * 512;
}
size_t tc_gradpass_kept_workspace(const rr_plan* pl, int64_t N, bool split_c) {
  const int64_t Dp = gp_dp(pl), Dk = gp_dk(pl);
  return (split_c ? 2 : 1) * (align_up((size_t)Dp * Dk * 2, 1024) + 1024 + 256) +
         align_up((size_t)N * 4, 256) + align_up((size_t)Dk * 4, 256) + 8192;
}

// The value pass of the same evaluation left Phi behind (tc3_suffstats, kept != NULL):
// residuals in one streaming pass over the image, then ONE persistent GEMM launch over
// all rows -- no second evaluation of the feature map, no per-chunk launches.
int tc_gradpass_kept(const rr_plan* pl, const float* X, const float* y, int64_t N,
                     const float* m, const float* C, double* R, double* sqerr,
                     const void* kept, void* ws, size_t wsb, bool split_c, cudaStream_t st) {
  const int Dp = gp_dp(pl), Dk = gp_dk(pl), FB = Dp / G2_TN, nkb = Dk / G2_KT;
  const uint8_t* PhT = static_cast<const uint8_t*>(kept);
  if (N >= ((int64_t)1 << 31) - 256 || (reinterpret_cast<uintptr_t>(PhT) & 1023) != 0) {
    set_error("kept gradient pass: too many rows for one launch, or image not 1024-byte aligned");
    return RR_ERR_INVALID;
  }
  Workspace W(ws, wsb);
  uint8_t* BtT = W.take<uint8_t>(align_up((size_t)Dp * Dk * 2, 1024) + 1024);
  uint8_t* BtL = split_c ? W.take<uint8_t>(align_up((size_t)Dp * Dk * 2, 1024) + 1024) : nullptr;
  float* err = W.take<float>((size_t)N);
  float* mi = W.take<float>((size_t)Dk);
  unsigned int* cmax = W.take<unsigned int>(1);
  if (!BtT || (split_c && !BtL) || !err || !mi || !cmax) {
    set_error("kept gradient pass workspace too small (need %zu bytes)",
              tc_gradpass_kept_workspace(pl, N, split_c));
    return RR_ERR_WORKSPACE;
  }
  BtT = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(BtT) + 1023) & ~(uintptr_t)1023);
  if (BtL) BtL = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(BtL) + 1023) & ~(uintptr_t)1023);
  RR_CUDA_CHECK(cudaMemsetAsync(cmax, 0, sizeof(unsigned int), st));
  mint_kernel<<<(Dk + 255) / 256, 256, 0, st>>>(*pl, m, Dp, Dk, mi);
  RR_LAUNCH_CHECK("mint_kernel");
  const int RB = (int)((N + G2_TM - 1) / G2_TM);
  {
    const size_t smem = (size_t)Dk * sizeof(float);
    if (smem <= 160 * 1024) {
      if (smem > 48 * 1024)
        RR_CUDA_CHECK(cudaFuncSetAttribute(fit16_kernel<true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      fit16_kernel<true><<<RB, 256, smem, st>>>(PhT, nkb, mi, y, N, err, sqerr);
    } else {
      fit16_kernel<false><<<RB, 256, 0, st>>>(PhT, nkb, mi, y, N, err, sqerr);
    }
    RR_LAUNCH_CHECK("fit16_kernel");
  }
  absmax_kernel<<<sm_count() * 4, 256, 0, st>>>(C, (int64_t)pl->D * pl->D, cmax);
  RR_LAUNCH_CHECK("absmax_kernel");
  {
    dim3 grid((Dk + 255) / 256, Dp);
    prep_c_kernel<<<grid, 256, 0, st>>>(*pl, C, Dp, Dk, cmax, BtT, 0);
    RR_LAUNCH_CHECK("prep_c_kernel");
    if (split_c) {
      prep_c_kernel<<<grid, 256, 0, st>>>(*pl, C, Dp, Dk, cmax, BtL, 1);
      RR_LAUNCH_CHECK("prep_c_kernel");
    }
  }
  int rc = launch_gp2_d(pl, X, err, (int)N, RB, FB, nkb, PhT, BtT, m, cmax, 1.0f, R, st);
  if (rc == RR_OK && split_c)
    rc = launch_gp2_d(pl, X, nullptr, (int)N, RB, FB, nkb, PhT, BtL, m, cmax, 1.0f / G2_LO_SCALE, R, st);
  return rc;
}

// Residuals only (value-only evaluations; any feature plan):
// sqerr += sum (y - Phi m)^2, optionally the N residuals.  fbuf: N floats of
// scratch (overwritten).
int phi_residual(const rr_plan* pl, const float* X, const float* y, int64_t N,
                 const float* m, float* err, double* sqerr, float* fbuf,
                 cudaStream_t st) {
  RR_CUDA_CHECK(cudaMemsetAsync(fbuf, 0, (size_t)N * sizeof(float), st));
  const int64_t CH = 1 << 20;
  for (int64_t s = 0; s < N; s += CH) {
    const int rows = (int)((N - s) < CH ? (N - s) : CH);
    int rc = launch_phi_fit(pl, X + s * pl->d, rows, rows, 0, 0, m, nullptr, fbuf + s, st);
    if (rc) return rc;
  }
  resid_finish_kernel<<<sm_count() * 4, 256, 0, st>>>(y, fbuf, N, err, sqerr);
  RR_LAUNCH_CHECK("resid_finish_kernel");
  return RR_OK;
}

// The Phi chunk of row chunk c+1 is generated on the context's helper stream while
// the GEMM of chunk c runs on the caller's stream (two scratch chunks, ping-pong).
// A phi_fit block needs 9 K registers and 1.6 KB of shared memory, so one fits on
// every SM next to the resident GEMM CTA; and with nothing queued between them,
// consecutive GEMM launches overlap the epilogue tail of one with the pipeline
// fill of the next.  Without a context everything runs on the caller's stream.
// error path: whatever the helper stream still has queued is joined back into the
// caller's stream before the status is returned
static int join_helper(int rc, bool overlap, cudaStream_t sp, cudaStream_t st, cudaEvent_t ev) {
  if (overlap && cudaEventRecord(ev, sp) == cudaSuccess) cudaStreamWaitEvent(st, ev, 0);
  return rc;
}

int tc_gradpass(const rr_plan* pl, const float* X, const float* y, int64_t N,
                const float* m, const float* C, double* R, double* sqerr, void* ws,
                size_t wsb, rr_context* ctx, bool split_c, cudaStream_t st) {
  const int Dp = gp_dp(pl), Dk = gp_dk(pl), FB = Dp / G2_TN, nkb = Dk / G2_KT;
  const int RBc = gp_chunk_blocks(pl, N);
  const int64_t RC = (int64_t)RBc * G2_TM;
  Workspace W(ws, wsb);
  uint8_t* PhTb[2];
  PhTb[0] = W.take<uint8_t>(align_up((size_t)RC * Dk * 2, 1024) + 1024);
  PhTb[1] = W.take<uint8_t>(align_up((size_t)RC * Dk * 2, 1024) + 1024);
  uint8_t* BtT = W.take<uint8_t>(align_up((size_t)Dp * Dk * 2, 1024) + 1024);
  uint8_t* BtL = split_c ? W.take<uint8_t>(align_up((size_t)Dp * Dk * 2, 1024) + 1024) : nullptr;
  float* err = W.take<float>((size_t)N);      // fitted values, then residuals
  unsigned int* cmax = W.take<unsigned int>(1);
  cudaStream_t sp = st;                                                   // Phi stream
  cudaEvent_t ev_fork = nullptr, ev_phi[2] = {nullptr, nullptr}, ev_gemm[2] = {nullptr, nullptr};
  if (ctx_aux(ctx, &sp, &ev_fork, ev_phi, ev_gemm) != RR_OK) return RR_ERR_CUDA;
  const bool overlap = sp != st;
  if (!PhTb[0] || !PhTb[1] || !BtT || (split_c && !BtL) || !err || !cmax) {
    set_error("tcgen05 gradpass workspace too small (need %zu bytes)",
              tc_gradpass_workspace(pl, N, split_c));
    return RR_ERR_WORKSPACE;
  }
  if (BtL) BtL = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(BtL) + 1023) & ~(uintptr_t)1023);
  // the bulk copies need 16-byte aligned images
  for (int i = 0; i < 2; ++i)
    PhTb[i] = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(PhTb[i]) + 1023) & ~(uintptr_t)1023);
  BtT = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(BtT) + 1023) & ~(uintptr_t)1023);
  RR_CUDA_CHECK(cudaMemsetAsync(cmax, 0, sizeof(unsigned int), st));
  RR_CUDA_CHECK(cudaMemsetAsync(err, 0, (size_t)N * sizeof(float), st));
  if (overlap) {
    RR_CUDA_CHECK(cudaEventRecord(ev_fork, st));          // inputs (m, y, X) are ready
    RR_CUDA_CHECK(cudaStreamWaitEvent(sp, ev_fork, 0));
  }
  absmax_kernel<<<sm_count() * 4, 256, 0, st>>>(C, (int64_t)pl->D * pl->D, cmax);
  RR_LAUNCH_CHECK("absmax_kernel");
  {
    dim3 grid((Dk + 255) / 256, Dp);
    prep_c_kernel<<<grid, 256, 0, st>>>(*pl, C, Dp, Dk, cmax, BtT, 0);
    RR_LAUNCH_CHECK("prep_c_kernel");
    if (split_c) {
      prep_c_kernel<<<grid, 256, 0, st>>>(*pl, C, Dp, Dk, cmax, BtL, 1);
      RR_LAUNCH_CHECK("prep_c_kernel");
    }
  }
  const int d = pl->d;
  int c = 0;
  for (int64_t s = 0; s < N; s += RC, ++c) {
    const int rows = (int)((N - s) < RC ? (N - s) : RC);
    const int RB = (rows + G2_TM - 1) / G2_TM;
    const int rows_pad = RB * G2_TM;
    const int buf = c & 1;
    uint8_t* PhT = PhTb[buf];
    // helper stream: Phi chunk c (after the GEMM of chunk c-2 released the buffer)
    if (overlap && c >= 2) RR_CUDA_CHECK(cudaStreamWaitEvent(sp, ev_gemm[buf], 0));
    int rc = launch_phi_fit(pl, X + s * d, rows, rows_pad, Dp, Dk, m, PhT, err + s, sp);
    if (rc) return join_helper(rc, overlap, sp, st, ev_fork);
    // fitted values -> residuals (in place) + their sum of squares
    resid_finish_kernel<<<(rows + 1023) / 1024, 256, 0, sp>>>(y + s, err + s, rows, err + s,
                                                             sqerr);
    RR_LAUNCH_CHECK("resid_finish_kernel");
    if (overlap) {
      RR_CUDA_CHECK(cudaEventRecord(ev_phi[buf], sp));
      // caller's stream: GEMM + epilogue of chunk c
      RR_CUDA_CHECK(cudaStreamWaitEvent(st, ev_phi[buf], 0));
    }
    const float* Xc = X + s * d;
    const float* ec = err + s;
    rc = launch_gp2_d(pl, Xc, ec, rows, RB, FB, nkb, PhT, BtT, m, cmax, 1.0f, R, st);
    if (rc == RR_OK && split_c)
      rc = launch_gp2_d(pl, Xc, nullptr, rows, RB, FB, nkb, PhT, BtL, m, cmax, 1.0f / G2_LO_SCALE, R, st);
    if (rc) return join_helper(rc, overlap, sp, st, ev_fork);
    if (overlap) RR_CUDA_CHECK(cudaEventRecord(ev_gemm[buf], st));
  }
  return RR_OK;
}

}  // namespace rr
