// Value pass of the SLM log marginal likelihood in exact 24-bit fixed point on
// the int8 tensor cores (tcgen05.mma.kind::i8):
//     G += Phi^T Phi,   p += Phi^T y,   Phi = amp * [cos | sin](2 pi X Wt) (+ affine columns).
//
// Every feature value v in [-1, 1] (a cosine, a sine, x / max|x| of an affine
// column, y / max|y|) is rounded ONCE to the integer I = rint(v * S),
// S = 2^23 - 2^16 - 2^9, and written as three balanced base-256 digits
//     I = d0 * 65536 + d1 * 256 + d2,    d0, d1, d2 in [-128, 127]
// i.e. three int8 planes.  The Gram matrix of the integers is then
//     sum_n I_a I_b = 2^32 S00 + 2^24 (S01 + S10) + 2^16 (S11 + S02 + S20) + [2^8 (S12 + S21) + S22]
// with Sij = sum_n di_a dj_b: six int8 x int8 -> int32 tensor-core products
// into three accumulators; the bracket (relative weight 2^-24) is dropped.
// Integer accumulation is EXACT (|ACC| < 2^31 for 32768-row chains), so the
// result does not depend on tile order, chunking or the number of ranks, and
// there is no accumulation bias to engineer around (the fp16 path needed a
// fixed-point head / remainder split for that).  Each chain is drained as
//     int64 v = ACC0 * 65536 + ACC1 * 256 + ACC2     (exact)
// and added (RED.F64, also exact: |sum| < 2^53) to a float64 image T of the
// upper triangle; a finalize kernel applies 2^16 / S^2 and the amplitudes.
//
// Phi itself is produced once per (row, frequency) by t3_digits_kernel -- tf32
// tensor-core projection on split operands (three products), exact range
// reduction in turns, polynomial sin/cos -- as a tile-major int8 image that the
// GEMM streams with bulk copies.  In the previous design
// the trigonometric generators sat inside the tensor-core kernel and
// re-evaluated every feature for each of the ~18 output tiles it takes part
// in; at 2 x 10^9 (row, frequency) pairs per pass that made the generators,
// not the tensor pipe, the bound.
//
// y rides along on the B side as THREE feature columns D, D+1, D+2 holding its
// digits one at a time, (y0, 0, 0), (y1, 0, 0), (y2, 0, 0).  With b1 = b2 = 0 none
// of a column's products is dropped, so sum_n I_a y_k is exact for each digit and
//     Phi^T y = V0 * 65536 + V1 * 256 + V2
// carries the FULL 24 x 24-bit products.  (With y as one ordinary column the
// dropped bracket of p is not a Gram-consistent perturbation: it moved the
// posterior mean of a 1-D, 256-frequency, N = 1000 problem by 1.5e-4; the columns
// are free, they sit in the zero padding of the last B tile.)
//
// Replaces: revrand/slm.py:145-146, :157 and basis_functions.py:859-864,
// 1622-1627 (BasisCat.transform), 468-485 (LinearBasis), 415-432 (BiasBasis).
#include <stdlib.h>
#include <string.h>

#include "rr_common.cuh"
#include "rr_tc.cuh"

namespace rr {

using namespace tc;

constexpr int S3_TM = 256;                        // A-side features per tile (128 per CTA)
constexpr int S3_TN = 160;                        // B-side features per tile (80 per CTA)
constexpr int S3_KB = 64;                         // rows of X per K block (one 64-byte line)
constexpr int S3_STAGES = 5;
constexpr int S3_A_PLANE = (S3_TM / 2) * S3_KB;   // 8192: one digit plane of a CTA's A half
constexpr int S3_B_PLANE = (S3_TN / 2) * S3_KB;   // 5120
constexpr int S3_STAGE_BYTES = 3 * (S3_A_PLANE + S3_B_PLANE);   // 39936
constexpr int S3_THREADS = 6 * 32;                // producer, MMA / relay, 4 epilogue warps
constexpr int S3_CHAIN = 32768;                   // rows per exact int32 accumulation chain
constexpr int S3_CHAIN_KB = S3_CHAIN / S3_KB;     // 512
constexpr float S3_SCALE = 8322560.0f;            // S = 2^23 - 2^16 - 2^9: |d0| <= 127
constexpr int S3_GROUP_CHAINS = 8;                // chains per generator / GEMM launch
constexpr int S3_FIRST_CHAINS = 1;                // ... of the first launch: its generator
                                                  // pass is the only one not hidden behind
                                                  // tensor-core work
constexpr size_t S3_GROUP_BYTES_MAX = (size_t)4 << 30;   // digit image budget per buffer

static_assert(S3_A_PLANE % 512 == 0 && S3_B_PLANE % 512 == 0, "64B-swizzle tiles are 512-byte aligned");
static_assert(3 * S3_TN <= 512, "three int32 accumulators in TMEM");
// worst case |ACC2| = 3 * 128 * 128 * S3_CHAIN
static_assert(3ll * 128 * 128 * S3_CHAIN < (1ll << 31), "int32 accumulators must not overflow");

// first column panel that reaches the upper triangle of row panel ib
__host__ __device__ __forceinline__ int s3_jmin(int ib) { return (S3_TM * ib) / S3_TN; }

// ---- scales ---------------------------------------------------------------------
// scales[i] = max |X[:, i]| (i < d, only when need_x), scales[d] = max |y|, as
// float bit patterns (non-negative floats order like unsigned integers).
__global__ void __launch_bounds__(256)
t3_scales_kernel(const float* __restrict__ X, const float* __restrict__ y, int64_t N, int d,
                 int need_x, unsigned int* __restrict__ scales) {
  extern __shared__ unsigned int smax[];   // d + 1
  for (int i = threadIdx.x; i <= d; i += blockDim.x) smax[i] = 0u;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (y) {
    float my = 0.0f;
    for (int64_t i = t0; i < N; i += stride) my = fmaxf(my, fabsf(y[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my = fmaxf(my, __shfl_xor_sync(0xffffffffu, my, o));
    if ((threadIdx.x & 31) == 0) atomicMax(&smax[d], __float_as_uint(my));
  }
  if (need_x) {
    // a thread walks one column: stride is a multiple of d away from its start
    const int64_t total = N * d;
    const int64_t step = stride * d;           // keeps (e % d) fixed per thread
    for (int64_t c0 = t0; c0 < stride * d && c0 < total; c0 += stride) {
      float mx = 0.0f;
      for (int64_t e = c0; e < total; e += step) mx = fmaxf(mx, fabsf(X[e]));
      atomicMax(&smax[(int)(c0 % d)], __float_as_uint(mx));
    }
  }
  __syncthreads();
  // (Row shards of one data set see different maxima: a sharded caller passes the
  // job-wide maxima in rr_plan.col_scale instead, which makes the ranks' quantised
  // statistics add up to the one-process result bit for bit.)
  for (int i = threadIdx.x; i <= d; i += blockDim.x)
    if (smax[i]) atomicMax(&scales[i], smax[i]);
}

// camp[f] = real amplitude of integer feature f (f <= D; f == D is the y column).
__global__ void t3_colamp_kernel(rr_plan plan, const unsigned int* __restrict__ scales,
                                 float* __restrict__ camp) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < plan.ktot) {
    const float a = plan.amp[t];
    if (plan.col_cos[t] >= 0) camp[plan.col_cos[t]] = a;
    if (plan.col_sin[t] >= 0) camp[plan.col_sin[t]] = a;
  }
  if (t < plan.next) {
    const int src = plan.ext_src[t];
    const int pw = plan.ext_pow ? plan.ext_pow[t] : 1;
    camp[plan.ext_col[t]] = src >= 0 ? ipowf(__uint_as_float(scales[src]), pw) : plan.ext_val[t];
  }
  if (t == 0) camp[plan.D] = __uint_as_float(scales[plan.d]);
}

// ---- digit image ------------------------------------------------------------------
// img[kb][plane][f][64 bytes]: K block kb = 64 consecutive rows, plane 0..2 = d0,
// d1, d2, feature f in Phi's own column order (f == D .. D+2: the digits of y,
// f > D + 2: zero padding up to Fp).  Inside a 64-byte line the 16-byte chunk of rows 16c .. 16c+15 sits at
// chunk position c ^ ((f >> 1) & 3): the SWIZZLE_64B image of a tile whose first
// feature is a multiple of 8, so ANY run of features (8-aligned) of one plane and
// K block is one contiguous, ready-to-use operand tile.
__device__ __forceinline__ int t3_quant(float v) {
  // (I + 0x8080) ^ 0x8080: byte 0 = d2, byte 1 = d1, byte 2 = d0 (two's complement)
  return (__float2int_rn(v * S3_SCALE) + 0x8080) ^ 0x8080;
}
// gather byte `b` of four quantised values into one word per plane
__device__ __forceinline__ void t3_pack4(int k0, int k1, int k2, int k3, uint32_t& p0,
                                         uint32_t& p1, uint32_t& p2) {
  const uint32_t t01 = __byte_perm((uint32_t)k0, (uint32_t)k1, 0x5140);
  const uint32_t t23 = __byte_perm((uint32_t)k2, (uint32_t)k3, 0x5140);
  const uint32_t u01 = __byte_perm((uint32_t)k0, (uint32_t)k1, 0x0062);
  const uint32_t u23 = __byte_perm((uint32_t)k2, (uint32_t)k3, 0x0062);
  p2 = __byte_perm(t01, t23, 0x5410);
  p1 = __byte_perm(t01, t23, 0x7632);
  p0 = __byte_perm(u01, u23, 0x5410);
}
__device__ __forceinline__ void t3_store_planes(uint8_t* __restrict__ img, int64_t kb, int Fp,
                                                int f, int chunk, const uint4& d0,
                                                const uint4& d1, const uint4& d2) {
  uint8_t* base = img + (((int64_t)kb * 3) * Fp + f) * 64 + (((uint32_t)chunk ^ (((uint32_t)f >> 1) & 3u)) << 4);
  const int64_t plane = (int64_t)Fp * 64;
  *reinterpret_cast<uint4*>(base) = d0;
  *reinterpret_cast<uint4*>(base + plane) = d1;
  *reinterpret_cast<uint4*>(base + 2 * plane) = d2;
}

constexpr int T3_FREQS = 128;    // frequencies per trigonometric block (16 per warp)
constexpr int T3_OTHER = 64;     // affine / y / padding features per "other" block

// Row <-> byte mapping inside a K block (any fixed bijection works: the Gram sum
// over rows does not care about their order, as long as EVERY feature uses the
// same one).  Byte b of chunk q holds row 8 (b >> 1) + 2 q + (b & 1): exactly the
// rows one lane of an m16n8k8 accumulator fragment owns across the eight row
// tiles, so a generator thread packs its sixteen values without any shuffle.
__device__ __forceinline__ int t3_row_of(int q, int b) { return 8 * (b >> 1) + 2 * q + (b & 1); }

// round an fp32 value to tf32 (10 explicit mantissa bits), ties away from zero
__device__ __forceinline__ float t3_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void t3_mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                            uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, "
      "{%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// sin and cos of 2 pi u (u in turns), |error| ~ 1e-7: exact reduction to
// v in [-1/8, 1/8] turns around the nearest quarter turn j, degree-9 / degree-8
// Taylor polynomials in 2 pi v (truncation 2e-9 / 2.4e-8 at |v| = 1/8), then the
// rotation by j quarter turns as a swap and two sign flips.  ~25 instructions on
// the FMA / ALU pipes (sincospif: ~100).
__device__ __forceinline__ void t3_sincos_turns(float u, float& s, float& c) {
  const float MAGIC = 12582912.0f;                 // 1.5 * 2^23
  const float r = u - ((u + MAGIC) - MAGIC);       // u - rint(u): exact, in [-0.5, 0.5]
  const float t = fmaf(r, 4.0f, MAGIC);            // low mantissa bits = rint(4 r)
  const int j = (int)__float_as_uint(t);           // two's complement in the low bits
  const float v = fmaf(t - MAGIC, -0.25f, r);      // exact, in [-1/8, 1/8]
  const float z = v * v;
  float ps = fmaf(z, 42.058693944897654f, -76.70585975306136f);     // (2pi)^9/9!, -(2pi)^7/7!
  ps = fmaf(ps, z, 81.60524927607504f);                             // (2pi)^5/5!
  ps = fmaf(ps, z, -41.341702240399755f);                           // -(2pi)^3/3!
  ps = fmaf(ps, z, 6.283185307179586f);
  const float sn = ps * v;
  float pc = fmaf(z, 60.24464137187666f, -85.45681720669373f);      // (2pi)^8/8!, -(2pi)^6/6!
  pc = fmaf(pc, z, 64.93939402266829f);                             // (2pi)^4/4!
  pc = fmaf(pc, z, -19.739208802178716f);                           // -(2pi)^2/2!
  const float cs = fmaf(pc, z, 1.0f);
  // rotate (cs, sn) by j quarter turns: j = 0: (cs, sn), 1: (-sn, cs), 2: (-cs, -sn), 3: (sn, -cs)
  const bool odd = (j & 1) != 0;
  const uint32_t cb = __float_as_uint(odd ? sn : cs), sb = __float_as_uint(odd ? cs : sn);
  c = __uint_as_float(cb ^ (((uint32_t)(j + 1) << 30) & 0x80000000u));
  s = __uint_as_float(sb ^ (((uint32_t)j << 30) & 0x80000000u));
}

// sixteen values of one feature (bytes 0..15 of one chunk) -> its three digit planes
__device__ __forceinline__ void t3_emit(uint8_t* __restrict__ img, int64_t kb, int Fp, int f,
                                        int chunk, const int (&k)[16]) {
  uint4 q0, q1, q2;
  uint32_t* w0 = reinterpret_cast<uint32_t*>(&q0);
  uint32_t* w1 = reinterpret_cast<uint32_t*>(&q1);
  uint32_t* w2 = reinterpret_cast<uint32_t*>(&q2);
#pragma unroll
  for (int g = 0; g < 4; ++g)
    t3_pack4(k[4 * g], k[4 * g + 1], k[4 * g + 2], k[4 * g + 3], w0[g], w1[g], w2[g]);
  t3_store_planes(img, kb, Fp, f, chunk, q0, q1, q2);
}

// Digit image of one K block (64 rows).  grid.y counts the trigonometric blocks
// (128 frequencies, 16 per warp) first, then the blocks of "other" features:
// affine columns, the three y digit columns and the zero padding (64 per block, thread =
// feature x chunk).
//
// Trigonometric block: the projection u = X Wt runs on the tensor cores
// (mma.sync m16n8k8 tf32, frequencies = M, rows = N, input dimensions = K) on
// operands split into tf32 parts.  EXACT = false (the default): two parts
// (11 + 11 bits) and the three products hi*hi + lo*hi + hi*lo, i.e. every product
// to ~2^-21 |x w|: a phase error of a few 1e-6 rad per unit of |u|, which moves
// the config-2 posterior by less than the digit rounding does (DESIGN.md section 4:
// the same 2.0e-5 worst case as an fp32 FMA chain) and costs 3.5 ms less per
// value pass than EXACT = true (compile with -DRR_T3_EXACT_PROJECTION=1): an exact
// three-part split (11 + 11 + 2 bits) and the six products down to 2^-22 |x w|,
// as accurate as the fp32 FMA chain.  A lane ends up with two frequencies x
// sixteen rows, i.e. one 16-byte chunk of each of the twelve digit lines it then
// fills.
#ifndef RR_T3_EXACT_PROJECTION
#define RR_T3_EXACT_PROJECTION 0
#endif
constexpr bool T3_EXACT = RR_T3_EXACT_PROJECTION != 0;
constexpr int T3_PARTS = T3_EXACT ? 3 : 2;

//
// KEEP: the block also leaves its values behind as fp16 in the tile-major operand
// image of the gradient pass (rr_tc_gradpass.cu: [row block of 256][k block of 64
// internal features] = one 32 KB SWIZZLE_128B image; internal order = blocks of
// [64 cos | 64 sin] per 64 frequencies, then the affine columns), so that the
// gradient pass of the same evaluation does not have to evaluate the feature map a
// second time.  A trigonometric block owns 64 rows x 4 k blocks of that image.  It
// stages first its cosines, then its sines (two k blocks = 16 KB each time) in
// shared memory in their final byte order and writes them out as contiguous 8 KB
// runs.  The staging area ALIASES the input slab (dead after the projection): with
// 16 KB of shared memory the block still fits next to a resident GEMM CTA (197 KB of
// the SM's 228 KB), which is what lets generation overlap the tensor-core work.
constexpr int T3_KEEP_BYTES = 2 * S3_KB * 128;    // staging: 2 k blocks x 64 rows x 128 B

template <bool KEEP>
__global__ void __launch_bounds__(256, 2)
t3_digits_kernel(rr_plan plan, const float* __restrict__ X, const float* __restrict__ y,
                 int64_t rows, int Fp, int gy_trig, const unsigned int* __restrict__ scales,
                 uint8_t* __restrict__ img, uint8_t* __restrict__ phi16, int nkb16) {
  extern __shared__ float xs[];            // [T3_PARTS][64][stride]: tf32 parts of the slab
  const int d = plan.d, ktot = plan.ktot;
  const int kp = (d + 7) & ~7;             // input dimensions padded to whole k-steps
  const int stride = kp + 4;               // (stride / 4 odd: fragment loads hit 32 banks)
  float* xh = xs;
  float* xm = xs + S3_KB * stride;
  float* xl = xs + (T3_PARTS - 1) * S3_KB * stride;    // EXACT only (aliases xm otherwise)
  uint8_t* stage = reinterpret_cast<uint8_t*>(xs);      // KEEP: aliases the slab
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t kb = blockIdx.x;
  const int64_t n0 = kb * S3_KB;
  const bool trig = (int)blockIdx.y < gy_trig;
  for (int e = tid; e < S3_KB * kp; e += 256) {
    const int r = e / kp, i = e - r * kp;
    const float x = (i < d && n0 + r < rows) ? X[(n0 + r) * d + i] : 0.0f;
    if (trig) {
      const float h = t3_tf32(x);
      const float m = t3_tf32(x - h);
      xh[r * stride + i] = h;
      xm[r * stride + i] = m;
      if (T3_EXACT) xl[r * stride + i] = (x - h) - m;    // <= 3 significant bits: exact in tf32
    } else {
      xh[r * stride + i] = x;
    }
  }
  __syncthreads();
  if (!trig) {
    // ---- affine columns, y, padding -------------------------------------------
    const int fl = tid >> 2, q = tid & 3;
    const int e = ((int)blockIdx.y - gy_trig) * T3_OTHER + fl;
    if (KEEP && e < (nkb16 - 4 * gy_trig) * 64) {
      // internal feature 256 gy_trig + e of the kept image: the raw column value (zero
      // in the padding up to a whole k block)
      const int fc = 256 * gy_trig + e;
      uint8_t* line = phi16 + ((int64_t)(n0 >> 8) * nkb16 + (fc >> 6)) * 32768 +
                      (int64_t)(n0 & 255) * 128 + (fc & 7) * 2;
      const uint32_t c4 = (uint32_t)((fc & 63) >> 3);
#pragma unroll
      for (int b = 0; b < 16; ++b) {
        const int r = t3_row_of(q, b);
        const float v = (e < plan.next && n0 + r < rows) ? ext_value(plan, e, xh + r * stride, 1) : 0.0f;
        *reinterpret_cast<__half*>(line + r * 128 + ((c4 ^ (uint32_t)(r & 7)) << 4)) = __float2half_rn(v);
      }
    }
    int f;
    int kq[16];
    if (e < plan.next) {
      f = plan.ext_col[e];
      const int src = plan.ext_src[e];
      const int pw = plan.ext_pow ? plan.ext_pow[e] : 1;
      float inv = 0.0f;
      if (src >= 0) {
        const float sc = __uint_as_float(scales[src]);
        inv = sc > 0.0f ? 1.0f / sc : 0.0f;
      }
#pragma unroll
      for (int b = 0; b < 16; ++b) {
        const int r = t3_row_of(q, b);
        // (x / max|x|)^p in [-1, 1]; the amplitude max|x|^p is applied at the end
        float v = src >= 0 ? ipowf(xh[r * stride + src] * inv, pw) : 1.0f;
        v = fminf(1.0f, fmaxf(-1.0f, v));
        kq[b] = t3_quant((n0 + r < rows) ? v : 0.0f);
      }
    } else if (e < plan.next + 3) {
      // y digit (e - next) alone in plane 0 of column D + (e - next)
      const int dig = e - plan.next;
      f = plan.D + dig;
      const float sc = __uint_as_float(scales[d]);
      const float inv = (y != nullptr && sc > 0.0f) ? 1.0f / sc : 0.0f;
#pragma unroll
      for (int b = 0; b < 16; ++b) {
        const int r = t3_row_of(q, b);
        const float v = (y != nullptr && n0 + r < rows) ? y[n0 + r] * inv : 0.0f;
        kq[b] = t3_quant(fminf(1.0f, fmaxf(-1.0f, v)));
      }
      uint4 q0, q1, q2;
      uint32_t* w0 = reinterpret_cast<uint32_t*>(&q0);
      uint32_t* w1 = reinterpret_cast<uint32_t*>(&q1);
      uint32_t* w2 = reinterpret_cast<uint32_t*>(&q2);
#pragma unroll
      for (int g = 0; g < 4; ++g)
        t3_pack4(kq[4 * g], kq[4 * g + 1], kq[4 * g + 2], kq[4 * g + 3], w0[g], w1[g], w2[g]);
      const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
      t3_store_planes(img, kb, Fp, f, q, dig == 0 ? q0 : (dig == 1 ? q1 : q2), zero, zero);
      return;
    } else {
      f = plan.D + (e - plan.next);
      if (f >= Fp) return;
#pragma unroll
      for (int b = 0; b < 16; ++b) kq[b] = 0;
    }
    t3_emit(img, kb, Fp, f, q, kq);
    return;
  }
  // ---- trigonometric features ------------------------------------------------------
  const int g = lane >> 2, q = lane & 3;
  const int f0 = (int)blockIdx.y * T3_FREQS + 16 * warp;     // first frequency of this warp
  if (!KEEP && f0 >= ktot) return;
  const int fA = f0 + g, fB = f0 + g + 8;
  float acc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[nt][j] = 0.0f;
  for (int ks = 0; ks < kp && f0 < ktot; ks += 8) {
    // A fragment (frequencies x input dimensions): a0 (g, q), a1 (g+8, q), a2 (g, q+4), a3 (g+8, q+4)
    uint32_t ah[4], am[4], al[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int fi = (j & 1) ? fB : fA;
      const int ki = ks + q + ((j & 2) ? 4 : 0);
      const float w = (fi < ktot && ki < d) ? __ldg(plan.Wt + (int64_t)ki * ktot + fi) : 0.0f;
      const float h = t3_tf32(w);
      const float m = t3_tf32(w - h);
      ah[j] = __float_as_uint(h);
      am[j] = __float_as_uint(m);
      al[j] = __float_as_uint((w - h) - m);
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      // B fragment (input dimensions x rows): b0 (k = q, n = g), b1 (k = q + 4, n = g)
      const int o = (8 * nt + g) * stride + ks + q;
      const uint32_t bh0 = __float_as_uint(xh[o]), bh1 = __float_as_uint(xh[o + 4]);
      const uint32_t bm0 = __float_as_uint(xm[o]), bm1 = __float_as_uint(xm[o + 4]);
      if (T3_EXACT) {
        // x w = (xh + xm + xl)(wh + wm + wl): every product down to 2^-22 |x w|, small first
        const uint32_t bl0 = __float_as_uint(xl[o]), bl1 = __float_as_uint(xl[o + 4]);
        t3_mma_tf32(acc[nt], am, bm0, bm1);
        t3_mma_tf32(acc[nt], al, bh0, bh1);
        t3_mma_tf32(acc[nt], ah, bl0, bl1);
      }
      t3_mma_tf32(acc[nt], am, bh0, bh1);
      t3_mma_tf32(acc[nt], ah, bm0, bm1);
      t3_mma_tf32(acc[nt], ah, bh0, bh1);
    }
  }
  if (KEEP) __syncthreads();      // every warp is done with the slab: staging may overwrite it
  __half2 skeep[2][8];            // KEEP: this lane's sines, until the cosines have left
  // accumulator fragment: c0 (g, 2q), c1 (g, 2q+1), c2 (g+8, 2q), c3 (g+8, 2q+1) of row tile nt
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int k = h ? fB : fA;
    const bool valid = k < ktot;
    if (!KEEP && !valid) continue;
    int kc[16], ksn[16];
    // KEEP: local frequency fl -> k block pair fl >> 6, 16-byte chunk (fl & 63) >> 3,
    // element g inside the chunk (16 warp + 8 h is a multiple of 8).  Every (row,
    // frequency) slot of the staging area has exactly one owner, valid or not, so
    // frequencies past ktot and rows past the end are written as zeros.
    const int fl = 16 * warp + 8 * h + g;
    uint8_t* kcos = stage + (fl >> 6) * (S3_KB * 128) + 2 * g;
    const uint32_t kchunk = (uint32_t)((fl & 63) >> 3);
    float sprev = 0.0f;
#pragma unroll
    for (int b = 0; b < 16; ++b) {
      float sv = 0.0f, cv = 0.0f;
      const int r = t3_row_of(q, b);
      const bool live = valid && n0 + r < rows;
      if (live) t3_sincos_turns(acc[b >> 1][2 * h + (b & 1)], sv, cv);
      kc[b] = t3_quant(cv);
      ksn[b] = t3_quant(sv);
      if (KEEP) {
        // (q differs -> r & 7 differs -> four distinct chunks: conflict-free)
        *reinterpret_cast<__half*>(kcos + r * 128 + ((kchunk ^ (uint32_t)(r & 7)) << 4)) =
            __float2half_rn(cv);
        if (b & 1) skeep[h][b >> 1] = __floats2half2_rn(sprev, sv);
        sprev = sv;
      }
    }
    if (valid) {
      const int fc = plan.col_cos[k], fs = plan.col_sin[k];
      if (fc >= 0) t3_emit(img, kb, Fp, fc, q, kc);
      if (fs >= 0) t3_emit(img, kb, Fp, fs, q, ksn);
    }
  }
  if (KEEP) {
    // rows n0 .. n0+63 of k blocks 4 y (cos, frequencies 0..63 of the block), 4 y + 1
    // (their sines), 4 y + 2, 4 y + 3 (frequencies 64..127): one contiguous 8 KB run each
    uint8_t* dst = phi16 + ((int64_t)(n0 >> 8) * nkb16 + 4 * (int)blockIdx.y) * 32768 +
                   (int64_t)(n0 & 255) * 128;
    const uint4* src = reinterpret_cast<const uint4*>(stage);
#pragma unroll
    for (int part = 0; part < 2; ++part) {       // 0: cosines, 1: sines
      if (part == 1) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int fl = 16 * warp + 8 * h + g;
          uint8_t* ksin = stage + (fl >> 6) * (S3_KB * 128) + 2 * g;
          const uint32_t kchunk = (uint32_t)((fl & 63) >> 3);
#pragma unroll
          for (int b = 0; b < 16; ++b) {
            const int r = t3_row_of(q, b);
            *reinterpret_cast<__half*>(ksin + r * 128 + ((kchunk ^ (uint32_t)(r & 7)) << 4)) =
                (b & 1) ? __high2half(skeep[h][b >> 1]) : __low2half(skeep[h][b >> 1]);
          }
        }
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < T3_KEEP_BYTES / 16 / 256; ++i) {
        const int e = i * 256 + tid;          // pair e >> 9, 16-byte piece e & 511 of its run
        *reinterpret_cast<uint4*>(dst + (int64_t)(2 * (e >> 9) + part) * 32768 + (e & 511) * 16) =
            src[e];
      }
      if (part == 0) __syncthreads();         // the cosines have left: sines may overwrite them
    }
  }
}

// ---- int8 SYRK ----------------------------------------------------------------------
struct S3Bars {
  uint64_t full[S3_STAGES];        // own bulk copies landed (complete_tx)
  uint64_t peer_full[S3_STAGES];   // leader only: the peer CTA's copies landed
  uint64_t empty[S3_STAGES];       // multicast commit: stage consumed by the MMAs
  uint64_t acc_full;               // multicast commit: the chain's accumulators are complete
  uint64_t acc_empty;              // leader waits; count 8 (epilogue warps of both CTAs)
  uint32_t tmem_base;
};

struct S3Item {
  int ib, jb;
  int kb0, nkb;      // K blocks of this chain inside the launch's image
};

__device__ __forceinline__ S3Item s3_decode(int item, int ntiles, int NIB, int NJB,
                                            int nkb_total) {
  S3Item it;
  const int chain = item / ntiles;
  int t = item - chain * ntiles;
  int ib = 0;
  for (; ib < NIB; ++ib) {
    const int cnt = NJB - s3_jmin(ib);
    if (t < cnt) break;
    t -= cnt;
  }
  it.ib = ib;
  it.jb = s3_jmin(ib) + t;
  it.kb0 = chain * S3_CHAIN_KB;
  const int left = nkb_total - it.kb0;
  it.nkb = left < S3_CHAIN_KB ? left : S3_CHAIN_KB;
  return it;
}

// T is (D + 3) x ldT float64 (rows D .. D+2: the y digit columns), T[fb][fa] for fb >= fa (fa fastest: a warp's 32
// accumulator lanes are 32 consecutive fa).
__global__ void __launch_bounds__(S3_THREADS, 1)
t3_syrk_kernel(const uint8_t* __restrict__ img, int Fp, int nkb_total, int D, int NIB, int NJB,
               int ntiles, int nitems, double* __restrict__ T, int64_t ldT) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ S3Bars sb;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t crank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (tid == 0) {
    for (int s = 0; s < S3_STAGES; ++s) {
      mbar_init(&sb.full[s], 1);
      mbar_init(&sb.peer_full[s], 1);
      mbar_init(&sb.empty[s], 1);
    }
    mbar_init(&sb.acc_full, 1);
    mbar_init(&sb.acc_empty, 8);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc_2cta(&sb.tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem = sb.tmem_base;
  const int64_t plane = (int64_t)Fp * 64;          // bytes of one digit plane of one K block

  if (warp == 0) {
    // ============================ producer (both CTAs) ============================
    if (elect_one()) {
      uint32_t g = 0;
      for (int item = pair; item < nitems; item += npairs) {
        const S3Item it = s3_decode(item, ntiles, NIB, NJB, nkb_total);
        const uint8_t* a_src = img + ((int64_t)it.kb0 * 3) * plane +
                               (int64_t)(S3_TM * it.ib + (S3_TM / 2) * (int)crank) * 64;
        const uint8_t* b_src = img + ((int64_t)it.kb0 * 3) * plane +
                               (int64_t)(S3_TN * it.jb + (S3_TN / 2) * (int)crank) * 64;
        for (int kb = 0; kb < it.nkb; ++kb, ++g) {
          const uint32_t s = g % S3_STAGES;
          mbar_wait_cl(&sb.empty[s], ((g / S3_STAGES) & 1) ^ 1);
          const uint32_t dst = smem_u32(smem + s * S3_STAGE_BYTES);
          mbar_expect_tx(&sb.full[s], S3_STAGE_BYTES);
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            bulk_g2s(dst + j * S3_A_PLANE, a_src + ((int64_t)kb * 3 + j) * plane, S3_A_PLANE,
                     &sb.full[s]);
            bulk_g2s(dst + 3 * S3_A_PLANE + j * S3_B_PLANE,
                     b_src + ((int64_t)kb * 3 + j) * plane, S3_B_PLANE, &sb.full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (crank == 0) {
      // ========================== MMA issuer (leader CTA) ==========================
      const uint32_t idesc = make_idesc_i8(S3_TM, S3_TN);
      const uint32_t acc0 = tmem, acc1 = tmem + S3_TN, acc2 = tmem + 2 * S3_TN;
      uint32_t g = 0, itc = 0;
      for (int item = pair; item < nitems; item += npairs, ++itc) {
        const S3Item it = s3_decode(item, ntiles, NIB, NJB, nkb_total);
        mbar_wait_cl(&sb.acc_empty, (itc & 1) ^ 1);
        tc_fence_after_sync();
        for (int kb = 0; kb < it.nkb; ++kb, ++g) {
          const uint32_t s = g % S3_STAGES;
          const uint32_t ph = (g / S3_STAGES) & 1;
          mbar_wait_cl(&sb.full[s], ph);
          mbar_wait_cl(&sb.peer_full[s], ph);
          tc_fence_after_sync();
          if (elect_one()) {
            const uint32_t a0 = smem_u32(smem + s * S3_STAGE_BYTES);
            const uint32_t b0 = a0 + 3 * S3_A_PLANE;
            const uint64_t da0 = make_desc_sw64(a0), da1 = make_desc_sw64(a0 + S3_A_PLANE),
                           da2 = make_desc_sw64(a0 + 2 * S3_A_PLANE);
            const uint64_t db0 = make_desc_sw64(b0), db1 = make_desc_sw64(b0 + S3_B_PLANE),
                           db2 = make_desc_sw64(b0 + 2 * S3_B_PLANE);
#pragma unroll
            for (int k = 0; k < S3_KB / 32; ++k) {
              const uint64_t adv = (uint64_t)(2 * k);
              const uint32_t acc = (kb | k) != 0;
              umma2_i8_ss(acc0, da0 + adv, db0 + adv, idesc, acc);
              umma2_i8_ss(acc1, da0 + adv, db1 + adv, idesc, acc);
              umma2_i8_ss(acc1, da1 + adv, db0 + adv, idesc, 1);
              umma2_i8_ss(acc2, da1 + adv, db1 + adv, idesc, acc);
              umma2_i8_ss(acc2, da0 + adv, db2 + adv, idesc, 1);
              umma2_i8_ss(acc2, da2 + adv, db0 + adv, idesc, 1);
            }
            umma2_commit_mc(&sb.empty[s]);
            if (kb == it.nkb - 1) umma2_commit_mc(&sb.acc_full);
          }
          __syncwarp();
        }
      }
    } else {
      // ===================== relay (peer CTA): my stage landed =====================
      uint32_t g = 0;
      for (int item = pair; item < nitems; item += npairs) {
        const S3Item it = s3_decode(item, ntiles, NIB, NJB, nkb_total);
        for (int kb = 0; kb < it.nkb; ++kb, ++g) {
          const uint32_t s = g % S3_STAGES;
          mbar_wait_cl(&sb.full[s], (g / S3_STAGES) & 1);
          if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&sb.peer_full[s]), 0));
          __syncwarp();
        }
      }
    }
  } else {
    // ============================ epilogue (warps 2..5) ============================
    const int q = warp & 3;               // TMEM lane quadrant this warp may read
    uint32_t itc = 0;
    for (int item = pair; item < nitems; item += npairs, ++itc) {
      const S3Item it = s3_decode(item, ntiles, NIB, NJB, nkb_total);
      const int fa = S3_TM * it.ib + (S3_TM / 2) * (int)crank + 32 * q + lane;
      const int fb0 = S3_TN * it.jb;
      mbar_wait_cl(&sb.acc_full, itc & 1);
      tc_fence_after_sync();
      const uint32_t tacc = tmem + ((uint32_t)(32 * q) << 16);
      double* Tcol = T + fa;
#pragma unroll 1
      for (int c0 = 0; c0 < S3_TN; c0 += 16) {
        int a0[16], a1[16], a2[16];
        tmem_ld16i_nowait(tacc + (uint32_t)c0, a0);
        tmem_ld16i_nowait(tacc + (uint32_t)(S3_TN + c0), a1);
        tmem_ld16i_nowait(tacc + (uint32_t)(2 * S3_TN + c0), a2);
        tmem_ld_wait();
        if (c0 + 16 >= S3_TN) {             // all TMEM reads of this chain are done
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&sb.acc_empty), 0));
        }
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const int fb = fb0 + c0 + r;
          if (fa < D && fb <= D + 2 && fb >= fa) {
            const long long v = (long long)a0[r] * 65536ll + (long long)a1[r] * 256ll +
                                (long long)a2[r];
            if (v != 0) atomicAdd(Tcol + (int64_t)fb * ldT, (double)v);
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2cta(tmem, 512);
}

// G[i][j] += kappa * camp_i * camp_j * T[max(i,j)][min(i,j)],
// p[i] += kappa * camp_i * camp_D * (T[D][i] + T[D+1][i] / 256 + T[D+2][i] / 65536).
__global__ void __launch_bounds__(256)
t3_finalize_kernel(const double* __restrict__ T, int64_t ldT, const float* __restrict__ camp,
                   double* __restrict__ G, double* __restrict__ p, int D, double kappa) {
  __shared__ double tile[32][33];
  const int bx = blockIdx.x, by = blockIdx.y;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  if (by == gridDim.y - 1) {
    // last grid row: the y column -> p
    if (p != nullptr) {
      const int i = bx * 32 + tx;
      if (ty == 0 && i < D)
        p[i] += kappa * (double)camp[i] * (double)camp[D] *
                (T[(int64_t)D * ldT + i] + T[(int64_t)(D + 1) * ldT + i] * (1.0 / 256.0) +
                 T[(int64_t)(D + 2) * ldT + i] * (1.0 / 65536.0));
    }
    return;
  }
  if (by >= bx) {
    // lower (or diagonal) block of G: rows by, columns bx -> T[i][j] directly
    for (int r = ty; r < 32; r += 8) {
      const int i = by * 32 + r, j = bx * 32 + tx;
      tile[r][tx] = (i < D && j < D) ? T[(int64_t)i * ldT + j] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      const int i = by * 32 + r, j = bx * 32 + tx;
      if (i < D && j < D) {
        const double v = (by > bx || r >= tx) ? tile[r][tx] : tile[tx][r];
        G[(int64_t)i * D + j] += kappa * (double)camp[i] * (double)camp[j] * v;
      }
    }
  } else {
    // upper block: G[i][j] = T[j][i]; read T's block (rows bx, columns by) coalesced
    for (int r = ty; r < 32; r += 8) {
      const int jj = bx * 32 + r, ii = by * 32 + tx;
      tile[r][tx] = (jj < D && ii < D) ? T[(int64_t)jj * ldT + ii] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      const int i = by * 32 + r, j = bx * 32 + tx;
      if (i < D && j < D)
        G[(int64_t)i * D + j] += kappa * (double)camp[i] * (double)camp[j] * tile[tx][r];
    }
  }
}

// ---------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------
struct S3Shape {
  int D, NIB, NJB, Fp, ntiles;
  int64_t ldT, group_rows;
  size_t group_bytes;
};

static S3Shape s3_shape(const rr_plan* pl, int64_t N) {
  S3Shape s;
  s.D = pl->D;
  s.NIB = (s.D + S3_TM - 1) / S3_TM;
  s.NJB = (s.D + 3 + S3_TN - 1) / S3_TN;
  const int fa = S3_TM * s.NIB, fb = S3_TN * s.NJB;
  s.Fp = fa > fb ? fa : fb;
  s.ntiles = 0;
  for (int ib = 0; ib < s.NIB; ++ib) s.ntiles += s.NJB - s3_jmin(ib);
  s.ldT = ((int64_t)s.D + 3) / 4 * 4;
  int64_t chains = S3_GROUP_CHAINS;
  const int64_t per_chain = (int64_t)3 * S3_CHAIN * s.Fp;
  while (chains > 1 && (size_t)(chains * per_chain) > S3_GROUP_BYTES_MAX) --chains;
  const int64_t need = (N + S3_CHAIN - 1) / S3_CHAIN;
  if (need < chains) chains = need < 1 ? 1 : need;
  s.group_rows = chains * S3_CHAIN;
  int64_t rows_buf = s.group_rows < N ? s.group_rows : ((N + S3_KB - 1) / S3_KB) * S3_KB;
  if (rows_buf < S3_KB) rows_buf = S3_KB;
  s.group_bytes = (size_t)3 * rows_buf * s.Fp;
  return s;
}

int tc3_suffstats_supported(const rr_plan* pl) {
  if (pl->d < 1 || pl->d > 128 || pl->D < 1 || pl->kind != nullptr) return 0;
  return pl->D == 2 * pl->ktot + pl->next ? 1 : 0;
}

size_t tc3_suffstats_workspace(const rr_plan* pl, int64_t N) {
  const S3Shape s = s3_shape(pl, N);
  return align_up((size_t)(s.D + 3) * s.ldT * sizeof(double), 256) +
         align_up((size_t)(s.D + 1) * sizeof(float), 256) +
         align_up((size_t)(pl->d + 1) * sizeof(unsigned int), 256) +
         2 * (align_up(s.group_bytes, 1024) + 1024) + 4096;
}

static int launch_syrk(const uint8_t* img, const S3Shape& s, int nkb, double* T, cudaStream_t st) {
  const size_t smem = (size_t)S3_STAGES * S3_STAGE_BYTES + 1024;
  RR_CUDA_CHECK(cudaFuncSetAttribute(t3_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
  const int nchains = (nkb + S3_CHAIN_KB - 1) / S3_CHAIN_KB;
  const int nitems = nchains * s.ntiles;
  int npairs = sm_count() / 2;
  if (nitems < npairs) npairs = nitems;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * npairs);
  cfg.blockDim = dim3(S3_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  RR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, t3_syrk_kernel, img, s.Fp, nkb, s.D, s.NIB, s.NJB,
                                   s.ntiles, nitems, T, s.ldT));
  RR_LAUNCH_CHECK("t3_syrk_kernel");
  return RR_OK;
}

static int tc3_groups(const rr_plan* pl, const S3Shape& s, const float* X, const float* y,
                      int64_t N, const unsigned int* scales, uint8_t* const imgs[2], double* T,
                      uint8_t* phi16, bool overlap, cudaStream_t sg, cudaStream_t st,
                      cudaEvent_t ev_gen[2], cudaEvent_t ev_mma[2]) {
  const int d = pl->d, D = s.D;
  const int gy_trig = (pl->ktot + T3_FREQS - 1) / T3_FREQS;
  const int nother = pl->next + 3 + (s.Fp - (D + 3));
  const int gy_other = (nother + T3_OTHER - 1) / T3_OTHER;
  const int nkb16 = (int)(kept_features_cols(pl) / 64);
  size_t dsmem = (size_t)T3_PARTS * S3_KB * (((d + 7) & ~7) + 4) * sizeof(float);
  if (phi16 && dsmem < (size_t)T3_KEEP_BYTES) dsmem = T3_KEEP_BYTES;   // staging aliases the slab
  if (dsmem > 48 * 1024)
    RR_CUDA_CHECK(cudaFuncSetAttribute(phi16 ? t3_digits_kernel<true> : t3_digits_kernel<false>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem));
  int g = 0;
  int64_t step = 0;
  for (int64_t r0 = 0; r0 < N; r0 += step, ++g) {
    // group 0 is short (little exposed generator time); a job that fits one group anyway
    // is not split
    step = (g == 0 && N > s.group_rows) ? (int64_t)S3_FIRST_CHAINS * S3_CHAIN : s.group_rows;
    if (step > s.group_rows) step = s.group_rows;
    const int64_t rows = (N - r0) < step ? (N - r0) : step;
    const int nkb = (int)((rows + S3_KB - 1) / S3_KB);
    const int buf = g & 1;
    if (overlap && g >= 2) RR_CUDA_CHECK(cudaStreamWaitEvent(sg, ev_mma[buf], 0));
    dim3 grid((unsigned)nkb, (unsigned)(gy_trig + gy_other));
    if (phi16)   // groups start on multiples of 32768 rows: whole 256-row images
      t3_digits_kernel<true><<<grid, 256, dsmem, sg>>>(
          *pl, X + r0 * d, y ? y + r0 : nullptr, rows, s.Fp, gy_trig, scales, imgs[buf],
          phi16 + (r0 >> 8) * (int64_t)nkb16 * 32768, nkb16);
    else
      t3_digits_kernel<false><<<grid, 256, dsmem, sg>>>(*pl, X + r0 * d, y ? y + r0 : nullptr,
                                                        rows, s.Fp, gy_trig, scales, imgs[buf],
                                                        nullptr, nkb16);
    RR_LAUNCH_CHECK("t3_digits_kernel");
    if (overlap) {
      RR_CUDA_CHECK(cudaEventRecord(ev_gen[buf], sg));
      RR_CUDA_CHECK(cudaStreamWaitEvent(st, ev_gen[buf], 0));
    }
    const int rc = launch_syrk(imgs[buf], s, nkb, T, st);
    if (rc) return rc;
    if (overlap) RR_CUDA_CHECK(cudaEventRecord(ev_mma[buf], st));
  }
  return RR_OK;
}

int tc3_suffstats(const rr_plan* pl, const float* X, const float* y, int64_t N, double* G,
                  double* p, void* ws, size_t ws_bytes, rr_context* ctx, cudaStream_t st,
                  void* kept) {
  const S3Shape s = s3_shape(pl, N);
  uint8_t* phi16 = static_cast<uint8_t*>(kept);
  if (phi16 != nullptr) {
    if (!tc_gradpass_supported(pl) || (reinterpret_cast<uintptr_t>(phi16) & 1023) != 0) {
      set_error("kept features: plan not supported by the tensor-core gradient pass, or the "
                "buffer is not 1024-byte aligned");
      return RR_ERR_INVALID;
    }
    if (N % 256 != 0) {
      // rows past the end of the last 256-row image: the K blocks of the generator
      // stop at the next multiple of 64
      const int64_t cols = kept_features_cols(pl);
      RR_CUDA_CHECK(cudaMemsetAsync(phi16 + (N / 256) * cols * 512, 0, (size_t)cols * 512, st));
    }
  }
  const int D = s.D, d = pl->d;
  Workspace W(ws, ws_bytes);
  double* T = W.take<double>((size_t)(D + 3) * s.ldT);
  float* camp = W.take<float>((size_t)D + 1);
  unsigned int* scales = W.take<unsigned int>((size_t)d + 1);
  uint8_t* imgs[2];
  imgs[0] = W.take<uint8_t>(align_up(s.group_bytes, 1024) + 1024);
  imgs[1] = W.take<uint8_t>(align_up(s.group_bytes, 1024) + 1024);
  if (!T || !camp || !scales || !imgs[0] || !imgs[1]) {
    set_error("int8 suffstats workspace too small (need %zu bytes)",
              tc3_suffstats_workspace(pl, N));
    return RR_ERR_WORKSPACE;
  }
  for (int i = 0; i < 2; ++i)
    imgs[i] = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(imgs[i]) + 1023) & ~(uintptr_t)1023);

  RR_CUDA_CHECK(cudaMemsetAsync(T, 0, (size_t)(D + 3) * s.ldT * sizeof(double), st));
  RR_CUDA_CHECK(cudaMemsetAsync(scales, 0, (size_t)(d + 1) * sizeof(unsigned int), st));
  RR_CUDA_CHECK(cudaMemsetAsync(camp, 0, (size_t)(D + 1) * sizeof(float), st));
  bool need_x = false;   // any affine column that copies an input column?  (host plan is a
                         // device image: decide from `next` alone)
  need_x = pl->next > 0;
  {
    if (pl->col_scale != nullptr) {
      // caller-provided (job-wide) scales: float bit patterns, as the kernel would write
      RR_CUDA_CHECK(cudaMemcpyAsync(scales, pl->col_scale, (size_t)(d + 1) * sizeof(float),
                                    cudaMemcpyDeviceToDevice, st));
    } else {
      const int blocks = sm_count() * 4;
      t3_scales_kernel<<<blocks, 256, (size_t)(d + 1) * sizeof(unsigned int), st>>>(
          X, (p != nullptr) ? y : nullptr, N, d, need_x ? 1 : 0, scales);
      RR_LAUNCH_CHECK("t3_scales_kernel");
    }
    const int nth = pl->ktot > pl->next ? pl->ktot : pl->next;
    t3_colamp_kernel<<<(nth + 256) / 256, 256, 0, st>>>(*pl, scales, camp);
    RR_LAUNCH_CHECK("t3_colamp_kernel");
  }

  cudaStream_t sg = st;                     // generator stream
  cudaEvent_t ev_fork = nullptr, ev_gen[2] = {nullptr, nullptr}, ev_mma[2] = {nullptr, nullptr};
  if (ctx != nullptr && ctx_aux(ctx, &sg, &ev_fork, ev_gen, ev_mma) != RR_OK) return RR_ERR_CUDA;
  const bool overlap = sg != st;
  if (overlap) {
    RR_CUDA_CHECK(cudaEventRecord(ev_fork, st));
    RR_CUDA_CHECK(cudaStreamWaitEvent(sg, ev_fork, 0));
  }

  {
    const int rc = tc3_groups(pl, s, X, (p != nullptr) ? y : nullptr, N, scales, imgs, T, phi16,
                              overlap, sg, st, ev_gen, ev_mma);
    if (rc) {   // join whatever the helper stream still has queued, then report
      if (overlap && cudaEventRecord(ev_fork, sg) == cudaSuccess) cudaStreamWaitEvent(st, ev_fork, 0);
      return rc;
    }
  }
  {
    const double kappa = 65536.0 / ((double)S3_SCALE * (double)S3_SCALE);
    dim3 fg((D + 31) / 32, (D + 31) / 32 + 1);
    t3_finalize_kernel<<<fg, 256, 0, st>>>(T, s.ldT, camp, G, p, D, kappa);
    RR_LAUNCH_CHECK("t3_finalize_kernel");
  }
  return RR_OK;
}

}  // namespace rr

// ---------------------------------------------------------------------------
// Self-test: the production GEMM kernel on a random digit image, bit-exact
// against a host integer reference (descriptor encodings, 64-byte swizzle,
// digit-product schedule, TMEM drain, triangular tile set).
// ---------------------------------------------------------------------------
extern "C" int rr_tcgen05_i8_selftest(int32_t kblocks, int64_t* mismatches) {
  using namespace rr;
  RR_REQUIRE(kblocks >= 1 && kblocks <= 4096, "kblocks out of range");
  rr_plan pl;
  memset(&pl, 0, sizeof(pl));
  pl.d = 1;
  pl.ktot = 150;
  pl.D = 300;
  const int64_t N = (int64_t)kblocks * S3_KB;
  S3Shape s = s3_shape(&pl, N);
  const int D = s.D, Fp = s.Fp;
  const size_t img_bytes = (size_t)3 * kblocks * Fp * 64;
  const size_t t_bytes = (size_t)(D + 3) * s.ldT * sizeof(double);
  int8_t* hd = (int8_t*)malloc((size_t)3 * N * Fp);        // [plane][feature f][row n]
  uint8_t* himg = (uint8_t*)calloc(img_bytes, 1);
  double* hT = (double*)malloc(t_bytes);
  uint8_t* dimg = nullptr;
  double* dT = nullptr;
  int rc = RR_OK;
  cudaError_t e;
  uint32_t seed = 777u;
  auto rnd = [&]() {
    seed = seed * 1664525u + 1013904223u;
    return (int)((seed >> 13) & 0xFF) - 128;
  };
  for (int j = 0; j < 3; ++j)
    for (int64_t n = 0; n < N; ++n)
      for (int f = 0; f < Fp; ++f) {
        const int8_t v = (int8_t)((f <= D) ? rnd() : 0);
        hd[((size_t)j * Fp + f) * N + n] = v;
        const int64_t kb = n / S3_KB;
        const uint32_t r = (uint32_t)(n % S3_KB);
        himg[(((size_t)kb * 3 + j) * Fp + f) * 64 + (((r >> 4) ^ (((uint32_t)f >> 1) & 3u)) << 4) + (r & 15u)] =
            (uint8_t)v;
      }
#define ST_CHECK(x) do { e = (x); if (e != cudaSuccess) { set_error("i8 selftest: %s: %s", #x, cudaGetErrorString(e)); rc = RR_ERR_CUDA; goto done; } } while (0)
  ST_CHECK(cudaMalloc(&dimg, img_bytes));
  ST_CHECK(cudaMalloc(&dT, t_bytes));
  ST_CHECK(cudaMemcpy(dimg, himg, img_bytes, cudaMemcpyHostToDevice));
  ST_CHECK(cudaMemset(dT, 0, t_bytes));
  rc = launch_syrk(dimg, s, kblocks, dT, 0);
  if (rc) goto done;
  ST_CHECK(cudaDeviceSynchronize());
  ST_CHECK(cudaMemcpy(hT, dT, t_bytes, cudaMemcpyDeviceToHost));
  {
    int64_t bad = 0;
    for (int fb = 0; fb <= D; ++fb)
      for (int fa = 0; fa < D; ++fa) {
        long long ref = 0;
        if (fb >= fa) {
          const int8_t *a0 = hd + ((size_t)0 * Fp + fa) * N, *a1 = hd + ((size_t)1 * Fp + fa) * N,
                       *a2 = hd + ((size_t)2 * Fp + fa) * N;
          const int8_t *b0 = hd + ((size_t)0 * Fp + fb) * N, *b1 = hd + ((size_t)1 * Fp + fb) * N,
                       *b2 = hd + ((size_t)2 * Fp + fb) * N;
          long long s00 = 0, s01 = 0, s11 = 0;
          for (int64_t n = 0; n < N; ++n) {
            s00 += (int)a0[n] * (int)b0[n];
            s01 += (int)a0[n] * (int)b1[n] + (int)a1[n] * (int)b0[n];
            s11 += (int)a1[n] * (int)b1[n] + (int)a0[n] * (int)b2[n] + (int)a2[n] * (int)b0[n];
          }
          ref = s00 * 65536ll + s01 * 256ll + s11;
        }
        if (hT[(size_t)fb * s.ldT + fa] != (double)ref) ++bad;
      }
    if (mismatches) *mismatches = bad;
    if (bad) {
      set_error("tcgen05 kind::i8 selftest: %lld accumulator entries differ from the host reference",
                (long long)bad);
      rc = RR_ERR_CUDA;
    }
  }
done:
#undef ST_CHECK
  cudaFree(dimg);
  cudaFree(dT);
  free(hd);
  free(himg);
  free(hT);
  return rc;
}
