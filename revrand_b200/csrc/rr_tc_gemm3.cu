// fp32-grade GEMM on the tcgen05 tensor cores: C (+)= alpha * A' B'^T with both
// operands split into two tf32 parts and three products per k-step,
//     a b ~= a_hi b_hi + a_lo b_hi + a_hi b_lo        (a_hi = tf32(a), a_lo = tf32(a - a_hi))
// i.e. every product to ~2^-21 |a b|, fp32 accumulation in TMEM -- the arithmetic
// of the mma.sync tf32x3 kernel this replaces, at tcgen05 rates.  It carries the
// three GEMM-shaped contractions of the GLM SVI step (glm.py:303 F = Phi Ws^T,
// :307 Edws = dF^T Phi, :310 EdPhi = dF Ws) and the latent draws of the GLM
// predictive sampler (glm.py:404-418).
//
//   g3_pack_kernel   fp32 operand with arbitrary (row, k) strides -> tile-major tf32
//                    images [row block of 256][k block of 32][hi | lo], every image a
//                    contiguous 32 KB K-major SWIZZLE_128B tile, zero padded
//   g3_gemm_kernel   persistent CTA pairs (cta_group::2, M = 256, N = 128), 4-stage
//                    ring of bulk copies (48 KB per CTA and stage), 12
//                    tcgen05.mma.kind::tf32 per stage, accumulators double-buffered in
//                    TMEM; K is cut into chains of 256 elements that the epilogue warps
//                    add up in fp32 registers (bounds the truncation bias of the
//                    in-TMEM accumulation); split-K over work items (fp32 atomics)
//                    only when the output has too few tiles for the 74 CTA pairs
#include "rr_common.cuh"
#include "rr_tc.cuh"

namespace rr {

using namespace tc;

constexpr int G3_TM = 256;                    // rows of A per tile (128 per CTA)
constexpr int G3_TN = 128;                    // rows of B per tile (64 per CTA)
constexpr int G3_KB = 32;                     // tf32 elements per k block (one 128-byte line)
constexpr int G3_IMG = 256 * 128;             // one operand image: 256 rows x 128 B = 32 KB
constexpr int G3_A_HALF = (G3_TM / 2) * 128;  // 16 KB
constexpr int G3_B_HALF = (G3_TN / 2) * 128;  // 8 KB
constexpr int G3_STAGES = 4;
constexpr int G3_STAGE_BYTES = 2 * G3_A_HALF + 2 * G3_B_HALF;   // A hi, A lo, B hi, B lo: 48 KB
constexpr int G3_THREADS = 6 * 32;            // producer, MMA / relay, 4 epilogue warps
constexpr int G3_CHAIN_KB = 8;                // k blocks per accumulation chain (see gemm3)

__device__ __forceinline__ float g3_tf32(float x) {
  // round to 10 explicit mantissa bits, low 13 bits zero (the tensor core truncates)
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// Block = 32 rows x 32 k of one (row block, k block); the source is read with the
// lane on whichever index is contiguous and written one 128-byte image line (32 k
// of one row) per warp instruction.
__global__ void __launch_bounds__(256)
g3_pack_kernel(const float* __restrict__ src, int64_t sR, int64_t sK, int R, int K, int nkb,
               uint8_t* __restrict__ img) {
  __shared__ float tile[32][33];
  const int kb = blockIdx.x;
  const int r0 = blockIdx.y * 32;             // first row of this block (global)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const bool k_contig = (sK == 1) || (sR != 1);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a = w + 8 * i;                  // the slow index of this pass
    const int r = k_contig ? r0 + a : r0 + lane;
    const int k = k_contig ? kb * G3_KB + lane : kb * G3_KB + a;
    const float v = (r < R && k < K) ? src[(int64_t)r * sR + (int64_t)k * sK] : 0.0f;
    if (k_contig) tile[a][lane] = v;
    else tile[lane][a] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rl = w + 8 * i;
    const int r = r0 + rl;
    const float v = tile[rl][lane];
    const float hi = g3_tf32(v);
    const float lo = g3_tf32(v - hi);
    const int rb = r >> 8, ri = r & 255;
    uint8_t* base = img + (((int64_t)rb * nkb + kb) * 2) * G3_IMG +
                    sw128_off((uint32_t)ri, (uint32_t)(lane >> 2)) + (lane & 3) * 4;
    *reinterpret_cast<float*>(base) = hi;
    *reinterpret_cast<float*>(base + G3_IMG) = lo;
  }
}

struct G3Bars {
  uint64_t full[G3_STAGES];
  uint64_t peer_full[G3_STAGES];
  uint64_t empty[G3_STAGES];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};

struct G3Item {
  int rb, nb, kb0, kb1;
};
__device__ __forceinline__ G3Item g3_decode(int item, int NB, int nsplit, int nkb, int ntiles) {
  G3Item it;
  // k range slowest: the CTA pairs running together work on the same k range of
  // different tiles and share its operand images in L2
  const int ks = item / ntiles, tile = item - ks * ntiles;
  it.rb = tile / NB;
  it.nb = tile - it.rb * NB;
  it.kb0 = (int)(((int64_t)ks * nkb) / nsplit);
  it.kb1 = (int)(((int64_t)(ks + 1) * nkb) / nsplit);
  return it;
}

// The tensor core adds every MMA's result into the TMEM accumulator with truncation
// (measured: 768 accumulating MMAs, K = 2048, leave a 1.5e-5 relative bias against
// float64).  So the k range of a work item is cut into CHAINS of G3_CHAIN_KB k blocks
// (96 MMAs: < 3e-6): the issuer alternates between two TMEM accumulators, one per
// chain, and the epilogue warps add each finished chain into fp32 REGISTERS (round to
// nearest, 128 per thread) while the next chain runs -- one write of C per item.
__global__ void __launch_bounds__(G3_THREADS, 1)
g3_gemm_kernel(const uint8_t* __restrict__ Aimg, const uint8_t* __restrict__ Bimg, int M, int N,
               int nkb, int NB, int nsplit, int nitems, float alpha, float* __restrict__ C,
               int64_t ldc, int atomic) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ G3Bars sb;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t crank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int ntiles = nitems / nsplit;

  if (tid == 0) {
    for (int s = 0; s < G3_STAGES; ++s) {
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
  if (warp == 1) tmem_alloc_2cta(&sb.tmem_base, 2 * G3_TN);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem = sb.tmem_base;

  if (warp == 0) {
    // ============================ producer (both CTAs) ============================
    if (elect_one()) {
      uint32_t g = 0;
      for (int item = pair; item < nitems; item += npairs) {
        const G3Item it = g3_decode(item, NB, nsplit, nkb, ntiles);
        // A: this CTA's 128 rows of the 256-row block; B: its 64 rows of the 128-row
        // tile, i.e. of half (nb & 1) of B's 256-row block nb >> 1
        const uint8_t* a_src = Aimg + ((int64_t)it.rb * nkb) * 2 * G3_IMG + (int64_t)crank * G3_A_HALF;
        const uint8_t* b_src = Bimg + ((int64_t)(it.nb >> 1) * nkb) * 2 * G3_IMG +
                               (int64_t)(it.nb & 1) * (2 * G3_B_HALF) + (int64_t)crank * G3_B_HALF;
        for (int kb = it.kb0; kb < it.kb1; ++kb, ++g) {
          const uint32_t s = g % G3_STAGES;
          mbar_wait_cl(&sb.empty[s], ((g / G3_STAGES) & 1) ^ 1);
          const uint32_t dst = smem_u32(smem + s * G3_STAGE_BYTES);
          mbar_expect_tx(&sb.full[s], G3_STAGE_BYTES);
          const uint8_t* ak = a_src + (int64_t)kb * 2 * G3_IMG;
          const uint8_t* bk = b_src + (int64_t)kb * 2 * G3_IMG;
          bulk_g2s(dst, ak, G3_A_HALF, &sb.full[s]);
          bulk_g2s(dst + G3_A_HALF, ak + G3_IMG, G3_A_HALF, &sb.full[s]);
          bulk_g2s(dst + 2 * G3_A_HALF, bk, G3_B_HALF, &sb.full[s]);
          bulk_g2s(dst + 2 * G3_A_HALF + G3_B_HALF, bk + G3_IMG, G3_B_HALF, &sb.full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (crank == 0) {
      // ========================== MMA issuer (leader CTA) ==========================
      const uint32_t idesc = make_idesc(2, G3_TM, G3_TN);     // tf32 x tf32 -> fp32
      uint32_t g = 0, cc = 0;                                  // stage and chain counters
      for (int item = pair; item < nitems; item += npairs) {
        const G3Item it = g3_decode(item, NB, nsplit, nkb, ntiles);
        for (int c0 = it.kb0; c0 < it.kb1; c0 += G3_CHAIN_KB, ++cc) {
          const int c1 = c0 + G3_CHAIN_KB < it.kb1 ? c0 + G3_CHAIN_KB : it.kb1;
          const uint32_t buf = cc & 1;
          mbar_wait_cl(&sb.acc_empty[buf], ((cc >> 1) & 1) ^ 1);
          tc_fence_after_sync();
          for (int kb = c0; kb < c1; ++kb, ++g) {
            const uint32_t s = g % G3_STAGES;
            const uint32_t ph = (g / G3_STAGES) & 1;
            mbar_wait_cl(&sb.full[s], ph);
            mbar_wait_cl(&sb.peer_full[s], ph);
            tc_fence_after_sync();
            if (elect_one()) {
              const uint32_t a0 = smem_u32(smem + s * G3_STAGE_BYTES);
              const uint32_t b0 = a0 + 2 * G3_A_HALF;
              const uint64_t dah = make_desc_sw128(a0), dal = make_desc_sw128(a0 + G3_A_HALF);
              const uint64_t dbh = make_desc_sw128(b0), dbl = make_desc_sw128(b0 + G3_B_HALF);
              const uint32_t acc = tmem + buf * G3_TN;
#pragma unroll
              for (int k = 0; k < G3_KB / 8; ++k) {
                const uint64_t adv = (uint64_t)(2 * k);
                // small terms first
                umma2_tf32_ss(acc, dal + adv, dbh + adv, idesc, (kb > c0) | (k != 0));
                umma2_tf32_ss(acc, dah + adv, dbl + adv, idesc, 1);
                umma2_tf32_ss(acc, dah + adv, dbh + adv, idesc, 1);
              }
              umma2_commit_mc(&sb.empty[s]);
              if (kb == c1 - 1) umma2_commit_mc(&sb.acc_full[buf]);
            }
            __syncwarp();
          }
        }
      }
    } else {
      // ===================== relay (peer CTA): my stage landed =====================
      uint32_t g = 0;
      for (int item = pair; item < nitems; item += npairs) {
        const G3Item it = g3_decode(item, NB, nsplit, nkb, ntiles);
        for (int kb = it.kb0; kb < it.kb1; ++kb, ++g) {
          const uint32_t s = g % G3_STAGES;
          mbar_wait_cl(&sb.full[s], (g / G3_STAGES) & 1);
          if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&sb.peer_full[s]), 0));
          __syncwarp();
        }
      }
    }
  } else {
    // ============================ epilogue (warps 2..5) ============================
    const int q = warp & 3;
    uint32_t cc = 0;
    const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    for (int item = pair; item < nitems; item += npairs) {
      const G3Item it = g3_decode(item, NB, nsplit, nkb, ntiles);
      const int row = it.rb * G3_TM + 128 * (int)crank + 32 * q + lane;
      const int col0 = it.nb * G3_TN;
      float sum[G3_TN];
#pragma unroll
      for (int j = 0; j < G3_TN; ++j) sum[j] = 0.0f;
      for (int c0 = it.kb0; c0 < it.kb1; c0 += G3_CHAIN_KB, ++cc) {
        const uint32_t buf = cc & 1;
        mbar_wait_cl(&sb.acc_full[buf], (cc >> 1) & 1);
        tc_fence_after_sync();
        const uint32_t tacc = tmem + ((uint32_t)(32 * q) << 16) + buf * G3_TN;
#pragma unroll
        for (int c8 = 0; c8 < G3_TN / 32; ++c8) {
          float v[32];
          tmem_ld32_nowait(tacc + (uint32_t)(32 * c8), v);
          tmem_ld_wait();
          if (c8 == G3_TN / 32 - 1) {       // all TMEM reads of this chain are done
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&sb.acc_empty[buf]), 0));
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[32 * c8 + j] += v[j];
        }
      }
      if (row < M) {
        float* crow = C + (int64_t)row * ldc;
#pragma unroll
        for (int c8 = 0; c8 < G3_TN / 32; ++c8) {
          const int c = col0 + 32 * c8;
          if (c >= N) break;
          if (atomic) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c + j < N) atomicAdd(crow + c + j, alpha * sum[32 * c8 + j]);
          } else if (vec_ok && c + 32 <= N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(crow + c + j) =
                  make_float4(alpha * sum[32 * c8 + j], alpha * sum[32 * c8 + j + 1],
                              alpha * sum[32 * c8 + j + 2], alpha * sum[32 * c8 + j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c + j < N) crow[c + j] = alpha * sum[32 * c8 + j];
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2cta(tmem, 2 * G3_TN);
}

// ---- host -------------------------------------------------------------------------
size_t gemm3_image_bytes(int64_t R, int64_t K) {
  const int64_t rbs = (R + 255) / 256, nkb = (K + G3_KB - 1) / G3_KB;
  return (size_t)(rbs * nkb * 2 * G3_IMG) + 1024;
}

// below this many multiply-adds the CUDA-core kernel is as fast as pack + tcgen05
bool gemm3_worthwhile(int M, int N, int K) {
  return (int64_t)M * N * K >= ((int64_t)1 << 26) && K >= 64;
}

static int g3_pack(const float* src, int64_t sR, int64_t sK, int R, int K, uint8_t* img,
                   cudaStream_t st) {
  const int rbs = (R + 255) / 256, nkb = (K + G3_KB - 1) / G3_KB;
  dim3 grid((unsigned)nkb, (unsigned)(rbs * (256 / 32)));
  g3_pack_kernel<<<grid, 256, 0, st>>>(src, sR, sK, R, K, nkb, img);
  RR_LAUNCH_CHECK("g3_pack_kernel");
  return RR_OK;
}

// C (+)= alpha * A' B', A'(m,k) = A[m*sAm + k*sAk], B'(k,n) = B[k*sBk + n*sBn] (the
// conventions of sgemm); imgA / imgB: scratch of gemm3_image_bytes(M, K) / (N, K).
int gemm3(int M, int N, int K, float alpha, const float* A, int64_t sAm, int64_t sAk,
          const float* B, int64_t sBk, int64_t sBn, float* C, int64_t ldc, int accumulate,
          uint8_t* imgA, uint8_t* imgB, cudaStream_t st) {
  if (M <= 0 || N <= 0) return RR_OK;
  imgA = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(imgA) + 1023) & ~(uintptr_t)1023);
  imgB = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(imgB) + 1023) & ~(uintptr_t)1023);
  int rc = g3_pack(A, sAm, sAk, M, K, imgA, st);
  if (rc) return rc;
  rc = g3_pack(B, sBn, sBk, N, K, imgB, st);
  if (rc) return rc;
  const int MB = (M + G3_TM - 1) / G3_TM, NB = (N + G3_TN - 1) / G3_TN;
  const int nkb = (K + G3_KB - 1) / G3_KB;
  int npairs = sm_count() / 2;
  const int tiles = MB * NB;
  // split K over work items (combined with fp32 atomics) only when the output has too
  // few tiles to fill the CTA pairs; every item keeps at least one whole chain
  int nsplit = 1;
  if (tiles < npairs) {
    nsplit = npairs / tiles;
    const int maxsplit = nkb / G3_CHAIN_KB > 0 ? nkb / G3_CHAIN_KB : 1;
    if (nsplit > maxsplit) nsplit = maxsplit;
    if (nsplit < 1) nsplit = 1;
  }
  const int atomic = (nsplit > 1 || accumulate) ? 1 : 0;
  if (atomic && !accumulate)
    RR_CUDA_CHECK(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float),
                                    (size_t)M, st));
  const int nitems = tiles * nsplit;
  if (nitems < npairs) npairs = nitems;
  const size_t smem = (size_t)G3_STAGES * G3_STAGE_BYTES + 1024;
  RR_CUDA_CHECK(cudaFuncSetAttribute(g3_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * npairs);
  cfg.blockDim = dim3(G3_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  RR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, g3_gemm_kernel, (const uint8_t*)imgA,
                                   (const uint8_t*)imgB, M, N, nkb, NB, nsplit, nitems, alpha, C,
                                   ldc, atomic));
  RR_LAUNCH_CHECK("g3_gemm_kernel");
  return RR_OK;
}

}  // namespace rr

// Diagnostic / test entry: C = alpha * A B^T for row-major fp32 device matrices
// A (M x K, leading dimension lda), B (N x K, ldb), C (M x N, ldc) through the
// tcgen05 tf32x3 kernel (transa / transb != 0: the operand is stored K x M / K x N).
extern "C" int rr_tcgen05_gemm3(int32_t M, int32_t N, int32_t K, float alpha, const float* A,
                                int64_t lda, int32_t transa, const float* B, int64_t ldb,
                                int32_t transb, float* C, int64_t ldc, int32_t accumulate,
                                void* workspace, size_t workspace_bytes, void* stream) {
  using namespace rr;
  RR_REQUIRE(A && B && C && workspace, "null pointer");
  RR_REQUIRE(M > 0 && N > 0 && K > 0, "empty product");
  Workspace W(workspace, workspace_bytes);
  uint8_t* ia = W.take<uint8_t>(gemm3_image_bytes(M, K));
  uint8_t* ib = W.take<uint8_t>(gemm3_image_bytes(N, K));
  if (!ia || !ib) {
    set_error("gemm3 workspace too small (need %zu bytes)",
              gemm3_image_bytes(M, K) + gemm3_image_bytes(N, K) + 512);
    return RR_ERR_WORKSPACE;
  }
  return gemm3(M, N, K, alpha, A, transa ? 1 : lda, transa ? lda : 1, B, transb ? ldb : 1,
               transb ? 1 : ldb, C, ldc, accumulate, ia, ib, (cudaStream_t)stream);
}
