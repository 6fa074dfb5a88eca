// Diagnostic probe of the tcgen05 fp32 accumulation behaviour.
//
// The fused Gram kernels accumulate long chains of fp16 x fp16 products in
// TMEM.  How the tensor core rounds when it adds a K=16 product group into an
// fp32 accumulator is undocumented, and it decides how many rows may be
// accumulated before the partial tile must be flushed to float64 (DESIGN.md,
// "accumulation").  This probe repeats D += A B^T over one 128 x 256 x 64 tile
// `reps` times so that the exact result (reps * A B^T, computable on the host
// from the fp16 inputs) is known, and returns the TMEM contents.
//
//   mode 0 : D += Ahi Bhi^T
//   mode 1 : D += Ahi Bhi^T + Alo Bhi^T + Ahi Blo^T      (one accumulator)
//   mode 2 : D += Ahi Bhi^T ; AUX += Alo Bhi^T + Ahi Blo^T (two accumulators)
#include "rr_common.cuh"
#include "rr_tc.cuh"

namespace rr {

using namespace tc;

constexpr int PB_A_BYTES = 128 * 128;
constexpr int PB_B_BYTES = 256 * 128;

__global__ void __launch_bounds__(128, 1)
tc_accum_probe_kernel(const __half* __restrict__ Ahi, const __half* __restrict__ Alo,
                      const __half* __restrict__ Bhi, const __half* __restrict__ Blo,
                      int reps, int mode, float* __restrict__ Dout,
                      float* __restrict__ Aux) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sah = base;
  uint8_t* sal = sah + PB_A_BYTES;
  uint8_t* sbh = sal + PB_A_BYTES;
  uint8_t* sbl = sbh + PB_B_BYTES;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  for (int c = 0; c < 8; ++c) {
    st_shared_v4(smem_u32(sah) + sw128_off(tid, c),
                 *reinterpret_cast<const uint4*>(Ahi + tid * 64 + c * 8));
    st_shared_v4(smem_u32(sal) + sw128_off(tid, c),
                 *reinterpret_cast<const uint4*>(Alo + tid * 64 + c * 8));
    for (int h = 0; h < 2; ++h) {
      const int row = tid + 128 * h;
      st_shared_v4(smem_u32(sbh) + sw128_off(row, c),
                   *reinterpret_cast<const uint4*>(Bhi + row * 64 + c * 8));
      st_shared_v4(smem_u32(sbl) + sw128_off(row, c),
                   *reinterpret_cast<const uint4*>(Blo + row * 64 + c * 8));
    }
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, 256);
    const uint64_t dah = make_desc_sw128(smem_u32(sah)), dal = make_desc_sw128(smem_u32(sal));
    const uint64_t dbh = make_desc_sw128(smem_u32(sbh)), dbl = make_desc_sw128(smem_u32(sbl));
    for (int r = 0; r < reps; ++r) {
      for (int k = 0; k < 4; ++k) {
        const uint64_t adv = (uint64_t)(2 * k);
        umma_f16_ss(tmem, dah + adv, dbh + adv, idesc, (r | k) != 0);
        if (mode == 1) {
          umma_f16_ss(tmem, dal + adv, dbh + adv, idesc, 1);
          umma_f16_ss(tmem, dah + adv, dbl + adv, idesc, 1);
        } else if (mode == 2) {
          umma_f16_ss(tmem + 256, dal + adv, dbh + adv, idesc, (r | k) != 0);
          umma_f16_ss(tmem + 256, dah + adv, dbl + adv, idesc, 1);
        }
      }
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  for (int c0 = 0; c0 < 256; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, v);
    for (int r = 0; r < 32; ++r) Dout[(32 * warp + lane) * 256 + c0 + r] = v[r];
    if (mode == 2) {
      tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)(256 + c0), v);
      for (int r = 0; r < 32; ++r) Aux[(32 * warp + lane) * 256 + c0 + r] = v[r];
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace rr

// Host pointers in and out (diagnostic entry point, not on the product path).
extern "C" int rr_tcgen05_accum_probe(const uint16_t* Ahi, const uint16_t* Alo,
                                      const uint16_t* Bhi, const uint16_t* Blo,
                                      int32_t reps, int32_t mode, float* D,
                                      float* Aux) {
  using namespace rr;
  RR_REQUIRE(Ahi && Alo && Bhi && Blo && D && Aux, "null pointer");
  RR_REQUIRE(reps > 0 && mode >= 0 && mode <= 2, "bad reps/mode");
  const size_t na = 128 * 64 * 2, nb = 256 * 64 * 2, nd = 128 * 256 * 4;
  uint8_t* dev = nullptr;
  RR_CUDA_CHECK(cudaMalloc(&dev, 2 * na + 2 * nb + 2 * nd));
  uint8_t *dah = dev, *dal = dah + na, *dbh = dal + na, *dbl = dbh + nb;
  float *dD = (float*)(dbl + nb), *dX = dD + 128 * 256;
  int rc = RR_OK;
  cudaError_t e;
#define PB_CHECK(x) do { e = (x); if (e != cudaSuccess) { set_error("probe: %s: %s", #x, cudaGetErrorString(e)); rc = RR_ERR_CUDA; goto done; } } while (0)
  PB_CHECK(cudaMemcpy(dah, Ahi, na, cudaMemcpyHostToDevice));
  PB_CHECK(cudaMemcpy(dal, Alo, na, cudaMemcpyHostToDevice));
  PB_CHECK(cudaMemcpy(dbh, Bhi, nb, cudaMemcpyHostToDevice));
  PB_CHECK(cudaMemcpy(dbl, Blo, nb, cudaMemcpyHostToDevice));
  PB_CHECK(cudaMemset(dD, 0, 2 * nd));
  {
    const size_t smem = 2 * PB_A_BYTES + 2 * PB_B_BYTES + 1024;
    PB_CHECK(cudaFuncSetAttribute(tc_accum_probe_kernel,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    tc_accum_probe_kernel<<<1, 128, smem>>>((const __half*)dah, (const __half*)dal,
                                            (const __half*)dbh, (const __half*)dbl,
                                            reps, mode, dD, dX);
    count_launch();
    PB_CHECK(cudaGetLastError());
    PB_CHECK(cudaDeviceSynchronize());
  }
  PB_CHECK(cudaMemcpy(D, dD, nd, cudaMemcpyDeviceToHost));
  PB_CHECK(cudaMemcpy(Aux, dX, nd, cudaMemcpyDeviceToHost));
done:
#undef PB_CHECK
  cudaFree(dev);
  return rc;
}
