// Self-test of the tcgen05 / TMEM building blocks shared by the tensor-core
// kernels (descriptor encodings, 128B-swizzled K-major operand layout written
// by threads, TMEM load mapping) against a host reference.
#include "rr_common.cuh"
#include "rr_tc.cuh"

namespace rr {

using namespace tc;

constexpr int TC_A_BYTES = 128 * 128;   // 128 rows x 64 fp16
constexpr int TC_B_BYTES = 256 * 128;   // 256 rows x 64 fp16

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
