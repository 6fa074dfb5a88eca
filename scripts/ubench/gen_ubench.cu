// Microbenchmark of the value-pass generator body in isolation: W warps per CTA
// (one CTA per SM) each run the 16-row trig + split + pack (+ optional smem
// stores) loop ITER times on register-resident inputs.  Prints cycles per
// "slab" (= one pass of every warp) for several code variants so that the
// SM sub-partition throughput of each instruction mix is known.
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint64_t f2_pack(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void f2_unpack(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t f2_sub(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint32_t f2_to_h2(uint64_t v) { float a, b; f2_unpack(v, a, b); __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ uint32_t pk_h2(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) { asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }

constexpr float RM = 12582912.0f, GM = 196608.0f, TP = 6.283185307179586f;

// MODE bit0: MUFU on; bit1: STS on; bit2: packed fp32x2 (else scalar); bit3: B tile (3 images)
template <int MODE, int SPIN>
__global__ void __launch_bounds__(1024, 1) gen_kernel(const float* __restrict__ in, float* __restrict__ out, int iters, long long* cyc, int gen_warps) {
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)), "r"(gen_warps * 32));
  }
  __syncthreads();
  if ((int)(threadIdx.x >> 5) >= gen_warps) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(&bar);
    uint32_t ok = 0;
    while (!ok) {
      if (SPIN == 2)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a), "r"(0), "r"(1000000u) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a), "r"(0) : "memory");
      if (SPIN == 3 && !ok) __nanosleep(500);
    }
    return;
  }
  extern __shared__ uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float u[16];
  for (int i = 0; i < 16; ++i) u[i] = in[(blockIdx.x * blockDim.x + tid) * 16 + i];
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t row = (uint32_t)((warp & 3) * 32 + lane);
  const uint32_t off0 = row * 128u + ((((warp >> 2) * 2 + 0) ^ (row & 7u)) << 4);
  const uint32_t off1 = row * 128u + ((((warp >> 2) * 2 + 1) ^ (row & 7u)) << 4);
  uint32_t acc = 0;
  asm volatile("bar.sync 1, %0;" ::"r"(gen_warps * 32));
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint4 hc, rc, hs, rs, cf, sf;
    uint32_t *hcp = (uint32_t*)&hc, *rcp = (uint32_t*)&rc, *hsp = (uint32_t*)&hs, *rsp = (uint32_t*)&rs, *cfp = (uint32_t*)&cf, *sfp = (uint32_t*)&sf;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float c0, s0, c1, s1;
      if (MODE & 4) {
        const uint64_t RM2 = f2_pack(RM, RM), TP2 = f2_pack(TP, TP);
        const uint64_t uu = f2_pack(u[2 * i], u[2 * i + 1]);
        const uint64_t kk = f2_sub(f2_add(uu, RM2), RM2);
        const uint64_t ang = f2_mul(f2_sub(uu, kk), TP2);
        float a0, a1;
        f2_unpack(ang, a0, a1);
        if (MODE & 1) { c0 = __cosf(a0); s0 = __sinf(a0); c1 = __cosf(a1); s1 = __sinf(a1); }
        else { c0 = a0 * 0.1f; s0 = a0 * 0.2f; c1 = a1 * 0.1f; s1 = a1 * 0.2f; }
        const uint64_t GM2 = f2_pack(GM, GM);
        const uint64_t c2 = f2_pack(c0, c1), s2 = f2_pack(s0, s1);
        const uint64_t h_c = f2_sub(f2_add(c2, GM2), GM2), h_s = f2_sub(f2_add(s2, GM2), GM2);
        hcp[i & 3] = f2_to_h2(h_c); hsp[i & 3] = f2_to_h2(h_s);
        rcp[i & 3] = f2_to_h2(f2_sub(c2, h_c)); rsp[i & 3] = f2_to_h2(f2_sub(s2, h_s));
        if (MODE & 8) { cfp[i & 3] = f2_to_h2(c2); sfp[i & 3] = f2_to_h2(s2); }
      } else {
        float a[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const float uu = u[2 * i + k];
          const float kk = __fsub_rn(__fadd_rn(uu, RM), RM);
          a[k] = __fsub_rn(uu, kk) * TP;
        }
        if (MODE & 1) { c0 = __cosf(a[0]); s0 = __sinf(a[0]); c1 = __cosf(a[1]); s1 = __sinf(a[1]); }
        else { c0 = a[0] * 0.1f; s0 = a[0] * 0.2f; c1 = a[1] * 0.1f; s1 = a[1] * 0.2f; }
        const float hc0 = __fsub_rn(__fadd_rn(c0, GM), GM), hc1 = __fsub_rn(__fadd_rn(c1, GM), GM);
        const float hs0 = __fsub_rn(__fadd_rn(s0, GM), GM), hs1 = __fsub_rn(__fadd_rn(s1, GM), GM);
        hcp[i & 3] = pk_h2(hc0, hc1); hsp[i & 3] = pk_h2(hs0, hs1);
        rcp[i & 3] = pk_h2(c0 - hc0, c1 - hc1); rsp[i & 3] = pk_h2(s0 - hs0, s1 - hs1);
        if (MODE & 8) { cfp[i & 3] = pk_h2(c0, c1); sfp[i & 3] = pk_h2(s0, s1); }
      }
      if ((i & 3) == 3) {
        const uint32_t o = sbase + ((i >> 2) ? off1 : off0);
        if (MODE & 2) {
          st_shared_v4(o, hc); st_shared_v4(o + 8192, hs); st_shared_v4(o + 32768, rc); st_shared_v4(o + 40960, rs);
          if (MODE & 8) { st_shared_v4(o + 65536, cf); st_shared_v4(o + 73728, sf); }
        } else {
          acc ^= hc.x ^ hc.y ^ hc.z ^ hc.w ^ hs.x ^ hs.y ^ hs.z ^ hs.w ^ rc.x ^ rc.y ^ rc.z ^ rc.w ^ rs.x ^ rs.y ^ rs.z ^ rs.w;
          if (MODE & 8) acc ^= cf.x ^ cf.y ^ cf.z ^ cf.w ^ sf.x ^ sf.y ^ sf.z ^ sf.w;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) u[i] += 0.37f;   // new inputs every pass
    if (SPIN == 4) asm volatile("bar.sync 1, %0;" ::"r"(gen_warps * 32));          // lock step
    if (SPIN == 5) asm volatile("bar.sync %0, 128;" ::"r"(2 + (warp >> 2)));         // per group of 4 warps
    if (SPIN == 6) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); asm volatile("bar.sync 1, %0;" ::"r"(gen_warps * 32)); }
  }
  const long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
  out[blockIdx.x * blockDim.x + tid] = __uint_as_float(acc) + u[3];
}

template <int MODE, int SPIN = 0>
void run(int warps, const float* in, float* out, long long* cyc, const char* name) {
  const int iters = 2000;
  const int tot = warps + ((SPIN >= 1 && SPIN <= 3) ? 4 : 0);
  cudaFuncSetAttribute(gen_kernel<MODE, SPIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  gen_kernel<MODE, SPIN><<<148, tot * 32, 96 * 1024>>>(in, out, 10, cyc, warps);
  gen_kernel<MODE, SPIN><<<148, tot * 32, 96 * 1024>>>(in, out, iters, cyc, warps);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < 148; ++i) s += (double)h[i];
  printf("%-34s warps=%2d  cycles/pass=%8.1f  (%s)\n", name, warps, s / 148 / iters, cudaGetErrorString(e));
}

int main() {
  float *in, *out;
  long long* cyc;
  cudaMalloc(&in, 148 * 1024 * 16 * 4);
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  cudaMemset(in, 0, 148 * 1024 * 16 * 4);
  run<4 | 1 | 2 | 8, 4>(16, in, out, cyc, "packed mufu sts B lock-step 16");
  run<4 | 1 | 2 | 8, 5>(16, in, out, cyc, "packed mufu sts B lock-step groups");
  run<4 | 1 | 2 | 8, 6>(16, in, out, cyc, "packed mufu sts B fence+lockstep");
  run<4 | 1 | 2 | 8, 4>(8, in, out, cyc, "packed mufu sts B lock-step 8w");
  for (int w : {16}) {
    run<4 | 1 | 2 | 8>(w, in, out, cyc, "packed mufu sts B");
    run<4 | 1 | 2>(w, in, out, cyc, "packed mufu sts A");
    run<4 | 1>(w, in, out, cyc, "packed mufu nosts A");
    run<4 | 2>(w, in, out, cyc, "packed nomufu sts A");
    run<4>(w, in, out, cyc, "packed nomufu nosts A");
    run<1 | 2>(w, in, out, cyc, "scalar mufu sts A");
    run<1>(w, in, out, cyc, "scalar mufu nosts A");
    run<0>(w, in, out, cyc, "scalar nomufu nosts A");
  }
  return 0;
}
