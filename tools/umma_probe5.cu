// Probe (round 2): what does a narrow conv tcgen05.mma cost when it shares the SM with the FIR MMAs and with shared-memory /
// TMEM traffic?  actconv_tc_kernel measures ~90-150 cycles per conv MMA (M = 128, K = 16, N = 32 / 48 / 96) where the
// single-stream floor is 40-56 (profiles/r02_umma_probe4.txt).  One CTA, selectable actors (bit mask `act`):
//   1  conv issuer   (warp 0): SS, A K-major [kchunk][row][8] with tap row shifts, B K-major weights, accumulate chain, N = arg
//   2  up issuer     (warp 1): SS, A MN-major bf16, B bf16 hi/lo, N = 64, groups of 6
//   4  down issuer   (warp 2): TS, A fp16 pairs in TMEM, B fp16, N = 32, groups of 6
//   8  smem writers  (warps 4-7): st.shared.v4 streams (the stmatrix / patch traffic)
//   16 TMEM readers  (warps 8-11): tcgen05.ld.x16 loops (snake / store / epilogue traffic)
//   32 all three issuers run from ONE warp, interleaved 6 up / 6 down / n conv (in-order single stream)
//   64 issuers wait for their own previous group (commit + mbarrier) before the next: latency-exposed issue
// Prints cycles per MMA per issuer.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/umma_probe5 tools/umma_probe5.cu
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
template <uint32_t AHI, uint32_t BHI>
__device__ __forceinline__ void umma_ss_u(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
  asm volatile(
      "{\n.reg .pred p, e;\n.reg .b64 da, db;\n"
      "mov.b64 da, {%1, %4};\nmov.b64 db, {%2, %5};\n"
      "setp.ne.b32 p, 1, 0;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n}\n"
      ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "n"(AHI), "n"(BHI) : "memory");
}
template <uint32_t BHI>
__device__ __forceinline__ void umma_ts_u(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc) {
  asm volatile(
      "{\n.reg .pred p, e;\n.reg .b64 db;\n"
      "mov.b64 db, {%2, %4};\n"
      "setp.ne.b32 p, 1, 0;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n}\n"
      ::"r"(d), "r"(a_tmem), "r"(b_lo), "r"(idesc), "n"(BHI) : "memory");
}
__device__ __forceinline__ void commit_elect(uint64_t* bar) {
  asm volatile("{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\n"
               "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int kSR = 144;        // A-stage rows per chunk panel
constexpr int kXR = 80;         // x-stage rows per chunk slot

__global__ void __launch_bounds__(512, 1) probe(long long* out, int act, int N, int groups, int nconv, volatile int* stop) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* Astage = smem;                          // 3 x 12 panels x 144 rows x 16 B = 82944
  uint8_t* W = smem + 82944;                       // 55296 (6 slots of 9216)
  uint8_t* X = W + 55296;                          // 4 x 15360 = 61440
  uint8_t* taps = X + 61440;                       // 18432
  uint64_t* bar = reinterpret_cast<uint64_t*>(taps + 18432);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 8);
  uint8_t* scratch = reinterpret_cast<uint8_t*>(tptr + 4);   // 4096 B for the smem writers
  const int tid = threadIdx.x;
  for (int i = tid; i < (82944 + 55296 + 61440 + 18432) / 4; i += 512) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = *tptr;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const bool single = act & 32, latency = act & 64;

  const uint32_t idesc_c = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t idesc_up = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t idesc_dn = (1u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t a_base = (smem_u32(Astage) >> 4) | ((uint32_t)kSR << 16);
  const uint32_t w_base = (smem_u32(W) >> 4) | ((uint32_t)N << 16);
  const uint32_t x_base = (smem_u32(X) >> 4) | (8u << 16);
  const uint32_t uph = (smem_u32(taps) >> 4) | (64u << 16), upl = uph + 384u;
  const uint32_t dnb = ((smem_u32(taps) + 12288) >> 4) | (32u << 16);

  auto conv_group = [&](int g) {                   // nconv MMAs: taps x k-steps over one accumulator
    const uint32_t d = tm + 320u + (uint32_t)((g & 1) * N);
    const uint32_t a_st = a_base + (uint32_t)((g % 3) * (12 * kSR));
    for (int m = 0; m < nconv; ++m) {
      const int tp = m >> 2, ks = m & 3;
      umma_ss_u<0x4008u, 0x4008u>(d, a_st + 5u + (uint32_t)tp + (uint32_t)(ks * 2 * kSR), w_base + (uint32_t)((m % 24) * 2 * N), idesc_c);
    }
  };
  auto up_group = [&](int g) {
    const uint32_t d = tm + (uint32_t)((g & 1) * 64);
    const uint32_t a0 = x_base + (uint32_t)((g & 3) * (12 * kXR)) + (uint32_t)((g & 1) * 32);
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      umma_ss_u<0x4000u | kXR, 0x4008u>(d, a0 + 16u * s, uph + 128u * s, idesc_up);
      umma_ss_u<0x4000u | kXR, 0x4008u>(d, a0 + 16u * s, upl + 128u * s, idesc_up);
    }
  };
  auto dn_group = [&](int g) {
    const uint32_t d = tm + 256u + (uint32_t)((g & 1) * 32);
#pragma unroll
    for (int s = 0; s < 6; ++s) umma_ts_u<0x4008u>(d, tm + 128u + (uint32_t)(((g + (s > 4)) & 3) * 32 + (s % 4) * 8), dnb + 64u * s, idesc_dn);
  };

  if (single && warp == 0) {
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      if (act & 2) up_group(g);
      if (act & 4) dn_group(g);
      if ((act & 1) && (g & 3) == 3) conv_group(g >> 2);
      if (latency) { commit_elect(&bar[0]); mbar_wait(&bar[0], g & 1); }
    }
    if (!latency) { commit_elect(&bar[0]); mbar_wait(&bar[0], 0); }
    if (tid == 0) { out[0] = clock64() - t0; *stop = 1; }
  } else if (!single && warp == 0 && (act & 1)) {
    const long long t0 = clock64();
    for (int g = 0; g < groups / 4; ++g) {
      conv_group(g);
      if (latency) { commit_elect(&bar[0]); mbar_wait(&bar[0], g & 1); }
    }
    if (!latency) { commit_elect(&bar[0]); mbar_wait(&bar[0], 0); }
    if (tid == 0) { out[0] = clock64() - t0; *stop = 1; }
  } else if (!single && warp == 1 && (act & 2)) {
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      up_group(g);
      if (latency) { commit_elect(&bar[1]); mbar_wait(&bar[1], g & 1); }
    }
    if (!latency) { commit_elect(&bar[1]); mbar_wait(&bar[1], 0); }
    if (tid == 32) { out[1] = clock64() - t0; if (!(act & 1)) *stop = 1; }
  } else if (!single && warp == 2 && (act & 4)) {
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      dn_group(g);
      if (latency) { commit_elect(&bar[2]); mbar_wait(&bar[2], g & 1); }
    }
    if (!latency) { commit_elect(&bar[2]); mbar_wait(&bar[2], 0); }
    if (tid == 64) { out[2] = clock64() - t0; if (!(act & 3)) *stop = 1; }
  } else if (warp >= 4 && warp < 8 && (act & 8)) {
    // smem writers: each warp streams 16-byte stores over a 1 KB window until the issuers are done
    uint4* p = reinterpret_cast<uint4*>(scratch) + (warp - 4) * 64 + (tid & 31);
    long long n = 0;
    while (!*stop) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { p[(i & 1) * 32] = make_uint4(i, i, i, i); }
      n += 16;
    }
    if ((tid & 31) == 0) out[4 + warp - 4] = n;
  } else if (warp >= 8 && warp < 12 && (act & 16)) {
    const uint32_t ta = tm + ((uint32_t)((warp & 3) * 32) << 16);
    long long n = 0;
    uint32_t acc = 0;
    while (!*stop) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(ta + (uint32_t)(i * 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += r[0] + r[15];
      }
      n += 4;
    }
    if ((tid & 31) == 0) out[8 + warp - 8] = n + (acc == 0x12345u);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

int main(int argc, char** argv) {
  long long* out;
  int* stop;
  cudaMalloc(&out, 16 * 8);
  cudaMalloc(&stop, 4);
  const size_t smem = 82944 + 55296 + 61440 + 18432 + 64 + 16 + 4096 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int groups = 2048;          // FIR groups (blocks); one conv group per 4 of them
  struct Cfg { int act, N, nconv; const char* name; };
  const Cfg cfgs[] = {
      {1, 96, 66, "conv alone N=96 (66 per tile)"}, {1, 48, 66, "conv alone N=48"}, {1, 32, 88, "conv alone N=32 (88 per tile)"},
      {2, 96, 0, "up alone"}, {4, 96, 0, "down alone"}, {6, 96, 0, "up + down, two warps"},
      {7, 96, 66, "up + down + conv N=96 k=11, three warps"}, {7, 96, 18, "up + down + conv N=96 k=3, three warps"},
      {7, 32, 88, "up + down + conv N=32 k=11, three warps"}, {7, 48, 66, "up + down + conv N=48 k=11, three warps"},
      {32 | 7, 96, 66, "up + down + conv N=96 k=11, ONE warp"}, {32 | 7, 96, 18, "up + down + conv N=96 k=3, ONE warp"},
      {32 | 6, 96, 0, "up + down, ONE warp"},
      {7 | 8, 96, 66, "three warps + smem writers"}, {7 | 16, 96, 66, "three warps + TMEM readers"}, {7 | 24, 96, 66, "three warps + both"},
      {1 | 8, 96, 66, "conv alone + smem writers"}, {1 | 16, 96, 66, "conv alone + TMEM readers"},
      {64 | 1, 96, 66, "conv alone, wait per tile"}, {64 | 2, 96, 0, "up alone, wait per block"}, {64 | 4, 96, 0, "down alone, wait per block"},
      {64 | 7, 96, 66, "three warps, each waits per group"},
  };
  for (const Cfg& c : cfgs) {
    cudaMemset(out, 0, 16 * 8);
    cudaMemset(stop, 0, 4);
    probe<<<1, 512, smem>>>(out, c.act, c.N, groups, c.nconv, stop);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
    long long h[16];
    cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
    const double nconv = (double)(groups / 4) * c.nconv, nfir = (double)groups * 6;
    printf("%-48s:", c.name);
    if (c.act & 32) {
      const double n = ((c.act & 1) ? nconv : 0) + ((c.act & 2) ? nfir : 0) + ((c.act & 4) ? nfir : 0);
      printf(" total %8lld cycles, %.1f per MMA, %.0f per tile (4 blocks)", h[0], h[0] / n, h[0] / (groups / 4.0));
    } else {
      if (c.act & 1) printf(" conv %.1f/MMA (%.0f per tile)", h[0] / nconv, h[0] / (groups / 4.0));
      if (c.act & 2) printf(" up %.1f/MMA (%.0f per 4 blocks)", h[1] / nfir, h[1] / (groups / 4.0));
      if (c.act & 4) printf(" down %.1f/MMA (%.0f per 4 blocks)", h[2] / nfir, h[2] / (groups / 4.0));
    }
    if (c.act & 8) printf("  [smem st.v4 warp-instr: %lld]", h[4] + h[5] + h[6] + h[7]);
    if (c.act & 16) printf("  [tmem ld.x16 warp-instr: %lld]", h[8] + h[9] + h[10] + h[11]);
    printf("\n");
  }
  return 0;
}
