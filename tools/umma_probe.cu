// Microbenchmark / semantics probe for tcgen05.mma shared-memory operand descriptors on sm_100a.
//   1. Does a K-major SWIZZLE_128B A operand work when its start address is shifted by r rows
//      (r*128 B, not a multiple of the 1024 B swizzle atom)?  With base_offset = 0 or (addr>>7)&7?
//   2. Cycles per tcgen05.mma for no-swizzle vs 128B-swizzle operand layouts at N = 32/96/256.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe tools/umma_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

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
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// layout: 0 = none (interleave), 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}

constexpr int AROWS = 160;   // rows available in the A region (128 + shifts)

// mode: 0 correctness (shift r, base_off variant), 1 timing
__global__ void __launch_bounds__(128) probe(float* out, long long* cycles, int r, int use_base_off, int layout, int N,
                                             int timing_iters, int arows = AROWS, int brows = 256, int alt = 0, int spin = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* A = smem;                       // up to 40 KB
  uint8_t* Bm = smem + 40 * 1024;          // 256 x 128 B (+ extra weight tiles behind it)
  uint64_t* bar = reinterpret_cast<uint64_t*>(Bm + 256 * 128 + 3 * 96 * 128);
  uint64_t* bar2 = bar + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x;
  // fill A[row][k] = small ints, B[n][k]
  for (int i = tid; i < AROWS * 64; i += 128) {
    const int row = i / 64, k = i % 64;
    const float v = (float)(((row * 7 + k * 3) % 17) - 8) * 0.125f;
    uint32_t off;
    if (layout == 2) off = row * 128 + (((k >> 3) ^ (row & 7)) << 4) + (k & 7) * 2;   // absolute-address 128B swizzle
    else off = (k >> 3) * (arows * 16) + row * 16 + (k & 7) * 2;                      // [kchunk][row][8]
    if (row < arows) *reinterpret_cast<__nv_bfloat16*>(A + off) = __float2bfloat16(v);
  }
  for (int i = tid; i < 256 * 64; i += 128) {
    const int row = i / 64, k = i % 64;
    const float v = (float)(((row * 5 + k * 11) % 13) - 6) * 0.25f;
    uint32_t off;
    if (layout == 2) off = row * 128 + (((k >> 3) ^ (row & 7)) << 4) + (k & 7) * 2;
    else off = (k >> 3) * (brows * 16) + row * 16 + (k & 7) * 2;
    if (row < brows) *reinterpret_cast<__nv_bfloat16*>(Bm + off) = __float2bfloat16(v);
  }
  if (tid == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = *tptr;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  if (tid == 0) {
    const uint32_t a0 = smem_u32(A), b0 = smem_u32(Bm);
    long long t0 = clock64();
    const int iters = timing_iters > 0 ? timing_iters : 1;
    for (int it = 0; it < iters; ++it) {
      for (int k2 = 0; k2 < 4; ++k2) {     // K = 64 = 4 x 16
        uint64_t ad, bd;
        if (layout == 2) {
          const uint32_t aaddr = a0 + r * 128 + k2 * 32;
          const uint32_t bo = use_base_off ? ((aaddr >> 7) & 7) : 0;
          ad = desc(aaddr, 16, 1024, 2, bo);
          bd = desc(b0 + k2 * 32, 16, 1024, 2, 0);
        } else {
          // spin >= 10: descriptors that change every MMA (shift and weight slot vary with `it`), like the conv kernel
          if (spin == 20) {
            // conv-like address stream: A = 258-row x 64-ch tile (4 k-steps), two 128-row sub-tiles, 3 taps;
            // B = one 96 x 64 weight tile per tap
            const int tap = it % 3, ms = (it / 3) & 1;
            ad = desc(a0 + (ms * 128 + tap) * 16 + k2 * 2 * 258 * 16, 258 * 16, 128, 0, 0);
            bd = desc(b0 + tap * (96 * 128) + k2 * 2 * 96 * 16, 96 * 16, 128, 0, 0);
          } else if (spin >= 10) {
            const int rr = r + (it & 3);
            const int bsh = (it & 1) * 16;
            ad = desc(a0 + rr * 16 + k2 * 2 * arows * 16, arows * 16, 128, 0, 0);
            bd = desc(b0 + bsh + k2 * 2 * brows * 16, brows * 16, 128, 0, 0);
          } else {
            ad = desc(a0 + r * 16 + k2 * 2 * arows * 16, arows * 16, 128, 0, 0);
            bd = desc(b0 + k2 * 2 * brows * 16, brows * 16, 128, 0, 0);
          }
        }
        // alt: alternate between two accumulators every 4 MMAs, like the conv kernel's two time sub-tiles
        umma(tm + ((alt == 1 && (it & 1)) ? 128u : 0u) + (alt >= 2 ? (uint32_t)alt : 0u), ad, bd, idesc, (it > 1 || k2) ? 1u : 0u);
      }
    }
    commit(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    if (cycles) *cycles = t1 - t0;
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar2)) : "memory");
  } else if (spin == 1 && tid >= 32) {
    mbar_wait(bar2, 0);                                   // 3 warps spin on try_wait like idle epilogue warps
  } else if (spin == 2 && tid >= 32 && (tid & 31) == 0) {
    mbar_wait(bar2, 0);                                   // one lane per warp spins
  } else if (spin == 3 && tid >= 32) {
    uint32_t done = 0;                                    // spin with nanosleep back-off
    while (!done) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}\n"
                   : "=r"(done) : "r"(smem_u32(bar2)), "r"(0) : "memory");
      if (!done) __nanosleep(200);
    }
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // D[m][n]: lane = m
  const int warp = tid >> 5, lane = tid & 31;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(tm + ((uint32_t)(warp * 32) << 16) + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 256 + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256));
}

static float bfr(float v) { return __bfloat162float(__float2bfloat16(v)); }

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 128 * 256 * 4);
  cudaMalloc(&cyc, 8);
  const size_t smem = 40 * 1024 + 256 * 128 + 3 * 96 * 128 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<float> h(128 * 256);
  printf("== correctness: D[m][n] = sum_k A[m+r][k] B[n][k], N=32, K=64\n");
  for (int layout : {0, 2})
    for (int bo = 0; bo < (layout == 2 ? 2 : 1); ++bo)
      for (int r = 0; r <= 9; ++r) {
        probe<<<1, 128, smem>>>(out, cyc, r, bo, layout, 32, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("layout %d bo %d r %d: CUDA error %s\n", layout, bo, r, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), out, h.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 32; ++n) {
            double ref = 0;
            for (int k = 0; k < 64; ++k)
              ref += (double)bfr((float)((((m + r) * 7 + k * 3) % 17) - 8) * 0.125f) * bfr((float)(((n * 5 + k * 11) % 13) - 6) * 0.25f);
            maxerr = fmax(maxerr, fabs(ref - h[m * 256 + n]));
          }
        printf("layout=%s base_off=%s shift r=%d : max err %.4g %s\n", layout == 2 ? "SW128" : "NONE ", bo ? "(addr>>7)&7" : "0",
               r, maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
      }
  printf("== timing: cycles per tcgen05.mma (M=128, K=16), 4000 MMAs back to back into one accumulator\n");
  for (int layout : {0, 2})
    for (int N : {32, 48, 96, 128, 256}) {
      probe<<<1, 128, smem>>>(out, cyc, 0, 0, layout, N, 1000);
      cudaDeviceSynchronize();
      long long c;
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("layout=%s N=%3d : %.1f cycles/MMA (ideal %d)\n", layout == 2 ? "SW128" : "NONE ", N, c / 4000.0, N / 2);
    }
  printf("== timing, NONE layout, conv-like geometry (A rows=arows -> LBO, shift r, B rows = N)\n");
  struct Cfg { int N, arows, r, alt; } cfgs[] = {{96, 160, 0, 0}, {96, 160, 1, 0}, {96, 160, 3, 0}, {96, 146, 0, 0}, {96, 146, 1, 0},
                                                 {96, 146, 1, 1}, {32, 146, 1, 1}, {256, 146, 1, 1}, {256, 146, 0, 0}};
  for (auto c : cfgs) {
    probe<<<1, 128, smem>>>(out, cyc, c.r, 0, 0, c.N, 1000, c.arows, c.N, c.alt);
    cudaDeviceSynchronize();
    long long cc;
    cudaMemcpy(&cc, cyc, 8, cudaMemcpyDeviceToHost);
    printf("N=%3d arows=%3d shift=%d alt_acc=%d : %.1f cycles/MMA\n", c.N, c.arows, c.r, c.alt, cc / 4000.0);
  }
  printf("== timing vs accumulator column offset (alt>=2 -> D starts at column alt)\n");
  for (int N : {96, 32, 48}) for (int col : {0, 32, 48, 64, 96, 128, 144, 192}) {
    if (col == 0) continue;
    probe<<<1, 128, smem>>>(out, cyc, 1, 0, 0, N, 1000, 146, N, col, 0);
    cudaDeviceSynchronize();
    long long cc;
    cudaMemcpy(&cc, cyc, 8, cudaMemcpyDeviceToHost);
    printf("N=%d D column offset %d : %.1f cycles/MMA\n", N, col, cc / 4000.0);
  }
  printf("== timing with per-MMA varying descriptors (N=96 / N=32 / N=256)\n");
  for (int N : {96, 32, 256}) {
    probe<<<1, 128, smem>>>(out, cyc, 1, 0, 0, N, 1000, 146, N, 1, 10);
    cudaDeviceSynchronize();
    long long cc;
    cudaMemcpy(&cc, cyc, 8, cudaMemcpyDeviceToHost);
    printf("varying descriptors N=%d : %.1f cycles/MMA\n", N, cc / 4000.0);
  }
  printf("== timing with a conv-like operand address stream (A 33 KB tile, 3 weight tiles)\n");
  for (int N : {96}) {
    probe<<<1, 128, smem>>>(out, cyc, 0, 0, 0, N, 1000, 258, N, 1, 20);
    cudaDeviceSynchronize();
    long long cc;
    cudaMemcpy(&cc, cyc, 8, cudaMemcpyDeviceToHost);
    printf("conv-like stream N=%d : %.1f cycles/MMA (err %s)\n", N, cc / 4000.0, cudaGetErrorString(cudaGetLastError()));
  }
  printf("== timing with 3 other warps spinning on an mbarrier (N=96 / N=32)\n");
  for (int sp = 0; sp <= 3; ++sp)
    for (int N : {96, 32}) {
      probe<<<1, 128, smem>>>(out, cyc, 1, 0, 0, N, 1000, 146, N, 1, sp);
      cudaDeviceSynchronize();
      long long cc;
      cudaMemcpy(&cc, cyc, 8, cudaMemcpyDeviceToHost);
      printf("spin mode %d (0 none, 1 all lanes try_wait, 2 one lane/warp, 3 try_wait+nanosleep) N=%d : %.1f cycles/MMA\n", sp, N, cc / 4000.0);
    }
  return 0;
}
