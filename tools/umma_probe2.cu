// Probe 2: tcgen05.mma issue/execute rate for table-driven descriptor streams (no ALU work between MMAs).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/umma_probe2 tools/umma_probe2.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__host__ __device__ inline uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

struct Prog { uint32_t a_off[64]; uint32_t b_off[64]; uint32_t d_col[64]; int n; uint32_t a_lbo, b_lbo; };

__global__ void __launch_bounds__(128) probe(const Prog P, int N, int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 200 * 1024);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  uint64_t* ta = reinterpret_cast<uint64_t*>(smem + 201 * 1024);
  uint64_t* tb = ta + 64;
  uint32_t* td = reinterpret_cast<uint32_t*>(tb + 64);
  const int tid = threadIdx.x;
  for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 0xff);
  if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = *tptr;
  const uint32_t base = smem_u32(smem);
  if (tid < P.n) {
    ta[tid] = desc(base + P.a_off[tid], P.a_lbo, 128);
    tb[tid] = desc(base + 100 * 1024 + P.b_off[tid], P.b_lbo, 128);
    td[tid] = tm + P.d_col[tid];
  }
  __syncthreads();
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  if (tid == 0) {
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int i = 0; i < P.n; ++i) umma(td[i], ta[i], tb[i], idesc, 1u);
    }
    commit(bar);
    mbar_wait(bar, 0);
    if (blockIdx.x == 0) *cycles = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

static int g_grid = 1;
static void run(const char* name, const Prog& P, int N, long long* cyc) {
  const size_t smem = 203 * 1024;
  probe<<<g_grid, 128, smem>>>(P, N, 200, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-64s N=%3d : %6.1f cycles/MMA %s\n", name, N, c / (200.0 * P.n), e == cudaSuccess ? "" : cudaGetErrorString(e));
}

// conv-like program: ntap taps x MT sub-tiles x nk k-steps; A tile XR rows/kchunk; B tiles of NB rows
static Prog conv_prog(int ntap, int MT, int nk, int XR, int NB, int dil, bool vary_b, bool vary_shift, bool vary_ms) {
  Prog P = {};
  int n = 0;
  for (int tp = 0; tp < ntap; ++tp)
    for (int ms = 0; ms < MT; ++ms)
      for (int k = 0; k < nk; ++k) {
        P.a_off[n] = ((vary_ms ? ms * 128 : 0) + (vary_shift ? tp * dil : 0)) * 16 + k * 2 * XR * 16;
        P.b_off[n] = (vary_b ? tp * NB * 128 : 0) + k * 2 * NB * 16;
        P.d_col[n] = ms * NB;
        ++n;
      }
  P.n = n;
  P.a_lbo = XR * 16;
  P.b_lbo = NB * 16;
  return P;
}

int main(int argc, char** argv) {
  if (argc > 1) g_grid = atoi(argv[1]);
  printf("grid = %d CTAs\n", g_grid);
  long long* cyc;
  cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 203 * 1024);
  run("same descriptors (1 entry)", conv_prog(1, 1, 1, 258, 96, 1, false, false, false), 96, cyc);
  run("k-steps only (4 entries)", conv_prog(1, 1, 4, 258, 96, 1, false, false, false), 96, cyc);
  run("k-steps + 2 sub-tiles", conv_prog(1, 2, 4, 258, 96, 1, false, false, true), 96, cyc);
  run("3 taps: shift only", conv_prog(3, 2, 4, 258, 96, 1, false, true, true), 96, cyc);
  run("3 taps: B tile only", conv_prog(3, 2, 4, 258, 96, 1, true, false, true), 96, cyc);
  run("3 taps: shift + B (conv-like, XR=258)", conv_prog(3, 2, 4, 258, 96, 1, true, true, true), 96, cyc);
  run("conv-like XR=256", conv_prog(3, 2, 4, 256, 96, 1, true, true, true), 96, cyc);
  run("conv-like XR=264", conv_prog(3, 2, 4, 264, 96, 1, true, true, true), 96, cyc);
  run("conv-like XR=272", conv_prog(3, 2, 4, 272, 96, 1, true, true, true), 96, cyc);
  run("conv-like XR=258 dil=3", conv_prog(3, 2, 4, 262, 96, 3, true, true, true), 96, cyc);
  run("conv-like XR=258 NB=32 MT=4 nk=2", conv_prog(3, 4, 2, 514, 32, 1, true, true, true), 32, cyc);
  run("conv-like XR=258 NB=256", conv_prog(3, 2, 4, 258, 256, 1, false, true, true), 256, cyc);
  run("conv-like XR=258 NB=128", conv_prog(3, 2, 4, 258, 128, 1, true, true, true), 128, cyc);
  return 0;
}
