// Semantics probe: tcgen05.mma (kind::f16) with an MN-major, no-swizzle B operand on sm_100a.
//
// Why: the c8t staging layout [chunk(8 channels)][time row][8 ch] is K-major for the conv (K = channels), but for a
// contraction over TIME (a banded Toeplitz FIR: U[rows_out, ch] = G[rows_out, t] * X[t, ch]) the same bytes are an
// MN-major B operand (N = channels contiguous).  If that works, the anti-alias FIRs could run on the tensor pipe
// straight from the staged tile.  The probe checks which of (LBO, SBO) is the stride between 8-element N groups and
// which the stride between 8-row K groups, with the instruction descriptor's b_major bit (bit 16) set.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/umma_probe3 tools/umma_probe3.cu
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
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;     // descriptor version, no swizzle, base offset 0
  return d;
}

constexpr int M = 128, N = 32, K = 16;
constexpr int BROWS = 24;          // K-row pitch of a channel chunk in the B staging (rows x 16 B), > K on purpose

// variant 0: LBO = stride between N groups (chunk pitch), SBO = stride between 8-row K groups (128 B)
// variant 1: the other way round
__global__ void __launch_bounds__(128) probe(float* out, int variant) {
  __shared__ __align__(128) uint8_t A[2 * M * 16];               // K-major no-swizzle: [kchunk][row][8]
  __shared__ __align__(128) uint8_t B[(N / 8) * BROWS * 16];     // MN-major: [n chunk][k row][8 n]
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int tid = threadIdx.x;
  for (int i = tid; i < M * K; i += 128) {
    const int row = i / K, k = i % K;
    const float v = (float)(((row * 7 + k * 3) % 17) - 8) * 0.125f;
    reinterpret_cast<__nv_bfloat16*>(A)[((k / 8) * M + row) * 8 + (k % 8)] = __float2bfloat16(v);
  }
  for (int i = tid; i < (N / 8) * BROWS * 8; i += 128) reinterpret_cast<__nv_bfloat16*>(B)[i] = __float2bfloat16(100.f);   // poison
  __syncthreads();
  for (int i = tid; i < K * N; i += 128) {
    const int k = i / N, n = i % N;
    const float v = (float)(((k * 5 + n * 11) % 13) - 6) * 0.25f;
    reinterpret_cast<__nv_bfloat16*>(B)[((n / 8) * BROWS + k) * 8 + (n % 8)] = __float2bfloat16(v);
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tptr;
  if (tid == 0) {
    // D fp32, A/B bf16, A K-major, B MN-major (bit 16), N, M
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint64_t ad = desc(smem_u32(A), M * 16, 128);
    const uint32_t chunk_pitch = BROWS * 16, kgroup = 128;
    const uint64_t bd = variant == 0 ? desc(smem_u32(B), chunk_pitch, kgroup) : desc(smem_u32(B), kgroup, chunk_pitch);
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[32];
  const uint32_t taddr = tmem + ((uint32_t)((tid >> 5) * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
      "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int n = 0; n < N; ++n) out[tid * N + n] = __uint_as_float(v[n]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
  }
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  float* d_out;
  cudaMalloc(&d_out, M * N * sizeof(float));
  std::vector<float> ref(M * N, 0.f), got(M * N);
  for (int row = 0; row < M; ++row)
    for (int n = 0; n < N; ++n) {
      float s = 0.f;
      for (int k = 0; k < K; ++k)
        s += bf((float)(((row * 7 + k * 3) % 17) - 8) * 0.125f) * bf((float)(((k * 5 + n * 11) % 13) - 6) * 0.25f);
      ref[row * N + n] = s;
    }
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(d_out, 0, M * N * sizeof(float));
    probe<<<1, 128>>>(d_out, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(got.data(), d_out, M * N * sizeof(float), cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int i = 0; i < M * N; ++i) maxerr = fmax(maxerr, fabs((double)got[i] - ref[i]));
    printf("MN-major B, no swizzle, variant %d (%s): max |D - ref| = %.4g  -> %s\n", variant,
           variant == 0 ? "LBO = N-group (chunk) pitch, SBO = K-group stride 128 B" : "LBO = K-group stride 128 B, SBO = N-group (chunk) pitch",
           maxerr, maxerr < 1e-3 ? "MATCH" : "mismatch");
  }
  return 0;
}
