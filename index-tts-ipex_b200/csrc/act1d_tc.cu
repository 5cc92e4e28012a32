// Activation1d on the tensor cores (c8t bf16 -> c8t bf16), sm_100a: both anti-alias FIRs as banded-Toeplitz
// tcgen05.mma over TIME, SnakeBeta on the CUDA cores between them.
//
// Reference semantics: alias_free_torch/act.py:24-29 = UpSample1d (resample.py:25-33) -> SnakeBeta
// (activations.py:109-122) -> DownSample1d (resample.py:46-49, filter.py:87-96); closed form in act1d.cu.
//
// Why: the 12 + 12-tap FIRs cost ~24 FMA per element on the FP32 pipe, which capped every stencil kernel of round 1 at
// ~2 elements / cycle / SM (0.35 of the HBM roofline with 2-byte I/O).  As matrix products over time they cost nothing
// on the CUDA cores; what is left there is the snake (MUFU-bound at 8 elements / cycle / SM) and data movement.
//
// Orientation (facts: profiles/r02_umma_probe4.txt).  TMEM lanes = channels, TMEM columns = time:
//   up    U^T[c, m]  = sum_t X^T[c, t] * Gup[t, m]    SS MMA: A = the staged c8t tile [lane-chunk][time row][8 ch] read as an
//                      MN-major operand (M = channels, K = time; LBO = 128 B, SBO = chunk pitch), B = Toeplitz taps (K-major,
//                      bf16 hi + lo parts: operand formats cannot be mixed, and bf16 taps alone would be a -48 dB filter error)
//   snake a = u + 1/(e^b + 1e-9) sin^2(e^a u), a thread owns one channel (lane) x 32 consecutive upsampled samples
//         (tcgen05.ld), per-channel constants in registers, result written back to TMEM as fp16 pairs (tcgen05.st)
//   down  Y^T[c, t]  = sum_m A^T[c, m] * Gdn[m, t]    TS MMA: A = the fp16 pairs in TMEM (never touches shared memory),
//                      B = Toeplitz taps fp16 (2^-12 relative: -70 dB), N = 32 output time steps
//   store Y^T (lane = channel, 32 time steps in registers) -> bf16 -> lane-pair shuffle -> [chunk][row][8 ch] tile in shared
//         memory (conflict-free 32-bit stores) -> TMA bulk store.
// Narrow tensors fill the 128 lanes with several time SEGMENTS of the same utterance (C = 24: 4 segments x 4 chunks).
// A CTA streams through a time range in 32-step blocks: up(i) | snake(i-1) | down(i-2) | store(i-3) run concurrently
// on the tensor pipe / 8 snake warps / 4 store warps; the activated signal of 4 consecutive blocks lives in a TMEM ring
// so that the down-FIR of block j reads its 5/6-sample halos from the neighbours (one extra block of up + snake at each
// end of a range is the whole halo cost).
//
// Edges.  Rows t < 0 / t >= T of a staged tile are overwritten with x[0] / x[T-1] in shared memory (replicate padding of
// the input, resample.py:28).  The replicate padding of the ACTIVATED signal (filter.py:90-92) only changes y[0..2] and
// y[T-3..T-1]; those six rows per utterance, the zero halo rows and the padding channels are (re)written afterwards by
// act1d_c8t_edge_kernel on the CUDA cores with the exact stencil of act1d_core.cuh.
#include <cuda_fp16.h>

#include <algorithm>

#include "act1d_core.cuh"
#include "bvg_common.cuh"
#include "umma.cuh"
#include "umma_ptx.cuh"

namespace bvg {
namespace {

constexpr int kTBlk = 32;                       // output time steps per block
constexpr int kTXRows = 144;                    // staged rows per lane-chunk and stage: 4 blocks + 8-row halo per side
constexpr int kTXStages = 2;
constexpr int kTOutPitch = 130;                 // out-stage row pitch per lane-chunk: == 2 (mod 8) -> conflict-free STS.32
constexpr int kTWStore0 = 0;                    // warps 0-3: store (TMEM lane quarter = warp % 4)
constexpr int kTWSnake0 = 4;                    // warps 4-11: snake (quarter = warp % 4, column half = (warp - 4) / 4)
constexpr int kTWProd = 12, kTWIssue = 13, kTWPatch = 14;
constexpr int kTThreads = 15 * 32;
constexpr uint32_t kTColU = 0, kTColA = 128, kTColY = 256;     // TMEM columns: U 2 x 64 | A ring 4 x 32 | Y 2 x 32

struct ActTcParams {
  const __nv_bfloat16* x; __nv_bfloat16* y;
  const float* alpha; const float* beta;
  int C, chunks, T, Tp, pad;
  int64_t bstride;
  int cps, sps, nseg, ntile;      // chunks per tile, lane-chunk slots per segment, segments per item, channel tiles
  int RL, NG;                     // rows per range, range groups per utterance
  int nitems;
};

struct TcItem { int b, tile, grp, nblk, cps_t; };
__device__ __forceinline__ TcItem tc_item(const ActTcParams& P, int item) {
  TcItem it;
  const int per_b = P.ntile * P.NG;
  it.b = item / per_b;
  const int r = item - it.b * per_b;
  it.tile = r / P.NG;
  it.grp = r - it.tile * P.NG;
  it.cps_t = min(P.cps, P.chunks - it.tile * P.cps);
  const int tr0 = it.grp * P.nseg * P.RL;                      // segment 0 (the longest of the item)
  it.nblk = (min(P.T, tr0 + P.RL) - tr0 + kTBlk - 1) / kTBlk;
  return it;
}

__device__ __forceinline__ uint32_t cvt_f16x2_sat(float lo, float hi) {
  // fp16 pair, finite-saturating (the reference itself runs this path under fp16 autocast, infer.py:194)
  lo = fminf(fmaxf(lo, -65504.f), 65504.f);
  hi = fminf(fmaxf(hi, -65504.f), 65504.f);
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

__global__ void __launch_bounds__(kTThreads, 1) act1d_tc_kernel(const ActTcParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  constexpr uint32_t kXStageBytes = 16u * kTXRows * 16u;          // 36864
  constexpr uint32_t kUpBytes = 6u * 64u * 16u;                   // one of (hi, lo): [kchunk 6][n 64][8] bf16
  constexpr uint32_t kDnBytes = 12u * 32u * 16u;                  // [kchunk 12][n 32][8] fp16
  constexpr uint32_t kOutBytes = 16u * kTOutPitch * 16u;          // 33280
  uint8_t* xsm = smem;
  uint8_t* up_hi = xsm + kTXStages * kXStageBytes;
  uint8_t* up_lo = up_hi + kUpBytes;
  uint8_t* dnm = up_lo + kUpBytes;
  uint8_t* osm = dnm + kDnBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(osm + 2 * kOutBytes);
  uint64_t* x_full = bars;            // [2]
  uint64_t* x_ready = bars + 2;       // [2]
  uint64_t* x_free = bars + 4;        // [2]
  uint64_t* u_full = bars + 6;        // [2]
  uint64_t* u_free = bars + 8;        // [2]
  uint64_t* a_full = bars + 10;       // [4]
  uint64_t* a_free = bars + 14;       // [4]
  uint64_t* y_full = bars + 18;       // [2]
  uint64_t* y_free = bars + 20;       // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 22);

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x_full[i], 1); mbar_init(&x_ready[i], 1); mbar_init(&x_free[i], 1);
      mbar_init(&u_full[i], 1); mbar_init(&u_free[i], 8);
      mbar_init(&y_full[i], 1); mbar_init(&y_free[i], 4);
    }
    for (int i = 0; i < 4; ++i) { mbar_init(&a_full[i], 8); mbar_init(&a_free[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kTWPatch) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  {
    // Toeplitz tap matrices (K-major, no swizzle: [kchunk][n][8]).
    //   up:   B[n = m_local][k]: coefficient of x[t0 - 8 + k] in u[2 t0 + m_local]; d = k - 8 - m_local / 2;
    //         even m: 2 f[5 - 2d] (d = -3..2), odd m: 2 f[6 - 2d] (d = -2..3)          (act1d.cu closed form)
    //   down: B[n = t_local][k]: coefficient of a[2 t0 - 16 + k] in y[t0 + t_local]: f[k - 2 n - 11]
    const float f[12] = {BVG_F0, BVG_F1, BVG_F2, BVG_F3, BVG_F4, BVG_F5, BVG_F5, BVG_F4, BVG_F3, BVG_F2, BVG_F1, BVG_F0};
    for (int idx = threadIdx.x; idx < 64 * 48; idx += kTThreads) {
      const int n = idx / 48, k = idx - n * 48;
      const int d = k - 8 - (n >> 1);
      const int ti = (n & 1) ? 6 - 2 * d : 5 - 2 * d;
      float g = 0.f;
#pragma unroll
      for (int q = 0; q < 12; ++q) if (q == ti) g = 2.f * f[q];
      const __nv_bfloat16 h = __float2bfloat16_rn(g);
      const __nv_bfloat16 l = __float2bfloat16_rn(g - __bfloat162float(h));
      const int off = ((k >> 3) * 64 + n) * 8 + (k & 7);
      reinterpret_cast<__nv_bfloat16*>(up_hi)[off] = h;
      reinterpret_cast<__nv_bfloat16*>(up_lo)[off] = l;
    }
    for (int idx = threadIdx.x; idx < 32 * 96; idx += kTThreads) {
      const int n = idx / 96, k = idx - n * 96;
      const int ti = k - 2 * n - 11;
      float g = 0.f;
#pragma unroll
      for (int q = 0; q < 12; ++q) if (q == ti) g = f[q];
      reinterpret_cast<__half*>(dnm)[((k >> 3) * 32 + n) * 8 + (k & 7)] = __float2half_rn(g);
    }
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == kTWProd) {
    // ===================== TMA producer: raw rows of every (segment, chunk) of the item, 128 + 16 rows per stage ==========
    if (lane == 0) {
      int xs = 0; uint32_t xph = 0;
      for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
        const TcItem it = tc_item(P, item);
        const int nstages = (it.nblk + 2 + 3) >> 2;
        const __nv_bfloat16* xb = P.x + (int64_t)it.b * P.bstride;
        for (int st = 0; st < nstages; ++st) {
          mbar_wait_relaxed(&x_free[xs], xph ^ 1);
          uint32_t total = 0;
          for (int s = 0; s < P.nseg; ++s) {
            const int tr0 = (it.grp * P.nseg + s) * P.RL;
            if (tr0 >= P.T) break;
            const int ts = tr0 + 128 * st - 40;
            const int lo = max(ts, 0), hi = min(ts + kTXRows, P.T);
            if (hi > lo) total += (uint32_t)(hi - lo) * 16u * (uint32_t)it.cps_t;
          }
          mbar_expect_tx(&x_full[xs], total);
          for (int s = 0; s < P.nseg; ++s) {
            const int tr0 = (it.grp * P.nseg + s) * P.RL;
            if (tr0 >= P.T) break;
            const int ts = tr0 + 128 * st - 40;
            const int lo = max(ts, 0), hi = min(ts + kTXRows, P.T);
            if (hi <= lo) continue;
            for (int cc = 0; cc < it.cps_t; ++cc)
              bulk_g2s(smem_u32(xsm + xs * kXStageBytes) + (uint32_t)(((s * P.sps + cc) * kTXRows + (lo - ts)) * 16),
                       xb + ((int64_t)(it.tile * P.cps + cc) * P.Tp + P.pad + lo) * 8, (uint32_t)(hi - lo) * 16u, &x_full[xs]);
          }
          if (++xs == kTXStages) { xs = 0; xph ^= 1; }
        }
      }
    }
  } else if (warp == kTWPatch) {
    // ===================== replicate padding of the input in the staged tile (rows t < 0 <- x[0], t >= T <- x[T-1]) =====
    int xs = 0; uint32_t xph = 0;
    for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
      const TcItem it = tc_item(P, item);
      const int nstages = (it.nblk + 2 + 3) >> 2;
      for (int st = 0; st < nstages; ++st) {
        mbar_wait_relaxed(&x_full[xs], xph);
        uint4* stage = reinterpret_cast<uint4*>(xsm + xs * kXStageBytes);
        for (int s = 0; s < P.nseg; ++s) {
          const int tr0 = (it.grp * P.nseg + s) * P.RL;
          if (tr0 >= P.T) break;
          const int ts = tr0 + 128 * st - 40;
          const int nlo = min(max(-ts, 0), kTXRows);                 // rows [0, nlo): t < 0
          const int rhi = min(max(P.T - ts, 0), kTXRows);            // rows [rhi, 144): t >= T
          if (nlo == 0 && rhi == kTXRows) continue;
          for (int cc = 0; cc < it.cps_t; ++cc) {
            uint4* base = stage + (s * P.sps + cc) * kTXRows;
            if (nlo > 0) {
              const uint4 v = nlo < kTXRows ? base[nlo] : make_uint4(0, 0, 0, 0);       // row of t = 0
              for (int r = lane; r < nlo; r += 32) base[r] = v;
            }
            if (rhi < kTXRows) {
              const uint4 v = rhi > 0 ? base[rhi - 1] : make_uint4(0, 0, 0, 0);         // row of t = T - 1
              __syncwarp();
              for (int r = rhi + lane; r < kTXRows; r += 32) base[r] = v;
            }
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&x_ready[xs]);
        if (++xs == kTXStages) { xs = 0; xph ^= 1; }
      }
    }
  } else if (warp == kTWIssue) {
    // ===================== MMA issuer (whole warp, warp-uniform operands, election inside the asm) =====================
    // D fp32 | A bf16 | B bf16 | A MN-major | N = 64 | M = 128          and   D fp32 | A fp16 | B fp16 | N = 32 | M = 128
    const uint32_t idesc_up = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc_dn = (1u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t uph_lo = (smem_u32(up_hi) >> 4) | (64u << 16);      // K-major B: LBO = 64 rows * 16 B, SBO = 128 B
    const uint32_t upl_lo = (smem_u32(up_lo) >> 4) | (64u << 16);
    const uint32_t dn_lo = (smem_u32(dnm) >> 4) | (32u << 16);
    const uint32_t xs_lo = (smem_u32(xsm) >> 4) | (8u << 16);          // MN-major A: LBO = 128 B, SBO = 144 rows (hi 0x4090)
    int xs = 0; uint32_t xph = 0;
    int ub = 0; uint32_t uph = 0;
    int yb = 0; uint32_t yph = 0;
    uint32_t afull_ph = 0;                                             // one parity bit per ring slot
    for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
      const TcItem it = tc_item(P, item);
      const int nup = it.nblk + 2;
      int next_a = -1;                                                 // first block whose activated samples were not awaited yet
      for (int tick = 0; tick <= it.nblk + 2; ++tick) {
        if (tick < nup) {
          // ---- up-FIR of block i = tick - 1: U[ub] = X rows [t0 - 8, t0 + 40) x taps (hi + lo)
          const int pos = tick & 3;
          if (pos == 0) { mbar_wait_backoff(&x_ready[xs], xph); }
          mbar_wait_backoff(&u_free[ub], uph ^ 1);
          tc_fence_after();
          const uint32_t d = tmem_base + kTColU + (uint32_t)ub * 64u;
          const uint32_t a0 = xs_lo + (uint32_t)xs * (kXStageBytes >> 4) + (uint32_t)pos * 32u;
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            umma_ss_elect<0x4090u, 0x4008u>(d, a0 + 16u * s, uph_lo + 128u * s, idesc_up, s > 0 ? 1u : 0u);
            umma_ss_elect<0x4090u, 0x4008u>(d, a0 + 16u * s, upl_lo + 128u * s, idesc_up, 1u);
          }
          umma_commit_elect(&u_full[ub]);
          if (pos == 3 || tick == nup - 1) {
            umma_commit_elect(&x_free[xs]);
            if (++xs == kTXStages) { xs = 0; xph ^= 1; }
          }
          if (++ub == 2) { ub = 0; uph ^= 1; }
        }
        const int j = tick - 3;
        if (j >= 0 && j < it.nblk) {
          // ---- down-FIR of block j: Y[yb] = A ring samples [64 j - 16, 64 j + 80) x taps
          while (next_a <= j + 1) {
            const int sl = next_a & 3;
            mbar_wait_backoff(&a_full[sl], (afull_ph >> sl) & 1u);
            afull_ph ^= 1u << sl;
            ++next_a;
          }
          mbar_wait_backoff(&y_free[yb], yph ^ 1);
          tc_fence_after();
          const uint32_t d = tmem_base + kTColY + (uint32_t)yb * 32u;
          const uint32_t ap = tmem_base + kTColA + (uint32_t)((j - 1) & 3) * 32u + 24u;
          const uint32_t ac = tmem_base + kTColA + (uint32_t)(j & 3) * 32u;
          const uint32_t an = tmem_base + kTColA + (uint32_t)((j + 1) & 3) * 32u;
          umma_ts_elect<0x4008u>(d, ap, dn_lo, idesc_dn, 0u);
#pragma unroll
          for (int s = 1; s < 5; ++s) umma_ts_elect<0x4008u>(d, ac + 8u * (s - 1), dn_lo + 64u * s, idesc_dn, 1u);
          umma_ts_elect<0x4008u>(d, an, dn_lo + 64u * 5, idesc_dn, 1u);
          umma_commit_elect(&y_full[yb]);
          umma_commit_elect(&a_free[(j - 1) & 3]);
          if (j == it.nblk - 1) {                                      // the blocks no later down-FIR of this item reads
            umma_commit_elect(&a_free[j & 3]);
            umma_commit_elect(&a_free[(j + 1) & 3]);
          }
          if (++yb == 2) { yb = 0; yph ^= 1; }
        }
      }
    }
  } else if (warp >= kTWSnake0) {
    // ===================== snake: U (fp32, TMEM) -> a = u + hb - hb cos(2 e^alpha u) -> fp16 pairs (TMEM ring) ==========
    const int q = warp & 3, h = (warp - kTWSnake0) >> 2;
    const int ln = q * 32 + lane;                                      // TMEM lane = channel slot
    const int L = ln >> 3;
    const int cc = L % P.sps;
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    int ub = 0; uint32_t uph = 0;
    uint32_t afree_ph = 0;
    // Cody-Waite constants: 2 pi = 6.28125 + 1.9353071795864769e-3 (the high part has 9 significant bits)
    const f32x2 kInv2Pi = pk2(0.15915494309189535f, 0.15915494309189535f);
    const f32x2 kMagic = pk2(12582912.f, 12582912.f), kNegMagic = pk2(-12582912.f, -12582912.f);
    const f32x2 kN2PiHi = pk2(-6.28125f, -6.28125f), kN2PiLo = pk2(-1.9353071795864769e-3f, -1.9353071795864769e-3f);
    for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
      const TcItem it = tc_item(P, item);
      float sc0 = 0.f, sc1 = 0.f;
      {
        const int ch = (it.tile * P.cps + cc) * 8 + (ln & 7);
        if (cc < it.cps_t && ch < P.C) snake_params<false>(P.alpha[ch], P.beta[ch], sc0, sc1);
      }
      const f32x2 SC0 = pk2(sc0, sc0), SC1 = pk2(sc1, sc1), NSC1 = pk2(-sc1, -sc1);
      for (int i = -1; i <= it.nblk; ++i) {
        const int sl = i & 3;
        mbar_wait(&u_full[ub], uph);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32_nowait(tq + kTColU + (uint32_t)ub * 64u + (uint32_t)h * 32u, v);
        mbar_wait(&a_free[sl], ((afree_ph >> sl) & 1u) ^ 1u);
        afree_ph ^= 1u << sl;
        tc_fence_after();
        tmem_ld_wait();
        uint32_t w[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const f32x2 u = pk2(__uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1]));
          const f32x2 z = mul2(u, SC0);
          const f32x2 kk = add2(fma2(z, kInv2Pi, kMagic), kNegMagic);             // rint(z / 2 pi)
          f32x2 r = fma2(kk, kN2PiHi, z);
          r = fma2(kk, kN2PiLo, r);
          float rx, ry;
          unpk2(r, rx, ry);
          const f32x2 a = fma2(NSC1, pk2(__cosf(rx), __cosf(ry)), add2(u, SC1));
          float ax, ay;
          unpk2(a, ax, ay);
          w[k] = cvt_f16x2_sat(ax, ay);
        }
        tmem_st16(tq + kTColA + (uint32_t)sl * 32u + (uint32_t)h * 16u, w);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&u_free[ub]); mbar_arrive(&a_full[sl]); }
        if (++ub == 2) { ub = 0; uph ^= 1; }
      }
    }
  } else {
    // ===================== store: Y (fp32, TMEM; lane = channel, 32 time steps) -> bf16 c8t rows -> TMA bulk store =======
    const int q = warp & 3;
    const int ln = q * 32 + lane;
    const int L = ln >> 3;
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool odd = lane & 1;
    const uint32_t sel = odd ? 0x7632u : 0x5410u;
    int yb = 0; uint32_t yph = 0;
    int ob = 0;
    for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
      const TcItem it = tc_item(P, item);
      __nv_bfloat16* ybase = P.y + (int64_t)it.b * P.bstride;
      const int ngroups = (it.nblk + 3) >> 2;
      for (int g = 0; g < ngroups; ++g) {
        if (warp == kTWStore0 && lane < 16) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        named_bar_sync(1, 128);
        uint8_t* obuf = osm + ob * kOutBytes;
        for (int jj = 0; jj < 4 && 4 * g + jj < it.nblk; ++jj) {
          mbar_wait(&y_full[yb], yph);
          tc_fence_after();
          uint32_t v[32];
          tmem_ld32_nowait(tq + kTColY + (uint32_t)yb * 32u, v);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&y_free[yb]);
          if (++yb == 2) { yb = 0; yph ^= 1; }
          // even lane: rows 2k get (own ch, partner ch); odd lane: rows 2k+1
          uint32_t* orow = reinterpret_cast<uint32_t*>(obuf + ((size_t)L * kTOutPitch + jj * 32 + (odd ? 1 : 0)) * 16) + ((ln & 7) >> 1);
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const uint32_t own = cvt_bf16x2(__uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1]));
            const uint32_t oth = __shfl_xor_sync(0xffffffffu, own, 1);
            uint32_t wv;
            asm("prmt.b32 %0, %1, %2, %3;" : "=r"(wv) : "r"(odd ? oth : own), "r"(odd ? own : oth), "r"(sel));
            orow[k * 8] = wv;                                          // 2 rows = 32 bytes = 8 words
          }
        }
        fence_async_smem();
        named_bar_sync(1, 128);
        if (warp == kTWStore0 && lane < 16) {
          const int s = lane / P.sps, cc = lane - s * P.sps;
          const int tr0 = (it.grp * P.nseg + s) * P.RL;
          if (s < P.nseg && cc < it.cps_t && tr0 < P.T) {
            const int row0 = tr0 + 128 * g;
            const int nrows = min(128, min(P.T, tr0 + P.RL) - row0);
            if (nrows > 0)
              bulk_s2g(ybase + ((int64_t)(it.tile * P.cps + cc) * P.Tp + P.pad + row0) * 8, obuf + (size_t)lane * kTOutPitch * 16,
                       (uint32_t)nrows * 16u);
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        ob ^= 1;
      }
    }
    if (warp == kTWStore0 && lane < 16) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kTWPatch) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// The rows the tensor-core pass cannot get right plus the tensor's zero frame (see the header): y[0..2], y[T-3..T-1] with
// the exact replicate-pad semantics, zero halo rows, zero padding channels.  One warp per (chunk, utterance).
__global__ void __launch_bounds__(32) act1d_c8t_edge_kernel(__nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ x,
                                                            const float* __restrict__ alpha_log, const float* __restrict__ beta_log,
                                                            int C, int chunks, int T, int Tp, int pad) {
  const int chunk = blockIdx.x, b = blockIdx.y, lane = threadIdx.x;
  const int64_t base = ((int64_t)b * chunks + chunk) * Tp;
  if (lane < 16) {
    constexpr int V = 8;
    const int side = lane >> 3, c8 = lane & 7;
    const int ch = chunk * 8 + c8;
    const int64_t tg = side ? T - V : 0;
    float xw[V + 16], yv[V];
#pragma unroll
    for (int i = 0; i < V + 16; ++i) {
      const int64_t t = tg - 8 + i;
      xw[i] = (t >= 0 && t < T) ? __bfloat162float(x[(base + pad + t) * 8 + c8]) : 0.f;
    }
    float sc0 = 0.f, sc1 = 0.f;
    if (ch < C) snake_params<false>(alpha_log[ch], beta_log[ch], sc0, sc1);
    act1d_window<V, false>(xw, yv, sc0, sc1, tg, (int64_t)T);
#pragma unroll
    for (int qq = 0; qq < 3; ++qq) {
      const int qi = side ? V - 3 + qq : qq;
      y[(base + pad + tg + qi) * 8 + c8] = __float2bfloat16_rn(ch < C ? yv[qi] : 0.f);
    }
  }
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (int r = lane; r < 2 * pad; r += 32) {
    const int row = r < pad ? r : T + r;                               // [0, pad) and [pad + T, Tp)
    *reinterpret_cast<uint4*>(y + (base + row) * 8) = z;
  }
}

}  // namespace

// Tensor-core Activation1d (see the header).  BVG_ERR_STATE (nothing launched) when the tensor does not qualify; the caller
// then takes the CUDA-core stencil kernel (act1d_c8t.cu).
int act1d_tc_launch(const C8T& y, const C8T& x, const float* alpha_log, const float* beta_log, int64_t B, cudaStream_t st) {
  if (x.T < 256 || x.pad != kC8tPad || x.chunks < 1) return BVG_ERR_STATE;
  ActTcParams P;
  P.x = x.p; P.y = y.p; P.alpha = alpha_log; P.beta = beta_log;
  P.C = x.C; P.chunks = x.chunks; P.T = x.T; P.Tp = x.Tp; P.pad = x.pad;
  P.bstride = x.batch_stride();
  const int ch = x.chunks;
  if (ch % 16 == 0 || ch > 16) { P.cps = 16; P.sps = 16; }
  else if (ch % 8 == 0) { P.cps = 8; P.sps = 8; }
  else if (ch % 4 == 0) { P.cps = 4; P.sps = 4; }
  else { P.cps = ch; P.sps = ch <= 4 ? 4 : ch <= 8 ? 8 : 16; }
  P.nseg = 16 / P.sps;
  P.ntile = (ch + P.cps - 1) / P.cps;
  int num_sms = 0;
  BVG_TRY(current_device_sms(&num_sms));
  // range length: multiples of 128 rows; minimise (rounds of items over the SMs) x (blocks per item incl. the 2 halo blocks)
  int64_t best_cost = -1;
  for (int rl = 128; rl <= 8192; rl += 128) {
    const int64_t nr = (x.T + rl - 1) / rl, ng = (nr + P.nseg - 1) / P.nseg;
    const int64_t items = B * P.ntile * ng;
    const int64_t cost = ((items + num_sms - 1) / num_sms) * (rl / kTBlk + 3);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; P.RL = rl; P.NG = (int)ng; }
    if (rl >= x.T) break;
  }
  const int64_t items = B * P.ntile * P.NG;
  BVG_CHECK_ARG(items < (1ll << 31), "act1d_tc: too many work items");
  P.nitems = (int)items;
  const size_t smem = (size_t)kTXStages * 16 * kTXRows * 16 + 2 * 6 * 64 * 16 + 12 * 32 * 16 + 2 * 16 * kTOutPitch * 16 + 24 * 8 + 16;
  static std::atomic<uint64_t> opted{0};
  BVG_TRY(smem_opt_in(act1d_tc_kernel, opted, (int)smem));
  {
    ProfScope prof(st, KC_ACT1D);
    act1d_tc_kernel<<<(unsigned)std::min<int64_t>(items, num_sms), kTThreads, smem, st>>>(P);
    BVG_LAUNCHED();
  }
  BVG_CHECK_ARG(B <= 65535, "act1d_tc: batch too large for the edge pass grid");
  {
    ProfScope prof(st, KC_ACT1D);
    act1d_c8t_edge_kernel<<<dim3((unsigned)x.chunks, (unsigned)B), 32, 0, st>>>(y.p, x.p, alpha_log, beta_log, x.C, x.chunks, x.T, x.Tp, x.pad);
    BVG_LAUNCHED();
  }
  return BVG_OK;
}

}  // namespace bvg
