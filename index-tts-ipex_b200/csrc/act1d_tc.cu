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
//   snake a = u + 1/(e^b + 1e-9) sin^2(e^a u), a thread owns one channel (lane) x 16 consecutive upsampled samples
//         (tcgen05.ld), per-channel constants in registers, result written back to TMEM as fp16 pairs (tcgen05.st)
//   down  Y^T[c, t]  = sum_m A^T[c, m] * Gdn[m, t]    TS MMA: A = the fp16 pairs in TMEM (never touches shared memory),
//                      B = Toeplitz taps fp16 (2^-12 relative: -70 dB), N = 32 output time steps
//   store Y^T (lane = channel, 32 time steps in registers) -> bf16 -> lane-pair shuffle -> [chunk][row][8 ch] tile in shared
//         memory (conflict-free 32-bit stores) -> TMA bulk store.
// Narrow tensors fill the 128 lanes with several time SEGMENTS of the same utterance (C = 24: 4 segments x 4 chunks).
// A CTA streams through a time range in 32-step blocks: up(i) | snake(i-1) | down(i-2) | store(i-3) run concurrently
// on the tensor pipe / 16 snake warps / 8 store warps (4 U and 4 Y accumulator buffers in TMEM); the activated signal of 4 consecutive blocks lives in a TMEM ring
// so that the down-FIR of block j reads its 5/6-sample halos from the neighbours (one extra block of up + snake at each
// end of a range is the whole halo cost).
//
// Edges.  Rows t < 0 / t >= T of a staged tile are overwritten with x[0] / x[T-1] in shared memory (replicate padding of
// the input, resample.py:28).  The replicate padding of the ACTIVATED signal (filter.py:90-92) only changes y[0..2] and
// y[T-3..T-1]; those six rows per utterance, the zero halo rows and the padding channels are written in the kernel's prologue
// (act1d_edge_unit, CUDA cores, the exact stencil of act1d_core.cuh); the streaming part never stores them.
#include <cuda_fp16.h>
#include <stdlib.h>

#include <algorithm>

#include "act1d_core.cuh"
#include "bvg_common.cuh"
#include "umma.cuh"
#include "umma_ptx.cuh"

namespace bvg {
namespace {

constexpr int kTBlk = 32;                       // output time steps per block
#ifndef BVG_TXBLK
#define BVG_TXBLK 4
#endif
#ifndef BVG_TXSTAGES
#define BVG_TXSTAGES 3
#endif
constexpr int kTXBlk = BVG_TXBLK;               // blocks per input stage
constexpr int kTXRows = kTXBlk * kTBlk + 16;    // staged rows per lane-chunk and stage: the blocks + 8-row halo per side
constexpr int kTXStages = BVG_TXSTAGES;         // 3 x 36 KB in flight per SM.  Measured (C = 96, B = 32, data movement only):
                                                // 16 copies of 2304 B per stage 142 us, of 1280 B 185 us, of 768 B 302 us --
                                                // the TMA unit pays ~100 cycles per bulk copy, so few large copies win
constexpr int kTPrefetch = 0;                   // stages between an L2 bulk prefetch of a row range and its TMA load; 0 = off:
                                                // measured 295 us against 228 us (C = 96, B = 32) -- the prefetches queue in front of the loads
constexpr int kTOutPitch = 128;                 // out-stage rows per lane-chunk (stmatrix writes whole 128-byte row groups)
// The warp scheduler favours high warp ids: control roles on top, then the (light) store warps, the throughput-bound
// snake warps take what is left.
// The snake is latency-bound per warp (MUFU and dependent packed FMAs): 4 snake warps per scheduler keep the MUFU pipe fed.
constexpr int kTWSnake0 = 0;                    // warps 0-15: snake, two groups of 8 taking alternate blocks
constexpr int kTNSnake = 16;
constexpr int kTWStore0 = 16;                   // warps 16-23: store (quarter = warp % 4, 16-column half = (warp / 4) % 2)
constexpr int kTWProd = 24, kTWIssueUp = 25, kTWIssueDn = 26, kTWPatch = 27;
constexpr int kTThreads = 28 * 32;              // launched with 72 registers per thread, re-balanced with setmaxnreg:
                                                // 16 snake warps x 88 + 8 store warps x 56 + 4 control warps x 40 = 28 x 72
constexpr int kTNU = 4, kTNY = 4;               // U / Y accumulator buffers (the issuer runs this far ahead of snake / store)
constexpr uint32_t kTAHi = 0x4000u | (uint32_t)kTXRows;       // A descriptor high word: SBO = chunk pitch (rows), version bit
constexpr uint32_t kTColU = 0, kTColA = 256, kTColY = 384;     // TMEM columns: U 4 x 64 | A ring 4 x 32 | Y 4 x 32

struct ActTcParams {
  const __nv_bfloat16* x; __nv_bfloat16* y;
  const float* alpha; const float* beta;
  int C, chunks, T, Tp, pad;
  int64_t bstride;
  int cps, sps, nseg, ntile;      // chunks per tile, lane-chunk slots per segment, segments per item, channel tiles
  int RL, NG;                     // rows per range, range groups per utterance
  int nitems, nb;
  const int* lens; int len_mul;   // ragged batch: utterance b has lens[b] * len_mul rows (else T)
  long long* dbg;                 // optional [grid][16] per-role cycle counters (bvg_debug_set_umma_counters)
  int dry;                        // BVG_DEBUG builds only: skip parts of the pipeline (bottleneck experiments, garbage results)
};

struct TcItem { int b, tile, grp, nblk, cps_t, T; };
__device__ __forceinline__ TcItem tc_item(const ActTcParams& P, int item) {
  TcItem it;
  const int per_b = P.ntile * P.NG;
  it.b = item / per_b;
  const int r = item - it.b * per_b;
  it.tile = r / P.NG;
  it.grp = r - it.tile * P.NG;
  it.cps_t = min(P.cps, P.chunks - it.tile * P.cps);
  const int tr0 = it.grp * P.nseg * P.RL;                      // segment 0 (the longest of the item)
  it.T = rows_of(P.lens, P.len_mul, it.b, P.T);
  it.nblk = max(0, (min(it.T, tr0 + P.RL) - tr0 + kTBlk - 1) / kTBlk);    // 0: the whole item lies past the utterance's end
  return it;
}

__device__ __forceinline__ uint32_t cvt_f16x2_sat(float lo, float hi) {
  // fp16 pair, finite-saturating (the reference itself runs this path under fp16 autocast, infer.py:194)
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// The rows the tensor-core pass cannot get right plus the tensor's zero frame (see the header): y[0..2], y[T-3..T-1] with
// the exact replicate-pad semantics, zero halo rows, zero padding channels, for one (chunk, utterance) by one warp.  The
// streaming part of the kernel never stores those six rows, so this runs in its prologue (no ordering needed, no second
// launch: 55 tiny launches per decode step were 12 % of the Activation1d time).
__device__ __forceinline__ void act1d_edge_unit(__nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ x,
                                                const float* __restrict__ alpha_log, const float* __restrict__ beta_log, int C,
                                                int chunks, int T, int Tp, int pad, int chunk, int b, int lane) {
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
      if (tg + qi >= 0) y[(base + pad + tg + qi) * 8 + c8] = __float2bfloat16_rn(ch < C ? yv[qi] : 0.f);
    }
  }
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (int r = lane; r < 2 * pad; r += 32) {
    const int row = r < pad ? r : T + r;                               // [0, pad) and [pad + T, Tp)
    *reinterpret_cast<uint4*>(y + (base + row) * 8) = z;
  }
}

__global__ void __launch_bounds__(kTThreads, 1) act1d_tc_kernel(const ActTcParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const long long dbg_entry = P.dbg ? clock64() : 0;

  constexpr uint32_t kXStageBytes = 16u * kTXRows * 16u;          // 20480
  constexpr uint32_t kUpBytes = 6u * 64u * 16u;                   // one of (hi, lo): [kchunk 6][n 64][8] bf16
  constexpr uint32_t kDnBytes = 12u * 32u * 16u;                  // [kchunk 12][n 32][8] fp16
  constexpr uint32_t kOutBytes = 16u * kTOutPitch * 16u;          // 33280
  uint8_t* xsm = smem;
  uint8_t* up_hi = xsm + kTXStages * kXStageBytes;
  uint8_t* up_lo = up_hi + kUpBytes;
  uint8_t* dnm = up_lo + kUpBytes;
  uint8_t* osm = dnm + kDnBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(osm + 2 * kOutBytes);
  uint64_t* x_full = bars;                       // [kTXStages]
  uint64_t* x_ready = x_full + kTXStages;        // [kTXStages]
  uint64_t* x_free = x_ready + kTXStages;        // [kTXStages]
  uint64_t* u_full = x_free + kTXStages;         // [4]
  uint64_t* u_free = u_full + 4;                 // [4]
  uint64_t* a_full = u_free + 4;                 // [4]
  uint64_t* a_free = a_full + 4;                 // [4]
  uint64_t* y_full = a_free + 4;                 // [4]
  uint64_t* y_free = y_full + 4;                 // [4]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(y_free + 4);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kTXStages; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_ready[i], 1); mbar_init(&x_free[i], 1); }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&u_full[i], 1); mbar_init(&u_free[i], kTNSnake / 2);      // one snake group per block
      mbar_init(&y_full[i], 1); mbar_init(&y_free[i], 8);
      mbar_init(&a_full[i], kTNSnake / 2); mbar_init(&a_free[i], 1);
    }
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
  // edge rows / zero frame of every (chunk, utterance): one unit per warp, grid-strided over all warps of the grid
  for (int u = blockIdx.x * (kTThreads / 32) + warp; u < P.nb * P.chunks; u += gridDim.x * (kTThreads / 32)) {
    const int b = u / P.chunks, chunk = u - b * P.chunks;
    act1d_edge_unit(P.y, P.x, P.alpha, P.beta, P.C, P.chunks, rows_of(P.lens, P.len_mul, b, P.T), P.Tp, P.pad, chunk, b, lane);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp >= kTWProd) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == kTWProd) {
    // ===================== TMA producer: raw rows of every (segment, chunk) of the item, 128 + 16 rows per stage ==========
    // lane L < 16 issues the copy of lane-chunk L (segment L / sps, chunk L % sps); lane 0 arms the barrier first
    int xs = 0; uint32_t xph = 0;
    const int ps = lane / P.sps, pcc = lane - ps * P.sps;
    // Pull the rows of stage `pst` of item `pit` towards L2 (no data returns to the SM): the TMA loads then see L2
    // latency instead of loaded-HBM latency, which the staging ring alone (~120 KB per SM) does not cover.
    auto prefetch_stage = [&](const TcItem& pit, int pst) {
      if (!(lane < 16 && ps < P.nseg && pcc < pit.cps_t)) return;
      const int tr0 = (pit.grp * P.nseg + ps) * P.RL;
      if (tr0 >= pit.T) return;
      const int lo = max(tr0 + kTXBlk * kTBlk * pst - 40 + (pst ? 16 : 0), 0);
      const int hi = min(tr0 + kTXBlk * kTBlk * pst - 40 + kTXRows, pit.T);
      if (hi > lo)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
                     ::"l"(P.x + (int64_t)pit.b * P.bstride + ((int64_t)(pit.tile * P.cps + pcc) * P.Tp + P.pad + lo) * 8),
                       "r"((uint32_t)(hi - lo) * 16u) : "memory");
    };
    bool first = true;
    for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
      const TcItem it = tc_item(P, item);
      if (it.nblk <= 0) continue;
      const int nstages = (it.nblk + 2 + kTXBlk - 1) / kTXBlk;
      const bool has_next = item + (int)gridDim.x < P.nitems;
      const TcItem nit = has_next ? tc_item(P, item + gridDim.x) : it;
      const int nnstages = (nit.nblk + 2 + kTXBlk - 1) / kTXBlk;
      const __nv_bfloat16* xb = P.x + (int64_t)it.b * P.bstride;
      if (first && kTPrefetch > 0) {
        for (int pst = 0; pst < kTPrefetch && pst < nstages; ++pst) prefetch_stage(it, pst);
        first = false;
      }
      for (int st = 0; st < nstages; ++st) {
        if (kTPrefetch > 0) {
          const int pst = st + kTPrefetch;
          if (pst < nstages) prefetch_stage(it, pst);
          else if (has_next && pst - nstages < nnstages) prefetch_stage(nit, pst - nstages);
        }
        mbar_wait_relaxed(&x_free[xs], xph ^ 1);
        if (lane == 0) {
          uint32_t total = 0;
          for (int s = 0; s < P.nseg; ++s) {
            const int tr0 = (it.grp * P.nseg + s) * P.RL;
            if (tr0 >= it.T) break;
            const int ts = tr0 + kTXBlk * kTBlk * st - 40;
            const int lo = max(ts, 0), hi = min(ts + kTXRows, it.T);
            if (hi > lo) total += (uint32_t)(hi - lo) * 16u * (uint32_t)it.cps_t;
          }
          mbar_expect_tx(&x_full[xs], total);
        }
        __syncwarp();
        if (lane < 16 && ps < P.nseg && pcc < it.cps_t) {
          const int tr0 = (it.grp * P.nseg + ps) * P.RL;
          const int ts = tr0 + kTXBlk * kTBlk * st - 40;
          const int lo = max(ts, 0), hi = min(ts + kTXRows, it.T);
          if (tr0 < it.T && hi > lo)
            bulk_g2s(smem_u32(xsm + xs * kXStageBytes) + (uint32_t)((lane * kTXRows + (lo - ts)) * 16),
                     xb + ((int64_t)(it.tile * P.cps + pcc) * P.Tp + P.pad + lo) * 8, (uint32_t)(hi - lo) * 16u, &x_full[xs]);
        }
        if (++xs == kTXStages) { xs = 0; xph ^= 1; }
      }
    }
  } else if (warp == kTWPatch) {
    // ===================== replicate padding of the input in the staged tile (rows t < 0 <- x[0], t >= T <- x[T-1]) =====
    int xs = 0; uint32_t xph = 0;
    for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
      const TcItem it = tc_item(P, item);
      if (it.nblk <= 0) continue;
      const int nstages = (it.nblk + 2 + kTXBlk - 1) / kTXBlk;
      for (int st = 0; st < nstages; ++st) {
        mbar_wait_relaxed(&x_full[xs], xph);
        uint4* stage = reinterpret_cast<uint4*>(xsm + xs * kXStageBytes);
        for (int s = 0; s < P.nseg; ++s) {
          const int tr0 = (it.grp * P.nseg + s) * P.RL;
          if (tr0 >= it.T) break;
          const int ts = tr0 + kTXBlk * kTBlk * st - 40;
          const int nlo = min(max(-ts, 0), kTXRows);                 // rows [0, nlo): t < 0
          const int rhi = min(max(it.T - ts, 0), kTXRows);            // rows [rhi, kTXRows): t >= T
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
  } else if (warp == kTWIssueUp) {
    // ===================== MMA issuer 1: up-FIRs (whole warp, warp-uniform operands, election inside the asm) ============
    // Two issuer warps, one per FIR: each blocks only on what its own MMAs need (a single issuer that waits for the snake
    // warps before a down-FIR would also hold back the up-FIRs those warps need next; polling instead costs ~150 cycles
    // per mbarrier test, more than the MMAs of a block take).
    // D fp32 | A bf16 | B bf16 | A MN-major | N = 64 | M = 128
    const uint32_t idesc_up = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t uph_lo = (smem_u32(up_hi) >> 4) | (64u << 16);      // K-major B: LBO = 64 rows * 16 B, SBO = 128 B
    const uint32_t upl_lo = (smem_u32(up_lo) >> 4) | (64u << 16);
    const uint32_t xs_lo = (smem_u32(xsm) >> 4) | (8u << 16);          // MN-major A: LBO = 128 B, SBO = kTXRows rows (descriptor high word)
    int xs = 0; uint32_t xph = 0;
    int ub = 0; uint32_t uph = 0;
    long long dbg_wx = 0, dbg_wu = 0;
    const long long dbg_start = P.dbg ? clock64() : 0;
    unsigned long long dbg_ns0 = 0;
    if (P.dbg) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_ns0));
    for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
      const TcItem it = tc_item(P, item);
      if (it.nblk <= 0) continue;
      const int nup = it.nblk + 2;
      for (int nu = 0; nu < nup; ++nu) {
        // ---- up-FIR of block i = nu - 1: U[ub] = X rows [t0 - 8, t0 + 40) x taps (hi + lo)
        const int pos = nu % kTXBlk;
        if (pos == 0) { DBG_T0(); mbar_wait(&x_ready[xs], xph); DBG_ADD(dbg_wx); }
        { DBG_T0(); mbar_wait(&u_free[ub], uph ^ 1); DBG_ADD(dbg_wu); }
        tc_fence_after();
        const uint32_t d = tmem_base + kTColU + (uint32_t)ub * 64u;
        const uint32_t a0 = xs_lo + (uint32_t)xs * (kXStageBytes >> 4) + (uint32_t)pos * 32u;
#ifdef BVG_DEBUG
        if (!(P.dry & 4))
#endif
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          umma_ss_elect<kTAHi, 0x4008u>(d, a0 + 16u * s, uph_lo + 128u * s, idesc_up, s > 0 ? 1u : 0u);
          umma_ss_elect<kTAHi, 0x4008u>(d, a0 + 16u * s, upl_lo + 128u * s, idesc_up, 1u);
        }
        umma_commit_elect(&u_full[ub]);
        if (pos == kTXBlk - 1 || nu == nup - 1) {
          umma_commit_elect(&x_free[xs]);
          if (++xs == kTXStages) { xs = 0; xph ^= 1; }
        }
        if (++ub == kTNU) { ub = 0; uph ^= 1; }
      }
    }
    if (P.dbg && lane == 0) {
      long long* d = P.dbg + blockIdx.x * 16;
      d[0] = dbg_wx; d[1] = dbg_wu; d[4] = clock64() - dbg_start; d[14] = dbg_start - dbg_entry;
      unsigned long long ns1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1)); d[13] = (long long)(ns1 - dbg_ns0);
    }
  } else if (warp == kTWIssueDn) {
    // ===================== MMA issuer 2: down-FIRs.  D fp32 | A fp16 (TMEM) | B fp16 | N = 32 | M = 128 ====================
    const uint32_t idesc_dn = (1u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t dn_lo = (smem_u32(dnm) >> 4) | (32u << 16);
    int yb = 0; uint32_t yph = 0;
    uint32_t afull_ph = 0;                                             // one parity bit per ring slot
    long long dbg_wa = 0, dbg_wy = 0;
    for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
      const TcItem it = tc_item(P, item);
      if (it.nblk <= 0) continue;
      int next_a = -1;                                                 // first block whose activated samples were not awaited yet
      for (int j = 0; j < it.nblk; ++j) {
        // ---- down-FIR of block j: Y[yb] = A ring samples [64 j - 16, 64 j + 80) x taps
        while (next_a <= j + 1) {
          const int sl = next_a & 3;
          { DBG_T0(); mbar_wait(&a_full[sl], (afull_ph >> sl) & 1u); DBG_ADD(dbg_wa); }
          afull_ph ^= 1u << sl;
          ++next_a;
        }
        { DBG_T0(); mbar_wait(&y_free[yb], yph ^ 1); DBG_ADD(dbg_wy); }
        tc_fence_after();
        const uint32_t d = tmem_base + kTColY + (uint32_t)yb * 32u;
        const uint32_t ap = tmem_base + kTColA + (uint32_t)((j - 1) & 3) * 32u + 24u;
        const uint32_t ac = tmem_base + kTColA + (uint32_t)(j & 3) * 32u;
        const uint32_t an = tmem_base + kTColA + (uint32_t)((j + 1) & 3) * 32u;
#ifdef BVG_DEBUG
        if (!(P.dry & 8))
#endif
        {
          umma_ts_elect<0x4008u>(d, ap, dn_lo, idesc_dn, 0u);
#pragma unroll
          for (int s = 1; s < 5; ++s) umma_ts_elect<0x4008u>(d, ac + 8u * (s - 1), dn_lo + 64u * s, idesc_dn, 1u);
          umma_ts_elect<0x4008u>(d, an, dn_lo + 64u * 5, idesc_dn, 1u);
        }
        umma_commit_elect(&y_full[yb]);
        umma_commit_elect(&a_free[(j - 1) & 3]);
        if (j == it.nblk - 1) {                                        // the blocks no later down-FIR of this item reads
          umma_commit_elect(&a_free[j & 3]);
          umma_commit_elect(&a_free[(j + 1) & 3]);
        }
        if (++yb == kTNY) { yb = 0; yph ^= 1; }
      }
    }
    if (P.dbg && lane == 0) { long long* d = P.dbg + blockIdx.x * 16; d[2] = dbg_wa; d[3] = dbg_wy; }
  } else if (warp < kTWStore0) {
    // ===================== snake: U (fp32, TMEM) -> a = u + hb - hb cos(2 e^alpha u) -> fp16 pairs (TMEM ring) ==========
    // The cosine is the MUFU approximation of the un-reduced argument z = 2 e^alpha u: its error is 1.3e-7 |z|
    // (profiles/r02_umma_probe4.txt; 1.4e-4 at |z| = 1024), far below the bf16 half-ulp of this path's output; a
    // Cody-Waite reduction costs 4 more packed instructions per pair on warps that are already issue-limited.
    // Two groups of 8 warps take alternate blocks: while one group is in its TMEM write-back / barrier round trip the
    // other keeps the MUFU pipe busy (all warps on the same block would run their MUFU phases, and then their
    // latencies, in lockstep).  Within a group: TMEM lane quarter = warp % 4, 32-column half = (warp / 4) % 2.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
    const int grp = warp >> 3;
    const int q = warp & 3, h = (warp >> 2) & 1;
    const int ln = q * 32 + lane;                                      // TMEM lane = channel slot
    const int L = ln >> 3;
    const int cc = L % P.sps;
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t nb = 0;                                                   // blocks seen so far (all items): U buffer = nb % 4
    uint32_t afree_ph = 0;
    long long dbg_su = 0, dbg_sa = 0;
    const long long dbg_sstart = P.dbg ? clock64() : 0;
    for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
      const TcItem it = tc_item(P, item);
      if (it.nblk <= 0) continue;
      float sc0 = 0.f, sc1 = 0.f;
      {
        const int ch = (it.tile * P.cps + cc) * 8 + (ln & 7);
        if (cc < it.cps_t && ch < P.C) snake_params<false>(P.alpha[ch], P.beta[ch], sc0, sc1);
      }
      const f32x2 SC0 = pk2(sc0, sc0), SC1 = pk2(sc1, sc1), NSC1 = pk2(-sc1, -sc1);
      for (int i = -1; i <= it.nblk; ++i, ++nb) {
        const int sl = i & 3;
        const uint32_t aph = (afree_ph >> sl) & 1u;
        afree_ph ^= 1u << sl;                                          // every block uses its slot once, whoever computes it
        if ((int)(nb & 1u) != grp) continue;
        const uint32_t ub = nb & 3u, uph = (nb >> 2) & 1u;
        { DBG_T0(); mbar_wait(&u_full[ub], uph); DBG_ADD(dbg_su); }
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32_nowait(tq + kTColU + ub * 64u + (uint32_t)h * 32u, v);
        { DBG_T0(); mbar_wait(&a_free[sl], aph ^ 1u); DBG_ADD(dbg_sa); }
        tc_fence_after();
        tmem_ld_wait();                                                // U is in registers: the buffer can be rewritten
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&u_free[ub]);
        uint32_t w[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const f32x2 u = pk2(__uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1]));
          float zx, zy;
          unpk2(mul2(u, SC0), zx, zy);
#ifdef BVG_DEBUG
          if (P.dry & 1) { w[k] = cvt_f16x2_sat(zx, zy); continue; }
#endif
          const f32x2 a = fma2(NSC1, pk2(__cosf(zx), __cosf(zy)), add2(u, SC1));
          float ax, ay;
          unpk2(a, ax, ay);
          w[k] = cvt_f16x2_sat(ax, ay);
        }
        tmem_st16(tq + kTColA + (uint32_t)sl * 32u + (uint32_t)h * 16u, w);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[sl]);
      }
    }
    if (P.dbg && threadIdx.x == kTWSnake0 * 32) {
      long long* d = P.dbg + blockIdx.x * 16;
      d[5] = dbg_su; d[6] = dbg_sa; d[8] = clock64() - dbg_sstart;
    }
  } else {
    // ===================== store: Y (fp32, TMEM; lane = channel, 32 time steps) -> bf16 c8t rows -> TMA bulk store =======
    // 8 warps: TMEM lane quarter q = warp % 4, 16-column (time) half h = (warp / 4) % 2 of every block.  The two warps of a
    // quarter fill the rows of its 4 lane-chunks in the out stage and synchronise on a 64-thread named barrier per quarter;
    // the next block's columns are fetched from TMEM while the current ones are packed (cur / nxt below).
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    const int q = warp & 3, h = (warp >> 2) & 1;
    // TMEM -> registers in the mma C-fragment layout (tcgen05.ld.16x256b: thread T holds 2 consecutive time steps of
    // channel-lane T/4 and T/4 + 8), which is exactly what stmatrix.trans needs to write 8 channels x 8 time steps as eight
    // 16-byte c8t rows: no shuffles, 2 loads + 8 packs + 2 stores per 32 lanes x 16 time steps.
    const uint32_t tqa = tmem_base + ((uint32_t)(q * 32) << 16) + kTColY + (uint32_t)h * 16u;      // lanes +0..15
    const uint32_t tqb = tqa + (16u << 16);                                                          // lanes +16..31
    // this thread's stmatrix row: matrix k = lane / 8 -> chunk (k & 1) [+2 for the second store], time (k >> 1) * 8 + lane % 8
    const int st_chunk = 4 * q + ((lane >> 3) & 1);
    const int st_row = 16 * h + (lane >> 4) * 8 + (lane & 7);
    const bool issuer = h == 0 && lane < 4;                            // lane l stores lane-chunk 4 q + l
    const int sL = 4 * q + lane;
    int yb = 0; uint32_t yph = 0;
    int ob = 0;
    long long dbg_ty = 0, dbg_tb = 0;
    const long long dbg_tstart = P.dbg ? clock64() : 0;
    for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
      const TcItem it = tc_item(P, item);
      if (it.nblk <= 0) continue;
      __nv_bfloat16* ybase = P.y + (int64_t)it.b * P.bstride;
      // one block: pack `ca`/`cb` (already requested from TMEM) into rows [32 (j % 4) + 16 h, +16) of the out stage
      auto step = [&](uint32_t (&ca)[8], uint32_t (&cb)[8], uint32_t (&na)[8], uint32_t (&nb)[8], int j, uint8_t* obuf) {
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&y_free[yb]);
        if (++yb == kTNY) { yb = 0; yph ^= 1; }
        if (j + 1 < it.nblk && ((j + 1) & 3) != 0) {                   // (a new store group first waits for its out buffer)
          { DBG_T0(); mbar_wait(&y_full[yb], yph); DBG_ADD(dbg_ty); }
          tc_fence_after();
          tmem_ld_16x256b_x2_nowait(tqa + (uint32_t)yb * 32u, na);
          tmem_ld_16x256b_x2_nowait(tqb + (uint32_t)yb * 32u, nb);
        }
#ifdef BVG_DEBUG
        if (P.dry & 2) return;
#endif
        uint8_t* row = obuf + ((size_t)st_chunk * kTOutPitch + (j & 3) * 32 + st_row) * 16;
        stmatrix_x4_trans(row, cvt_bf16x2(__uint_as_float(ca[0]), __uint_as_float(ca[1])),
                          cvt_bf16x2(__uint_as_float(ca[2]), __uint_as_float(ca[3])),
                          cvt_bf16x2(__uint_as_float(ca[4]), __uint_as_float(ca[5])),
                          cvt_bf16x2(__uint_as_float(ca[6]), __uint_as_float(ca[7])));
        stmatrix_x4_trans(row + (size_t)2 * kTOutPitch * 16, cvt_bf16x2(__uint_as_float(cb[0]), __uint_as_float(cb[1])),
                          cvt_bf16x2(__uint_as_float(cb[2]), __uint_as_float(cb[3])),
                          cvt_bf16x2(__uint_as_float(cb[4]), __uint_as_float(cb[5])),
                          cvt_bf16x2(__uint_as_float(cb[6]), __uint_as_float(cb[7])));
      };
      const int ngroups = (it.nblk + 3) >> 2;
      for (int g = 0; g < ngroups; ++g) {
        { DBG_T0();
          if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          named_bar_sync(1 + q, 64);
          DBG_ADD(dbg_tb); }
        uint8_t* obuf = osm + ob * kOutBytes;
        uint32_t va[8], vb[8], wa[8], wb[8];
        { DBG_T0(); mbar_wait(&y_full[yb], yph); DBG_ADD(dbg_ty); }
        tc_fence_after();
        tmem_ld_16x256b_x2_nowait(tqa + (uint32_t)yb * 32u, va);
        tmem_ld_16x256b_x2_nowait(tqb + (uint32_t)yb * 32u, vb);
        const int j0 = 4 * g;
        step(va, vb, wa, wb, j0, obuf);
        if (j0 + 1 < it.nblk) step(wa, wb, va, vb, j0 + 1, obuf);
        if (j0 + 2 < it.nblk) step(va, vb, wa, wb, j0 + 2, obuf);
        if (j0 + 3 < it.nblk) step(wa, wb, va, vb, j0 + 3, obuf);
        fence_async_smem();
        named_bar_sync(1 + q, 64);
        if (issuer) {
          const int s = sL / P.sps, cc = sL - s * P.sps;
          const int tr0 = (it.grp * P.nseg + s) * P.RL;
          if (s < P.nseg && cc < it.cps_t && tr0 < it.T) {
            const int row0 = tr0 + 128 * g;
            // (rows 0..2 and T-3..T-1 belong to the exact edge pass of the prologue)
            const int lo = max(row0, 3), hi = min(min(row0 + 128, min(it.T, tr0 + P.RL)), it.T - 3);
            if (hi > lo)
              bulk_s2g(ybase + ((int64_t)(it.tile * P.cps + cc) * P.Tp + P.pad + lo) * 8,
                       obuf + ((size_t)sL * kTOutPitch + (lo - row0)) * 16, (uint32_t)(hi - lo) * 16u);
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        ob ^= 1;
      }
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (P.dbg && threadIdx.x == kTWStore0 * 32) {
      long long* d = P.dbg + blockIdx.x * 16;
      d[9] = dbg_ty; d[10] = dbg_tb; d[11] = clock64() - dbg_tstart; d[15] = clock64() - dbg_entry;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kTWPatch) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
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
  P.lens = x.lens; P.len_mul = x.len_mul;
  const int ch = x.chunks;
  // the largest tile of 16 / 8 / 4 chunks that divides the chunk count fills all 128 TMEM lanes (C = 192: 24 chunks = 3 tiles
  // of 8 chunks x 2 time segments, not 16 + a half-empty 8); otherwise 16-chunk tiles with a ragged last one
  if (ch % 16 == 0) { P.cps = 16; P.sps = 16; }
  else if (ch % 8 == 0 && BVG_ENV_ONCE("BVG_ACT_TC_SPLIT8", 1)) { P.cps = 8; P.sps = 8; }
  else if (ch > 16) { P.cps = 16; P.sps = 16; }
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
  BVG_CHECK_ARG(B * (int64_t)x.chunks < (1ll << 31), "act1d_tc: too many (chunk, utterance) units");
  P.nb = (int)B;
  P.dbg = g_dbg_buf;
  P.dry = 0;
#ifdef BVG_DEBUG
  if (const char* e = getenv("BVG_ACT_TC_DRY")) P.dry = atoi(e);
#endif
  const size_t smem = (size_t)kTXStages * 16 * kTXRows * 16 + 2 * 6 * 64 * 16 + 12 * 32 * 16 + 2 * 16 * kTOutPitch * 16 + (3 * kTXStages + 24) * 8 + 16;
  static std::atomic<uint64_t> opted{0};
  BVG_TRY(smem_opt_in(act1d_tc_kernel, opted, (int)smem));
  {
    ProfScope prof(st, KC_ACT1D);
    act1d_tc_kernel<<<(unsigned)std::min<int64_t>(items, num_sms), kTThreads, smem, st>>>(P);
    BVG_LAUNCHED();
  }
  return BVG_OK;
}

}  // namespace bvg
