// Fused Activation1d -> Conv1d for the narrow generator stages (C <= 128), tcgen05 / TMEM, sm_100a.
//
// AMPBlock1 always feeds its convs with an Activation1d output (models.py:65-74 of the reference:
// xt = c1(a1(x)); xt = c2(a2(xt)); x = xt + x).  For C <= 96 both ops are memory/issue bound, so here the activated
// tensor never exists in HBM: the TMA producer stages the RAW input tile (conv rows + 8-row FIR halo), the worker
// warps run the packed-fp32x2 stencil of act1d_core.cuh on it and write the result straight into the K-major UMMA
// A-operand tile in shared memory (zero rows where the conv pads), the MMA issuers run the taps as row-shifted
// descriptors exactly as in conv_umma_kernel, and the same worker warps drain the accumulators one tile later
// (+bias, +residual(s), *scale -> bf16).  HBM traffic per act+conv pair drops from 5 tensor passes to 2-3.
//
// Warp roles (640 threads = 5 warpgroups):
//   warp 0      TMA producer, raw input rows           warp 3      TMA producer, weights
//   warps 1-2   MMA issuers (accumulators dealt out)   warps 4-19  16 workers: stencil of tile i, then epilogue of
//                                                                   tile i-1 (TMEM lane quarter = warp % 4)
// The FIR stencil is FMA-pipe bound and is the critical resource, so
//   * the register file is re-balanced with setmaxnreg (control warpgroup 64, workers 104; 4*64 + 16*104 = 20*96,
//     the budget the CTA is launched with -- a larger sum makes setmaxnreg.inc spin forever);
//   * stencil work is cut into lane-units (one 32-bit channel-pair word x 16 rows) that are dealt to the 512 worker
//     lanes round-robin with a running offset that carries across stages and tiles, so no lane is systematically
//     the one with an extra unit;
//   * the A tile is rounded up to whole 16-row groups (no partial groups on the slow path) and only real channel
//     words are computed (padding words of the A stages are zeroed once at start-up and never written again).
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "act1d_core.cuh"
#include "bvg_common.cuh"
#include "umma.cuh"
#include "umma_ptx.cuh"

namespace bvg {
namespace {

constexpr int kFThreads = 640;
constexpr int kFWorker0 = 4;                  // first worker warp
constexpr int kFWorkers = 16;
constexpr int kFLanes = kFWorkers * 32;       // 512 worker lanes
constexpr int kFAStages = 2;                  // A-operand stages
constexpr int kFMaxRaw = 4;                   // raw-input stages (P.x_stages of them are used)
constexpr int kFMaxW = 32;
constexpr int kFV = 16;                       // rows per lane-unit

__host__ __device__ inline size_t fused_fixed_smem(int NB, int Cin_p) {
  return (size_t)(2 * kFMaxRaw + 2 * kFAStages + 2 * kFMaxW + 4) * 8 + 16 + (size_t)NB * 4 + (size_t)Cin_p * 8;
}

__global__ void __launch_bounds__(kFThreads, 1) conv_umma_fused_kernel(const UmmaConvParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int nraw = P.x_stages, kWStages = P.w_stages;
  const int XRAW = P.XR + 16;                                      // staged raw rows per chunk (8-row FIR halo each side)
  const uint32_t raw_stage_bytes = (uint32_t)XRAW * P.kc_max * 16u;
  const uint32_t x_stage_bytes = (uint32_t)P.XR * P.kc_max * 16u;
  const uint32_t w_stage_bytes = (uint32_t)P.NB * P.kc_max * 16u;
  uint8_t* rsm = smem;
  uint8_t* xsm = rsm + nraw * raw_stage_bytes;
  uint8_t* wsm = xsm + kFAStages * x_stage_bytes;
  uint8_t* tail = wsm + kWStages * w_stage_bytes;
  uint64_t* full_raw = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_raw = full_raw + kFMaxRaw;
  uint64_t* full_x = empty_raw + kFMaxRaw;
  uint64_t* empty_x = full_x + kFAStages;
  uint64_t* full_w = empty_x + kFAStages;
  uint64_t* empty_w = full_w + kFMaxW;
  uint64_t* tmem_full = empty_w + kFMaxW;                          // [2]
  uint64_t* tmem_empty = tmem_full + 2;                            // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);   // [4] (keeps what follows 16-byte aligned)
  float* bias_s = reinterpret_cast<float*>(tmem_ptr + 4);          // [NB]
  float4* snk = reinterpret_cast<float4*>(bias_s + P.NB);          // [Cin_p/2] (sc0, sc1) of a channel pair

  const int nacc = P.MT;                                           // one accumulator per 128-row time sub-tile
  const int acc_cols = nacc * P.NB;
  const int ntiles = P.tiles_per_batch * P.B;
  const int n_iss = P.n_issuers;

  if (threadIdx.x == 0) {
    for (int i = 0; i < nraw; ++i) { mbar_init(&full_raw[i], 1); mbar_init(&empty_raw[i], kFWorkers); }
    for (int i = 0; i < kFAStages; ++i) { mbar_init(&full_x[i], kFWorkers); mbar_init(&empty_x[i], n_iss); }
    for (int i = 0; i < kWStages; ++i) { mbar_init(&full_w[i], 1); mbar_init(&empty_w[i], n_iss); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], n_iss); mbar_init(&tmem_empty[i], kFWorkers); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(P.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  {
    // the A stages start as zeros: padding channel words are never written afterwards
    uint4* xz = reinterpret_cast<uint4*>(xsm);
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (uint32_t i = threadIdx.x; i < kFAStages * (x_stage_bytes >> 4); i += kFThreads) xz[i] = z;
    for (int i = threadIdx.x; i < P.NB; i += kFThreads) bias_s[i] = (P.bias && i < P.Cout) ? P.bias[i] : 0.f;
    for (int i = threadIdx.x; i < (P.Cin_p >> 1); i += kFThreads) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      if (2 * i < P.Cin) snake_params<false>(P.act_alpha[2 * i], P.act_beta[2 * i], s.x, s.y);
      if (2 * i + 1 < P.Cin) snake_params<false>(P.act_alpha[2 * i + 1], P.act_beta[2 * i + 1], s.z, s.w);
      snk[i] = s;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < kFWorker0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 0) {
      // ===================== TMA producer 1: raw input rows (real channel chunks only) =====================
      if (lane == 0) {
        int rs = 0;
        uint32_t rph = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
          const int mt = tile % P.tiles_per_batch;
          const int b = tile / P.tiles_per_batch;
          const int q0 = mt * P.MT * 128;
          const __nv_bfloat16* xb = P.x + (int64_t)b * P.x_bstride;
          // raw tile row i <-> padded-space row r0 + i, r0 = x_row0 + q0 - lo - 8; clipped to the chunk [0, x_tp)
          // (rows outside only ever feed replicate-padded positions, which the stencil overrides)
          const int r0 = P.x_row0 + q0 - P.lo - 8;
          const int lo_r = max(r0, 0), hi_r = min(r0 + XRAW, P.x_tp);
          const uint32_t nbytes = (uint32_t)(hi_r - lo_r) * 16u;
          for (int cb = 0; cb < P.n_ci_blk; ++cb) {
            const int kcr = min(8, (P.Cin - cb * 64 + 7) >> 3);
            mbar_wait_relaxed(&empty_raw[rs], rph ^ 1);
            mbar_expect_tx(&full_raw[rs], nbytes * kcr);
            for (int kc = 0; kc < kcr; ++kc)
              bulk_g2s(smem_u32(rsm + rs * raw_stage_bytes) + (kc * XRAW + (lo_r - r0)) * 16,
                       xb + ((int64_t)(cb * 8 + kc) * P.x_tp + lo_r) * 8, nbytes, &full_raw[rs]);
            if (++rs == nraw) { rs = 0; rph ^= 1; }
          }
        }
      }
    } else if (warp == 3) {
      // ===================== TMA producer 2: weights (own warp: a full weight ring must not delay the raw rows) =====
      if (lane == 0) {
        int ws = 0;
        uint32_t wph = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
          if (P.w_resident && tile != (int)blockIdx.x) break;          // resident weights are fetched once
          for (int cb = 0; cb < P.n_ci_blk; ++cb) {
            const int kcn = min(8, (P.Cin_p - cb * 64) >> 3);
            const uint32_t wbytes = (uint32_t)P.NB * kcn * 16u;
            const __nv_bfloat16* wsrc = P.w + (int64_t)cb * 64 * P.NB * P.ntaps;
            for (int tp = 0; tp < P.ntaps; ++tp) {
              const int slot = P.w_resident ? cb * P.ntaps + tp : ws;
              if (!P.w_resident) mbar_wait_relaxed(&empty_w[ws], wph ^ 1);
              mbar_expect_tx(&full_w[slot], wbytes);
              bulk_g2s(smem_u32(wsm + slot * w_stage_bytes), wsrc + (int64_t)tp * kcn * 8 * P.NB, wbytes, &full_w[slot]);
              if (!P.w_resident && ++ws == kWStages) { ws = 0; wph ^= 1; }
            }
          }
        }
      }
    } else {
      // ===================== MMA issuers (as in conv_umma_kernel, conv taps only) =====================
      const int ii = warp - 1;
      if (ii < n_iss) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P.NB >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a_lbo = (uint32_t)P.XR << 16, b_lbo = (uint32_t)P.NB << 16;
        const uint32_t astep = 2u * P.XR, bstep = 2u * P.NB;
        const uint32_t xsb16 = x_stage_bytes >> 4, wslot16 = w_stage_bytes >> 4;
        const uint32_t x_base = (smem_u32(xsm) >> 4) | a_lbo, w_base = (smem_u32(wsm) >> 4) | b_lbo;
        const int first_tile = blockIdx.x;
        int xs = 0, ws = 0, as = 0;
        uint32_t xph = 0, wph = 0, aph = 0;
        long long dbg_wx = 0, dbg_wt = 0;
        const long long dbg_start = P.dbg ? clock64() : 0;
        for (int tile = first_tile; tile < ntiles; tile += gridDim.x) {
          { DBG_T0(); mbar_wait(&tmem_empty[as], aph ^ 1); DBG_ADD(dbg_wt); }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t dbase = tmem_base + (uint32_t)(as * acc_cols);
          for (int cb = 0; cb < P.n_ci_blk; ++cb) {
            const int nk = min(8, (P.Cin_p - cb * 64) >> 3) >> 1;
            { DBG_T0(); mbar_wait(&full_x[xs], xph); DBG_ADD(dbg_wx); }   // the workers filled this A stage
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_cb = x_base + (uint32_t)xs * xsb16;
            for (int tp = 0; tp < P.ntaps; ++tp) {
              const int slot = P.w_resident ? cb * P.ntaps + tp : ws;
              if (!P.w_resident || tile == first_tile) {
                mbar_wait(&full_w[slot], P.w_resident ? 0u : wph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              }
              const uint32_t b_lo = w_base + (uint32_t)slot * wslot16;
              const uint32_t accum0 = (cb > 0 || tp > 0) ? 1u : 0u;
              if (lane == 0) {
                const uint32_t a_tp = a_cb + (uint32_t)(tp * P.dil);
                for (int ms = ii; ms < P.MT; ms += n_iss) {
                  const uint32_t d = dbase + (uint32_t)(ms * P.NB);
                  uint32_t am = a_tp + (uint32_t)(ms * 128), bm = b_lo;
                  umma_bf16_imm(d, am, bm, idesc, accum0);
                  for (int k = 1; k < nk; ++k) {
                    am += astep; bm += bstep;
                    umma_bf16_imm(d, am, bm, idesc, 1u);
                  }
                }
              }
              if (!P.w_resident) {
                umma_commit_elect(&empty_w[ws]);
                if (++ws == kWStages) { ws = 0; wph ^= 1; }
              }
            }
            umma_commit_elect(&empty_x[xs]);
            if (++xs == kFAStages) { xs = 0; xph ^= 1; }
          }
          umma_commit_elect(&tmem_full[as]);
          if (++as == 2) { as = 0; aph ^= 1; }
        }
        if (P.dbg && lane == 0 && ii == 0) { long long* d = P.dbg + blockIdx.x * 8; d[3] = dbg_wx; d[4] = dbg_wt; d[5] = clock64() - dbg_start; }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ===================== workers: stencil of tile i, then epilogue of tile i-1 =====================
    constexpr int V = kFV;
    const int wk = warp - kFWorker0;                                // 0..15
    const uint32_t L = (uint32_t)(wk * 32 + lane);                  // worker lane 0..511
    const int wq = warp & 3;                                        // TMEM lane quarter this warp may access
    const int esub = wk >> 2;                                       // which of the quarter's four warps
    const int r = wq * 32 + lane;                                   // accumulator row
    const int ngroups = P.XR / V;                                   // 16-row groups of the A tile (XR is a multiple of 16)
    const int ngrp16 = P.NB >> 4;
    const int nit = nacc * ngrp16;                                  // epilogue items (accumulator, 16 columns) <= 16
    uint32_t off = 0;                                               // running lane-unit offset (same in every lane)
    int rs = 0, xs = 0;
    uint32_t rph = 0, xph = 0;
    long long dbg_araw = 0, dbg_ax = 0, dbg_abusy = 0, dbg_ewait = 0, dbg_ebusy = 0;
    const int my_tiles = ntiles > (int)blockIdx.x ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    for (int seq = 0; seq <= my_tiles; ++seq) {
      if (seq < my_tiles) {
        const int tile = blockIdx.x + seq * gridDim.x;
        const int mt = tile % P.tiles_per_batch;
        const int q0 = mt * P.MT * 128;
        const int64_t t_first = (int64_t)q0 - P.lo;                 // time index of A-tile row 0
        for (int cb = 0; cb < P.n_ci_blk; ++cb) {
          const int nwords = min(32, (P.Cin - cb * 64 + 1) >> 1);   // real channel-pair words per row in this block
          const int U = ngroups * nwords;
          { DBG_T0(); mbar_wait_relaxed(&full_raw[rs], rph); DBG_ADD(dbg_araw); }
          { DBG_T0(); mbar_wait_relaxed(&empty_x[xs], xph ^ 1); DBG_ADD(dbg_ax); }
          DBG_T0();
          const uint32_t* raw = reinterpret_cast<const uint32_t*>(rsm + rs * raw_stage_bytes);
          uint32_t* xo = reinterpret_cast<uint32_t*>(xsm + xs * x_stage_bytes);
          for (int e = (int)((L - off) & (kFLanes - 1)); e < U; e += kFLanes) {
            const int rg = e / nwords;
            const int wrd = e - rg * nwords;
            const int cg = wrd >> 2, pp = wrd & 3;
            const int j0 = rg * V;                                  // first A-tile row of this unit
            const int64_t t0 = t_first + j0;
            const int chA = (cb * 8 + cg) * 8 + 2 * pp;
            const uint32_t* inw = raw + (size_t)(cg * XRAW + j0) * 4 + pp;      // window row i <-> raw row j0 + i
            uint32_t* outw = xo + (size_t)(cg * P.XR + j0) * 4 + pp;
            const float4 sp = snk[chA >> 1];
            const bool interior = (t0 - 5 >= 0) && (t0 + V + 4 <= (int64_t)P.Tout - 1) && (chA + 1 < P.Cin);
            if (interior) {
              act1d_window2<V>([&](int j) { return unpack_bf16x2(inw[j * 4]); },
                               [&](int q, float ya, float yb) { outw[q * 4] = pack2(ya, yb); },
                               pk2(sp.x, sp.z), pk2(sp.y, sp.w));
            } else if ((t0 + V - 1 >= 0) && (t0 < P.Tout)) {
              uint32_t wd[V + 16];
#pragma unroll
              for (int j = 0; j < V + 16; ++j) wd[j] = inw[j * 4];
              float ylo[V];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                float yv[V];
                if (chA + h < P.Cin) {
                  float xw[V + 16];
#pragma unroll
                  for (int j = 0; j < V + 16; ++j)
                    xw[j] = h ? __uint_as_float(wd[j] & 0xffff0000u) : __uint_as_float(wd[j] << 16);
                  act1d_window<V, false>(xw, yv, h ? sp.z : sp.x, h ? sp.w : sp.y, t0, (int64_t)P.Tout);
#pragma unroll
                  for (int q = 0; q < V; ++q)
                    if (t0 + q < 0 || t0 + q >= P.Tout) yv[q] = 0.f;       // the conv's zero padding
                } else {
#pragma unroll
                  for (int q = 0; q < V; ++q) yv[q] = 0.f;
                }
                if (h == 0) {
#pragma unroll
                  for (int q = 0; q < V; ++q) ylo[q] = yv[q];
                } else {
#pragma unroll
                  for (int q = 0; q < V; ++q) outw[q * 4] = pack2(ylo[q], yv[q]);
                }
              }
            } else {
#pragma unroll
              for (int q = 0; q < V; ++q) outw[q * 4] = 0u;
            }
          }
          off = (off + (uint32_t)U) & (kFLanes - 1);
          // generic-proxy writes of the A tile -> visible to the tensor core's async-proxy reads
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) { mbar_arrive(&full_x[xs]); mbar_arrive(&empty_raw[rs]); }
          DBG_ADD(dbg_abusy);
          if (++rs == nraw) { rs = 0; rph ^= 1; }
          if (++xs == kFAStages) { xs = 0; xph ^= 1; }
        }
      }
      if (seq > 0) {
        // ---- epilogue of the previous tile (its MMAs ran while this warp did the stencil above) ----
        const int pseq = seq - 1;
        const int tile = blockIdx.x + pseq * gridDim.x;
        const int as = pseq & 1;
        const uint32_t aph = (uint32_t)(pseq >> 1) & 1u;
        const int mt = tile % P.tiles_per_batch;
        const int b = tile / P.tiles_per_batch;
        const int q0 = mt * P.MT * 128;
        __nv_bfloat16* yb = P.y + (int64_t)b * P.y_bstride;
        const __nv_bfloat16* r1 = P.res1 ? P.res1 + (int64_t)b * P.y_bstride : nullptr;
        const __nv_bfloat16* r2 = P.res2 ? P.res2 + (int64_t)b * P.y_bstride : nullptr;
        DBG_T0();
        // residual rows do not depend on the MMAs: get all of this warp's loads in flight before waiting for them
        uint4 e1[4][2], e2[4][2];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
          for (int g = 0; g < 2; ++g) { e1[k][g] = make_uint4(0, 0, 0, 0); e2[k][g] = make_uint4(0, 0, 0, 0); }
          const int it = esub + 4 * k;
          if (it < nit && (r1 || r2)) {
            const int a = it / ngrp16, c0 = (it - a * ngrp16) << 4;
            const int64_t t = (int64_t)q0 + a * 128 + r;
            if (t < P.Tout) {
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                const int co = c0 + 8 * g;
                if (co < P.Cout) {
                  const int64_t o = (int64_t)(co >> 3) * P.y_tp * 8 + ((int64_t)P.y_row0 + t) * 8;
                  if (r1) e1[k][g] = *reinterpret_cast<const uint4*>(r1 + o);
                  if (r2) e2[k][g] = *reinterpret_cast<const uint4*>(r2 + o);
                }
              }
            }
          }
        }
        mbar_wait_relaxed(&tmem_full[as], aph);
        DBG_ADD(dbg_ewait);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
          DBG_T0();
          const uint32_t tbase = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(as * acc_cols);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int it = esub + 4 * k;
            if (it < nit) {
              const int a = it / ngrp16, c0 = (it - a * ngrp16) << 4;
              const int64_t t = (int64_t)q0 + a * 128 + r;
              uint32_t v[16];
              tmem_ld16(tbase + (uint32_t)(a * P.NB + c0), v);
              if (t < P.Tout) {
                const int64_t rowoff = ((int64_t)P.y_row0 + t) * 8;
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                  const int co = c0 + 8 * g;
                  if (co >= P.y_chunks * 8) continue;                 // (padding channels inside the tensor become zeros)
                  float f[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[8 * g + j]) + bias_s[co + j];
                  if (r1) { float ee[8]; unpack8(e1[k][g], ee);
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] += ee[j]; }
                  if (r2) { float ee[8]; unpack8(e2[k][g], ee);
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] += ee[j]; }
                  uint4 o;
                  o.x = pack2(f[0] * P.scale, f[1] * P.scale); o.y = pack2(f[2] * P.scale, f[3] * P.scale);
                  o.z = pack2(f[4] * P.scale, f[5] * P.scale); o.w = pack2(f[6] * P.scale, f[7] * P.scale);
                  *reinterpret_cast<uint4*>(yb + (int64_t)(co >> 3) * P.y_tp * 8 + rowoff) = o;
                }
              }
            }
          }
          // this warp's TMEM reads of the stage are complete: hand it back to the MMA issuers
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[as]);
          if (P.zero_pads) {
            // rows [-PAD, 0) by the first tile, [Tout, Tout+PAD) by the last: keeps the c8t zero halo intact
            const int et = threadIdx.x - kFWorker0 * 32;
            const int chn = P.y_chunks;
            const uint4 z = make_uint4(0, 0, 0, 0);
            if (mt == 0)
              for (int i = et; i < chn * P.y_row0; i += kFLanes)
                *reinterpret_cast<uint4*>(yb + ((int64_t)(i / P.y_row0) * P.y_tp + (i % P.y_row0)) * 8) = z;
            if (mt == P.tiles_per_batch - 1)
              for (int i = et; i < chn * P.y_row0; i += kFLanes)
                *reinterpret_cast<uint4*>(yb + ((int64_t)(i / P.y_row0) * P.y_tp + P.y_row0 + P.Tout + (i % P.y_row0)) * 8) = z;
          }
          DBG_ADD(dbg_ebusy);
        }
      }
    }
    if (P.dbg && threadIdx.x == kFWorker0 * 32) {
      long long* d = P.dbg + blockIdx.x * 8;
      d[0] = dbg_araw; d[1] = dbg_ax; d[2] = dbg_abusy; d[6] = dbg_ewait; d[7] = dbg_ebusy;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(P.tmem_cols));
  }
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}

}  // namespace

// Activation1d(x) -> Conv1d fused (see conv_umma_fused_kernel).  Returns BVG_ERR_STATE without launching when the
// layer does not qualify (caller falls back to act1d_c8t_launch + conv_umma_launch).
int conv_umma_fused_launch(const UmmaLayer& L, const C8T& x, const float* act_alpha, const float* act_beta,
                           const C8T& y, const UmmaEpilogue& ep, int64_t B, cudaStream_t st) {
  if (L.transposed || !act_alpha || !act_beta) return BVG_ERR_STATE;
  BVG_CHECK_ARG(L.w && x.p && y.p, "conv_umma_fused: null pointer");
  BVG_CHECK_ARG(x.C == L.Cin && y.C == L.Cout && y.T == x.T, "conv_umma_fused: shape mismatch");
  UmmaConvParams P;
  memset(&P, 0, sizeof P);
  int n_nblk = 1;
  P.NPH = 1;
  umma_choose_nb(L.Cout, 1, &P.NB, &n_nblk);
  if (n_nblk != 1 || P.NB > 128 || ep.cond || ep.relu || ep.post_scale || ep.act) return BVG_ERR_STATE;
  P.ntaps = L.K;
  BVG_CHECK_ARG(L.K <= 16, "conv_umma_fused: at most 16 taps");
  const int halo = L.dil * (L.K - 1);
  P.lo = halo / 2;
  BVG_CHECK_ARG(P.lo <= x.pad, "conv_umma_fused: conv padding %d exceeds the c8t halo %d", P.lo, x.pad);
  P.dil = L.dil;
  P.u = 1; P.p = 0;
  P.Tout = y.T;
  P.Cin = L.Cin;
  P.Cin_p = (L.Cin + 15) / 16 * 16;
  BVG_CHECK_ARG(x.chunks * 8 >= P.Cin_p, "conv_umma_fused: input tensor must carry channel padding to a multiple of 16");
  P.n_ci_blk = (P.Cin_p + 63) / 64;
  P.Cout = L.Cout;
  P.x = x.p; P.x_bstride = (int64_t)x.chunks * x.Tp * 8; P.x_tp = x.Tp; P.x_row0 = x.pad;
  P.y = y.p; P.y_bstride = (int64_t)y.chunks * y.Tp * 8; P.y_tp = y.Tp; P.y_row0 = y.pad; P.y_chunks = y.chunks;
  P.w = L.w;
  P.bias = ep.bias; P.scale = ep.scale; P.res1 = ep.res1; P.res2 = ep.res2; P.zero_pads = ep.zero_pads;
  P.act_alpha = act_alpha; P.act_beta = act_beta;
  P.dbg = ep.dbg;
  P.acc_stages = 2;
  P.n_nblk = 1;
  P.B = (int)B;
  P.kc_max = std::min(8, P.Cin_p / 8);
  // Tile plan: the largest MT (128-row accumulators per tile; fewer halo rows recomputed per output) whose two raw
  // stages, two A stages and weights fit shared memory with two accumulator stages in TMEM; leftover shared
  // memory deepens the raw-input ring.  BVG_FUSE_MT / BVG_FUSE_RAW / BVG_FUSE_WRES override for experiments.
  const size_t budget = 227 * 1024 - fused_fixed_smem(P.NB, P.Cin_p);
  const size_t wsb = (size_t)P.NB * P.kc_max * 16;
  const int wslots = P.ntaps * P.n_ci_blk;
  const int mt_env = env_int("BVG_FUSE_MT", 0), raw_env = env_int("BVG_FUSE_RAW", 0), wres_env = env_int("BVG_FUSE_WRES", -1);
  size_t rsb = 0, xsb = 0;
  int mt_pick = 0;
  for (int mt = mt_env > 0 ? mt_env : 4; mt >= 1; --mt) {
    if (2 * mt * P.NB > 512 || mt * (P.NB >> 4) > 16) { if (mt_env > 0) break; continue; }
    const int xr = (mt * 128 + halo + kFV - 1) / kFV * kFV;
    rsb = (size_t)(xr + 16) * P.kc_max * 16;
    xsb = (size_t)xr * P.kc_max * 16;
    const size_t wneed = (wslots <= kFMaxW && wslots * wsb <= 80 * 1024) ? wslots * wsb : 2 * wsb;
    if (2 * rsb + kFAStages * xsb + wneed <= budget) { mt_pick = mt; P.XR = xr; break; }
    if (mt_env > 0) break;
  }
  if (!mt_pick) return BVG_ERR_STATE;
  P.MT = mt_pick;
  P.tiles_per_batch = (y.T + P.MT * 128 - 1) / (P.MT * 128);
  int pw = 32;
  while (pw < 2 * P.MT * P.NB) pw <<= 1;
  P.tmem_cols = pw;
  P.n_issuers = P.MT >= 2 ? 2 : 1;
  const size_t base_bytes = 2 * rsb + kFAStages * xsb;
  bool resident = wslots <= kFMaxW && base_bytes + wslots * wsb <= budget;
  if (wres_env == 0) resident = false;
  size_t used = base_bytes;
  if (resident) {
    P.w_resident = 1;
    P.w_stages = wslots;
    used += wslots * wsb;
  } else {
    P.w_resident = 0;
    P.w_stages = (int)std::min<size_t>({(size_t)kFMaxW, (budget - base_bytes) / wsb, (size_t)std::max(2, std::min(wslots, 6))});
    if (P.w_stages < 2) return BVG_ERR_STATE;
    used += P.w_stages * wsb;
  }
  P.x_stages = 2 + (int)std::min<size_t>(kFMaxRaw - 2, (budget - used) / rsb);
  if (raw_env >= 2) P.x_stages = std::min(P.x_stages, raw_env);
  const size_t smem = (size_t)P.x_stages * rsb + kFAStages * xsb + (size_t)P.w_stages * wsb + fused_fixed_smem(P.NB, P.Cin_p);
  static bool attr_set = false;
  if (!attr_set) {
    BVG_CUDA(cudaFuncSetAttribute(conv_umma_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int64_t ntiles = (int64_t)P.tiles_per_batch * B;
  BVG_CHECK_ARG(ntiles < (1ll << 31), "conv_umma_fused: too many tiles");
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    BVG_CUDA(cudaGetDevice(&dev));
    BVG_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  dim3 grid((unsigned)std::min<int64_t>(ntiles, num_sms));
  ProfScope prof(st, KC_CONV);
  conv_umma_fused_kernel<<<grid, kFThreads, smem, st>>>(P);
  BVG_LAUNCHED();
  return BVG_OK;
}

}  // namespace bvg
