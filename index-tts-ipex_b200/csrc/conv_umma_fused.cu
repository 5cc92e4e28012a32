// Fused Activation1d -> Conv1d for the narrow generator stages (C <= 128), tcgen05 / TMEM, sm_100a.
//
// AMPBlock1 always feeds its convs with an Activation1d output (models.py:65-74 of the reference:
// xt = c1(a1(x)); xt = c2(a2(xt)); x = xt + x).  For C <= 96 both ops are memory/issue bound, so here the activated
// tensor never exists in HBM: the TMA producer stages the RAW input tile (conv rows + 8-row FIR halo), the stencil
// warps run the packed-fp32x2 stencil of act1d_core.cuh on it and write the result straight into the K-major UMMA
// A-operand tile in shared memory (zero rows where the conv pads), the MMA issuers run the taps as row-shifted
// descriptors exactly as in conv_umma_kernel, and the epilogue warps drain the accumulators (+bias, +residual(s),
// *scale -> bf16).  HBM traffic per act+conv pair drops from 5 tensor passes to 2-3.
//
// Warp roles (640 threads = 5 warpgroups):
//   warps 0-11  stencil (384 lanes)                    warps 12-15  epilogue (one per TMEM lane quarter)
//   warp 16     TMA producer, raw input rows           warps 17-18  MMA issuers (accumulators dealt out)
//   warp 19     TMA producer, weights
// (the warp scheduler favours high warp ids: the latency-critical roles get them, the stencil fills the rest)
// The FIR stencil is FMA-pipe bound and is the critical resource, so the geometry is built around it:
//   * a lane-unit is one 32-bit channel-pair word x 16 rows (21 stencil iterations for 16 outputs); input channels
//     are processed in blocks of <= 48 (24 words) and an A stage holds XR = 16 * (384 / words) rows INCLUDING the
//     conv halo -- 256 rows for C = 96 / 48, 512 rows for C = 24 -- so every stage is exactly one unit per stencil
//     lane, no lane ever waits for a neighbour with a second unit.  A tile therefore yields XR - halo output
//     rows; the MMAs still run on whole 128-row accumulators and the epilogue drops the rows past the tile.
//   * only real channel words are computed (padding words of the A stages are zeroed once at start-up);
//   * the register file is re-balanced with setmaxnreg: control 48, epilogue 96, stencil 112
//     (4*48 + 4*96 + 12*112 = 20*96, the budget the CTA is launched with -- a larger sum makes
//     setmaxnreg.inc spin forever);
//   * the epilogue warps keep two accumulator slices and the next two residual slices in flight, so the
//     TMEM / global latencies overlap instead of adding up.
// Accumulation order is (channel block of 48, tap, 16-channel step); for Cin <= 48 that is the order of
// conv_umma_kernel and the result is bit-identical to the two-kernel path.
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
// The warp scheduler favours high warp ids, so the latency-critical roles sit at the top: control warpgroup 16-19,
// epilogue 12-15, and the throughput-bound stencil warps 0-11 soak up whatever issue slots are left.
constexpr int kFStWarps = 12;                 // stencil warps 0..11
constexpr int kFEpiWarp0 = 12;                // epilogue warps 12..15 (TMEM lane quarter = warp % 4)
constexpr int kFCtlWarp0 = 16;                // 16: raw-row producer, 17-18: MMA issuers, 19: weight producer
constexpr int kFLanes = kFStWarps * 32;       // 384 stencil lanes
constexpr int kFMaxRaw = 4;                   // raw-input stages (P.x_stages of them are used)
constexpr int kFMaxA = 4;                     // A-operand stages (P.a_stages)
constexpr int kFMaxW = 32;
constexpr int kFV = 16;                       // rows per lane-unit
constexpr int kFBlk = 48;                     // input channels per block (6 chunks of 8)

__host__ __device__ inline size_t fused_fixed_smem(int NB, int Cin_p) {
  return (size_t)(2 * kFMaxRaw + 2 * kFMaxA + 2 * kFMaxW + 4) * 8 + 16 + (size_t)NB * 4 + (size_t)Cin_p * 8;
}

// Epilogue of one tile for one epilogue warp (TMEM lane quarter wq): accumulator slices of NCH 8-channel chunks
// -> (+bias) (+res1) (+res2) (*scale) -> bf16 -> 16-byte stores.  Residual slices are register-prefetched one slice
// ahead (their rows were pulled into L2 a tile ahead by the caller), the first slice before the wait for the MMAs.
// HAS_R2 (the last conv of resblocks 2.. adds the running sum as a second residual) uses 16-column slices so that two
// residual streams fit the register budget.
template <int NCH, bool HAS_R2>
__device__ __forceinline__ void fused_epilogue_tile(const UmmaConvParams& P, __nv_bfloat16* yb, const __nv_bfloat16* r1,
                                                    const __nv_bfloat16* r2, const float* bias_s, uint32_t tbase, int q0, int r,
                                                    int nacc, uint64_t* tmem_full_bar, uint32_t aph, long long& dbg_ewait, int Tout_b) {
  constexpr int W = 8 * NCH;                                        // columns per slice
  const int nsl = (P.NB + W - 1) / W;                               // slices per accumulator (the last may be narrower)
  const int cs = P.y_tp * 8;                                        // elements between channel chunks
  const int cout8 = (P.Cout + 7) & ~7, ych8 = P.y_chunks * 8;
  const float scale = P.scale;
  // element offset (within the batch element) of accumulator row `a`, chunk 0; -1 if the row is not an output
  auto row_off = [&](int a) -> int {
    const int row = a * 128 + r;
    return (a < nacc && row < P.rows_out && q0 + row < Tout_b) ? (P.y_row0 + q0 + row) * 8 : -1;
  };
  auto load_res = [&](const __nv_bfloat16* rp, int a, int si, uint4 (&e)[NCH]) {
    const int ro = rp ? row_off(a) : -1;
#pragma unroll
    for (int g = 0; g < NCH; ++g) {
      const int co = si * W + 8 * g;
      e[g] = (ro >= 0 && co < cout8) ? *reinterpret_cast<const uint4*>(rp + ro + (co >> 3) * cs) : make_uint4(0, 0, 0, 0);
    }
  };
  uint4 cur1[NCH], nxt1[NCH], cur2[HAS_R2 ? NCH : 1], nxt2[HAS_R2 ? NCH : 1];
  load_res(r1, 0, 0, cur1);
  if (HAS_R2) load_res(r2, 0, 0, *reinterpret_cast<uint4(*)[NCH]>(&cur2[0]));
  { DBG_T0(); mbar_wait_relaxed(tmem_full_bar, aph); DBG_ADD(dbg_ewait); }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int a = 0, si = 0; a < nacc;) {
    int na = a, nsi = si + 1;
    if (nsi == nsl) { nsi = 0; ++na; }
    load_res(r1, na, nsi, nxt1);                                    // next slice's residuals overlap this slice's TMEM reads
    if (HAS_R2) load_res(r2, na, nsi, *reinterpret_cast<uint4(*)[NCH]>(&nxt2[0]));
    const int c0 = si * W;
    const bool wide = NCH == 4 && c0 + 16 < P.NB;                   // a 32-column slice may be cut to 16 at the end
    uint32_t v[W];
    tmem_ld16_nowait(tbase + (uint32_t)(a * P.NB + c0), *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
    if (NCH == 4 && wide) tmem_ld16_nowait(tbase + (uint32_t)(a * P.NB + c0 + 16), *reinterpret_cast<uint32_t(*)[16]>(&v[W - 16]));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int ro = row_off(a);
    if (ro >= 0) {
#pragma unroll
      for (int g = 0; g < NCH; ++g) {
        const int co = c0 + 8 * g;
        if ((g >= 2 && !wide) || co >= ych8) continue;              // (padding channels inside the tensor become zeros)
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[8 * g + j]) + bias_s[co + j];
        if (r1) { float ee[8]; unpack8(cur1[g], ee);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] += ee[j]; }
        if (HAS_R2) { float ee[8]; unpack8(cur2[HAS_R2 ? g : 0], ee);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] += ee[j]; }
        uint4 o;
        o.x = pack2(f[0] * scale, f[1] * scale); o.y = pack2(f[2] * scale, f[3] * scale);
        o.z = pack2(f[4] * scale, f[5] * scale); o.w = pack2(f[6] * scale, f[7] * scale);
        *reinterpret_cast<uint4*>(yb + ro + (co >> 3) * cs) = o;
      }
    }
#pragma unroll
    for (int g = 0; g < NCH; ++g) {
      cur1[g] = nxt1[g];
      if (HAS_R2) cur2[HAS_R2 ? g : 0] = nxt2[HAS_R2 ? g : 0];
    }
    a = na; si = nsi;
  }
}

__global__ void __launch_bounds__(kFThreads, 1) conv_umma_fused_kernel(const UmmaConvParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  // (the shuffle tells the compiler the warp index is warp-uniform: role-dependent values such as the MMA descriptors
  // can then stay in uniform registers instead of being broadcast out of a lane before every tcgen05.mma)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  const int nraw = P.x_stages, nast = P.a_stages, kWStages = P.w_stages;
  // Row pitches of a chunk inside a stage are == 1 (mod 8): consecutive chunks then start 4 banks apart and the
  // per-lane 32-bit accesses of the stencil (lane = channel-pair word) spread over all 32 banks.
  const int XLOAD = P.XR + 16;                                     // raw rows staged per chunk (8-row FIR halo each side)
  const int XRAW = P.XR + 17;                                      // raw row pitch
  const int XA = P.XR + 1;                                         // A-operand row pitch (= the descriptors' LBO)
  const uint32_t raw_stage_bytes = (uint32_t)XRAW * P.kc_max * 16u;
  const uint32_t x_stage_bytes = (uint32_t)XA * P.kc_max * 16u;
  const uint32_t w_stage_bytes = (uint32_t)P.NB * P.kc_max * 16u;
  uint8_t* rsm = smem;
  uint8_t* xsm = rsm + nraw * raw_stage_bytes;
  uint8_t* wsm = xsm + nast * x_stage_bytes;
  uint8_t* tail = wsm + kWStages * w_stage_bytes;
  uint64_t* full_raw = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_raw = full_raw + kFMaxRaw;
  uint64_t* full_x = empty_raw + kFMaxRaw;
  uint64_t* empty_x = full_x + kFMaxA;
  uint64_t* full_w = empty_x + kFMaxA;
  uint64_t* empty_w = full_w + kFMaxW;
  uint64_t* tmem_full = empty_w + kFMaxW;                          // [2]
  uint64_t* tmem_empty = tmem_full + 2;                            // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);   // [4] (keeps what follows 16-byte aligned)
  float* bias_s = reinterpret_cast<float*>(tmem_ptr + 4);          // [NB]
  float4* snk = reinterpret_cast<float4*>(bias_s + P.NB);          // [Cin_p/2] (sc0, sc1) of a channel pair

  const int nacc = P.MT;                                           // 128-row accumulators per tile
  const int acc_cols = nacc * P.NB;
  const int ntiles = P.tiles_per_batch * P.B;
  const int n_iss = P.n_issuers;
  const int nblk = (P.Cin_p + kFBlk - 1) / kFBlk;                  // input-channel blocks

  if (threadIdx.x == 0) {
    for (int i = 0; i < nraw; ++i) { mbar_init(&full_raw[i], 1); mbar_init(&empty_raw[i], kFStWarps); }
    for (int i = 0; i < nast; ++i) { mbar_init(&full_x[i], kFStWarps); mbar_init(&empty_x[i], n_iss); }
    for (int i = 0; i < kWStages; ++i) { mbar_init(&full_w[i], 1); mbar_init(&empty_w[i], n_iss); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], n_iss); mbar_init(&tmem_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kFCtlWarp0 + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(P.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  {
    // the A stages start as zeros: padding channel words are never written afterwards
    uint4* xz = reinterpret_cast<uint4*>(xsm);
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (uint32_t i = threadIdx.x; i < nast * (x_stage_bytes >> 4); i += kFThreads) xz[i] = z;
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

  if (warp >= kFCtlWarp0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    if (warp == kFCtlWarp0) {
      // ===================== TMA producer 1: raw input rows (real channel chunks only) =====================
      if (lane == 0) {
        int rs = 0;
        uint32_t rph = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
          const int mt = tile % P.tiles_per_batch;
          const int b = tile / P.tiles_per_batch;
          const int q0 = mt * P.rows_out;
          if (P.lens && q0 >= __ldg(P.lens + b) * P.len_mul) continue;          // ragged batch: tile past the utterance's end
          const __nv_bfloat16* xb = P.x + (int64_t)b * P.x_bstride;
          // raw tile row i <-> padded-space row r0 + i, r0 = x_row0 + q0 - lo - 8; clipped to the chunk [0, x_tp)
          // (rows outside only ever feed replicate-padded positions, which the stencil overrides)
          const int r0 = P.x_row0 + q0 - P.lo - 8;
          const int lo_r = max(r0, 0), hi_r = min(r0 + XLOAD, P.x_tp);
          const uint32_t nbytes = (uint32_t)(hi_r - lo_r) * 16u;
          for (int blk = 0; blk < nblk; ++blk) {
            const int kcr = min(kFBlk / 8, (P.Cin - blk * kFBlk + 7) >> 3);
            mbar_wait_relaxed(&empty_raw[rs], rph ^ 1);
            mbar_expect_tx(&full_raw[rs], nbytes * kcr);
            for (int kc = 0; kc < kcr; ++kc)
              bulk_g2s(smem_u32(rsm + rs * raw_stage_bytes) + (kc * XRAW + (lo_r - r0)) * 16,
                       xb + ((int64_t)(blk * (kFBlk / 8) + kc) * P.x_tp + lo_r) * 8, nbytes, &full_raw[rs]);
            if (++rs == nraw) { rs = 0; rph ^= 1; }
          }
        }
      }
    } else if (warp == kFCtlWarp0 + 3) {
      // ===================== TMA producer 2: weights (own warp: a full weight ring must not delay the raw rows) =====
      // A slot holds one (channel block, tap) tile [chunk][NB][8].  The packed weights are grouped in 64-channel
      // blocks ([cb64][tap][kchunk][NB][8]); a 48-channel block is one or two contiguous runs of that layout.
      if (lane == 0) {
        int ws = 0;
        uint32_t wph = 0;
        bool w_loaded = false;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
          if (P.lens && (tile % P.tiles_per_batch) * P.rows_out >= __ldg(P.lens + tile / P.tiles_per_batch) * P.len_mul) continue;
          if (P.w_resident && w_loaded) break;                         // resident weights are fetched once
          w_loaded = true;
          for (int blk = 0; blk < nblk; ++blk) {
            const int g0 = blk * (kFBlk / 8);                            // first global chunk of the block
            const int nc = min(kFBlk / 8, (P.Cin_p >> 3) - g0);          // chunks the MMAs of this block read
            const int cbA = g0 >> 3, kcA = g0 & 7;
            const int kcnA = min(8, (P.Cin_p - cbA * 64) >> 3);
            const int n1 = min(nc, kcnA - kcA), n2 = nc - n1;
            const int kcnB = n2 > 0 ? min(8, (P.Cin_p - (cbA + 1) * 64) >> 3) : 0;
            for (int tp = 0; tp < P.ntaps; ++tp) {
              const int slot = P.w_resident ? blk * P.ntaps + tp : ws;
              if (!P.w_resident) mbar_wait_relaxed(&empty_w[ws], wph ^ 1);
              mbar_expect_tx(&full_w[slot], (uint32_t)P.NB * nc * 16u);
              const uint32_t dst = smem_u32(wsm + slot * w_stage_bytes);
              bulk_g2s(dst, P.w + ((int64_t)cbA * 64 * P.ntaps + (int64_t)tp * kcnA * 8 + kcA * 8) * P.NB,
                       (uint32_t)P.NB * n1 * 16u, &full_w[slot]);
              if (n2 > 0)
                bulk_g2s(dst + (uint32_t)P.NB * n1 * 16u, P.w + ((int64_t)(cbA + 1) * 64 * P.ntaps + (int64_t)tp * kcnB * 8) * P.NB,
                         (uint32_t)P.NB * n2 * 16u, &full_w[slot]);
              if (!P.w_resident && ++ws == kWStages) { ws = 0; wph ^= 1; }
            }
          }
        }
      }
    } else {
      // ===================== MMA issuers (as in conv_umma_kernel, conv taps only) =====================
      const int ii = warp - (kFCtlWarp0 + 1);
      if (ii < n_iss) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P.NB >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a_lbo = (uint32_t)XA << 16, b_lbo = (uint32_t)P.NB << 16;
        const uint32_t astep = 2u * XA, bstep = 2u * P.NB;
        const uint32_t xsb16 = x_stage_bytes >> 4, wslot16 = w_stage_bytes >> 4;
        const uint32_t x_base = (smem_u32(xsm) >> 4) | a_lbo, w_base = (smem_u32(wsm) >> 4) | b_lbo;
        const int first_tile = blockIdx.x;
        int xs = 0, ws = 0, as = 0;
        uint32_t xph = 0, wph = 0, aph = 0;
        long long dbg_wx = 0, dbg_wt = 0, dbg_ww = 0;
        const long long dbg_start = P.dbg ? clock64() : 0;
        unsigned long long dbg_ns0 = 0;
        if (P.dbg) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_ns0));
        bool w_waited = false;
        for (int tile = first_tile; tile < ntiles; tile += gridDim.x) {
          if (P.lens && (tile % P.tiles_per_batch) * P.rows_out >= rows_of(P.lens, P.len_mul, tile / P.tiles_per_batch, P.Tout)) continue;
          { DBG_T0(); mbar_wait_backoff(&tmem_empty[as], aph ^ 1); DBG_ADD(dbg_wt); }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t dbase = tmem_base + (uint32_t)(as * acc_cols);
          for (int blk = 0; blk < nblk; ++blk) {
            const int nk = min(kFBlk / 8, (P.Cin_p >> 3) - blk * (kFBlk / 8)) >> 1;   // 16-channel steps (1..3)
            { DBG_T0(); mbar_wait_backoff(&full_x[xs], xph); DBG_ADD(dbg_wx); }   // the stencil warps filled this A stage
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_cb = x_base + (uint32_t)xs * xsb16;
            for (int tp = 0; tp < P.ntaps; ++tp) {
              const int slot = P.w_resident ? blk * P.ntaps + tp : ws;
              if (!P.w_resident || !w_waited) {
                DBG_T0();
                mbar_wait(&full_w[slot], P.w_resident ? 0u : wph);
                DBG_ADD(dbg_ww);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              }
              const uint32_t b_lo = w_base + (uint32_t)slot * wslot16;
              const uint32_t accum0 = (blk > 0 || tp > 0) ? 1u : 0u;
              {
                // the whole (converged) warp runs the loop on warp-uniform values; one elected lane issues
                const uint32_t a_tp = a_cb + (uint32_t)(tp * P.dil);
                for (int ms = ii; ms < nacc; ms += n_iss) {
                  const uint32_t d = dbase + (uint32_t)(ms * P.NB);
                  const uint32_t am = a_tp + (uint32_t)(ms * 128);
                  umma_bf16_imm_elect(d, am, b_lo, idesc, accum0);
                  if (nk > 1) umma_bf16_imm_elect(d, am + astep, b_lo + bstep, idesc, 1u);
                  if (nk > 2) umma_bf16_imm_elect(d, am + 2 * astep, b_lo + 2 * bstep, idesc, 1u);
                }
              }
              if (!P.w_resident) {
                umma_commit_elect(&empty_w[ws]);
                if (++ws == kWStages) { ws = 0; wph ^= 1; }
              }
            }
            umma_commit_elect(&empty_x[xs]);
            if (++xs == nast) { xs = 0; xph ^= 1; }
          }
          umma_commit_elect(&tmem_full[as]);
          w_waited = true;
          if (++as == P.acc_stages) { as = 0; aph ^= 1; }
        }
        if (P.dbg && lane == 0 && ii == 0) { long long* d = P.dbg + blockIdx.x * 16; d[3] = dbg_wx; d[4] = dbg_wt; d[5] = clock64() - dbg_start; d[8] = dbg_ww;
          unsigned long long ns1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1)); d[9] = (long long)(ns1 - dbg_ns0); }
      }
    }
  } else if (warp >= kFEpiWarp0) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
    // ===================== epilogue (4 warps, one per TMEM lane quarter) =====================
    const int wq = warp & 3;
    const int r = wq * 32 + lane;                                   // accumulator row
    const int cs = P.y_tp * 8;                                      // elements between channel chunks
    const int cout8 = (P.Cout + 7) & ~7;
    int as = 0;
    uint32_t aph = 0;
    long long dbg_ewait = 0, dbg_ebusy = 0;
    // pull the residual rows of a tile into L2 (one prefetch per 16-byte chunk row of this lane's accumulator rows)
    auto l2_prefetch_tile = [&](int t) {
      if (t >= ntiles || !(P.res1 || P.res2)) return;
      const int nq0 = (t % P.tiles_per_batch) * P.rows_out;
      const int64_t boff = (int64_t)(t / P.tiles_per_batch) * P.y_bstride;
      for (int a = 0; a < nacc; ++a) {
        const int row = a * 128 + r;
        if (row < P.rows_out && nq0 + row < P.Tout) {                  // (ragged batches: a few prefetches past the end, harmless)
          const int64_t ro = boff + (P.y_row0 + nq0 + row) * 8;
          for (int co = 0; co < cout8; co += 8) {
            if (P.res1) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.res1 + ro + (co >> 3) * cs));
            if (P.res2) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.res2 + ro + (co >> 3) * cs));
          }
        }
      }
    };
    l2_prefetch_tile(blockIdx.x);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int mt = tile % P.tiles_per_batch;
      const int b = tile / P.tiles_per_batch;
      const int q0 = mt * P.rows_out;
      const int Tout_b = rows_of(P.lens, P.len_mul, b, P.Tout);
      if (q0 >= Tout_b) continue;                                     // ragged batch: tile past the utterance's end
      __nv_bfloat16* yb = P.y + (int64_t)b * P.y_bstride;
      const __nv_bfloat16* r1 = P.res1 ? P.res1 + (int64_t)b * P.y_bstride : nullptr;
      const __nv_bfloat16* r2 = P.res2 ? P.res2 + (int64_t)b * P.y_bstride : nullptr;
      // the residual rows of this CTA's NEXT tile go to L2 a whole tile ahead, so that the one-slice-ahead register
      // prefetch only has to cover an L2 hit, not an HBM miss
      l2_prefetch_tile(tile + gridDim.x);
      DBG_T0();
      const long long ew0 = dbg_ewait;
      const uint32_t tbase = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(as * acc_cols);
      if (r2) fused_epilogue_tile<2, true>(P, yb, r1, r2, bias_s, tbase, q0, r, nacc, &tmem_full[as], aph, dbg_ewait, Tout_b);
      else fused_epilogue_tile<4, false>(P, yb, r1, nullptr, bias_s, tbase, q0, r, nacc, &tmem_full[as], aph, dbg_ewait, Tout_b);
      // all of this warp's TMEM reads for the stage are complete: hand it back to the MMA issuers
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      if (++as == P.acc_stages) { as = 0; aph ^= 1; }
      if (P.zero_pads) {
        // rows [-PAD, 0) by the first tile, [Tout, Tout+PAD) by the last: keeps the c8t zero halo intact
        const int et = threadIdx.x - kFEpiWarp0 * 32;
        const int chn = P.y_chunks;
        const uint4 z = make_uint4(0, 0, 0, 0);
        if (mt == 0)
          for (int i = et; i < chn * P.y_row0; i += 128)
            *reinterpret_cast<uint4*>(yb + ((int64_t)(i / P.y_row0) * P.y_tp + (i % P.y_row0)) * 8) = z;
        if (mt == (Tout_b - 1) / P.rows_out)
          for (int i = et; i < chn * P.y_row0; i += 128)
            *reinterpret_cast<uint4*>(yb + ((int64_t)(i / P.y_row0) * P.y_tp + P.y_row0 + Tout_b + (i % P.y_row0)) * 8) = z;
      }
      DBG_ADD(dbg_ebusy);
      dbg_ebusy -= dbg_ewait - ew0;
    }
    if (P.dbg && threadIdx.x == kFEpiWarp0 * 32) { P.dbg[blockIdx.x * 16 + 6] = dbg_ewait; P.dbg[blockIdx.x * 16 + 7] = dbg_ebusy; }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    // ===================== stencil warps: raw tile -> Activation1d -> A-operand tile =====================
    constexpr int V = kFV;
    const int L = warp * 32 + lane;                   // stencil lane 0..383
    const int ngroups = P.XR / V;                                   // 16-row groups of the A tile (XR is a multiple of 16)
    int rs = 0, xs = 0;
    uint32_t rph = 0, xph = 0;
    long long dbg_araw = 0, dbg_ax = 0, dbg_abusy = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int mt = tile % P.tiles_per_batch;
      const int q0 = mt * P.rows_out;
      const int Tout_b = P.lens ? __ldg(P.lens + tile / P.tiles_per_batch) * P.len_mul : P.Tout;
      if (q0 >= Tout_b) continue;                                   // ragged batch: tile past the utterance's end
      const int64_t t_first = (int64_t)q0 - P.lo;                   // time index of A-tile row 0
      for (int blk = 0; blk < nblk; ++blk) {
        const int nwords = min(kFBlk / 2, (P.Cin - blk * kFBlk + 1) >> 1);   // real channel-pair words per row in this block
        const int U = ngroups * nwords;                             // lane-units of this stage (== 384 for the model's shapes)
        { DBG_T0(); mbar_wait_relaxed(&full_raw[rs], rph); DBG_ADD(dbg_araw); }
        { DBG_T0(); mbar_wait_relaxed(&empty_x[xs], xph ^ 1); DBG_ADD(dbg_ax); }
        DBG_T0();
        const uint32_t* raw = reinterpret_cast<const uint32_t*>(rsm + rs * raw_stage_bytes);
        uint32_t* xo = reinterpret_cast<uint32_t*>(xsm + xs * x_stage_bytes);
        for (int e = L; e < (P.dry ? 0 : U); e += kFLanes) {
          const int rg = e / nwords;
          const int wrd = e - rg * nwords;
          const int cg = wrd >> 2, pp = wrd & 3;
          const int j0 = rg * V;                                    // first A-tile row of this unit
          const int64_t t0 = t_first + j0;
          const int chA = blk * kFBlk + cg * 8 + 2 * pp;
          const uint32_t* inw = raw + (size_t)(cg * XRAW + j0) * 4 + pp;      // window row i <-> raw row j0 + i
          uint32_t* outw = xo + (size_t)(cg * XA + j0) * 4 + pp;
          const float4 sp = snk[chA >> 1];
          const bool interior = (t0 - 5 >= 0) && (t0 + V + 4 <= (int64_t)Tout_b - 1) && (chA + 1 < P.Cin);
          if (interior) {
            act1d_window2<V>([&](int j) { return unpack_bf16x2(inw[j * 4]); },
                             [&](int q, float ya, float yb) { outw[q * 4] = pack2(ya, yb); },
                             pk2(sp.x, sp.z), pk2(sp.y, sp.w));
          } else if ((t0 + V - 1 >= 0) && (t0 < Tout_b)) {
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
                act1d_window<V, false>(xw, yv, h ? sp.z : sp.x, h ? sp.w : sp.y, t0, (int64_t)Tout_b);
#pragma unroll
                for (int q = 0; q < V; ++q)
                  if (t0 + q < 0 || t0 + q >= Tout_b) yv[q] = 0.f;       // the conv's zero padding
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
        // generic-proxy writes of the A tile -> visible to the tensor core's async-proxy reads
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) { mbar_arrive(&full_x[xs]); mbar_arrive(&empty_raw[rs]); }
        DBG_ADD(dbg_abusy);
        if (++rs == nraw) { rs = 0; rph ^= 1; }
        if (++xs == nast) { xs = 0; xph ^= 1; }
      }
    }
    if (P.dbg && threadIdx.x == 0) { long long* d = P.dbg + blockIdx.x * 16; d[0] = dbg_araw; d[1] = dbg_ax; d[2] = dbg_abusy; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kFCtlWarp0 + 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(P.tmem_cols));
  }
}

}  // namespace

// Activation1d(x) -> Conv1d fused (see conv_umma_fused_kernel).  Returns BVG_ERR_STATE without launching when the
// layer does not qualify (caller falls back to act1d_c8t_launch + conv_umma_launch).
int conv_umma_fused_launch(const UmmaLayer& L, const C8T& x, const float* act_alpha, const float* act_beta,
                           const C8T& y, const UmmaEpilogue& ep, int64_t B, cudaStream_t st, int max_nb) {
  if (L.transposed || !act_alpha || !act_beta) return BVG_ERR_STATE;
  if (max_nb <= 0) max_nb = BVG_ENV_ONCE("BVG_FUSE_MAX_NB", 128);
  BVG_CHECK_ARG(L.w && x.p && y.p, "conv_umma_fused: null pointer");
  BVG_CHECK_ARG(x.C == L.Cin && y.C == L.Cout && y.T == x.T, "conv_umma_fused: shape mismatch");
  UmmaConvParams P;
  memset(&P, 0, sizeof P);
  int n_nblk = 1;
  P.NPH = 1;
  umma_choose_nb(L.Cout, 1, &P.NB, &n_nblk);
  if (n_nblk != 1 || P.NB > max_nb || ep.cond || ep.relu || ep.post_scale || ep.act) return BVG_ERR_STATE;
  P.ntaps = L.K;
  BVG_CHECK_ARG(L.K <= 16, "conv_umma_fused: at most 16 taps");
  // Wide-and-long layers (C = 96, k = 11) are bound by the number of narrow tcgen05.mma instructions (~100-150 cycles
  // each whatever N is) rather than by the stencil; in isolation the two-kernel path is faster there, inside the
  // decode the fused kernel still wins (one tensor pass less through HBM), so the cut-off is off by default.
  if ((int64_t)((L.Cin + 15) / 16 * 16) * L.K > BVG_ENV_ONCE("BVG_FUSE_MAX_CK", 1 << 30)) return BVG_ERR_STATE;
  const int halo = L.dil * (L.K - 1);
  P.lo = halo / 2;
  BVG_CHECK_ARG(P.lo <= x.pad, "conv_umma_fused: conv padding %d exceeds the c8t halo %d", P.lo, x.pad);
  P.dil = L.dil;
  P.u = 1; P.p = 0;
  P.Tout = y.T;
  P.Cin = L.Cin;
  P.Cin_p = (L.Cin + 15) / 16 * 16;
  BVG_CHECK_ARG(x.chunks * 8 >= P.Cin_p, "conv_umma_fused: input tensor must carry channel padding to a multiple of 16");
  P.n_ci_blk = (P.Cin_p + kFBlk - 1) / kFBlk;
  P.Cout = L.Cout;
  P.x = x.p; P.x_bstride = (int64_t)x.chunks * x.Tp * 8; P.x_tp = x.Tp; P.x_row0 = x.pad;
  P.y = y.p; P.y_bstride = (int64_t)y.chunks * y.Tp * 8; P.y_tp = y.Tp; P.y_row0 = y.pad; P.y_chunks = y.chunks;
  P.w = L.w;
  P.bias = ep.bias; P.scale = ep.scale; P.res1 = ep.res1; P.res2 = ep.res2; P.zero_pads = ep.zero_pads;
  P.act_alpha = act_alpha; P.act_beta = act_beta;
  P.lens = y.lens; P.len_mul = y.len_mul;
  P.dbg = ep.dbg;
#ifdef BVG_DEBUG
  P.dry = BVG_ENV_ONCE("BVG_FUSE_DRY", 0);     // debug builds only: skip the stencil math (timing experiments, results are garbage)
#endif
  P.acc_stages = 2;
  P.n_nblk = 1;
  P.B = (int)B;
  P.kc_max = std::min(kFBlk / 8, P.Cin_p / 8);
  // Tile plan: XR rows per A stage such that one stage is one lane-unit per stencil lane (see the header), capped by
  // 512 rows and by two accumulator stages in TMEM.  BVG_FUSE_XR overrides for experiments.
  const int nwords = (std::min(L.Cin, kFBlk) + 1) / 2;
  int xr = std::min(512, std::max(1, kFLanes / nwords) * kFV);
#ifdef BVG_DEBUG
  if (const int e = BVG_ENV_ONCE("BVG_FUSE_XR", 0)) xr = e / kFV * kFV;   // debug builds only: untested tile geometries
#endif
  auto nacc_of = [&](int rows) { return (rows - halo + 127) / 128; };
  // two accumulator stages (the epilogue of tile i overlaps the MMAs of tile i+1) when they fit the 512 TMEM columns
  // at full stencil occupancy; wide layers (C = 192) keep the full tile and run with one stage instead
  if (2 * nacc_of(xr) * P.NB > 512 && nacc_of(xr) * P.NB <= 512 && P.NB > 128) P.acc_stages = 1;
  while (xr > halo + kFV && P.acc_stages * nacc_of(xr) * P.NB > 512) xr -= kFV;
  if (xr < halo + kFV || P.acc_stages * nacc_of(xr) * P.NB > 512) return BVG_ERR_STATE;
  P.XR = xr;
  P.rows_out = xr - halo;
  P.MT = nacc_of(xr);
  P.tiles_per_batch = (y.T + P.rows_out - 1) / P.rows_out;
  int pw = 32;
  while (pw < P.acc_stages * P.MT * P.NB) pw <<= 1;
  P.tmem_cols = pw;
  P.n_issuers = P.MT >= 2 ? 2 : 1;
  // shared memory: raw ring + A ring + weights (resident when they fit next to 2+2 stages, else a ring of slots)
  const size_t budget = 227 * 1024 - fused_fixed_smem(P.NB, P.Cin_p);
  const size_t rsb = (size_t)(xr + 17) * P.kc_max * 16, xsb = (size_t)(xr + 1) * P.kc_max * 16, wsb = (size_t)P.NB * P.kc_max * 16;
  const int wslots = P.ntaps * P.n_ci_blk;
  if (2 * rsb + 2 * xsb + 2 * wsb > budget) return BVG_ERR_STATE;
  size_t used = 2 * rsb + 2 * xsb;
  bool resident = wslots <= kFMaxW && used + wslots * wsb <= budget;
  if (BVG_ENV_ONCE("BVG_FUSE_WRES", 1) == 0) resident = false;
  if (resident) {
    P.w_resident = 1;
    P.w_stages = wslots;
  } else {
    P.w_resident = 0;
    // (with one accumulator stage the A ring, not the weight ring, is what keeps the stencil running: cap at 4 slots)
    P.w_stages = (int)std::min<size_t>({(size_t)kFMaxW, (budget - used) / wsb,
                                        (size_t)std::max(2, std::min(wslots, P.acc_stages == 1 ? 4 : 8))});
  }
  used += P.w_stages * wsb;
  // leftover: deepen the A ring first (lets the stencil run ahead of the MMAs), then the raw ring
  const int max_a = std::min(kFMaxA, std::max(2, BVG_ENV_ONCE("BVG_FUSE_A", 3)));
  const int max_raw = std::min(kFMaxRaw, std::max(2, BVG_ENV_ONCE("BVG_FUSE_RAW", 3)));
  P.a_stages = 2; P.x_stages = 2;
  for (bool grew = true; grew;) {
    grew = false;
    if (P.acc_stages == 1 && P.a_stages < max_a && used + xsb <= budget) { ++P.a_stages; used += xsb; grew = true; }
    if (P.x_stages < max_raw && used + rsb <= budget) { ++P.x_stages; used += rsb; grew = true; }
    if (P.acc_stages == 2 && P.a_stages < max_a && used + xsb <= budget) { ++P.a_stages; used += xsb; grew = true; }
  }
  const size_t smem = used + fused_fixed_smem(P.NB, P.Cin_p);
  static std::atomic<uint64_t> opted{0};
  BVG_TRY(smem_opt_in(conv_umma_fused_kernel, opted, 227 * 1024));
  const int64_t ntiles = (int64_t)P.tiles_per_batch * B;
  BVG_CHECK_ARG(ntiles < (1ll << 31), "conv_umma_fused: too many tiles");
  int num_sms = 0;
  BVG_TRY(current_device_sms(&num_sms));
  dim3 grid((unsigned)std::min<int64_t>(ntiles, num_sms));
  ProfScope prof(st, KC_ACTCONV);
  conv_umma_fused_kernel<<<grid, kFThreads, smem, st>>>(P);
  BVG_LAUNCHED();
  return BVG_OK;
}

}  // namespace bvg
