// tcgen05 / TMEM implicit-GEMM Conv1d and ConvTranspose1d for sm_100a (bf16 operands, fp32 accumulate).
//
// Replaces the cuDNN calls behind torch.nn.Conv1d / ConvTranspose1d in the generator
// (models.py:25-42 AMPBlock1 convs, :149 conv_pre, :155-161 ups) on the bf16 throughput path.
//
// Layout ("c8t"): activations live in HBM as [B][C/8][Tp][8] bf16 -- 8 channels (16 bytes)
// innermost, then time, then channel-chunk -- with PAD zero rows before t=0 and after t=T-1
// (Tp = T + 2*PAD).  With that layout
//   * a [rows x 64 ch] input tile is 8 contiguous runs of rows*16 B  -> 8 TMA bulk copies
//     (cp.async.bulk, UBLKCP) land it in shared memory as [kchunk][row][8], which is exactly the
//     canonical K-major no-swizzle UMMA operand layout (core matrix = 8 rows x 16 B contiguous,
//     SBO = 128 B between 8-row groups, LBO = rows*16 B between K chunks);
//   * a conv tap is a ROW OFFSET of the A operand: the same staged tile serves all K taps by
//     moving the descriptor start address by tap*dilation*16 B -- no re-load, no im2col;
//   * conv zero padding is the PAD rows; out-of-tile rows only ever feed discarded outputs.
// GEMM view: D[t, co] (+)= X[t + shift_k, ci] * W_k[co, ci]   M = 128 time rows per accumulator,
// N = NB <= 256 output channels, K = 16 input channels per tcgen05.mma.  Accumulators stay in
// TMEM; a CTA owns MT*NPH of them (conv: 2 time sub-tiles sharing every weight load;
// ConvTranspose: one accumulator per output phase).  Epilogue: tcgen05.ld -> +bias +cond
// +residual(s), *scale -> bf16 -> 16-byte coalesced stores (lane = time row).
//
// Warp roles (416 threads, persistent CTAs): warp 0 = TMA producer, warps 1..4 = MMA issuers (accumulators dealt
// out among them, descriptors in uniform registers), warps 5..12 = epilogue (TMEM lane quarter x column half).
// The narrow stages of the decode do not come here: see conv_umma_fused.cu.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "act1d_core.cuh"
#include "bvg_common.cuh"
#include "umma.cuh"
#include "umma_ptx.cuh"

namespace bvg {
namespace {

constexpr int kMaxXStages = 4;
constexpr int kMaxWStages = 32;
constexpr int kEpiWarps = 8;
constexpr int kMaxIssuers = 4;                         // MMA-issuing warps (one tcgen05.mma stream each)
constexpr int kEpiWarp0 = 1 + kMaxIssuers;             // first epilogue warp (kEpiWarp0 % 4 == 1)
constexpr int kThreads = (kEpiWarp0 + kEpiWarps) * 32;


// Persistent kernel: each CTA walks tiles (m_tile, n_block, batch) with a static stride; the smem rings and
// the TMEM accumulator stages keep flowing across tiles, so the epilogue of tile i overlaps the
// mainloop of tile i+1 whenever two accumulator stages fit in TMEM.
__global__ void __launch_bounds__(kThreads, 1) conv_umma_kernel(const UmmaConvParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  // (the shuffle tells the compiler the warp index is warp-uniform: role-dependent values such as the MMA descriptors
  // then stay in uniform registers instead of being broadcast out of a lane before every tcgen05.mma)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  const int kXStages = P.x_stages, kWStages = P.w_stages;
  const uint32_t x_stage_bytes = (uint32_t)P.XR * P.kc_max * 16u;  // kc_max kchunks x XR rows x 16 B
  const uint32_t w_stage_bytes = (uint32_t)P.NB * P.kc_max * 16u;  // kc_max kchunks x NB rows x 16 B
  uint8_t* xsm = smem;
  uint8_t* wsm = smem + kXStages * x_stage_bytes;
  uint8_t* tail = wsm + kWStages * w_stage_bytes;
  uint64_t* full_x = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_x = full_x + kMaxXStages;
  uint64_t* full_w = empty_x + kMaxXStages;
  uint64_t* empty_w = full_w + kMaxWStages;
  uint64_t* tmem_full = empty_w + kMaxWStages;                    // [2]
  uint64_t* tmem_empty = tmem_full + 2;                           // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* bias_s = reinterpret_cast<float*>(tmem_ptr + 4);         // [n_nblk * NB] (+32 pad), 16-byte aligned

  const int nacc = P.MT * P.NPH;
  const int acc_cols = nacc * P.NB;
  const int ntiles = P.tiles_per_batch * P.n_nblk * P.B;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kXStages; ++i) { mbar_init(&full_x[i], 1); mbar_init(&empty_x[i], P.n_issuers); }
    for (int i = 0; i < kWStages; ++i) { mbar_init(&full_w[i], 1); mbar_init(&empty_w[i], P.n_issuers); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], P.n_issuers); mbar_init(&tmem_empty[i], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(P.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp >= kEpiWarp0) {
    for (int i = threadIdx.x - kEpiWarp0 * 32; i < P.n_nblk * P.NB; i += kEpiWarps * 32)
      bias_s[i] = (P.bias && i < P.Cout) ? P.bias[i] : 0.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && !P.dry) {
      int xs = 0, ws = 0;
      uint32_t xph = 0, wph = 0;
      long long dbg_prod_wait = 0;
      bool w_loaded = false;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int nb = tile % P.n_nblk;
        const int mt = (tile / P.n_nblk) % P.tiles_per_batch;
        const int b = tile / (P.n_nblk * P.tiles_per_batch);
        const int q0 = mt * P.MT * 128;
        if (P.lens && q0 * P.u - P.p >= __ldg(P.lens + b) * P.len_mul) continue;    // ragged batch: tile past the utterance's end
        const __nv_bfloat16* xb = P.x + (int64_t)b * P.x_bstride;
        const int64_t row_start = (int64_t)P.x_row0 + q0 - P.lo;    // first staged row within a chunk
        for (int cb = 0; cb < P.n_ci_blk; ++cb) {
          const int kcn = min(8, (P.Cin_p - cb * 64) >> 3);
          { DBG_T0(); mbar_wait_relaxed(&empty_x[xs], xph ^ 1); DBG_ADD(dbg_prod_wait); }
          mbar_expect_tx(&full_x[xs], (uint32_t)kcn * P.XR * 16u);
          for (int kc = 0; kc < kcn; ++kc)
            bulk_g2s(smem_u32(xsm + xs * x_stage_bytes) + kc * P.XR * 16,
                     xb + ((int64_t)(cb * 8 + kc) * P.x_tp + row_start) * 8, (uint32_t)P.XR * 16u, &full_x[xs]);
          if (++xs == kXStages) { xs = 0; xph ^= 1; }
          const uint32_t wbytes = (uint32_t)P.NB * kcn * 16u;
          const __nv_bfloat16* wsrc = P.w + ((int64_t)nb * P.Cin_p + (int64_t)cb * 64) * P.NB * P.ntaps;
          if (P.w_resident) {
            // the whole layer's weights stay in shared memory: fetched once, on this CTA's first tile
            if (!w_loaded)
              for (int tp = 0; tp < P.ntaps; ++tp) {
                const int slot = cb * P.ntaps + tp;
                mbar_expect_tx(&full_w[slot], wbytes);
                bulk_g2s(smem_u32(wsm + slot * w_stage_bytes), wsrc + (int64_t)tp * kcn * 8 * P.NB, wbytes, &full_w[slot]);
              }
          } else {
            for (int tp = 0; tp < P.ntaps; ++tp) {
              { DBG_T0(); mbar_wait_relaxed(&empty_w[ws], wph ^ 1); DBG_ADD(dbg_prod_wait); }
              mbar_expect_tx(&full_w[ws], wbytes);
              bulk_g2s(smem_u32(wsm + ws * w_stage_bytes), wsrc + (int64_t)tp * kcn * 8 * P.NB, wbytes, &full_w[ws]);
              if (++ws == kWStages) { ws = 0; wph ^= 1; }
            }
          }
        }
        w_loaded = true;                                            // (first PROCESSED tile: ragged batches skip tiles)
      }
      if (P.dbg) P.dbg[blockIdx.x * 8 + 0] = dbg_prod_wait;
    }
  } else if (warp < kEpiWarp0) {
    // ===================== MMA issuers =====================
    // A single warp sustains roughly one tcgen05.mma per ~100 cycles, twice the execution time of a narrow
    // (N <= 96) MMA, so the accumulators are dealt out to n_issuers warps: issuer ii owns accumulators
    // acc % n_issuers == ii and runs its own tcgen05.mma / tcgen05.commit stream; every ring barrier that the
    // MMAs release therefore expects n_issuers arrivals.
    // The whole warp runs this loop with warp-uniform values (descriptor words stay in uniform registers and
    // advance with uniform adds); only the tcgen05.mma / tcgen05.commit instructions sit under elect_one.
    // The issue path between two MMAs must stay well under the ~54-cycle execution time of a small-N MMA.
    const int ii = warp - 1;
    if (ii < P.n_issuers) {
      // The whole warp runs this loop on warp-uniform values: descriptor words live in uniform registers, advance with
      // uniform adds and feed tcgen05.mma directly (no per-lane broadcast); one elected lane issues.
      // instruction descriptor: D fp32, A/B bf16, both K-major, N = NB, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P.NB >> 3) << 17) | ((128u >> 4) << 24);
      // smem descriptors (K-major, no swizzle): lo word = start>>4 | (LBO>>4)<<16, hi word = SBO>>4 | version
      const uint32_t a_lbo = (uint32_t)P.XR << 16, b_lbo = (uint32_t)P.NB << 16;
      const uint32_t astep = 2u * P.XR, bstep = 2u * P.NB;            // 16 input channels further along K
      const int NB = P.NB, MT = P.MT, NPH = P.NPH, dil = P.dil, n_iss = P.n_issuers, ntaps = P.ntaps;
      const int n_ci_blk = P.n_ci_blk, Cin_p = P.Cin_p, resident = P.w_resident, dry = P.dry, transposed = P.transposed;
      const int n_xst = kXStages, n_wst = kWStages, n_ast = P.acc_stages, acols = acc_cols, tap_mod = P.tap_mod;
      const int hi_chunks = P.split_hi_chunks;
      const uint32_t xsb16 = x_stage_bytes >> 4, wslot16 = w_stage_bytes >> 4;
      const uint32_t x_base = (smem_u32(xsm) >> 4) | a_lbo, w_base = (smem_u32(wsm) >> 4) | b_lbo;
      const int first_tile = blockIdx.x;
      int xs = 0, ws = 0, as = 0;
      uint32_t xph = 0, wph = 0, aph = 0;
      long long dbg_wx = 0, dbg_ww = 0, dbg_wt = 0, dbg_issue = 0;   // (dbg_issue spans a whole channel block, weight waits included)
      const long long dbg_start = P.dbg ? clock64() : 0;
      bool w_waited = false;
      for (int tile = first_tile; tile < ntiles; tile += gridDim.x) {
        if (P.lens) {                                               // ragged batch: skip tiles past the utterance's end (as the other roles do)
          const int mt_ = (tile / P.n_nblk) % P.tiles_per_batch, b_ = tile / (P.n_nblk * P.tiles_per_batch);
          if (mt_ * P.MT * 128 * P.u - P.p >= rows_of(P.lens, P.len_mul, b_, P.Tout)) continue;
        }
        if (!dry) { DBG_T0(); mbar_wait_backoff(&tmem_empty[as], aph ^ 1); DBG_ADD(dbg_wt); }   // epilogue drained this accumulator stage
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t dbase = tmem_base + (uint32_t)(as * acols);
        for (int cb = 0; cb < n_ci_blk; ++cb) {
          const int nk = min(8, (Cin_p - cb * 64) >> 3) >> 1;       // MMAs along K in this channel block (1..4)
          if (!dry) { DBG_T0(); mbar_wait_backoff(&full_x[xs], xph); DBG_ADD(dbg_wx); }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          DBG_T0();
          const uint32_t a_cb = x_base + (uint32_t)xs * xsb16;
          for (int tp = 0; tp < ntaps; ++tp) {
            const int slot = resident ? cb * ntaps + tp : ws;
            if (!dry && (!resident || !w_waited)) {                 // resident weights are waited for once
              const long long ww0 = P.dbg ? clock64() : 0;
              mbar_wait(&full_w[slot], resident ? 0u : wph);
              if (P.dbg) dbg_ww += clock64() - ww0;
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const uint32_t b_lo = w_base + (uint32_t)slot * wslot16;
            const uint32_t accum0 = (cb > 0 || tp >= NPH) ? 1u : 0u;  // taps 0..NPH-1 are the first to touch their accumulator
            // split convs: the w_lo taps (tp >= tap_mod) multiply the lo half of the input by zero weights -- skip those MMAs
            const int nk_t = (hi_chunks > 0 && tp >= tap_mod) ? max(0, min(nk, (hi_chunks - cb * 8 + 1) >> 1)) : nk;
            if (!transposed) {
              // conv: tap = row shift tp*dil; issuer ii owns time sub-tiles ms = ii, ii + n_issuers, ...
              const uint32_t a_tp = a_cb + (uint32_t)((tp >= tap_mod ? tp - tap_mod : tp) * dil);
              for (int ms = ii; ms < MT; ms += n_iss) {
                const uint32_t d = dbase + (uint32_t)(ms * NB);
                uint32_t am = a_tp + (uint32_t)(ms * 128), bm = b_lo;
                if (nk_t > 0) umma_bf16_imm_elect(d, am, bm, idesc, accum0);
                for (int k = 1; k < nk_t; ++k) {
                  am += astep; bm += bstep;
                  umma_bf16_imm_elect(d, am, bm, idesc, 1u);
                }
              }
            } else if ((P.tap_acc[tp] & (n_iss - 1)) == ii) {
              // ConvTranspose: one accumulator per output phase, issuer ii owns phases == ii (mod n_issuers)
              const uint32_t d = dbase + (uint32_t)(P.tap_acc[tp] * NB);
              uint32_t am = a_cb + (uint32_t)P.tap_shift[tp], bm = b_lo;
              if (nk_t > 0) umma_bf16_imm_elect(d, am, bm, idesc, accum0);
              for (int k = 1; k < nk_t; ++k) {
                am += astep; bm += bstep;
                umma_bf16_imm_elect(d, am, bm, idesc, 1u);
              }
            }
            if (!resident && !dry) {
              umma_commit_elect(&empty_w[ws]);        // weight slot reusable once these MMAs retire
              if (++ws == n_wst) { ws = 0; wph ^= 1; }
            }
          }
          DBG_ADD(dbg_issue);
          if (!dry) umma_commit_elect(&empty_x[xs]);
          if (++xs == n_xst) { xs = 0; xph ^= 1; }
        }
        if (!dry) umma_commit_elect(&tmem_full[as]);
        w_waited = true;
        if (++as == n_ast) { as = 0; aph ^= 1; }
      }
      if (dry) {                         // drain: wait for every issued MMA so the total is inclusive
        umma_commit_elect(&tmem_full[0]);
        mbar_wait(&tmem_full[0], 0);
      }
      if (P.dbg && lane == 0 && ii == 0) {
        long long* d = P.dbg + blockIdx.x * 8;
        d[1] = dbg_wx; d[2] = dbg_ww; d[3] = dbg_wt; d[4] = dbg_issue; d[5] = clock64() - dbg_start;
      }
    }
  } else {
    // ===================== epilogue (kEpiWarps warps: TMEM lane quarter x column half) =====================
    const int wq = warp & 3;                                       // TMEM lane quarter this warp may access
    const int half = (warp - kEpiWarp0) >> 2;                      // which 32-column groups this warp takes
    const int r = wq * 32 + lane;
    int as = 0;
    uint32_t aph = 0;
    long long dbg_ewait = 0, dbg_ebusy = 0;
    for (int tile = blockIdx.x; tile < (P.dry ? 0 : ntiles); tile += gridDim.x) {
      const int nb = tile % P.n_nblk;
      const int mt = (tile / P.n_nblk) % P.tiles_per_batch;
      const int b = tile / (P.n_nblk * P.tiles_per_batch);
      const int q0 = mt * P.MT * 128;
      const int Tout_b = rows_of(P.lens, P.len_mul, b, P.Tout);
      if (q0 * P.u - P.p >= Tout_b) continue;                           // ragged batch: tile past the utterance's end
      __nv_bfloat16* yb = P.y + (int64_t)b * P.y_bstride;
      const __nv_bfloat16* r1 = P.res1 ? P.res1 + (int64_t)b * P.y_bstride : nullptr;
      const __nv_bfloat16* r2 = P.res2 ? P.res2 + (int64_t)b * P.y_bstride : nullptr;
      const float* cond = P.cond ? P.cond + (int64_t)(P.cond_B == 1 ? 0 : b) * P.Cout : nullptr;
      const float* bs = bias_s + nb * P.NB;
      { DBG_T0(); mbar_wait_relaxed(&tmem_full[as], aph); DBG_ADD(dbg_ewait); }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      DBG_T0();
      const uint32_t tbase = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(as * acc_cols);
      // work items = (accumulator, 32-column group); the two warps of a TMEM lane quarter take alternate items
      const int ngrp = (P.NB + 31) >> 5;
      const bool plain = !cond && !P.relu && !P.post_scale && !P.act && !P.yf;   // generator AMP convs: bias, residuals, scale only
      float* yfb = P.yf ? P.yf + (int64_t)b * P.Cout * P.Tout : nullptr;           // fp32 plain output (split convs)
      const float* r1f = P.r1f ? P.r1f + (int64_t)b * P.Cout * P.Tout : nullptr;
      const float* r2f = P.r2f ? P.r2f + (int64_t)b * P.Cout * P.Tout : nullptr;
      const int cs = P.y_tp * 8;                                        // elements between channel chunks
      for (int a = 0; a < nacc; ++a) {
        const int ms = a / P.NPH, s = a - ms * P.NPH;
        const int64_t q = (int64_t)q0 + ms * 128 + r;
        const int64_t t = q * P.u + s - P.p;
        const bool valid = (t >= 0) && (t < Tout_b);
        const int64_t rowoff = ((int64_t)P.y_row0 + (valid ? t : 0)) * 8;
        for (int grp = (half - a * ngrp) & 1; grp < ngrp; grp += 2) {
          // columns [c0, c0+32) of accumulator a (the last group may be 16 wide)
          const int c0 = grp << 5;
          const int ng = (c0 + 16 < P.NB) ? 4 : 2;                  // 8-channel chunks in this group
          const int cobase = nb * P.NB + c0;
          const int nok = max(0, min(ng, (P.y_chunks * 8 - cobase + 7) >> 3));   // chunks that exist in the output tensor
          const int64_t off0 = (int64_t)(cobase >> 3) * cs + rowoff;
          uint4 e1[4], e2[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            e1[g] = make_uint4(0, 0, 0, 0); e2[g] = make_uint4(0, 0, 0, 0);
            if (valid && g < nok) {
              if (r1) e1[g] = *reinterpret_cast<const uint4*>(r1 + off0 + g * cs);
              if (r2) e2[g] = *reinterpret_cast<const uint4*>(r2 + off0 + g * cs);
            }
          }
          uint32_t v[32];
          tmem_ld16_nowait(tbase + (uint32_t)(a * P.NB + c0), *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
          if (ng == 4) tmem_ld16_nowait(tbase + (uint32_t)(a * P.NB + c0 + 16), *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (!valid) continue;
          if (yfb) {
            // fp32 [Cout][T] rows: for a fixed channel the 32 lanes write 32 consecutive time steps
            const float scale = P.scale;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int co = cobase + 8 * g;
              if (g >= ng || co >= P.Cout) continue;
              const float4 b0 = *reinterpret_cast<const float4*>(bs + c0 + 8 * g), b1 = *reinterpret_cast<const float4*>(bs + c0 + 8 * g + 4);
              float f[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              const int nj = min(8, P.Cout - co);                     // (8 for every layer of the generator)
              const int64_t o = (int64_t)co * P.Tout + t;
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] += __uint_as_float(v[8 * g + j]);
              if (cond) {
#pragma unroll
                for (int j = 0; j < 8; ++j) if (j < nj) f[j] += cond[co + j];
              }
              if (P.relu) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
              }
              if (P.post_scale) {                                     // eval-BatchNorm folded to an affine (ECAPA TDNN)
#pragma unroll
                for (int j = 0; j < 8; ++j) if (j < nj) f[j] = fmaf(f[j], P.post_scale[co + j], P.post_shift[co + j]);
              }
              if (P.act == 1) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = tanhf(f[j]);
              }
              if (r1f) {
#pragma unroll
                for (int j = 0; j < 8; ++j) if (j < nj) f[j] += r1f[o + (int64_t)j * P.Tout];
              }
              if (r2f) {
#pragma unroll
                for (int j = 0; j < 8; ++j) if (j < nj) f[j] += r2f[o + (int64_t)j * P.Tout];
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) if (j < nj) yfb[o + (int64_t)j * P.Tout] = f[j] * scale;
              if (P.y && g < nok) {                                   // both forms of the output (speaker encoder: fp32 for the
                uint4 oc;                                             // statistics, c8t bf16 for the next 1x1 GEMM)
                oc.x = pack2(f[0] * scale, f[1] * scale); oc.y = pack2(f[2] * scale, f[3] * scale);
                oc.z = pack2(f[4] * scale, f[5] * scale); oc.w = pack2(f[6] * scale, f[7] * scale);
                *reinterpret_cast<uint4*>(yb + off0 + g * cs) = oc;
              }
            }
            continue;
          }
          if (plain) {
            // straight-line code for the four chunks (predicated stores only): the branches of the general path
            // below cost more than the arithmetic
            const float scale = P.scale;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float4 b0 = *reinterpret_cast<const float4*>(bs + c0 + 8 * g), b1 = *reinterpret_cast<const float4*>(bs + c0 + 8 * g + 4);
              float f[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] += __uint_as_float(v[8 * g + j]);
              float ea[8], eb[8];
              unpack8(e1[g], ea); unpack8(e2[g], eb);               // zeros when there is no residual
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = ((f[j] + ea[j]) + eb[j]) * scale;
              uint4 o;
              o.x = pack2(f[0], f[1]); o.y = pack2(f[2], f[3]); o.z = pack2(f[4], f[5]); o.w = pack2(f[6], f[7]);
              if (g < nok) *reinterpret_cast<uint4*>(yb + off0 + g * cs) = o;
            }
            continue;
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int co = cobase + 8 * g;
            if (g >= nok) continue;                                   // padding channels inside the tensor are written as zeros
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[8 * g + j]) + bs[c0 + 8 * g + j];
            if (cond) {
#pragma unroll
              for (int j = 0; j < 8; ++j) if (co + j < P.Cout) f[j] += cond[co + j];
            }
            if (P.relu) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if (P.post_scale) {                                       // eval-BatchNorm folded to an affine (ECAPA TDNN)
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (co + j < P.Cout) f[j] = fmaf(f[j], P.post_scale[co + j], P.post_shift[co + j]);
            }
            if (P.act == 1) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = tanhf(f[j]);
            }
            if (r1) { float e[8]; unpack8(e1[g], e);
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] += e[j]; }
            if (r2) { float e[8]; unpack8(e2[g], e);
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] += e[j]; }
            uint4 o;
            o.x = pack2(f[0] * P.scale, f[1] * P.scale); o.y = pack2(f[2] * P.scale, f[3] * P.scale);
            o.z = pack2(f[4] * P.scale, f[5] * P.scale); o.w = pack2(f[6] * P.scale, f[7] * P.scale);
            *reinterpret_cast<uint4*>(yb + off0 + g * cs) = o;
          }
        }
      }
      // all of this warp's TMEM reads for the stage are complete: hand it back to the MMA issuer
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      DBG_ADD(dbg_ebusy);
      if (++as == P.acc_stages) { as = 0; aph ^= 1; }
      if (P.zero_pads) {
        // rows [-PAD, 0) by the first tile, [Tout, Tout+PAD) by the last: keeps the c8t zero halo intact
        const int et = threadIdx.x - kEpiWarp0 * 32;
        const int ch0 = (nb * P.NB) >> 3, chn = min(P.NB >> 3, P.y_chunks - ch0);
        const uint4 z = make_uint4(0, 0, 0, 0);
        if (mt == 0)
          for (int i = et; i < chn * P.y_row0; i += kEpiWarps * 32)
            *reinterpret_cast<uint4*>(yb + ((int64_t)(ch0 + i / P.y_row0) * P.y_tp + (i % P.y_row0)) * 8) = z;
        if (mt == (Tout_b - 1) / (P.MT * 128))                          // (zero_pads is only used by plain convs: u = 1)
          for (int i = et; i < chn * P.y_row0; i += kEpiWarps * 32)
            *reinterpret_cast<uint4*>(yb + ((int64_t)(ch0 + i / P.y_row0) * P.y_tp + P.y_row0 + Tout_b + (i % P.y_row0)) * 8) = z;
      }
    }
    if (P.dbg && threadIdx.x == kEpiWarp0 * 32) { P.dbg[blockIdx.x * 8 + 6] = dbg_ewait; P.dbg[blockIdx.x * 8 + 7] = dbg_ebusy; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(P.tmem_cols));
  }
}

// torch-layout fp32 weight -> packed bf16 tiles [n_blk][ci_blk][tap][kchunk][NB][8], zero padded.
// conv: src [Cout][Cin][K]; transposed: src [Cin][Cout][K].
// split_cp > 0: the 3-term split pack (UmmaLayer::split): Cin / K here are the EFFECTIVE sizes (2*Cp channels, 2*K0 taps),
// src holds the real [Cout][Cin0][K0] weight.
__global__ void pack_umma_kernel(__nv_bfloat16* __restrict__ dst, const float* __restrict__ src, int Cout, int Cin,
                                 int K, int transposed, int NB, int n_nblk, int Cin_p, int split_cp = 0, int Cin0 = 0,
                                 int K0 = 0) {
  const int64_t total = (int64_t)n_nblk * Cin_p * NB * K;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // decode i -> (nb, cb, tap, kc, row, e)
    const int64_t per_nb = (int64_t)Cin_p * NB * K;
    const int nb = (int)(i / per_nb);
    int64_t rem = i - (int64_t)nb * per_nb;
    const int cb = (int)(rem / ((int64_t)64 * NB * K));
    rem -= (int64_t)cb * 64 * NB * K;
    const int kcn = min(8, (Cin_p - cb * 64) >> 3);
    const int per_tap = kcn * 8 * NB;
    const int tap = (int)(rem / per_tap);
    rem -= (int64_t)tap * per_tap;
    const int kc = (int)(rem / (NB * 8));
    rem -= (int64_t)kc * NB * 8;
    const int row = (int)(rem >> 3), e = (int)(rem & 7);
    const int co = nb * NB + row, ci = cb * 64 + kc * 8 + e;
    float v = 0.f;
    if (split_cp > 0) {
      const int lo_in = ci >= split_cp, cil = lo_in ? ci - split_cp : ci;   // which half of the input the channel reads
      const int part = tap / K0, k = tap - part * K0;                        // 0: w_hi taps, 1: w_lo taps
      if (co < Cout && cil < Cin0 && part < 2) {
        const float w = transposed ? src[((int64_t)cil * Cout + co) * K0 + k] : src[((int64_t)co * Cin0 + cil) * K0 + k];
        const float hi = __bfloat162float(__float2bfloat16_rn(w));
        v = part == 0 ? hi : (lo_in ? 0.f : w - hi);
      }
    } else if (co < Cout && ci < Cin && tap < K) {
      v = transposed ? src[((int64_t)ci * Cout + co) * K + tap] : src[((int64_t)co * Cin + ci) * K + tap];
    }
    dst[i] = __float2bfloat16_rn(v);
  }
}

// plain [B,C,T] (fp32 or bf16) <-> c8t bf16
template <typename TS>
__global__ void to_c8t_kernel(__nv_bfloat16* __restrict__ dst, const TS* __restrict__ src, int64_t sb, int64_t sc,
                              int64_t st_, int C, int chunks, int T, int Tp, int pad, const int* __restrict__ lens, int len_mul,
                              int reflect) {
  // one thread per (b, chunk, row) 16-byte vector, pads and padding channels zeroed
  const int64_t n = (int64_t)gridDim.y * chunks * Tp;
  (void)n;
  const int b = blockIdx.y;
  if (lens) T = lens[b] * len_mul;                                  // ragged batch: rows past the utterance's end are zero
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)chunks * Tp; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i / Tp), row = (int)(i % Tp);
    int t = row - pad;
    // `reflect` halo rows on each side mirror the signal (speechbrain's "same" reflect padding, nnet/CNN.py:458-488), so
    // that the zero-padding tcgen05 conv computes the reflect-padded one; all other halo rows are zero
    if (t < 0 && t >= -reflect) t = -t;
    else if (t >= T && t < T + reflect) t = 2 * (T - 1) - t;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = ch * 8 + j;
      f[j] = (t >= 0 && t < T && c < C) ? to_f<TS>(src[b * sb + c * sc + t * st_]) : 0.f;
    }
    uint4 o;
    o.x = pack2(f[0], f[1]); o.y = pack2(f[2], f[3]); o.z = pack2(f[4], f[5]); o.w = pack2(f[6], f[7]);
    *reinterpret_cast<uint4*>(dst + ((int64_t)b * chunks * Tp + i) * 8) = o;
  }
}
// fp32 -> [hi | lo] bf16 halves in c8t (UmmaLayer::split): chunks [0, Cp/8) hold bf16(x), chunks [Cp/8, 2Cp/8) hold
// bf16(x - bf16(x)); halo rows and padding channels are zero
__global__ void split_to_c8t_kernel(__nv_bfloat16* __restrict__ dst, const float* __restrict__ src, int64_t sb, int64_t sc,
                                    int64_t st_, int C, int Cp8, int chunks, int T, int Tp, int pad) {
  const int b = blockIdx.y;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)chunks * Tp; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i / Tp), row = (int)(i % Tp);
    const int t = row - pad;
    const bool lo = ch >= Cp8;
    const int c0 = (lo ? ch - Cp8 : ch) * 8;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      const float x = (t >= 0 && t < T && c < C && ch < 2 * Cp8) ? src[b * sb + c * sc + t * st_] : 0.f;
      f[j] = lo ? x - __bfloat162float(__float2bfloat16_rn(x)) : x;
    }
    uint4 o;
    o.x = pack2(f[0], f[1]); o.y = pack2(f[2], f[3]); o.z = pack2(f[4], f[5]); o.w = pack2(f[6], f[7]);
    *reinterpret_cast<uint4*>(dst + ((int64_t)b * chunks * Tp + i) * 8) = o;
  }
}
template <typename TD>
__global__ void from_c8t_kernel(TD* __restrict__ dst, const __nv_bfloat16* __restrict__ src, int C, int chunks, int T,
                                int Tp, int pad) {
  const int b = blockIdx.y;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)C * T; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i / T), t = (int)(i % T);
    dst[(int64_t)b * C * T + i] =
        from_f<TD>(__bfloat162float(src[(((int64_t)b * chunks + (c >> 3)) * Tp + pad + t) * 8 + (c & 7)]));
  }
}

}  // namespace

static size_t umma_fixed_smem(int NB, int n_nblk) {
  return (size_t)(2 * kMaxXStages + 2 * kMaxWStages + 4) * 8 + 16 + (size_t)NB * n_nblk * 4 + 128 + 128;   // (+32 floats: bias over-read)
}

void umma_choose_nb(int Cout, int nph, int* NB, int* n_nblk) {
  // NPH * NB (x MT) must fit 512 TMEM columns.  For 256 < Cout <= 512 (stage 1, C = 384) three 128-column blocks beat
  // two 192-column ones: two accumulator stages fit, so the epilogue of tile i overlaps the MMAs of tile i+1 (measured
  // 1026-1485 against 963-1396 TFLOP/s); for C = 768 and C = 192 the extra passes over x cancel that gain.
  // BVG_UMMA_NB_CAP overrides (experiments).
  const int cap_env = BVG_ENV_ONCE("BVG_UMMA_NB_CAP", 0);
  const int cap = nph >= 4 ? 128 : cap_env > 0 ? cap_env : (Cout > 256 && Cout <= 512) ? 128 : 256;
  const int n = (Cout + cap - 1) / cap;
  int nb = ((Cout + n - 1) / n + 15) / 16 * 16;
  *NB = nb;
  *n_nblk = (Cout + nb - 1) / nb;
}

int64_t umma_pack_elems(int Cout, int Cin, int K, int nph) {
  int NB, nn;
  umma_choose_nb(Cout, nph, &NB, &nn);
  const int Cin_p = (Cin + 15) / 16 * 16;
  return (int64_t)nn * Cin_p * NB * K;
}

int umma_pack_launch(__nv_bfloat16* dst, const float* src_torch_layout, int Cout, int Cin, int K, int transposed,
                     int nph, cudaStream_t st) {
  int NB, nn;
  umma_choose_nb(Cout, nph, &NB, &nn);
  const int Cin_p = (Cin + 15) / 16 * 16;
  const int64_t total = (int64_t)nn * Cin_p * NB * K;
  int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
  pack_umma_kernel<<<blocks, 256, 0, st>>>(dst, src_torch_layout, Cout, Cin, K, transposed, NB, nn, Cin_p);
  BVG_LAUNCHED();
  return BVG_OK;
}

int64_t umma_pack_split_elems(int Cout, int Cin, int K, int nph) {
  return umma_pack_elems(Cout, 2 * ((Cin + 7) / 8 * 8), 2 * K, nph);
}

int umma_pack_split_launch(__nv_bfloat16* dst, const float* src_torch_layout, int Cout, int Cin, int K, int transposed,
                           int nph, cudaStream_t st) {
  int NB, nn;
  umma_choose_nb(Cout, nph, &NB, &nn);
  const int Cp = (Cin + 7) / 8 * 8, Ce = 2 * Cp, Ke = 2 * K;
  const int Cin_p = (Ce + 15) / 16 * 16;
  const int64_t total = (int64_t)nn * Cin_p * NB * Ke;
  int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
  pack_umma_kernel<<<blocks, 256, 0, st>>>(dst, src_torch_layout, Cout, Ce, Ke, transposed, NB, nn, Cin_p, Cp, Cin, K);
  BVG_LAUNCHED();
  return BVG_OK;
}

int split_to_c8t_launch(const C8T& dst, const float* src, int64_t sb, int64_t sc, int64_t st_, int C, int64_t B,
                        cudaStream_t st) {
  if (B == 0) return BVG_OK;
  const int Cp = (C + 7) / 8 * 8;
  BVG_CHECK_ARG(dst.C == 2 * Cp, "split_to_c8t: destination must have 2*roundup8(C) channels");
  dim3 grid((unsigned)std::min<int64_t>(((int64_t)dst.chunks * dst.Tp + 255) / 256, 4096), (unsigned)B);
  ProfScope prof(st, KC_OTHER);
  split_to_c8t_kernel<<<grid, 256, 0, st>>>(dst.p, src, sb, sc, st_, C, Cp / 8, dst.chunks, dst.T, dst.Tp, dst.pad);
  BVG_LAUNCHED();
  return BVG_OK;
}

int conv_umma_launch(const UmmaLayer& L, const C8T& x, const C8T& y, const UmmaEpilogue& ep, int64_t B,
                     cudaStream_t st) {
  BVG_CHECK_ARG(L.w && x.p && (y.p || ep.yf32), "conv_umma: null pointer");
  // split convs (UmmaLayer::split) see 2*roundup8(Cin) input channels and 2*K taps
  const int Cin_eff = L.split ? 2 * ((L.Cin + 7) / 8 * 8) : L.Cin;
  const int K_eff = L.split ? 2 * L.K : L.K;
  BVG_CHECK_ARG(!L.split || ep.yf32, "conv_umma: split weights need the fp32 output epilogue");
  BVG_CHECK_ARG(!ep.yf32 || (!ep.res1 && !ep.res2 && !ep.zero_pads && ep.act <= 1),
                "conv_umma: the fp32 output epilogue takes bias / cond / ReLU / affine / tanh / fp32 residuals / scale only");
  BVG_CHECK_ARG(x.C == Cin_eff && y.C == L.Cout, "conv_umma: channel mismatch (x.C=%d Cin=%d y.C=%d Cout=%d)", x.C, Cin_eff, y.C, L.Cout);
  UmmaConvParams P;
  memset(&P, 0, sizeof P);
  const int u = L.transposed ? L.stride : 1;
  const int J = L.transposed ? L.K / u : 1;
  P.NPH = L.transposed ? u : 1;
  int n_nblk = 1;
  umma_choose_nb(L.Cout, P.NPH, &P.NB, &n_nblk);
  P.ntaps = K_eff;
  P.tap_mod = L.split ? L.K : K_eff;
  P.split_hi_chunks = L.split ? (L.Cin + 7) / 8 : 0;
  BVG_CHECK_ARG(L.K <= 16 && K_eff <= 32, "conv_umma: at most 16 taps (32 with split weights)");
  int num_sms = 0;
  BVG_TRY(current_device_sms(&num_sms));
  // time sub-tiles per CTA: every weight tile is shared by MT*128 output rows; halved while the grid would leave SMs
  // idle (single-utterance latency: stage 0 of a 10 s utterance has 4 x 3 tiles at MT = 2)
  P.MT = L.transposed ? 1 : (P.NB <= 64 ? 4 : 2);
  while (P.MT > 1 && (int64_t)((y.T + P.MT * 128 - 1) / (P.MT * 128)) * B * n_nblk < num_sms) P.MT >>= 1;
  int halo;
  if (!L.transposed) {
    BVG_CHECK_ARG(y.T == x.T, "conv_umma: conv keeps the length");
    const int pad = L.dil * (L.K - 1) / 2;
    BVG_CHECK_ARG(pad <= x.pad, "conv_umma: conv padding %d exceeds the c8t halo %d", pad, x.pad);
    for (int k = 0; k < L.K; ++k) { P.tap_shift[k] = (int16_t)(k * L.dil); P.tap_acc[k] = 0; }
    P.lo = pad;
    halo = L.dil * (L.K - 1);
    P.u = 1; P.p = 0;
    P.Tout = y.T;
    P.tiles_per_batch = (y.T + P.MT * 128 - 1) / (P.MT * 128);
  } else {
    BVG_CHECK_ARG(L.K % u == 0 && (L.K - u) % 2 == 0 && J <= 2, "conv_umma: unsupported ConvTranspose1d K=%d stride=%d", L.K, u);
    BVG_CHECK_ARG(y.T == x.T * u, "conv_umma: ConvTranspose1d output length");
    BVG_CHECK_ARG(K_eff <= 16, "conv_umma: ConvTranspose1d tap tables hold 16 entries");
    for (int t = 0; t < K_eff; ++t) { const int k = t % L.K; P.tap_shift[t] = (int16_t)((J - 1) - k / u); P.tap_acc[t] = (int16_t)(k % u); }
    P.lo = J - 1;
    halo = J - 1;
    P.u = u; P.p = (L.K - u) / 2;
    P.Tout = y.T;
    const int qn = x.T + (P.p + u - 1) / u;                     // coarse rows that produce a valid output
    P.tiles_per_batch = (qn + 127) / 128;
    BVG_CHECK_ARG(x.pad >= 2, "conv_umma: c8t halo too small");
  }
  P.XR = P.MT * 128 + halo;
  P.Cin_p = (Cin_eff + 15) / 16 * 16;
  BVG_CHECK_ARG(x.chunks * 8 >= P.Cin_p, "conv_umma: input tensor must carry channel padding to a multiple of 16");
  P.n_ci_blk = (P.Cin_p + 63) / 64;
  P.Cout = L.Cout;
  P.x = x.p; P.x_bstride = x.batch_stride(); P.x_tp = x.Tp; P.x_row0 = x.pad;
  P.y = y.p; P.y_bstride = y.batch_stride(); P.y_tp = y.Tp; P.y_row0 = y.pad; P.y_chunks = y.chunks;
  P.lens = y.lens; P.len_mul = y.len_mul;
  BVG_CHECK_ARG(!y.lens || !ep.yf32, "conv_umma: ragged batches are a bf16 c8t feature");
  P.w = L.w;
  P.bias = ep.bias; P.cond = ep.cond; P.cond_B = (int)ep.cond_B; P.scale = ep.scale;
  P.res1 = ep.res1; P.res2 = ep.res2; P.zero_pads = ep.zero_pads;
  P.yf = ep.yf32; P.r1f = ep.res1_f32; P.r2f = ep.res2_f32;
  P.dbg = ep.dbg;
  P.dry = ep.dry;
  P.relu = ep.relu; P.post_scale = ep.post_scale; P.post_shift = ep.post_shift; P.act = ep.act;
  P.transposed = L.transposed;
  P.dil = L.dil;
  const int acc_cols = P.MT * P.NPH * P.NB;
  P.acc_stages = (2 * acc_cols <= 512) ? 2 : 1;
  int cols = acc_cols * P.acc_stages, pw = 32;
  while (pw < cols) pw <<= 1;
  BVG_CHECK_ARG(pw <= 512, "conv_umma: accumulators exceed TMEM");
  P.tmem_cols = pw;
  P.n_nblk = n_nblk;
  P.B = (int)B;
  P.n_issuers = 1;
  while (P.n_issuers * 2 <= std::min(kMaxIssuers, P.MT * P.NPH)) P.n_issuers *= 2;   // power of two
  if (const int e = BVG_ENV_ONCE("BVG_UMMA_ISSUERS", 0)) P.n_issuers = std::max(1, std::min(P.n_issuers, e));
  // shared-memory plan: keep the whole layer's weights resident when they fit next to >= 2 input stages,
  // otherwise stream them through as deep a ring as fits
  P.kc_max = std::min(8, P.Cin_p / 8);
  const size_t budget = 227 * 1024 - umma_fixed_smem(P.NB, n_nblk);
  const size_t xsb = (size_t)P.XR * P.kc_max * 16, wsb = (size_t)P.NB * P.kc_max * 16;
  const int wslots = P.ntaps * P.n_ci_blk;
  BVG_CHECK_ARG(2 * xsb + 2 * wsb <= budget, "conv_umma: tile does not fit shared memory");
  if (n_nblk == 1 && wslots <= kMaxWStages && wslots * wsb + 2 * xsb <= budget) {
    P.w_resident = 1;
    P.w_stages = wslots;
    P.x_stages = (int)std::min<size_t>(kMaxXStages, (budget - wslots * wsb) / xsb);
  } else {
    P.w_resident = 0;
    P.x_stages = 2;
    P.w_stages = (int)std::min<size_t>(kMaxWStages, (budget - 2 * xsb) / wsb);
  }
  const size_t smem = (size_t)P.x_stages * xsb + (size_t)P.w_stages * wsb + umma_fixed_smem(P.NB, n_nblk);
  static std::atomic<uint64_t> opted{0};
  BVG_TRY(smem_opt_in(conv_umma_kernel, opted, 227 * 1024));
  const int64_t ntiles = (int64_t)P.tiles_per_batch * B * n_nblk;
  BVG_CHECK_ARG(ntiles < (1ll << 31), "conv_umma: too many tiles");
  dim3 grid((unsigned)std::min<int64_t>(ntiles, num_sms));
  ProfScope prof(st, ep.prof_other ? KC_OTHER : (L.transposed ? KC_CONVTR : KC_CONV));
  conv_umma_kernel<<<grid, kThreads, smem, st>>>(P);
  BVG_LAUNCHED();
  return BVG_OK;
}


int to_c8t_launch(const C8T& dst, const void* src, int64_t sb, int64_t sc, int64_t st_, int src_dtype, int64_t B,
                  cudaStream_t st, int reflect) {
  if (B == 0) return BVG_OK;
  BVG_CHECK_ARG(reflect >= 0 && reflect <= dst.pad && reflect < dst.T && !(reflect && dst.lens), "to_c8t: bad reflect halo");
  dim3 grid((unsigned)std::min<int64_t>(((int64_t)dst.chunks * dst.Tp + 255) / 256, 4096), (unsigned)B);
  ProfScope prof(st, KC_OTHER);
  if (src_dtype == BVG_F32)
    to_c8t_kernel<float><<<grid, 256, 0, st>>>(dst.p, (const float*)src, sb, sc, st_, dst.C, dst.chunks, dst.T, dst.Tp, dst.pad, dst.lens, dst.len_mul, reflect);
  else if (src_dtype == BVG_BF16)
    to_c8t_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(dst.p, (const __nv_bfloat16*)src, sb, sc, st_, dst.C, dst.chunks, dst.T, dst.Tp, dst.pad, dst.lens, dst.len_mul, reflect);
  else if (src_dtype == BVG_F16)
    to_c8t_kernel<__half><<<grid, 256, 0, st>>>(dst.p, (const __half*)src, sb, sc, st_, dst.C, dst.chunks, dst.T, dst.Tp, dst.pad, dst.lens, dst.len_mul, reflect);
  else { set_error("to_c8t: unsupported dtype"); return BVG_ERR_INVALID; }
  BVG_LAUNCHED();
  return BVG_OK;
}

int from_c8t_launch(void* dst, const C8T& src, int dst_dtype, int64_t B, cudaStream_t st) {
  if (B == 0) return BVG_OK;
  dim3 grid((unsigned)std::min<int64_t>(((int64_t)src.C * src.T + 255) / 256, 4096), (unsigned)B);
  ProfScope prof(st, KC_OTHER);
  if (dst_dtype == BVG_F32)
    from_c8t_kernel<float><<<grid, 256, 0, st>>>((float*)dst, src.p, src.C, src.chunks, src.T, src.Tp, src.pad);
  else if (dst_dtype == BVG_BF16)
    from_c8t_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((__nv_bfloat16*)dst, src.p, src.C, src.chunks, src.T, src.Tp, src.pad);
  else { set_error("from_c8t: unsupported dtype"); return BVG_ERR_INVALID; }
  BVG_LAUNCHED();
  return BVG_OK;
}

}  // namespace bvg
