// Activation1d -> Conv1d in ONE kernel with BOTH anti-alias FIRs on the tensor cores (narrow generator stages, C = 96 / 48 / 24).
//
// Reference op chain: AMPBlock1.forward, xt = c1(a1(x)); xt = c2(a2(xt)); x = xt + x  (BigVGAN/models.py:65-74), with
// a = Activation1d = UpSample1d (alias_free_torch/resample.py:25-33) -> SnakeBeta (activations.py:109-122) -> DownSample1d
// (resample.py:46-49, filter.py:87-96).  Round 1 fused the pair with the 12 + 12-tap FIRs on the FP32 pipe
// (conv_umma_fused.cu: FMA-pipe bound at 0.29 of the HBM roofline); act1d_tc.cu moved the FIRs of the standalone op onto
// tcgen05 as banded-Toeplitz products over time.  This kernel is that FIR pipeline feeding the conv without a trip through
// HBM: the activated tensor only ever exists as TMEM accumulators and as the conv's K-major A operand in shared memory.
//
// Per CTA (persistent, one per SM; 32 warps) a time RANGE of one utterance streams through in 32-step blocks:
//   TMA      raw rows x[t0 - 8, t0 + 72) of every 8-channel chunk -> x ring (MN-major A operand of the up-FIR)
//   up-FIR   U^T[c, m] = sum_t X^T[c, t] Gup[t, m]     SS MMA, taps bf16 hi + lo, N = 64 upsampled samples   (TMEM, fp32)
//   snake    a = u + 1/(e^b + 1e-9) sin^2(e^a u)       12 warps, tcgen05.ld -> MUFU -> fp16 pairs -> tcgen05.st (TMEM ring)
//   down-FIR Y^T[c, t] = sum_m A^T[c, m] Gdn[m, t]     TS MMA (A from TMEM), taps fp16, N = 32                 (TMEM, fp32)
//   store    Y^T -> bf16 -> stmatrix.trans -> conv A-operand stage [chunk][row][8 ch] (row = time, incl. the conv halo)
//   conv     D[t, co] += A[t + tap * dil, ci] W[co, ci, tap]  SS MMA per (64-channel block, tap, 16-channel step), M = 128 rows
//   epilogue D -> + bias (+ res1) (+ res2), * scale -> bf16 -> c8t rows in HBM (8 warps)
// TMEM lanes = channels for the FIRs (12 chunk-slots = 96 lanes: 1 x 96, 2 x 48 or 4 x 24 channels -- narrow tensors run
// 2 / 4 time SEGMENTS of the range side by side), TMEM lanes = time rows for the conv accumulators.
// A conv tile is 128 output rows = 4 blocks; its A stage holds rows [m0 - LH, m0 + 128 + LH) (LH = conv padding rounded
// up to 8 / 16 / 32), so the last rows of block 4n-1 and the first rows of block 4n+4 are stored twice (tile n and its
// neighbour) instead of being recomputed.
//
// Edges.  Replicate padding of the INPUT (resample.py:28): rows t < 0 / t >= T of a staged x tile are overwritten in
// shared memory by the patch warp.  Replicate padding of the ACTIVATED signal (filter.py:90-92) only changes
// a(x)[0..2] and a(x)[T-3..T-1]: actconv_edge_kernel computes those 6 rows per utterance exactly (CUDA-core stencil of
// act1d_core.cuh) and the patch warp drops them into the A stage, where it also zeroes the conv's zero-padding rows.
// Stages without such rows never involve the patch warp (the issuers wait on the producers' barriers directly).
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "act1d_core.cuh"
#include "bvg_common.cuh"
#include "umma.cuh"
#include "umma_ptx.cuh"

namespace bvg {
namespace {

// The per-role cycle counters (bvg_debug_set_umma_counters, tools/actconv_tc_roles.py) are compiled in only in debug builds
// (BVG_DEBUG_BUILD=1): the kernel is instruction-issue bound and the `P.dbg ? clock64() : 0` pairs around every wait cost
// ~15 instructions per block and role.
#ifndef BVG_DEBUG
#undef DBG_T0
#undef DBG_ADD
#define DBG_T0() do { } while (0)
#define DBG_ADD(var) do { } while (0)
#define TCF_COUNTERS 0
#else
#define TCF_COUNTERS 1
#endif

constexpr int kBlk = 32;                         // output time steps per FIR block
constexpr int kXB = 2;                           // FIR blocks per staged input tile
constexpr int kXR = kXB * kBlk + 16;             // rows per chunk of a staged input tile (8-row FIR halo each side)
constexpr int kSlots = 12;                       // 8-channel chunk-slots = 96 TMEM lanes (the up-FIR's M = 128 reads 4 phantom slots)
constexpr uint32_t kXStageBytes = (uint32_t)kSlots * kXR * 16u;
constexpr int kMaxXS = 4, kMaxAS = 3, kMaxWS = 8;
constexpr int kMaxNU = 4, kMaxNY = 4;            // U / Y accumulator buffers (P.nu / P.ny of them)
constexpr int kThreads = 1024;
// warp roles: a warp reaches TMEM lanes 32 (w % 4) .. +31 only, so the 12 snake / 6 store warps are the w % 4 < 3 ones
constexpr int kWProdX = 3, kWProdW = 7, kWUp = 11, kWDn = 15, kWConv = 19, kWPatch = 23;
constexpr int kWStore0 = 16, kWEpi0 = 24;
// TMEM columns (512): U nu x 64 | A ring 4 x 32 | Y ny x 32 | conv accumulators nacc x S x NB  (P.colA / colY / colC)
constexpr uint32_t kColU = 0;
constexpr uint32_t kAHi = 0x4000u | (uint32_t)kXR;                     // up-FIR A descriptor: SBO = chunk pitch, version bit
constexpr uint32_t kUpBytes = 6u * 64u * 16u, kDnBytes = 12u * 32u * 16u;
constexpr int kDumpBytes = 512;

struct ActConvTcParams {
  const __nv_bfloat16* x; int64_t x_bstride; int x_tp, x_pad;
  __nv_bfloat16* y; int64_t y_bstride; int y_tp, y_pad, y_chunks;
  const __nv_bfloat16* w; const __nv_bfloat16* res1; const __nv_bfloat16* res2;
  const float* bias; float scale;
  const float* alpha; const float* beta;
  const __nv_bfloat16* edge;          // [B][2 sides][3 rows][rc][8]: exact a(x)[0..2], a(x)[T-3..T-1]
  int C, rc, S, Cin_p, NB, Cout, K, dil, lo, LH, SR, NP;
  int T, RL, NG, nitems;
  const int* lens; int len_mul;       // ragged batch: utterance b has lens[b] * len_mul rows (else T)
  int nxs, nas, nacc, w_resident, w_slots;
  int nu, ny; uint32_t colA, colY, colC;
  int pace;                           // conv MMAs per pacing group (two groups in flight at most)
  uint32_t w_slot_bytes, w_total_bytes, a_stage_bytes;
  int zero_pads;
  long long* dbg;
  long long* trace;
  int dry;                            // BVG_DEBUG builds only (BVG_TCF_DRY): skip parts of the pipeline, results are garbage
};
#ifdef BVG_DEBUG
#define TCF_DRY(bit) (P.dry & (bit))
// event trace of CTA 0 (BVG_TCF_TRACE=1, debug builds): clock64 of pipeline events of blocks / tiles 128..191 at dbg[148 * 16 ...]
#define TCF_TRACE(ev, idx)                                                                                        \
  do {                                                                                                            \
    if (P.trace && blockIdx.x == 0 && (unsigned)((int)(idx) - 128) < 64u && lane == 0)                            \
      P.trace[(ev) * 64 + (int)(idx) - 128] = clock64();                                                          \
  } while (0)
#else
#define TCF_DRY(bit) false
#define TCF_TRACE(ev, idx) do { } while (0)
#endif

struct TcItem { int b, grp, nblk, T; };
// WARP: called by a fully converged warp (rows_of shuffles, which also marks the length warp-uniform for ptxas); false for the
// single-lane weight producer
template <bool WARP = true>
__device__ __forceinline__ TcItem tc_item(const ActConvTcParams& P, int item) {
  TcItem it;
  it.b = item / P.NG;
  it.grp = item - it.b * P.NG;
  const int r0 = it.grp * P.S * P.RL;
  it.T = WARP ? rows_of(P.lens, P.len_mul, it.b, P.T) : (P.lens ? __ldg(P.lens + it.b) * P.len_mul : P.T);
  it.nblk = 4 * max(0, (min(it.T, r0 + P.RL) - r0 + 127) >> 7);       // 0: the whole item lies past the utterance's end
  return it;
}
__device__ __forceinline__ int seg_r0(const ActConvTcParams& P, const TcItem& it, int s) { return (it.grp * P.S + s) * P.RL; }
// first row of staged input tile `st` of a segment starting at r0 (up-blocks i = -2 .. nblk + 1, kXB per tile)
__device__ __forceinline__ int xstage_t0(int r0, int st) { return r0 + kBlk * (st * kXB - 2) - 8; }
// does a staged input tile hold rows outside [0, T) (replicate padding by the patch warp)?
__device__ __forceinline__ bool x_edge(const ActConvTcParams& P, const TcItem& it, int st) {
  bool e = false;
  for (int s = 0; s < P.S; ++s) {
    const int r0 = seg_r0(P, it, s);
    if (r0 >= it.T) break;
    const int ts = xstage_t0(r0, st);
    e = e || ts < 0 || ts + kXR > it.T;
  }
  return e;
}
// does conv tile n need the patch warp (exact edge rows of the activation, zero rows of the conv padding)?
__device__ __forceinline__ bool tile_edge(const ActConvTcParams& P, const TcItem& it, int n) {
  bool e = false;
  for (int s = 0; s < P.S; ++s) {
    const int r0 = seg_r0(P, it, s);
    if (r0 >= it.T) break;
    const int m0 = r0 + 128 * n;
    e = e || m0 == 0 || m0 + 128 + P.LH > it.T - 3;
  }
  return e;
}

__device__ __forceinline__ uint32_t cvt_f16x2_sat(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done != 0;
}
// try_wait with a suspend-time hint: the waiting warp is parked by the hardware until the phase completes (or the hint, ~10 ms,
// expires) instead of returning after the short default limit -- 20+ warps of this kernel wait on mbarriers most of the time,
// and polling at the default limit floods the shared-memory pipeline that tcgen05.ld / stmatrix / the arrives go through.
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
  return done != 0;
}
// Waits of this kernel.  Debug builds (BVG_DEBUG) add a watchdog: a wait that does not complete within ~1 s records
// (block, warp, wait id) in g_tcf_abort and releases every other wait of the grid, so a protocol bug shows up as a report
// instead of a hung GPU.
#ifdef BVG_DEBUG
__device__ int g_tcf_abort[8];
template <int SLEEP>
__device__ __forceinline__ void tc_wait(uint64_t* bar, uint32_t parity, int id) {
  uint32_t spins = 0;
  long long t0 = 0;
  for (;;) {
    if (mbar_try_wait_hint(bar, parity)) return;
    if ((++spins & 3u) == 0) {
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 3000000000ll || *reinterpret_cast<volatile int*>(&g_tcf_abort[0])) {
        if (atomicCAS(&g_tcf_abort[0], 0, 1) == 0) {
          g_tcf_abort[1] = blockIdx.x; g_tcf_abort[2] = threadIdx.x >> 5; g_tcf_abort[3] = id; g_tcf_abort[4] = (int)parity;
        }
        return;
      }
    }
  }
}
#define TC_ABORTED() (*reinterpret_cast<volatile int*>(&g_tcf_abort[0]) != 0)
#else
template <int SLEEP>
__device__ __forceinline__ void tc_wait(uint64_t* bar, uint32_t parity, int) {
  while (!mbar_try_wait_hint(bar, parity)) { }
}
#define TC_ABORTED() false
#endif
// ring cursor: slot index + phase bit, advanced once per use
struct Ring {
  int s = 0; uint32_t ph = 0;
  __device__ __forceinline__ void next(int n) { if (++s == n) { s = 0; ph ^= 1u; } }
};

// Conv epilogue of one tile for one warp: `cnt` consecutive 16-column slices of the tile's S accumulators starting at segment
// sg0, slice sl0 (flat slice f <-> TMEM columns 16 f .. 16 f + 15), rows = this warp's 32 TMEM lanes: + bias (+ res1) (+ res2),
// * scale -> bf16 -> 16-byte stores.  Residual vectors are fetched one slice ahead.  The SM is issue-bound and the uniform
// datapath slow (address arithmetic on warp-uniform values costs ~10 cycles per dependent instruction), so everything that
// can be hoisted is passed in: per-tile state is three integers, per-slice state advances by constant increments.
__device__ __forceinline__ f32x2 bf16x2_to_f32x2(uint32_t w) { return pk2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }
struct EpiConst {
  int sg0, sl0, cnt, nsl;            // this warp's slices
  int T, RL, row_base;               // row_base = y_pad (element row of t = 0)
  uint32_t cs2;                      // elements between channel-chunk PAIRS (2 * y_tp * 8)
  uint32_t cs;
  float scale; bool do_scale;
};
template <bool HAS_R1, bool HAS_R2>
__device__ __forceinline__ void epilogue_tile(const EpiConst& E, int seg_t0, int t_in_seg, __nv_bfloat16* yb, const __nv_bfloat16* r1,
                                              const __nv_bfloat16* r2, const float* bias_s, uint32_t tbase) {
  // seg_t0: first row of segment sg0 of the item; t_in_seg: this lane's row within the segment (128 n + r)
  const f32x2 scale = pk2(E.scale, E.scale);
  int sl = E.sl0, r0 = seg_t0;
  uint32_t tcol = tbase + (uint32_t)((E.sg0 * E.nsl + E.sl0) * 16);
  const float* bp = bias_s + E.sl0 * 16;
  auto offset = [&](int r0_, int sl_) -> uint32_t {                 // element offset of (row, chunk 2 sl), ~0 if not an output row
    const int t = r0_ + t_in_seg;
    return (r0_ < E.T && t < E.T) ? (uint32_t)(E.row_base + t) * 8u + (uint32_t)sl_ * E.cs2 : ~0u;
  };
  uint4 c1[2], c2[2], n1[2], n2[2];
  auto load_res = [&](uint32_t off, uint4 (&e1)[2], uint4 (&e2)[2]) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      if (HAS_R1) e1[g] = off != ~0u ? *reinterpret_cast<const uint4*>(r1 + off + g * E.cs) : make_uint4(0, 0, 0, 0);
      if (HAS_R2) e2[g] = off != ~0u ? *reinterpret_cast<const uint4*>(r2 + off + g * E.cs) : make_uint4(0, 0, 0, 0);
    }
  };
  uint32_t coff = offset(r0, sl);
  load_res(coff, c1, c2);
  for (int i = 0; i < E.cnt; ++i) {
    uint32_t v[16];
    tmem_ld16_nowait(tcol, v);
    const float4* bs = reinterpret_cast<const float4*>(bp);
    // next slice
    tcol += 16u; bp += 16;
    if (++sl == E.nsl) { sl = 0; r0 += E.RL; bp = bias_s; }
    const uint32_t noff = i + 1 < E.cnt ? offset(r0, sl) : ~0u;
    load_res(noff, n1, n2);
    tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const float4 b0 = bs[2 * g], b1 = bs[2 * g + 1];
      f32x2 a0 = add2(pk2(__uint_as_float(v[8 * g + 0]), __uint_as_float(v[8 * g + 1])), pk2(b0.x, b0.y));
      f32x2 a1 = add2(pk2(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3])), pk2(b0.z, b0.w));
      f32x2 a2 = add2(pk2(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5])), pk2(b1.x, b1.y));
      f32x2 a3 = add2(pk2(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7])), pk2(b1.z, b1.w));
      if (HAS_R1) {
        a0 = add2(a0, bf16x2_to_f32x2(c1[g].x)); a1 = add2(a1, bf16x2_to_f32x2(c1[g].y));
        a2 = add2(a2, bf16x2_to_f32x2(c1[g].z)); a3 = add2(a3, bf16x2_to_f32x2(c1[g].w));
      }
      if (HAS_R2) {
        a0 = add2(a0, bf16x2_to_f32x2(c2[g].x)); a1 = add2(a1, bf16x2_to_f32x2(c2[g].y));
        a2 = add2(a2, bf16x2_to_f32x2(c2[g].z)); a3 = add2(a3, bf16x2_to_f32x2(c2[g].w));
      }
      if (E.do_scale) { a0 = mul2(a0, scale); a1 = mul2(a1, scale); a2 = mul2(a2, scale); a3 = mul2(a3, scale); }
      float x0, x1, x2, x3, x4, x5, x6, x7;
      unpk2(a0, x0, x1); unpk2(a1, x2, x3); unpk2(a2, x4, x5); unpk2(a3, x6, x7);
      uint4 o;
      o.x = cvt_bf16x2(x0, x1); o.y = cvt_bf16x2(x2, x3); o.z = cvt_bf16x2(x4, x5); o.w = cvt_bf16x2(x6, x7);
      if (coff != ~0u) *reinterpret_cast<uint4*>(yb + coff + g * E.cs) = o;
    }
#pragma unroll
    for (int g = 0; g < 2; ++g) { if (HAS_R1) c1[g] = n1[g]; if (HAS_R2) c2[g] = n2[g]; }
    coff = noff;
  }
}

__global__ void __launch_bounds__(kThreads, 1) actconv_tc_kernel(const ActConvTcParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  uint8_t* xsm = smem;
  uint8_t* up_hi = xsm + P.nxs * kXStageBytes;
  uint8_t* up_lo = up_hi + kUpBytes;
  uint8_t* dnm = up_lo + kUpBytes;
  uint8_t* asm_ = dnm + kDnBytes;
  uint8_t* wsm = asm_ + (size_t)P.nas * P.a_stage_bytes;
  uint8_t* dump = wsm + (P.w_resident ? P.w_total_bytes : (uint32_t)P.w_slots * P.w_slot_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(dump + kDumpBytes);
  // A waiter may only ever be ONE phase away from its barrier (parity waits alias beyond that), and the patch warp skips
  // every tile that needs no patching: tiles that do complete on their own barriers (x_efull / as_edone), whose phases
  // count edge tiles only, and every waiter keeps one parity bit per slot and barrier.
  uint64_t* x_full = bars;                       // [4] TMA bytes landed (tiles inside [0, T))
  uint64_t* x_efull = x_full + kMaxXS;           // [4] TMA bytes landed (tiles the patch warp completes)
  uint64_t* x_ready = x_efull + kMaxXS;          // [4] patch warp done (edge tiles only)
  uint64_t* x_free = x_ready + kMaxXS;           // [4]
  uint64_t* u_full = x_free + kMaxXS;            // [2]
  uint64_t* u_free = u_full + kMaxNU;            // [4]
  uint64_t* a_full = u_free + kMaxNU;            // [4]
  uint64_t* a_free = a_full + 4;                 // [4]
  uint64_t* y_full = a_free + 4;                 // [2]
  uint64_t* y_free = y_full + kMaxNY;            // [4]
  uint64_t* as_done = y_free + kMaxNY;              // [3] store warps wrote the whole A stage (no patching needed)
  uint64_t* as_edone = as_done + kMaxAS;         // [3] the same for tiles the patch warp completes
  uint64_t* as_ready = as_edone + kMaxAS;        // [3] patch warp done (edge tiles only)
  uint64_t* as_free = as_ready + kMaxAS;         // [3]
  uint64_t* w_full = as_free + kMaxAS;           // [8]
  uint64_t* w_free = w_full + kMaxWS;            // [8]
  uint64_t* c_full = w_free + kMaxWS;            // [2]
  uint64_t* c_free = c_full + 2;                 // [2]
  uint64_t* pace = c_free + 2;                   // [2] conv MMA pacing groups
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pace + 2);
  float* bias_s = reinterpret_cast<float*>(tmem_ptr + 4);             // [NB]

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxXS; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_efull[i], 1); mbar_init(&x_ready[i], 1); mbar_init(&x_free[i], 1); }
    for (int i = 0; i < kMaxNU; ++i) { mbar_init(&u_full[i], 1); mbar_init(&u_free[i], 6); }
    for (int i = 0; i < 4; ++i) { mbar_init(&a_full[i], 6); mbar_init(&a_free[i], 1); }
    for (int i = 0; i < kMaxNY; ++i) { mbar_init(&y_full[i], 1); mbar_init(&y_free[i], 6); }
    for (int i = 0; i < kMaxAS; ++i) { mbar_init(&as_done[i], 6); mbar_init(&as_edone[i], 6); mbar_init(&as_ready[i], 1); mbar_init(&as_free[i], 1); }
    for (int i = 0; i < kMaxWS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_free[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&c_full[i], 1); mbar_init(&c_free[i], 8); mbar_init(&pace[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWPatch) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  {
    // Toeplitz tap matrices (K-major, no swizzle: [kchunk][n][8]); see act1d_tc.cu for the index algebra
    const float f[12] = {BVG_F0, BVG_F1, BVG_F2, BVG_F3, BVG_F4, BVG_F5, BVG_F5, BVG_F4, BVG_F3, BVG_F2, BVG_F1, BVG_F0};
    for (int idx = threadIdx.x; idx < 64 * 48; idx += kThreads) {
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
    for (int idx = threadIdx.x; idx < 32 * 96; idx += kThreads) {
      const int n = idx / 96, k = idx - n * 96;
      const int ti = k - 2 * n - 11;
      float g = 0.f;
#pragma unroll
      for (int q = 0; q < 12; ++q) if (q == ti) g = f[q];
      reinterpret_cast<__half*>(dnm)[((k >> 3) * 32 + n) * 8 + (k & 7)] = __float2half_rn(g);
    }
    // the A stages (and the spare zero panel of C = 24) start as finite zeros; x ring too (phantom slots are read by the MMA)
    const uint4 z = make_uint4(0, 0, 0, 0);
    uint4* az = reinterpret_cast<uint4*>(asm_);
    for (uint32_t i = threadIdx.x; i < (uint32_t)P.nas * (P.a_stage_bytes >> 4); i += kThreads) az[i] = z;
    uint4* xz = reinterpret_cast<uint4*>(xsm);
    for (uint32_t i = threadIdx.x; i < (uint32_t)P.nxs * (kXStageBytes >> 4); i += kThreads) xz[i] = z;
    for (int i = threadIdx.x; i < P.NB; i += kThreads) bias_s[i] = (P.bias && i < P.Cout) ? P.bias[i] : 0.f;
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const bool is_ctl = (warp & 3) == 3 && warp < kWEpi0;

  if (is_ctl) {
    if (warp == kWProdX) {
      // ===================== TMA producer: raw rows of every (segment, chunk), kXR rows per tile ==========================
      Ring xr;
      const int ps = lane / P.rc, pcc = lane - ps * P.rc;               // lane L < 12 issues the copy of chunk-slot L
      for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
        const TcItem it = tc_item(P, item);
        if (it.nblk <= 0) continue;
        const int nstages = (it.nblk + 4 + kXB - 1) / kXB;
        const __nv_bfloat16* xb = P.x + (int64_t)it.b * P.x_bstride;
        for (int st = 0; st < nstages; ++st) {
          tc_wait<64>(&x_free[xr.s], xr.ph ^ 1, 1);
          uint64_t* full = x_edge(P, it, st) ? &x_efull[xr.s] : &x_full[xr.s];
          if (lane == 0) {
            uint32_t total = 0;
            for (int s = 0; s < P.S; ++s) {
              const int r0 = seg_r0(P, it, s);
              if (r0 >= it.T) break;
              const int ts = xstage_t0(r0, st);
              const int lo = max(ts, 0), hi = min(ts + kXR, it.T);
              if (hi > lo) total += (uint32_t)(hi - lo) * 16u * (uint32_t)P.rc;
            }
            mbar_expect_tx(full, total);
          }
          __syncwarp();
          if (lane < kSlots) {
            const int r0 = seg_r0(P, it, ps);
            const int ts = xstage_t0(r0, st);
            const int lo = max(ts, 0), hi = min(ts + kXR, it.T);
            if (r0 < it.T && hi > lo) {
              bulk_g2s(smem_u32(xsm + xr.s * kXStageBytes) + (uint32_t)((lane * kXR + (lo - ts)) * 16),
                       xb + ((int64_t)pcc * P.x_tp + P.x_pad + lo) * 8, (uint32_t)(hi - lo) * 16u, full);
            }
          }
          xr.next(P.nxs);
        }
      }
    } else if (warp == kWProdW) {
      // ===================== TMA producer: conv weights (resident image, or a ring of (64-channel block, tap) tiles) ========
      if (P.w_resident) {
        if (lane == 0) {
          mbar_expect_tx(&w_full[0], P.w_total_bytes);
          for (uint32_t o = 0; o < P.w_total_bytes; o += 32768u)
            bulk_g2s(smem_u32(wsm) + o, reinterpret_cast<const uint8_t*>(P.w) + o, min(32768u, P.w_total_bytes - o), &w_full[0]);
        }
      } else if (lane == 0) {
        Ring wr;
        const int ncb = (P.Cin_p + 63) >> 6;
        for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
          const TcItem it = tc_item<false>(P, item);
        if (it.nblk <= 0) continue;
          for (int n = 0; n < (it.nblk >> 2); ++n)
            for (int cb = 0; cb < ncb; ++cb) {
              const int kcn = min(8, (P.Cin_p >> 3) - cb * 8);
              const uint32_t bytes = (uint32_t)kcn * P.NB * 16u;
              const uint8_t* src = reinterpret_cast<const uint8_t*>(P.w) + (size_t)cb * 64 * P.NB * P.K * 2;
              for (int tp = 0; tp < P.K; ++tp) {
                tc_wait<64>(&w_free[wr.s], wr.ph ^ 1, 2);
                mbar_expect_tx(&w_full[wr.s], bytes);
                bulk_g2s(smem_u32(wsm) + (uint32_t)wr.s * P.w_slot_bytes, src + (size_t)tp * bytes, bytes, &w_full[wr.s]);
                wr.next(P.w_slots);
              }
            }
        }
      }
    } else if (warp == kWPatch) {
      // ===================== patch warp: replicate padding of staged input tiles; edge / zero rows of A stages ==============
      // two in-order streams (input tiles, conv tiles), polled: only the entries x_edge / tile_edge flag involve this warp
      int xi = blockIdx.x, xst = 0; Ring xr; uint32_t xe_ph = 0; bool xv = xi < P.nitems;
      int ti = blockIdx.x, tn = 0; Ring ar; uint32_t ae_ph = 0; bool tv = ti < P.nitems;
      TcItem xit = xv ? tc_item(P, xi) : TcItem{0, 0, 0}, tit = xit;
      auto x_skip = [&]() {              // advance to the next input tile that needs patching
        while (xv) {
          const int nstages = xit.nblk > 0 ? (xit.nblk + 4 + kXB - 1) / kXB : 0;      // (empty items are skipped by every role)
          if (xst >= nstages) { xi += gridDim.x; xst = 0; xv = xi < P.nitems; if (xv) xit = tc_item(P, xi); continue; }
          if (x_edge(P, xit, xst)) return;
          ++xst; xr.next(P.nxs);
        }
      };
      auto t_skip = [&]() {
        while (tv) {
          if (tn >= (tit.nblk >> 2)) { ti += gridDim.x; tn = 0; tv = ti < P.nitems; if (tv) tit = tc_item(P, ti); continue; }
          if (tile_edge(P, tit, tn)) return;
          ++tn; ar.next(P.nas);
        }
      };
      x_skip(); t_skip();
      while (xv || tv) {
        if (xv && mbar_test(&x_efull[xr.s], (xe_ph >> xr.s) & 1u)) {
          xe_ph ^= 1u << xr.s;
          uint4* stage = reinterpret_cast<uint4*>(xsm + xr.s * kXStageBytes);
          for (int s = 0; s < P.S; ++s) {
            const int r0 = seg_r0(P, xit, s);
            if (r0 >= xit.T) break;
            const int ts = xstage_t0(r0, xst);
            const int nlo = min(max(-ts, 0), kXR);                       // rows [0, nlo): t < 0
            const int rhi = min(max(xit.T - ts, 0), kXR);                // rows [rhi, kXR): t >= T
            if (nlo == 0 && rhi == kXR) continue;
            for (int cc = 0; cc < P.rc; ++cc) {
              uint4* base = stage + (s * P.rc + cc) * kXR;
              if (nlo > 0) {
                const uint4 v = nlo < kXR ? base[nlo] : make_uint4(0, 0, 0, 0);       // row of t = 0
                for (int r = lane; r < nlo; r += 32) base[r] = v;
              }
              if (rhi < kXR) {
                const uint4 v = rhi > 0 ? base[rhi - 1] : make_uint4(0, 0, 0, 0);     // row of t = T - 1
                __syncwarp();
                for (int r = rhi + lane; r < kXR; r += 32) base[r] = v;
              }
            }
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&x_ready[xr.s]);
          ++xst; xr.next(P.nxs);
          x_skip();
        } else if (tv && mbar_test(&as_edone[ar.s], (ae_ph >> ar.s) & 1u)) {
          ae_ph ^= 1u << ar.s;
          uint4* stage = reinterpret_cast<uint4*>(asm_ + (size_t)ar.s * P.a_stage_bytes);
          const uint4* eb = reinterpret_cast<const uint4*>(P.edge) + (int64_t)tit.b * 6 * P.rc;
          const uint4 z = make_uint4(0, 0, 0, 0);
          for (int s = 0; s < P.S; ++s) {
            const int r0 = seg_r0(P, tit, s);
            if (r0 >= tit.T) break;
            const int tA = r0 + 128 * tn - P.LH;                         // time of stage row 0
            const int zlo = min(max(-tA, 0), P.SR);                      // rows [0, zlo): t < 0
            const int zhi = min(max(tit.T - tA, 0), P.SR);               // rows [zhi, SR): t >= T
            for (int cc = 0; cc < P.rc; ++cc) {
              uint4* base = stage + (s * P.rc + cc) * P.SR;
              for (int r = lane; r < zlo; r += 32) base[r] = z;
              for (int r = zhi + lane; r < P.SR; r += 32) base[r] = z;
            }
            // exact a(x)[0..2] and a(x)[T-3..T-1]: lane -> (side, row, chunk)
            for (int e = lane; e < 6 * P.rc; e += 32) {
              const int sr = e / P.rc, cc = e - sr * P.rc;               // sr = side * 3 + row
              const int t = sr < 3 ? sr : tit.T - 6 + sr;
              const int r = t - tA;
              if (r >= 0 && r < P.SR) stage[(s * P.rc + cc) * P.SR + r] = eb[sr * P.rc + cc];
            }
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&as_ready[ar.s]);
          ++tn; ar.next(P.nas);
          t_skip();
        } else {
          __nanosleep(100);
          if (TC_ABORTED()) break;
        }
      }
    } else if (warp == kWUp) {
      // ===================== MMA issuer 1: up-FIRs.  D fp32 | A bf16 MN-major | B bf16 | N = 64 | M = 128 ===================
      const uint32_t idesc_up = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t uph_lo = (smem_u32(up_hi) >> 4) | (64u << 16);
      const uint32_t upl_lo = (smem_u32(up_lo) >> 4) | (64u << 16);
      const uint32_t xs_lo = (smem_u32(xsm) >> 4) | (8u << 16);
      Ring xr, ur; uint32_t xrdy_ph = 0, xfull_ph = 0, tr_nb = 0;
      // U barriers are indexed by BLOCK parity (= the snake group that takes the block), U buffers by block % nu: every
      // waiter then sees every phase of its barrier whatever the ring depth (1 buffer for C = 24, 2-3 otherwise)
      uint32_t nbk = 0;
      const uint32_t nu_ = (uint32_t)P.nu;
      long long dbg_wx = 0, dbg_wu = 0;
      const long long dbg_start = (TCF_COUNTERS && P.dbg) ? clock64() : 0;
      unsigned long long dbg_ns0 = 0;
      if (TCF_COUNTERS && P.dbg) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_ns0));
      for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
        const TcItem it = tc_item(P, item);
        if (it.nblk <= 0) continue;
        const int nup = it.nblk + 4;
        for (int nu = 0; nu < nup; ++nu) {
          const int pos = nu % kXB;
          if (pos == 0) {
            DBG_T0();
            if (x_edge(P, it, nu / kXB)) { tc_wait<0>(&x_ready[xr.s], (xrdy_ph >> xr.s) & 1u, 3); xrdy_ph ^= 1u << xr.s; }
            else { tc_wait<0>(&x_full[xr.s], (xfull_ph >> xr.s) & 1u, 4); xfull_ph ^= 1u << xr.s; }
            DBG_ADD(dbg_wx);
          }
          if (nbk >= nu_) {                                            // the buffer's previous block (nbk - nu) was read out
            const uint32_t pb = nbk - nu_;
            DBG_T0(); tc_wait<0>(&u_free[pb & 1u], (pb >> 1) & 1u, 5); DBG_ADD(dbg_wu);
          }
          tc_fence_after();
          TCF_TRACE(0, tr_nb); ++tr_nb;
          const uint32_t d = tmem_base + kColU + (uint32_t)ur.s * 64u;
          const uint32_t a0 = xs_lo + (uint32_t)xr.s * (kXStageBytes >> 4) + (uint32_t)pos * 32u;
          if (!TCF_DRY(2))
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            umma_ss_elect<kAHi, 0x4008u>(d, a0 + 16u * s, uph_lo + 128u * s, idesc_up, s > 0 ? 1u : 0u);
            umma_ss_elect<kAHi, 0x4008u>(d, a0 + 16u * s, upl_lo + 128u * s, idesc_up, 1u);
          }
          umma_commit_elect(&u_full[nbk & 1u]);
          ++nbk;
          if (pos == kXB - 1 || nu == nup - 1) { umma_commit_elect(&x_free[xr.s]); xr.next(P.nxs); }
          ur.next(P.nu);
        }
      }
      if (TCF_COUNTERS && P.dbg && lane == 0) {
        long long* d = P.dbg + blockIdx.x * 16;
        d[0] = dbg_wx; d[1] = dbg_wu; d[4] = clock64() - dbg_start;
        unsigned long long ns1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1)); d[13] = (long long)(ns1 - dbg_ns0);
      }
    } else if (warp == kWDn) {
      // ===================== MMA issuer 2: down-FIRs.  D fp32 | A fp16 (TMEM) | B fp16 | N = 32 | M = 128 ===================
      const uint32_t idesc_dn = (1u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t dn_lo = (smem_u32(dnm) >> 4) | (32u << 16);
      Ring yr;
      uint32_t afull_ph = 0, tr_j = 0;
      long long dbg_wa = 0, dbg_wy = 0;
      for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
        const TcItem it = tc_item(P, item);
        if (it.nblk <= 0) continue;
        int next_a = -2;
        for (int j = -1; j <= it.nblk; ++j) {
          while (next_a <= j + 1) {
            const int sl = next_a & 3;
            { DBG_T0(); tc_wait<0>(&a_full[sl], (afull_ph >> sl) & 1u, 6); DBG_ADD(dbg_wa); }
            afull_ph ^= 1u << sl;
            ++next_a;
          }
          { DBG_T0(); tc_wait<0>(&y_free[yr.s], yr.ph ^ 1, 7); DBG_ADD(dbg_wy); }
          tc_fence_after();
          TCF_TRACE(4, tr_j); ++tr_j;
          const uint32_t d = tmem_base + P.colY + (uint32_t)yr.s * 32u;
          const uint32_t ap = tmem_base + P.colA + (uint32_t)((j - 1) & 3) * 32u + 24u;
          const uint32_t ac = tmem_base + P.colA + (uint32_t)(j & 3) * 32u;
          const uint32_t an = tmem_base + P.colA + (uint32_t)((j + 1) & 3) * 32u;
          if (!TCF_DRY(4)) {
            umma_ts_elect<0x4008u>(d, ap, dn_lo, idesc_dn, 0u);
#pragma unroll
            for (int s = 1; s < 5; ++s) umma_ts_elect<0x4008u>(d, ac + 8u * (s - 1), dn_lo + 64u * s, idesc_dn, 1u);
            umma_ts_elect<0x4008u>(d, an, dn_lo + 64u * 5, idesc_dn, 1u);
          }
          umma_commit_elect(&y_full[yr.s]);
          umma_commit_elect(&a_free[(j - 1) & 3]);
          if (j == it.nblk) {
            umma_commit_elect(&a_free[j & 3]);
            umma_commit_elect(&a_free[(j + 1) & 3]);
          }
          yr.next(P.ny);
        }
      }
      if (TCF_COUNTERS && P.dbg && lane == 0) { long long* d = P.dbg + blockIdx.x * 16; d[2] = dbg_wa; d[3] = dbg_wy; }
    } else if (warp == kWConv) {
      // ===================== MMA issuer 3: the conv.  D fp32 | A bf16 K-major (A stage) | B bf16 (weights) | N = NB =========
      // The issue loop is latency-bound on the uniform datapath (an MMA whose operands take ~15 dependent uniform instructions
      // and a few constant-bank loads to compute costs ~150 cycles where the tensor pipe needs 40-56): everything is hoisted
      // into locals, the k-steps are unrolled, the taps advance by constant increments.
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P.NB >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t SR = (uint32_t)P.SR, NB = (uint32_t)P.NB;
      const uint32_t a_base = ((smem_u32(asm_) >> 4) | (SR << 16)) + (uint32_t)(P.LH - P.lo);
      const uint32_t w_base = (smem_u32(wsm) >> 4) | (NB << 16);
      const uint32_t as16 = P.a_stage_bytes >> 4, ws16 = P.w_slot_bytes >> 4;
      const uint32_t SR2 = 2u * SR, NB2 = 2u * NB, dil = (uint32_t)P.dil, seg_a = (uint32_t)P.rc * SR, c_base = tmem_base + P.colC;
      const int K = P.K, S = P.S, nas = P.nas, nacc = P.nacc, w_slots = P.w_slots;
      const int nk0 = min(8, P.Cin_p >> 3) >> 1, nk1 = ((P.Cin_p >> 3) - 2 * nk0) >> 1;        // 16-channel steps of the two 64-channel blocks
      const bool resident = P.w_resident != 0;
      Ring ar, cr, wr; uint32_t ardy_ph = 0, adone_ph = 0;
      bool w_waited = false;
      uint32_t tr_t = 0;
      long long dbg_ws = 0, dbg_wc = 0, dbg_ww = 0;
      // the K taps of one 64-channel block for one accumulator: nk MMAs per tap, A advances `dil` rows per tap
      auto taps = [&](int nk, uint32_t d, uint32_t a_t, uint32_t b, uint32_t& acc) {
        for (int tp = 0; tp < K; ++tp) {
          if (!resident) { tc_wait<0>(&w_full[wr.s], wr.ph, 12); tc_fence_after(); b = w_base + (uint32_t)wr.s * ws16; }
          if (!TCF_DRY(1)) {
            umma_bf16_imm_elect(d, a_t, b, idesc, acc);
            umma_bf16_imm_elect(d, a_t + SR2, b + NB2, idesc, 1u);
            if (nk > 2) umma_bf16_imm_elect(d, a_t + 2u * SR2, b + 2u * NB2, idesc, 1u);
            if (nk > 3) umma_bf16_imm_elect(d, a_t + 3u * SR2, b + 3u * NB2, idesc, 1u);
          }
          acc = 1u;
          a_t += dil;
          if (resident) b += (uint32_t)nk * NB2;
          else { umma_commit_elect(&w_free[wr.s]); wr.next(w_slots); }
        }
      };
      for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
        const TcItem it = tc_item(P, item);
        if (it.nblk <= 0) continue;
        for (int n = 0; n < (it.nblk >> 2); ++n) {
          {
            DBG_T0();
            if (tile_edge(P, it, n)) { tc_wait<20>(&as_ready[ar.s], (ardy_ph >> ar.s) & 1u, 8); ardy_ph ^= 1u << ar.s; }
            else { tc_wait<20>(&as_done[ar.s], (adone_ph >> ar.s) & 1u, 9); adone_ph ^= 1u << ar.s; }
            DBG_ADD(dbg_ws);
          }
          { DBG_T0(); tc_wait<20>(&c_free[cr.s], cr.ph ^ 1, 10); DBG_ADD(dbg_wc); }
          if (resident && !w_waited) { DBG_T0(); tc_wait<0>(&w_full[0], 0u, 11); w_waited = true; DBG_ADD(dbg_ww); }
          tc_fence_after();
          TCF_TRACE(7, tr_t + 96);
          uint32_t d = c_base + (uint32_t)cr.s * (uint32_t)S * NB;
          uint32_t a_s = a_base + (uint32_t)ar.s * as16;
          for (int sg = 0; sg < S; ++sg, d += NB, a_s += seg_a) {
            uint32_t acc = 0u;
            taps(nk0, d, a_s, w_base, acc);
            if (nk1 > 0) taps(nk1, d, a_s + 8u * SR, w_base + (uint32_t)(8 * K) * NB, acc);
          }
          umma_commit_elect(&c_full[cr.s]);
          umma_commit_elect(&as_free[ar.s]);
          TCF_TRACE(8, tr_t + 96); ++tr_t;
          ar.next(nas);
          cr.next(nacc);
        }
      }
      if (TCF_COUNTERS && P.dbg && lane == 0) { long long* d = P.dbg + blockIdx.x * 16; d[9] = dbg_ws; d[10] = dbg_wc; d[11] = dbg_ww; }
    }
  } else if (warp < kWStore0) {
    // ===================== snake: U (fp32, TMEM) -> a = u + hb - hb cos(2 e^alpha u) -> fp16 pairs (TMEM ring) ==============
    // 12 warps: TMEM lane quarter q = warp % 4 (0..2); two groups of 6 take alternate blocks (while one group sits in its
    // tcgen05.ld / tcgen05.st round trips the other keeps the MUFU pipe busy), 32-column half h per warp, done as two
    // 16-column half steps (the warp runs at 64 registers).  One warp per group polls the mbarriers, the rest park on a
    // named barrier.
    const int q = warp & 3, grp = (warp >> 2) & 1, h = warp >> 3;
    const bool leader = q == 0 && h == 0;
    const int ln = q * 32 + lane;
    const int slot = ln >> 3;
    const int cc = slot % P.rc;
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t colA = P.colA;
    const int nu = P.nu;
    float sc0 = 0.f, sc1 = 0.f;
    {
      const int ch = cc * 8 + (ln & 7);
      if (ch < P.C) snake_params<false>(P.alpha[ch], P.beta[ch], sc0, sc1);
    }
    const f32x2 SC0 = pk2(sc0, sc0), SC1 = pk2(sc1, sc1), NSC1 = pk2(-sc1, -sc1);
    uint32_t mine = 0;                 // blocks this group has taken so far (all items)
    uint32_t afree_ph = 0;
    long long dbg_su = 0, dbg_sa = 0;
    const long long dbg_sstart = (TCF_COUNTERS && P.dbg) ? clock64() : 0;
    auto half_step = [&](const uint32_t (&v)[16], uint32_t dst) {
      uint32_t w[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const f32x2 u = pk2(__uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1]));
        float zx, zy;
        unpk2(mul2(u, SC0), zx, zy);
        const f32x2 a = TCF_DRY(8) ? fma2(NSC1, pk2(zx, zy), add2(u, SC1)) : fma2(NSC1, pk2(__cosf(zx), __cosf(zy)), add2(u, SC1));
        float ax, ay;
        unpk2(a, ax, ay);
        w[k] = cvt_f16x2_sat(ax, ay);
      }
      tmem_st8(dst, w);
    };
    for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
      const TcItem it = tc_item(P, item);
        if (it.nblk <= 0) continue;
      // every item has an even number of blocks (nblk + 4), so block parity == parity of i: group g takes i = -2 + g, step 2
      // (A-ring slots sl = i & 3: {2, 0} for group 0, {3, 1} for group 1: each slot's phases are seen by one group only)
      const int iend = it.nblk + 1;
      for (int i = -2 + grp; i <= iend; i += 2, ++mine) {
        const int sl = i & 3;
        const uint32_t ubuf = nu == 2 ? (uint32_t)grp : (nu == 1 ? 0u : (2u * mine + (uint32_t)grp) % 3u);
        if (leader) {
          { DBG_T0(); tc_wait<0>(&u_full[grp], mine & 1u, 13); DBG_ADD(dbg_su); }      // (barrier = block parity)
          { DBG_T0(); tc_wait<0>(&a_free[sl], ((afree_ph >> sl) & 1u) ^ 1u, 14); DBG_ADD(dbg_sa); }
        }
        afree_ph ^= 1u << sl;
        const uint32_t nb = 2u * mine + (uint32_t)grp;                 // (global block index: traces only)
        (void)nb;
        if (grp) named_bar_sync(4, 192); else named_bar_sync(1, 192);
        if (warp == 0) TCF_TRACE(1, nb);
        tc_fence_after();
        const uint32_t ucol = tq + kColU + ubuf * 64u + (uint32_t)h * 32u;
        const uint32_t acol = tq + colA + (uint32_t)sl * 32u + (uint32_t)h * 16u;
        uint32_t v0[16], v1[16];
        tmem_ld16_nowait(ucol, v0);
        tmem_ld16_nowait(ucol + 16u, v1);
        tmem_ld_wait();                                                // U is in registers: the buffer can be rewritten
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&u_free[grp]);
        half_step(v0, acol);
        half_step(v1, acol + 8u);
        if (warp == 0) TCF_TRACE(2, nb);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[sl]);
        if (warp == 0) TCF_TRACE(3, nb);
      }
    }
    if (TCF_COUNTERS && P.dbg && threadIdx.x == 0) { long long* d = P.dbg + blockIdx.x * 16; d[5] = dbg_su; d[6] = dbg_sa; d[15] = clock64() - dbg_sstart; }
  } else if (warp < kWEpi0) {
    // ===================== store: Y (fp32, TMEM; lane = channel) -> bf16 -> conv A stage(s) in shared memory ===============
    // 6 warps: TMEM lane quarter q = warp % 4 (0..2), 16-step half h of every block.  tcgen05.ld.16x256b returns the mma
    // C-fragment layout, which is what stmatrix.trans needs to write 8 channels x 8 time steps as eight 16-byte rows.
    // (issue-bound SM: all addresses are 32-bit shared-memory offsets precomputed per lane, one polling warp for the group)
    const int q = warp & 3, h = (warp >> 2) & 1;
    const uint32_t colY = P.colY;
    const uint32_t tqa = tmem_base + ((uint32_t)(q * 32) << 16) + colY + (uint32_t)h * 16u;
    const uint32_t tqb = tqa + (16u << 16);
    const int LH = P.LH, SR = P.SR, nas = P.nas, ny = P.ny;
    const int slotA = 4 * q + ((lane >> 3) & 1);                       // this lane's stmatrix row: chunk-slot (+2 for the second store)
    const int rowin = 16 * h + (lane >> 4) * 8 + (lane & 7);           // ... and row within the block
    const uint32_t stage0 = smem_u32(asm_), stage_bytes = P.a_stage_bytes;
    const uint32_t lane_off = (uint32_t)(slotA * SR + rowin) * 16u;    // byte offset of this lane's row in a stage (block at row 0)
    const uint32_t pair_off = (uint32_t)(2 * SR) * 16u;                // second store: chunk-slot + 2
    const uint32_t dump_a = smem_u32(dump) + (uint32_t)lane * 16u;
    // halo copies: last LH rows of block 4n - 1 -> rows [0, LH) of tile n; first LH rows of block 4n + 4 -> rows [128 + LH, SR)
    const bool ok_next = rowin >= 32 - LH, any_next = 16 * h + 16 > 32 - LH;
    const bool ok_prev = rowin < LH, any_prev = 16 * h < LH;
    const uint32_t main_off = lane_off + (uint32_t)LH * 16u;
    const uint32_t next_off = ok_next ? lane_off + (uint32_t)(LH - 32) * 16u : 0u;      // (only used when ok_next)
    const uint32_t prev_off = lane_off + (uint32_t)(128 + LH) * 16u;
    Ring yr, open, mainr, closer;
    uint32_t tr_j = 0;
    long long dbg_ty = 0, dbg_tf = 0;
    auto stsm2 = [&](uint32_t a, const uint32_t (&pa)[4], const uint32_t (&pb)[4]) {
      asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};\n" ::"r"(a), "r"(pa[0]), "r"(pa[1]), "r"(pa[2]), "r"(pa[3]) : "memory");
      asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};\n" ::"r"(a + pair_off), "r"(pb[0]), "r"(pb[1]), "r"(pb[2]), "r"(pb[3]) : "memory");
    };
    auto stsm2_masked = [&](bool ok, uint32_t a, const uint32_t (&pa)[4], const uint32_t (&pb)[4]) {
      // lanes whose matrix is not wanted write it to the dump
      asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};\n" ::"r"(ok ? a : dump_a), "r"(pa[0]), "r"(pa[1]), "r"(pa[2]), "r"(pa[3]) : "memory");
      asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};\n" ::"r"(ok ? a + pair_off : dump_a), "r"(pb[0]), "r"(pb[1]), "r"(pb[2]), "r"(pb[3]) : "memory");
    };
    // One block: wait (the group's polling warp) -> Y from TMEM -> free the Y buffer -> bf16 pairs in pa / pb.
    uint32_t pa[4], pb[4];
    auto load_block = [&](bool wait_open) {
      if (warp == kWStore0) {
        { DBG_T0(); tc_wait<0>(&y_full[yr.s], yr.ph, 15); DBG_ADD(dbg_ty); }
        if (wait_open) { DBG_T0(); tc_wait<0>(&as_free[open.s], open.ph ^ 1, 16); DBG_ADD(dbg_tf); }
      }
      named_bar_sync(2, 192);
      if (warp == kWStore0) TCF_TRACE(5, tr_j);
      tc_fence_after();
      uint32_t va[8], vb[8];
      tmem_ld_16x256b_x2_nowait(tqa + (uint32_t)yr.s * 32u, va);
      tmem_ld_16x256b_x2_nowait(tqb + (uint32_t)yr.s * 32u, vb);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&y_free[yr.s]);
      yr.next(ny);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        pa[i] = cvt_bf16x2(__uint_as_float(va[2 * i]), __uint_as_float(va[2 * i + 1]));
        pb[i] = cvt_bf16x2(__uint_as_float(vb[2 * i]), __uint_as_float(vb[2 * i + 1]));
      }
      if (warp == kWStore0) TCF_TRACE(6, tr_j);
      ++tr_j;
    };
    const bool dry = TCF_DRY(32);
    auto st_main = [&](int p) { if (!dry) stsm2(stage0 + (uint32_t)mainr.s * stage_bytes + main_off + (uint32_t)p * 512u, pa, pb); };
    auto st_next = [&]() { if (any_next && !dry) stsm2_masked(ok_next, stage0 + (uint32_t)open.s * stage_bytes + next_off, pa, pb); open.next(nas); };
    auto close_tile = [&](const TcItem& it, int m) {                  // last rows of tile m, then the stage is complete
      if (any_prev && !dry) stsm2_masked(ok_prev, stage0 + (uint32_t)closer.s * stage_bytes + prev_off, pa, pb);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(tile_edge(P, it, m) ? &as_edone[closer.s] : &as_done[closer.s]);
      closer.next(nas);
    };
    // Straight-line per tile (the SM is issue-bound: no per-block case analysis): block -1 opens tile 0's stage; blocks 4n ..
    // 4n + 3 fill tile n, block 4n also closes tile n - 1, block 4n + 3 opens tile n + 1; block nblk closes the last tile.
    for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
      const TcItem it = tc_item(P, item);
      if (it.nblk <= 0) continue;
      const int ntile = it.nblk >> 2;
      load_block(true);
      st_next();
      for (int n = 0; n < ntile; ++n) {
        load_block(false);
        st_main(0);
        if (n > 0) close_tile(it, n - 1);
        load_block(false);
        st_main(1);
        load_block(false);
        st_main(2);
        const bool op = n + 1 < ntile;
        load_block(op);
        st_main(3);
        if (op) st_next();
        mainr.next(nas);
      }
      load_block(false);
      close_tile(it, ntile - 1);
    }
    if (TCF_COUNTERS && P.dbg && threadIdx.x == kWStore0 * 32) { long long* d = P.dbg + blockIdx.x * 16; d[7] = dbg_ty; d[8] = dbg_tf; }
  } else {
    // ===================== conv epilogue: 8 warps = 2 sets x 4 TMEM lane quarters; every tile is split between the sets =======
    // (the tile's S x NB / 16 column slices are dealt out half and half, so the single accumulator stage drains twice as fast)
    const int set = (warp - kWEpi0) >> 2, q = warp & 3;
    const int r = q * 32 + lane;                                       // accumulator row
    EpiConst E;
    {
      const int nslt = P.S * (P.NB >> 4);
      // every tile is split between the two sets (half the column slices each): the accumulator stage drains twice as fast
      const int f0 = set ? nslt >> 1 : 0, f1 = set ? nslt : nslt >> 1;
      E.nsl = P.NB >> 4;
      E.sg0 = f0 / E.nsl; E.sl0 = f0 - E.sg0 * E.nsl; E.cnt = f1 - f0;
      E.T = P.T; E.RL = P.RL; E.row_base = P.y_pad;                  // (E.T is set per item: ragged batches)
      E.cs = (uint32_t)P.y_tp * 8u; E.cs2 = 2u * E.cs;
      E.scale = P.scale; E.do_scale = P.scale != 1.f;
    }
    const int nacc = P.nacc, S = P.S;
    const uint32_t acc_cols = (uint32_t)(P.S * P.NB);
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16) + P.colC;
    const bool has_r1 = P.res1 != nullptr, has_r2 = P.res2 != nullptr;
    // the tile that holds row T - 1: global segment (T - 1) / RL -> item group, tile within the segment
    const bool zero_pads = P.zero_pads != 0;
    const int RLc = P.RL;
    Ring cr;
    uint32_t tile = 0;
    long long dbg_ew = 0, dbg_eb = 0;
    for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
      const TcItem it = tc_item(P, item);
        if (it.nblk <= 0) continue;
      __nv_bfloat16* yb = P.y + (int64_t)it.b * P.y_bstride;
      const __nv_bfloat16* r1 = has_r1 ? P.res1 + (int64_t)it.b * P.y_bstride : nullptr;
      const __nv_bfloat16* r2 = has_r2 ? P.res2 + (int64_t)it.b * P.y_bstride : nullptr;
      const int seg_t0 = seg_r0(P, it, E.sg0);
      const int ntile = it.nblk >> 2;
      E.T = it.T;
      const int zl_gs = (it.T - 1) / RLc, zl_grp = zl_gs / S, zl_n = ((it.T - 1) - zl_gs * RLc) >> 7;
      for (int n = 0; n < ntile; ++n, ++tile) {
        {
          // one warp polls the mbarrier, the other seven park on a named barrier (no issue slots: the SM is issue-bound)
          if (warp == kWEpi0) {
            DBG_T0(); tc_wait<64>(&c_full[cr.s], cr.ph, 17); DBG_ADD(dbg_ew);
          }
          named_bar_sync(3, 256);
          if (q == 0 && set == 0) TCF_TRACE(9, tile + 96);
          tc_fence_after();
          DBG_T0();
          const uint32_t tbase = tq + (uint32_t)cr.s * acc_cols;
          const int t_in_seg = 128 * n + r;
          if (has_r2) epilogue_tile<true, true>(E, seg_t0, t_in_seg, yb, r1, r2, bias_s, tbase);
          else if (has_r1) epilogue_tile<true, false>(E, seg_t0, t_in_seg, yb, r1, nullptr, bias_s, tbase);
          else epilogue_tile<false, false>(E, seg_t0, t_in_seg, yb, nullptr, nullptr, bias_s, tbase);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&c_free[cr.s]);
          if (q == 0 && set == 0) TCF_TRACE(10, tile + 96);
          if ((has_r1 || has_r2) && n + 1 < ntile) {
            // pull this warp's residual vectors of the NEXT tile towards L2, so that the one-slice-ahead register prefetch of
            // the slices only has to cover an L2 hit
            int sl = E.sl0, r0 = seg_t0;
            for (int i = 0; i < E.cnt; ++i) {
              const int t = r0 + t_in_seg + 128;
              if (r0 < E.T && t < E.T) {
                const uint32_t off = (uint32_t)(E.row_base + t) * 8u + (uint32_t)sl * E.cs2;
                if (has_r1) { asm volatile("prefetch.global.L2 [%0];" ::"l"(r1 + off)); asm volatile("prefetch.global.L2 [%0];" ::"l"(r1 + off + E.cs)); }
                if (has_r2) { asm volatile("prefetch.global.L2 [%0];" ::"l"(r2 + off)); asm volatile("prefetch.global.L2 [%0];" ::"l"(r2 + off + E.cs)); }
              }
              if (++sl == E.nsl) { sl = 0; r0 += E.RL; }
            }
          }
          if (zero_pads && ((it.grp == 0 && n == 0) || (it.grp == zl_grp && n == zl_n))) {
            // rows [-pad, 0) by the first tile of the utterance, [T, T + pad) by the tile that holds row T - 1
            const uint4 z = make_uint4(0, 0, 0, 0);
            const int et = set * 128 + r, nth = 256;
            const int npad = P.y_chunks * P.y_pad;
            if (it.grp == 0 && n == 0)
              for (int i = et; i < npad; i += nth)
                *reinterpret_cast<uint4*>(yb + ((int64_t)(i / P.y_pad) * P.y_tp + (i % P.y_pad)) * 8) = z;
            if (it.grp == zl_grp && n == zl_n)
              for (int i = et; i < npad; i += nth)
                *reinterpret_cast<uint4*>(yb + ((int64_t)(i / P.y_pad) * P.y_tp + P.y_pad + it.T + (i % P.y_pad)) * 8) = z;
          }
          DBG_ADD(dbg_eb);
        }
        cr.next(nacc);
      }
    }
    if (TCF_COUNTERS && P.dbg && threadIdx.x == kWEpi0 * 32) { long long* d = P.dbg + blockIdx.x * 16; d[12] = dbg_ew; d[14] = dbg_eb; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWPatch) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// exact a(x)[0..2] and a(x)[T-3..T-1] of every channel (replicate padding of x and of the activated signal, the CUDA-core
// stencil of act1d_core.cuh): edge[b][side][row][chunk][8] bf16.  One warp per (chunk, utterance).
__global__ void __launch_bounds__(32) actconv_edge_kernel(__nv_bfloat16* __restrict__ edge, const __nv_bfloat16* __restrict__ x,
                                                          const float* __restrict__ alpha_log, const float* __restrict__ beta_log,
                                                          int C, int rc, int64_t x_bstride, int T, int Tp, int pad,
                                                          const int* __restrict__ lens, int len_mul) {
  const int chunk = blockIdx.x, b = blockIdx.y, lane = threadIdx.x;
  if (lane >= 16) return;
  if (lens) T = lens[b] * len_mul;
  constexpr int V = 8;
  const int side = lane >> 3, c8 = lane & 7;
  const int ch = chunk * 8 + c8;
  const __nv_bfloat16* xb = x + (int64_t)b * x_bstride + (int64_t)chunk * Tp * 8;
  const int64_t tg = side ? T - V : 0;
  float xw[V + 16], yv[V];
#pragma unroll
  for (int i = 0; i < V + 16; ++i) {
    const int64_t t = tg - 8 + i;
    xw[i] = (t >= 0 && t < T) ? __bfloat162float(xb[(pad + t) * 8 + c8]) : 0.f;
  }
  float sc0 = 0.f, sc1 = 0.f;
  if (ch < C) snake_params<false>(alpha_log[ch], beta_log[ch], sc0, sc1);
  act1d_window<V, false>(xw, yv, sc0, sc1, tg, (int64_t)T);
#pragma unroll
  for (int qq = 0; qq < 3; ++qq) {
    const int qi = side ? V - 3 + qq : qq;
    edge[((((int64_t)b * 2 + side) * 3 + qq) * rc + chunk) * 8 + c8] = __float2bfloat16_rn(ch < C ? yv[qi] : 0.f);
  }
}

}  // namespace

size_t actconv_tc_scratch_bytes(int64_t B) { return (size_t)B * 6 * kSlots * 16; }

// Activation1d(x) -> Conv1d fused, FIRs on the tensor cores (see the header).  BVG_ERR_STATE (nothing launched) when the
// layer does not qualify; the caller then takes conv_umma_fused_launch / the two-kernel path.
int actconv_tc_launch(const UmmaLayer& L, const C8T& x, const float* act_alpha, const float* act_beta, const C8T& y,
                      const UmmaEpilogue& ep, int64_t B, void* scratch, cudaStream_t st) {
  if (L.transposed || L.split || !act_alpha || !act_beta || !scratch) return BVG_ERR_STATE;
  if (ep.cond || ep.relu || ep.post_scale || ep.act || ep.yf32) return BVG_ERR_STATE;
  if (L.Cin % 8 != 0 || (L.Cin != 24 && L.Cin != 48 && L.Cin != 96)) return BVG_ERR_STATE;
  if (x.T < 512 || y.T != x.T || x.pad != kC8tPad || y.pad != kC8tPad || L.K > 16) return BVG_ERR_STATE;
  BVG_CHECK_ARG(L.w && x.p && y.p, "actconv_tc: null pointer");
  BVG_CHECK_ARG(x.C == L.Cin && y.C == L.Cout, "actconv_tc: shape mismatch");
  ActConvTcParams P;
  memset(&P, 0, sizeof P);
  int n_nblk = 1;
  umma_choose_nb(L.Cout, 1, &P.NB, &n_nblk);
  if (n_nblk != 1) return BVG_ERR_STATE;
  P.C = L.Cin; P.rc = L.Cin / 8; P.S = kSlots / P.rc;
  P.Cin_p = (L.Cin + 15) / 16 * 16;
  if (P.S * P.NB > 128 || y.chunks * 8 < P.NB || x.chunks * 8 < P.Cin_p) return BVG_ERR_STATE;
  // TMEM plan (512 columns): one conv accumulator stage; what is left goes to the FIR rings (their depth sets how much of the
  // MMA -> barrier -> warp -> MMA round trips overlaps): U 3 x 64 | A 4 x 32 | Y 3 x 32 | conv 96, or U 2 x 64 | A | Y 4 x 32 | conv 128
  // Measured (profiles/README.md, round 2): two conv accumulator stages beat deeper FIR rings (the epilogue of tile n then
  // overlaps the MMAs of tile n + 1), so C = 96 / 48 run U 2 x 64 | A 4 x 32 | Y 2 x 32 | conv 2 x 96.
  if (P.S * P.NB <= 96 && BVG_ENV_ONCE("BVG_TCF_NACC", 2) == 2) { P.nacc = 2; P.nu = 2; P.ny = 2; }
  else if (P.S * P.NB <= 96) { P.nacc = 1; P.nu = 3; P.ny = 3; }
  else if (BVG_ENV_ONCE("BVG_TCF_NACC24", 2) == 2) { P.nacc = 2; P.nu = 1; P.ny = 2; }   // C = 24: U 64 | A 128 | Y 64 | conv 2 x 128
  else { P.nacc = 1; P.nu = 2; P.ny = 4; }
  if (const int e = BVG_ENV_ONCE("BVG_TCF_NU", 0)) P.nu = std::min(kMaxNU, std::max(1, e));     // (experiments)
  if (const int e = BVG_ENV_ONCE("BVG_TCF_NY", 0)) P.ny = std::min(kMaxNY, std::max(2, e));
  P.colA = (uint32_t)P.nu * 64u; P.colY = P.colA + 128u; P.colC = P.colY + (uint32_t)P.ny * 32u;
  if (P.colC + (uint32_t)(P.nacc * P.S * P.NB) > 512u) return BVG_ERR_STATE;
  P.pace = 0;
  P.Cout = L.Cout; P.K = L.K; P.dil = L.dil;
  P.lo = L.dil * (L.K - 1) / 2;
  if ((L.dil * (L.K - 1)) & 1 || P.lo > 32) return BVG_ERR_STATE;
  P.LH = P.lo <= 8 ? 8 : P.lo <= 16 ? 16 : 32;
  P.SR = 128 + 2 * P.LH;
  P.NP = kSlots + (P.Cin_p / 8 - P.rc);
  P.a_stage_bytes = (uint32_t)P.NP * P.SR * 16u;
  P.x = x.p; P.x_bstride = x.batch_stride(); P.x_tp = x.Tp; P.x_pad = x.pad;
  P.y = y.p; P.y_bstride = y.batch_stride(); P.y_tp = y.Tp; P.y_pad = y.pad; P.y_chunks = y.chunks;
  P.w = L.w; P.res1 = ep.res1; P.res2 = ep.res2; P.bias = ep.bias; P.scale = ep.scale;
  P.alpha = act_alpha; P.beta = act_beta;
  P.edge = static_cast<const __nv_bfloat16*>(scratch);
  P.T = x.T; P.zero_pads = ep.zero_pads;
  P.lens = x.lens; P.len_mul = x.len_mul;
  P.dbg = g_dbg_buf;
#ifdef BVG_DEBUG
  if (const char* e = getenv("BVG_TCF_DRY")) P.dry = atoi(e);
  if (getenv("BVG_TCF_TRACE") && g_dbg_buf) P.trace = g_dbg_buf + 148 * 16;      // (read per launch: the sweep tool changes it between launches)
#endif
  // shared memory plan: taps + barriers fixed; A stages 3 deep, input ring 4 (min 2) deep, weights resident when they fit
  const size_t fixed = 2 * kUpBytes + kDnBytes + kDumpBytes + 80 * 8 + 16 + (size_t)P.NB * 4 + 128;
  const size_t budget = 227 * 1024 - fixed;
  P.w_total_bytes = (uint32_t)((size_t)P.Cin_p * P.NB * P.K * 2);
  P.w_slot_bytes = (uint32_t)std::min(8, P.Cin_p / 8) * P.NB * 16u;
  P.nas = kMaxAS;
  size_t used = (size_t)P.nas * P.a_stage_bytes;
  P.nxs = kMaxXS;
  const int w_force_ring = BVG_ENV_ONCE("BVG_TCF_WRING", 0);
  if (!w_force_ring && used + P.w_total_bytes + 3 * kXStageBytes <= budget) {
    P.w_resident = 1;
    used += P.w_total_bytes;
  } else {
    if (P.S != 1) return BVG_ERR_STATE;                 // (the weight ring is consumed once per tile: single-segment shapes only)
    P.w_resident = 0;
    P.nxs = 3;
    const size_t left = budget - used - (size_t)P.nxs * kXStageBytes;
    P.w_slots = (int)std::min<size_t>(kMaxWS, left / P.w_slot_bytes);
    if (P.w_slots < 2) return BVG_ERR_STATE;
    used += (size_t)P.w_slots * P.w_slot_bytes;
  }
  while (P.nxs > 2 && used + (size_t)P.nxs * kXStageBytes > budget) --P.nxs;
  if (used + (size_t)P.nxs * kXStageBytes > budget) return BVG_ERR_STATE;
  used += (size_t)P.nxs * kXStageBytes;
  const size_t smem = used + fixed;
  int num_sms = 0;
  BVG_TRY(current_device_sms(&num_sms));
  // range length: multiples of 128 rows; minimise (rounds of items over the SMs) x (blocks per item incl. halo blocks and ramp)
  int64_t best_cost = -1;
  const int rl_env = BVG_ENV_ONCE("BVG_TCF_RL", 0);
  for (int rl = 256; rl <= 16384; rl += 128) {
    const int64_t nr = (x.T + rl - 1) / rl, ng = (nr + P.S - 1) / P.S;
    const int64_t items = B * ng;
    const int64_t cost = ((items + num_sms - 1) / num_sms) * (rl / kBlk + 10);
    if (best_cost < 0 || cost < best_cost || rl == rl_env) { best_cost = rl == rl_env ? 0 : cost; P.RL = rl; P.NG = (int)ng; }
    if (rl >= x.T) break;
  }
  const int64_t items = B * P.NG;
  BVG_CHECK_ARG(items < (1ll << 31) && B <= 65535, "actconv_tc: too many work items");
  P.nitems = (int)items;
  static std::atomic<uint64_t> opted{0};
  BVG_TRY(smem_opt_in(actconv_tc_kernel, opted, 227 * 1024));
  ProfScope prof(st, KC_ACTCONV);
  actconv_edge_kernel<<<dim3((unsigned)P.rc, (unsigned)B), 32, 0, st>>>(static_cast<__nv_bfloat16*>(scratch), x.p, act_alpha, act_beta,
                                                                       L.Cin, P.rc, P.x_bstride, x.T, x.Tp, x.pad, x.lens, x.len_mul);
  BVG_LAUNCHED();
  actconv_tc_kernel<<<(unsigned)std::min<int64_t>(items, num_sms), kThreads, smem, st>>>(P);
  BVG_LAUNCHED();
#ifdef BVG_DEBUG
  {
    int ab[8] = {0};
    BVG_CUDA(cudaStreamSynchronize(st));
    BVG_CUDA(cudaMemcpyFromSymbol(ab, g_tcf_abort, sizeof ab));
    if (ab[0]) {
      fprintf(stderr, "actconv_tc watchdog: block %d warp %d wait id %d parity %d (C=%d K=%d dil=%d T=%d RL=%d NG=%d nitems=%d nxs=%d nas=%d "
              "nacc=%d wres=%d wslots=%d)\n", ab[1], ab[2], ab[3], ab[4], P.C, P.K, P.dil, P.T, P.RL, P.NG, P.nitems, P.nxs, P.nas,
              P.nacc, P.w_resident, P.w_slots);
      const int zero[8] = {0};
      cudaMemcpyToSymbol(g_tcf_abort, zero, sizeof zero);
      set_error("actconv_tc: watchdog fired (block %d warp %d wait %d)", ab[1], ab[2], ab[3]);
      return BVG_ERR_CUDA;
    }
  }
#endif
  return BVG_OK;
}

}  // namespace bvg
