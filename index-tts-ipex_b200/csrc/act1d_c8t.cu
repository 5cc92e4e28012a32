// Activation1d and the conv_post tail in the channel-chunked "c8t" layout of the bf16 path
// ([B][C/8][Tp][8] bf16, see conv_umma.cu).  Same arithmetic as act1d.cu (alias_free_torch/act.py:24-29
// of the reference, PyTorch edge semantics); only the data movement differs:
//   * a CTA owns NCH channel-chunks x TR rows; per chunk the rows are one contiguous run, staged by
//     one TMA bulk copy (cp.async.bulk, mbarrier completion) including the 8-row FIR halo;
//   * a thread owns one 32-bit word (2 channels) x 16 consecutive rows: 32 LDS give it the whole
//     window (16 outputs + 8-row halo each side), the two channels run through the shared stencil
//     one after the other (2.6 activated intermediates per output instead of 3.25 at 8 rows/thread),
//     results are re-packed in shared memory and leave through one TMA bulk store per chunk
//     (cp.async.bulk.global.shared::cta).  128 registers -> 2 CTAs/SM, so one CTA's TMA load/store
//     overlaps the other's FIR math.
//   * the kernel also (re)writes the zero halo rows and zero padding channels that the tcgen05 conv
//     reads as its zero padding.
#include "act1d_core.cuh"
#include "bvg_common.cuh"
#include "umma.cuh"
#include "umma_ptx.cuh"

namespace bvg {
namespace {

// Thread = one 32-bit word (2 channels) x 16 consecutive rows; warp = 4 words x NCH chunks x (8/NCH) row groups.
template <int NCH>
__global__ void __launch_bounds__(256, 3)
act1d_c8t_kernel(__nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ x,
                 const float* __restrict__ alpha_log, const float* __restrict__ beta_log, int C, int chunks, int T,
                 int Tp, int pad, const int* __restrict__ lens, int len_mul) {
  if (lens) T = lens[blockIdx.z] * len_mul;                       // ragged batch: this utterance's own length
  constexpr int V = 16;
  constexpr int RGW = 8 / NCH;            // row groups per warp
  constexpr int TR = 8 * RGW * V;         // rows per CTA (128 / 256 / 512)
  constexpr int RIN = TR + 16 + 1;        // staged rows per chunk (+1: odd pitch)
  constexpr int ROUT = TR + 1;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint4* in = reinterpret_cast<uint4*>(smem_raw);                 // [NCH][RIN]
  uint4* out = in + NCH * RIN;                                    // [NCH][ROUT]
  uint64_t* bar = reinterpret_cast<uint64_t*>(out + NCH * ROUT);

  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * TR;                                 // first (padded-space) row of the tile
  const int ch0 = blockIdx.y * NCH;
  const int b = blockIdx.z;
  const int nch = min(NCH, chunks - ch0);
  const int64_t base = (int64_t)b * chunks * Tp;

  const int lo = max(r0 - 8, 0), hi = min(r0 + TR + 8, Tp);
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)(hi - lo) * 16u;
    mbar_expect_tx(bar, bytes * nch);
    for (int c = 0; c < nch; ++c)
      bulk_g2s(in + c * RIN + (lo - (r0 - 8)), x + (base + (int64_t)(ch0 + c) * Tp + lo) * 8, bytes, bar);
  }
  const int lane = tid & 31, warp = tid >> 5;
  const int pp = lane & 3;
  const int cg = (lane >> 2) % NCH;
  const int rg = warp * RGW + (lane >> 2) / NCH;                  // 16-row group within the tile
  // snake parameters first: their global-load latency overlaps the TMA wait
  float pa0 = 0.f, pb0 = 0.f, pa1 = 0.f, pb1 = 0.f;
  {
    const int chP = (ch0 + cg) * 8 + 2 * pp;
    if (cg < nch && chP < C) snake_params<false>(alpha_log[chP], beta_log[chP], pa0, pb0);
    if (cg < nch && chP + 1 < C) snake_params<false>(alpha_log[chP + 1], beta_log[chP + 1], pa1, pb1);
  }
  mbar_wait(bar, 0);

  if (cg < nch) {
    const int64_t t0 = (int64_t)r0 + V * rg - pad;                // time index of this thread's first output
    const uint32_t* inw = reinterpret_cast<const uint32_t*>(in) + (size_t)(cg * RIN + V * rg) * 4 + pp;
    uint32_t* outw = reinterpret_cast<uint32_t*>(out) + (size_t)(cg * ROUT + V * rg) * 4 + pp;
    const int chA = (ch0 + cg) * 8 + 2 * pp;
    const bool any_valid = (t0 + V - 1 >= 0) && (t0 < T) && (chA < C);
    // interior: both channels real, no replicate padding inside the window -> packed fp32x2 fast path
    const bool interior = (t0 - 5 >= 0) && (t0 + V + 4 <= T - 1) && (chA + 1 < C);
    if (interior) {
      uint32_t wd[V + 16];
#pragma unroll
      for (int j = 3; j < V + 13; ++j) wd[j] = inw[j * 4];
      const float a0 = pa0, b0 = pb0, a1 = pa1, b1 = pb1;
      act1d_window2<V>([&](int j) { return unpack_bf16x2(wd[j]); },
                       [&](int q, float ya, float yb) { outw[q * 4] = pack2(ya, yb); },
                       pk2(a0, a1), pk2(b0, b1));
    } else if (any_valid) {
      uint32_t wd[V + 16];
#pragma unroll
      for (int j = 0; j < V + 16; ++j) wd[j] = inw[j * 4];
      float ylo[V];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int ch = chA + h;
        float yv[V];
        if (ch < C) {
          float xw[V + 16];
#pragma unroll
          for (int j = 0; j < V + 16; ++j)
            xw[j] = h ? __uint_as_float(wd[j] & 0xffff0000u) : __uint_as_float(wd[j] << 16);
          act1d_window<V, false>(xw, yv, h ? pa1 : pa0, h ? pb1 : pb0, t0, (int64_t)T);
#pragma unroll
          for (int q = 0; q < V; ++q)
            if (t0 + q < 0 || t0 + q >= T) yv[q] = 0.f;           // zero halo rows
        } else {
#pragma unroll
          for (int q = 0; q < V; ++q) yv[q] = 0.f;                // zero padding channel
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
      for (int q = 0; q < V; ++q) outw[q * 4] = 0u;               // zero halo rows / padding channels
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> async-proxy reads
  __syncthreads();
  if (tid < nch) {
    const int nrows = min(TR, Tp - r0);
    bulk_s2g(y + (base + (int64_t)(ch0 + tid) * Tp + r0) * 8, out + tid * ROUT, (uint32_t)nrows * 16u);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

// activation_post output (c8t) -> conv_post (Cin -> 1, K taps, zero pad) -> tanh -> fp32 wav / int16 pcm
// (models.py:246-248; infer.py:206-212,234 for the pcm epilogue)
// One thread = four consecutive samples: the 4 + K - 1 input rows of a chunk are loaded once (16 bytes each) and every row
// feeds up to four outputs; the weights sit in shared memory as [chunk][k][8 channels] so that one (chunk, tap) is two
// 16-byte loads, and each packed fp32x2 FMA advances the even- and the odd-channel partial sum of an output (the
// one-sample-per-thread form was instruction bound at 1.3 TB/s: one LDS and two ALU ops per FMA).
constexpr int kPostMaxK = 7;
__global__ void __launch_bounds__(256)
conv_post_c8t_kernel(float* __restrict__ wav, int16_t* __restrict__ pcm, const __nv_bfloat16* __restrict__ x,
                     const float* __restrict__ w, const float* __restrict__ bias, int Cin, int chunks, int T, int Tp,
                     int pad, int K, int64_t s_lo, int64_t s_hi, const int* __restrict__ lens, int len_mul) {
  extern __shared__ __align__(16) float wsm[];   // [nchunk][K][8], zero for channels >= Cin
  const int nchunk = (Cin + 7) >> 3;
  for (int i = threadIdx.x; i < nchunk * K * 8; i += blockDim.x) {
    const int j = i & 7, k = (i >> 3) % K, ch = (i >> 3) / K;
    const int c = ch * 8 + j;
    wsm[i] = c < Cin ? w[c * K + k] : 0.f;
  }
  __syncthreads();
  const int b = blockIdx.y;
  const int64_t Tout = T - s_lo - s_hi;                             // row pitch of the output (longest utterance)
  const int64_t to0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (to0 >= Tout) return;
  const int64_t Tb = lens ? (int64_t)lens[b] * len_mul : T;         // ragged batch: samples past this utterance's end are zero
  const int64_t t0 = to0 + s_lo;
  const int hp = (K - 1) / 2;
  const float b0 = bias ? bias[0] : 0.f;
  f32x2 acc[4];
#pragma unroll
  for (int o = 0; o < 4; ++o) acc[o] = pk2(0.f, 0.f);
  for (int ch = 0; ch < nchunk; ++ch) {
    const __nv_bfloat16* xr = x + (((int64_t)b * chunks + ch) * Tp + pad + t0 - hp) * 8;   // halo rows are zero
    const float4* wk = reinterpret_cast<const float4*>(wsm + ch * K * 8);
#pragma unroll
    for (int r = 0; r < 4 + kPostMaxK - 1; ++r) {
      if (r >= 4 + K - 1) break;
      const uint4 v = *reinterpret_cast<const uint4*>(xr + (int64_t)r * 8);
      const f32x2 x01 = pk2(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u));
      const f32x2 x23 = pk2(__uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
      const f32x2 x45 = pk2(__uint_as_float(v.z << 16), __uint_as_float(v.z & 0xffff0000u));
      const f32x2 x67 = pk2(__uint_as_float(v.w << 16), __uint_as_float(v.w & 0xffff0000u));
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const int k = r - o;
        if (k < 0 || k >= K) continue;
        const float4 wa = wk[2 * k], wb = wk[2 * k + 1];
        acc[o] = fma2(pk2(wa.x, wa.y), x01, acc[o]);
        acc[o] = fma2(pk2(wa.z, wa.w), x23, acc[o]);
        acc[o] = fma2(pk2(wb.x, wb.y), x45, acc[o]);
        acc[o] = fma2(pk2(wb.z, wb.w), x67, acc[o]);
      }
    }
  }
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    const int64_t to = to0 + o;
    if (to >= Tout) break;
    float e, od;
    unpk2(acc[o], e, od);
    float yv = tanhf((e + od) + b0);
    if (to + s_lo >= Tb - s_hi) yv = 0.f;
    if (wav) wav[(int64_t)b * Tout + to] = yv;
    if (pcm) pcm[(int64_t)b * Tout + to] = (int16_t)fminf(fmaxf(32767.f * yv, -32767.f), 32767.f);
  }
}

template <int NCH>
int launch_act(const C8T& y, const C8T& x, const float* a, const float* b_, int64_t B, cudaStream_t st) {
  constexpr int TR = 1024 / NCH;
  const size_t smem = (size_t)NCH * (TR + 17) * 16 + (size_t)NCH * (TR + 1) * 16 + 16;
  static std::atomic<uint64_t> opted{0};
  BVG_TRY(smem_opt_in(act1d_c8t_kernel<NCH>, opted, (int)smem));
  dim3 grid((unsigned)((x.Tp + TR - 1) / TR), (unsigned)((x.chunks + NCH - 1) / NCH), (unsigned)B);
  ProfScope prof(st, KC_ACT1D);
  act1d_c8t_kernel<NCH><<<grid, 256, smem, st>>>(y.p, x.p, a, b_, x.C, x.chunks, x.T, x.Tp, x.pad, x.lens, x.len_mul);
  BVG_LAUNCHED();
  return BVG_OK;
}

}  // namespace

int act1d_c8t_launch(const C8T& y, const C8T& x, const float* alpha_log, const float* beta_log, int64_t B,
                     cudaStream_t st, int impl) {
  BVG_CHECK_ARG(y.p && x.p && y.p != x.p, "act1d_c8t: bad buffers");
  BVG_CHECK_ARG(y.C == x.C && y.T == x.T && y.chunks == x.chunks && y.pad == x.pad, "act1d_c8t: geometry mismatch");
  BVG_CHECK_ARG(B >= 1 && B <= 65535, "act1d_c8t: bad batch");
  // tensor-core FIRs (act1d_tc.cu) for everything but short tensors; BVG_ACT_TC=0 keeps the CUDA-core stencil (A/B runs)
  // (small problems -- a single utterance -- stay on the stencil kernel: the persistent tensor-core kernel needs ~25 us
  // whatever the size, the stencil 13-22 us for one 10 s utterance; profiles/r02_act1d_tc_shapes.txt)
  const bool big = B * (int64_t)x.C * x.T >= (int64_t)tc_min_melems() * 1000000;
  if (impl == 2 || (impl == 0 && big && BVG_ENV_ONCE("BVG_ACT_TC", 1))) {
    const int rc = act1d_tc_launch(y, x, alpha_log, beta_log, B, st);
    if (rc != BVG_ERR_STATE) return rc;
    BVG_CHECK_ARG(impl != 2, "act1d_c8t: this tensor does not qualify for the tensor-core kernel (T >= 256 required)");
  }
  if (x.chunks % 8 == 0 || x.chunks > 16) return launch_act<8>(y, x, alpha_log, beta_log, B, st);
  if (x.chunks % 4 == 0) return launch_act<4>(y, x, alpha_log, beta_log, B, st);
  return launch_act<2>(y, x, alpha_log, beta_log, B, st);
}

int conv_post_c8t_launch(float* wav, int16_t* pcm, const C8T& x, const float* w, const float* bias, int K,
                         int64_t s_lo, int64_t s_hi, int64_t B, cudaStream_t st) {
  BVG_CHECK_ARG(wav || pcm, "conv_post: no output buffer");
  BVG_CHECK_ARG((K - 1) / 2 <= x.pad && x.C * K * 4 <= 48 * 1024, "conv_post: unsupported shape");
  const int64_t Tout = x.T - s_lo - s_hi;
  BVG_CHECK_ARG(Tout >= 0 && s_lo >= 0 && s_hi >= 0, "conv_post: bad crop");
  if (Tout == 0) return BVG_OK;
  BVG_CHECK_ARG(K <= kPostMaxK, "conv_post: at most %d taps", kPostMaxK);
  dim3 grid((unsigned)((Tout + 1023) / 1024), (unsigned)B);
  ProfScope prof(st, KC_OTHER);
  conv_post_c8t_kernel<<<grid, 256, (size_t)((x.C + 7) / 8) * 8 * K * 4, st>>>(wav, pcm, x.p, w, bias, x.C, x.chunks, x.T, x.Tp, x.pad, K,
                                                              s_lo, s_hi, x.lens, x.len_mul);
  BVG_LAUNCHED();
  return BVG_OK;
}

}  // namespace bvg
