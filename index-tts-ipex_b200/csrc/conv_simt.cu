// CUDA-core (fp32 FMA) dense Conv1d / ConvTranspose1d for sm_100a.
//
// This is the fp32 parity path (BASELINE config 2: max-abs <= 1e-4 vs the reference fp32
// forward needs true fp32 products; single-pass TF32/bf16 tensor-core operands do not hold that
// gate) and the path for layers too small for a tensor-core tile (ECAPA,
// cond matvecs).  The bf16 throughput path lives in conv_umma.cu (tcgen05/TMEM).
//
// Restates torch.nn.Conv1d / ConvTranspose1d as used by models.py:25-42,149,155-161,184 and
// nnet/CNN.py:411-456 (reflect "same" padding, :458-488).
//
// Tiling: one CTA = 128 output channels x 128 time steps of one batch element, 256 threads,
// 8x8 accumulators per thread.  Input channels are consumed 16 at a time: the x tile (with the
// (K-1)*dil halo) is staged once per chunk, the [16 x 128] weight slab once per tap.  Threads
// own time steps tx+16j, so x reads are conflict-free scalar LDS and global stores coalesce.
#include "bvg_common.cuh"

namespace bvg {
namespace {

constexpr int kTT = 128;     // time steps per CTA
constexpr int kTC = 128;     // output channels per CTA
constexpr int kCK = 16;      // input channels per chunk
constexpr int kMaxHalo = 64; // (K-1)*dil <= 64  (generator max: (11-1)*5 = 50)

struct EpiDev {
  const float* bias;
  const void* res1;
  const void* res2;
  float scale;
  const float* cond;
  int64_t cond_B;
  int relu;
  const float* post_scale;
  const float* post_shift;
  int act;
};

__device__ __forceinline__ float epilogue_apply(float v, const EpiDev& ep, int64_t b, int64_t co, int64_t Cout) {
  if (ep.bias) v += ep.bias[co];
  if (ep.cond) v += ep.cond[(ep.cond_B == 1 ? 0 : b) * Cout + co];
  if (ep.relu) v = fmaxf(v, 0.f);
  if (ep.post_scale) v = fmaf(v, ep.post_scale[co], ep.post_shift[co]);
  return v;
}
__device__ __forceinline__ float final_act(float v, int act) {
  if (act == 1) return tanhf(v);
  if (act == 2) return 1.f / (1.f + expf(-v));
  return v;
}

// NI x NJ accumulators per thread (output channels x time steps): tile = 16*NI channels x 16*NJ time steps.
// 8 x 8 is the 128 x 128 tile above; 4 x 4 a 64 x 64 tile for the small layers of the speaker encoder (Res2Net
// 64 -> 64 convs over 281 frames: four times as many CTAs, a quarter of the serial work); 6 / 3 / 2 x 8 fit the
// narrow generator stages (C = 96 / 48 / 24) without computing padding channels.
template <typename TI, typename TO, int NI, int NJ>
__global__ void __launch_bounds__(256)
conv1d_simt_kernel(TO* __restrict__ dst, int64_t dsb, const TI* __restrict__ src, const TI* __restrict__ src2,
                   int64_t sb, int64_t sc, int64_t st_, const float* __restrict__ w, EpiDev ep,
                   int64_t Cin, int64_t Cout, int64_t T, int K, int dil, int pad_mode) {
  constexpr int kTT = 16 * NJ, kTC = 16 * NI, kXW = kTT + kMaxHalo;
  __shared__ float xs[kCK][kXW];
  // two weight slabs: the [16 x TC] slab of the next tap streams in with cp.async while this tap's FMAs run
  __shared__ __align__(16) float ws[2][kCK][kTC];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t t0 = (int64_t)blockIdx.x * kTT;
  const int64_t co0 = (int64_t)blockIdx.y * kTC;
  const int64_t b = blockIdx.z;
  const int pad = dil * (K - 1) / 2;
  const int xw = kTT + dil * (K - 1);

  float acc[NI][NJ];
#pragma unroll
  for (int i = 0; i < NI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;

  const TI* sbase = src + b * sb;
  const TI* sbase2 = src2 ? src2 + b * sb : nullptr;

  // weight slab of iteration (ci0, k) -> ws[buf]: 16-byte cp.async per thread, zero-filled outside [Cin) x [Cout)
  // (Cout % 4 == 0 and a 16-byte aligned weight pointer: whole 4-float groups are in or out); scalar loads otherwise
  const bool vec_w = (Cout % 4 == 0) && ((reinterpret_cast<uintptr_t>(w) & 15) == 0);
  auto issue_w = [&](int64_t ci0, int k, int buf) {
    const float* wk = w + ((int64_t)k * Cin + ci0) * Cout + co0;
    if (vec_w) {
      for (int idx = tid; idx < kCK * (kTC / 4); idx += 256) {
        const int ci = idx / (kTC / 4), c = (idx % (kTC / 4)) * 4;
        const bool in = (ci0 + ci < Cin) && (co0 + c < Cout);
        const float* srcp = in ? wk + (int64_t)ci * Cout + c : w;
        const uint32_t dsts = static_cast<uint32_t>(__cvta_generic_to_shared(&ws[buf][ci][c]));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dsts), "l"(srcp), "r"(in ? 16 : 0) : "memory");
      }
    } else {
      for (int idx = tid; idx < kCK * kTC; idx += 256) {
        const int ci = idx / kTC, c = idx % kTC;
        float v = 0.f;
        if (ci0 + ci < Cin && co0 + c < Cout) v = wk[(int64_t)ci * Cout + c];
        ws[buf][ci][c] = v;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int buf = 0;
  issue_w(0, 0, 0);
  for (int64_t ci0 = 0; ci0 < Cin; ci0 += kCK) {
    for (int k = 0; k < K; ++k) {
      __syncthreads();   // everyone is done with the previous iteration: xs (at k == 0) and ws[buf ^ 1] may be overwritten
      const bool last = (k + 1 == K) && (ci0 + kCK >= Cin);
      if (!last) issue_w(k + 1 < K ? ci0 : ci0 + kCK, k + 1 < K ? k + 1 : 0, buf ^ 1);
      if (k == 0) {
        for (int idx = tid; idx < kCK * xw; idx += 256) {
          const int ci = idx / xw, p = idx - ci * xw;
          int64_t t = t0 - pad + p;
          float v = 0.f;
          if (ci0 + ci < Cin) {
            if (pad_mode == 1) {  // reflect
              if (t < 0) t = -t;
              if (t >= T) t = 2 * (T - 1) - t;
              if (t < 0) t = 0;   // (only reachable when T <= pad; guarded on the host)
            }
            if (t >= 0 && t < T) {
              const int64_t off = (ci0 + ci) * sc + t * st_;
              v = to_f<TI>(sbase[off]);
              if (sbase2) v += to_f<TI>(sbase2[off]);
            }
          }
          xs[ci][p] = v;
        }
      }
      if (last) asm volatile("cp.async.wait_group 0;" ::: "memory");
      else asm volatile("cp.async.wait_group 1;" ::: "memory");
      __syncthreads();   // this iteration's slab (and the x tile) are visible to all threads
      const int shift = k * dil + tx;
#pragma unroll 4
      for (int ci = 0; ci < kCK; ++ci) {
        float wv[NI];
        if (NI % 4 == 0) {
#pragma unroll
          for (int i = 0; i + 3 < NI; i += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(&ws[buf][ci][ty * NI + i]);
            wv[i] = w4.x; wv[i + 1] = w4.y; wv[i + 2] = w4.z; wv[i + 3] = w4.w;
          }
        } else if (NI % 2 == 0) {
#pragma unroll
          for (int i = 0; i + 1 < NI; i += 2) {
            const float2 w2 = *reinterpret_cast<const float2*>(&ws[buf][ci][ty * NI + i]);
            wv[i] = w2.x; wv[i + 1] = w2.y;
          }
        } else {
#pragma unroll
          for (int i = 0; i < NI; ++i) wv[i] = ws[buf][ci][ty * NI + i];
        }
        float xv[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) xv[j] = xs[ci][shift + 16 * j];
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
      }
      buf ^= 1;
    }
  }

  TO* dbase = dst + b * dsb;
  const TO* r1 = ep.res1 ? static_cast<const TO*>(ep.res1) + b * dsb : nullptr;
  const TO* r2 = ep.res2 ? static_cast<const TO*>(ep.res2) + b * dsb : nullptr;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int64_t co = co0 + ty * NI + i;
    if (co >= Cout) continue;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int64_t t = t0 + tx + 16 * j;
      if (t >= T) continue;
      float v = epilogue_apply(acc[i][j], ep, b, co, Cout);
      const int64_t o = co * T + t;
      if (r1) v += to_f<TO>(r1[o]);
      if (r2) v += to_f<TO>(r2[o]);
      v *= ep.scale;
      dbase[o] = from_f<TO>(final_act(v, ep.act));
    }
  }
}

// ConvTranspose1d: y[co, u*i - p + k] += x[ci, i] * w[ci, co, k],  p = (K-u)/2  (models.py:155-161)
// Output-stationary: for output t only taps k == (t+p) mod u contribute.  Threads own t = tx+16j
// and u | 16, so a thread's active taps are the same for all its j.
template <typename TT>
__global__ void __launch_bounds__(256)
convtr1d_simt_kernel(TT* __restrict__ dst, const TT* __restrict__ src, const float* __restrict__ w, EpiDev ep,
                     int64_t Cin, int64_t Cout, int64_t Tin, int K, int u) {
  constexpr int kXI = 96;   // input steps staged: 128/u + K/u + 2 <= 96 for u >= 2
  __shared__ float xs[kCK][kXI];
  __shared__ __align__(16) float ws[kCK][kTC];

  constexpr int NT = 8;     // 8 x 8 accumulators per thread (128 x 128 tile)
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t Tout = Tin * u;
  const int64_t t0 = (int64_t)blockIdx.x * kTT;
  const int64_t co0 = (int64_t)blockIdx.y * kTC;
  const int64_t b = blockIdx.z;
  const int p = (K - u) / 2;
  // input index range touched by this tile
  int64_t n_lo = t0 + p - (K - 1);
  const int64_t i_lo = (n_lo >= 0) ? n_lo / u : -((-n_lo + u - 1) / u);
  const int64_t i_hi = (t0 + kTT - 1 + p) / u;
  const int ni = (int)(i_hi - i_lo + 1);

  float acc[NT][NT];
#pragma unroll
  for (int i = 0; i < NT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j] = 0.f;

  const TT* sbase = src + b * Cin * Tin;
  for (int64_t ci0 = 0; ci0 < Cin; ci0 += kCK) {
    __syncthreads();
    for (int idx = tid; idx < kCK * ni; idx += 256) {
      const int ci = idx / ni, q = idx - ci * ni;
      const int64_t i = i_lo + q;
      float v = 0.f;
      if (ci0 + ci < Cin && i >= 0 && i < Tin) v = to_f<TT>(sbase[(ci0 + ci) * Tin + i]);
      xs[ci][q] = v;
    }
    for (int k = 0; k < K; ++k) {
      if (k) __syncthreads();
      const float* wk = w + ((int64_t)k * Cin + ci0) * Cout + co0;
      for (int idx = tid; idx < kCK * kTC; idx += 256) {
        const int ci = idx / kTC, c = idx % kTC;
        float v = 0.f;
        if (ci0 + ci < Cin && co0 + c < Cout) v = wk[(int64_t)ci * Cout + c];
        ws[ci][c] = v;
      }
      __syncthreads();
      const int64_t n0 = t0 + tx + p - k;           // n_j = n0 + 16 j
      const int64_t r = ((n0 % u) + u) % u;
      if (r == 0) {
        int q[8];
        bool ok[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int64_t n = n0 + 16 * j;
          const int64_t i = (n >= 0) ? n / u : -1;
          ok[j] = (n >= 0) && (i < Tin);
          q[j] = ok[j] ? (int)(i - i_lo) : 0;
        }
#pragma unroll 4
        for (int ci = 0; ci < kCK; ++ci) {
          const float4 w0 = *reinterpret_cast<const float4*>(&ws[ci][ty * 8]);
          const float4 w1 = *reinterpret_cast<const float4*>(&ws[ci][ty * 8 + 4]);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          float xv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) xv[j] = ok[j] ? xs[ci][q[j]] : 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
        }
      }
    }
  }
  TT* dbase = dst + b * Cout * Tout;
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const int64_t co = co0 + ty * NT + i;
    if (co >= Cout) continue;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int64_t t = t0 + tx + 16 * j;
      if (t >= Tout) continue;
      float v = epilogue_apply(acc[i][j], ep, b, co, Cout);
      dbase[co * Tout + t] = from_f<TT>(v * ep.scale);
    }
  }
}

// Same op with all lanes busy: a thread owns 8/U coarse positions q = tx + 16*jq and ALL U output phases of each
// (t = t0 + U*q + ph).  Tap k only feeds phase ph = (k - p) mod U, so with the phase loop outside (compile-time
// accumulator index) every tap is a plain strided-free FMA pass for the whole warp, where the kernel above keeps
// 1/U of the lanes busy per tap.  U = stride in {2, 4} (the generator's), 128 x 128 tile, same weight layout.
template <typename TT, int U>
__global__ void __launch_bounds__(256)
convtr1d_simt_phase_kernel(TT* __restrict__ dst, const TT* __restrict__ src, const float* __restrict__ w, EpiDev ep,
                           int64_t Cin, int64_t Cout, int64_t Tin, int K) {
  constexpr int kXI = 96;   // input steps staged: 128/U + K/U + 2 <= 96
  constexpr int NT = 8, NQ = NT / U;
  __shared__ float xs[kCK][kXI];
  __shared__ __align__(16) float ws[kCK][kTC];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t Tout = Tin * U;
  const int64_t t0 = (int64_t)blockIdx.x * kTT;
  const int64_t co0 = (int64_t)blockIdx.y * kTC;
  const int64_t b = blockIdx.z;
  const int p = (K - U) / 2;
  // input index range touched by this tile
  const int64_t n_lo = t0 + p - (K - 1);
  const int64_t i_lo = (n_lo >= 0) ? n_lo / U : -((-n_lo + U - 1) / U);
  const int64_t i_hi = (t0 + kTT - 1 + p) / U;
  const int ni = (int)(i_hi - i_lo + 1);

  float acc[NT][NT];        // acc[i][jq * U + ph]
#pragma unroll
  for (int i = 0; i < NT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j] = 0.f;

  const TT* sbase = src + b * Cin * Tin;
  for (int64_t ci0 = 0; ci0 < Cin; ci0 += kCK) {
    __syncthreads();
    for (int idx = tid; idx < kCK * ni; idx += 256) {
      const int ci = idx / ni, q = idx - ci * ni;
      const int64_t i = i_lo + q;
      float v = 0.f;
      if (ci0 + ci < Cin && i >= 0 && i < Tin) v = to_f<TT>(sbase[(ci0 + ci) * Tin + i]);
      xs[ci][q] = v;
    }
    bool first = true;
#pragma unroll
    for (int ph = 0; ph < U; ++ph) {
      for (int k = (ph + p) % U; k < K; k += U) {
        if (!first) __syncthreads();
        first = false;
        const float* wk = w + ((int64_t)k * Cin + ci0) * Cout + co0;
        for (int idx = tid; idx < kCK * kTC; idx += 256) {
          const int ci = idx / kTC, c = idx % kTC;
          float v = 0.f;
          if (ci0 + ci < Cin && co0 + c < Cout) v = wk[(int64_t)ci * Cout + c];
          ws[ci][c] = v;
        }
        __syncthreads();
        // input index of coarse position q for this tap: i = t0/U + q + (ph + p - k)/U  (exact division)
        const int64_t ib = t0 / U + (ph + p - k) / U;
        int qi[NQ];
        bool ok[NQ];
#pragma unroll
        for (int jq = 0; jq < NQ; ++jq) {
          const int64_t i = ib + tx + 16 * jq;
          ok[jq] = (i >= 0) && (i < Tin);
          qi[jq] = ok[jq] ? (int)(i - i_lo) : 0;
        }
#pragma unroll 4
        for (int ci = 0; ci < kCK; ++ci) {
          const float4 w0 = *reinterpret_cast<const float4*>(&ws[ci][ty * 8]);
          const float4 w1 = *reinterpret_cast<const float4*>(&ws[ci][ty * 8 + 4]);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          float xv[NQ];
#pragma unroll
          for (int jq = 0; jq < NQ; ++jq) xv[jq] = ok[jq] ? xs[ci][qi[jq]] : 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int jq = 0; jq < NQ; ++jq) acc[i][jq * U + ph] = fmaf(wv[i], xv[jq], acc[i][jq * U + ph]);
        }
      }
    }
  }
  TT* dbase = dst + b * Cout * Tout;
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const int64_t co = co0 + ty * NT + i;
    if (co >= Cout) continue;
#pragma unroll
    for (int jq = 0; jq < NQ; ++jq)
#pragma unroll
      for (int ph = 0; ph < U; ++ph) {
        const int64_t t = t0 + (int64_t)U * (tx + 16 * jq) + ph;
        if (t >= Tout) continue;
        float v = epilogue_apply(acc[i][jq * U + ph], ep, b, co, Cout);
        dbase[co * Tout + t] = from_f<TT>(v * ep.scale);
      }
  }
}

__global__ void repack_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t Cout, int64_t Cin,
                              int K, int transposed) {
  // dst[k][ci][co]
  const int64_t n = Cout * Cin * K;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t co = i % Cout;
    const int64_t ci = (i / Cout) % Cin;
    const int64_t k = i / (Cout * Cin);
    dst[i] = transposed ? src[(ci * Cout + co) * K + k] : src[(co * Cin + ci) * K + k];
  }
}

EpiDev to_dev(const ConvEpilogue& e) {
  EpiDev d;
  d.bias = e.bias; d.res1 = e.res1; d.res2 = e.res2; d.scale = e.scale; d.cond = e.cond;
  d.cond_B = e.cond_B; d.relu = e.relu; d.post_scale = e.post_scale; d.post_shift = e.post_shift;
  d.act = e.act;
  return d;
}

}  // namespace

int conv1d_simt_launch(void* dst, int64_t dsb, const void* src, const void* src2, int64_t sb, int64_t sc,
                       int64_t st_, const float* weight_kic, const ConvEpilogue& ep, int64_t B, int64_t Cin,
                       int64_t Cout, int64_t T, int K, int dil, int pad_mode, int in_dtype, int out_dtype,
                       cudaStream_t st) {
  BVG_CHECK_ARG(dst && src && weight_kic, "conv1d: null pointer");
  BVG_CHECK_ARG(K >= 1 && (K & 1) && dil >= 1 && (K - 1) * dil <= kMaxHalo,
                "conv1d: unsupported kernel K=%d dilation=%d (need odd K, (K-1)*dil <= %d)", K, dil, kMaxHalo);
  BVG_CHECK_ARG(B >= 0 && Cin > 0 && Cout > 0 && T >= 0, "conv1d: bad shape");
  if (pad_mode == 1)
    BVG_CHECK_ARG(T > dil * (K - 1) / 2, "conv1d: reflect padding needs T > pad (T=%lld pad=%d)", (long long)T,
                  dil * (K - 1) / 2);
  if (B == 0 || T == 0) return BVG_OK;
  // tile choice (fp32 path): channel-exact tiles for the narrow generator stages, 64 x 64 tiles for small problems
  // (few output channels, or too few 128 x 128 tiles to fill the GPU), 128 x 128 otherwise
  const bool f32 = in_dtype == BVG_F32 && out_dtype == BVG_F32;
  const int64_t tiles128 = ((T + 127) / 128) * ((Cout + 127) / 128) * B;
  int ni = 8, nj = 8;
  if (f32 && T >= 4096 && (Cout == 96 || Cout == 48 || (Cout > 16 && Cout <= 32))) ni = (int)(Cout + 15) / 16;
  else if (f32 && (Cout <= 64 || tiles128 < 148)) ni = nj = 4;
  const int tile_c = 16 * ni, tile_t = 16 * nj;
  BVG_CHECK_ARG(B <= 65535 && (Cout + tile_c - 1) / tile_c <= 65535, "conv1d: batch/channel grid too large");
  dim3 grid((unsigned)((T + tile_t - 1) / tile_t), (unsigned)((Cout + tile_c - 1) / tile_c), (unsigned)B);
  EpiDev e = to_dev(ep);
  ProfScope prof(st, (T >= 64 && !ep.prof_other) ? KC_CONV : KC_OTHER);
#define BVG_CONV_F32(NI_, NJ_)                                                                                        \
  conv1d_simt_kernel<float, float, NI_, NJ_><<<grid, 256, 0, st>>>((float*)dst, dsb, (const float*)src,                \
                                                                   (const float*)src2, sb, sc, st_, weight_kic, e, Cin, \
                                                                   Cout, T, K, dil, pad_mode)
  if (f32 && ni == 6) BVG_CONV_F32(6, 8);
  else if (f32 && ni == 3) BVG_CONV_F32(3, 8);
  else if (f32 && ni == 2) BVG_CONV_F32(2, 8);
  else if (f32 && ni == 4) BVG_CONV_F32(4, 4);
  else if (f32) BVG_CONV_F32(8, 8);
#undef BVG_CONV_F32
  else if (in_dtype == BVG_BF16 && out_dtype == BVG_BF16)
    conv1d_simt_kernel<__nv_bfloat16, __nv_bfloat16, 8, 8><<<grid, 256, 0, st>>>(
        (__nv_bfloat16*)dst, dsb, (const __nv_bfloat16*)src, (const __nv_bfloat16*)src2, sb, sc, st_, weight_kic, e,
        Cin, Cout, T, K, dil, pad_mode);
  else if (in_dtype == BVG_F32 && out_dtype == BVG_BF16)
    conv1d_simt_kernel<float, __nv_bfloat16, 8, 8><<<grid, 256, 0, st>>>((__nv_bfloat16*)dst, dsb, (const float*)src,
                                                                   (const float*)src2, sb, sc, st_, weight_kic, e,
                                                                   Cin, Cout, T, K, dil, pad_mode);
  else {
    set_error("conv1d: unsupported dtype pair (%d -> %d)", in_dtype, out_dtype);
    return BVG_ERR_INVALID;
  }
  BVG_LAUNCHED();
  return BVG_OK;
}

int convtr1d_simt_launch(void* dst, const void* src, const float* weight_kic, const ConvEpilogue& ep, int64_t B,
                         int64_t Cin, int64_t Cout, int64_t Tin, int K, int stride, int dtype, cudaStream_t st) {
  BVG_CHECK_ARG(dst && src && weight_kic, "convtr1d: null pointer");
  BVG_CHECK_ARG(stride >= 2 && 16 % stride == 0 && K >= stride && K <= 4 * stride && (K - stride) % 2 == 0,
                "convtr1d: unsupported stride=%d K=%d (need stride in {2,4,8,16}, stride<=K<=4*stride, K-stride even)",
                stride, K);
  BVG_CHECK_ARG(B >= 0 && Cin > 0 && Cout > 0 && Tin >= 0, "convtr1d: bad shape");
  if (B == 0 || Tin == 0) return BVG_OK;
  const int64_t Tout = Tin * stride;
  BVG_CHECK_ARG(B <= 65535, "convtr1d: batch too large");
  dim3 grid((unsigned)((Tout + kTT - 1) / kTT), (unsigned)((Cout + kTC - 1) / kTC), (unsigned)B);
  EpiDev e = to_dev(ep);
  ProfScope prof(st, KC_CONVTR);
  if (dtype == BVG_F32 && stride == 4)
    convtr1d_simt_phase_kernel<float, 4><<<grid, 256, 0, st>>>((float*)dst, (const float*)src, weight_kic, e, Cin, Cout, Tin, K);
  else if (dtype == BVG_F32 && stride == 2)
    convtr1d_simt_phase_kernel<float, 2><<<grid, 256, 0, st>>>((float*)dst, (const float*)src, weight_kic, e, Cin, Cout, Tin, K);
  else if (dtype == BVG_BF16 && stride == 4)
    convtr1d_simt_phase_kernel<__nv_bfloat16, 4><<<grid, 256, 0, st>>>((__nv_bfloat16*)dst, (const __nv_bfloat16*)src,
                                                                       weight_kic, e, Cin, Cout, Tin, K);
  else if (dtype == BVG_BF16 && stride == 2)
    convtr1d_simt_phase_kernel<__nv_bfloat16, 2><<<grid, 256, 0, st>>>((__nv_bfloat16*)dst, (const __nv_bfloat16*)src,
                                                                       weight_kic, e, Cin, Cout, Tin, K);
  else if (dtype == BVG_F32)
    convtr1d_simt_kernel<float><<<grid, 256, 0, st>>>((float*)dst, (const float*)src, weight_kic, e, Cin, Cout, Tin, K,
                                                      stride);
  else if (dtype == BVG_BF16)
    convtr1d_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((__nv_bfloat16*)dst, (const __nv_bfloat16*)src,
                                                              weight_kic, e, Cin, Cout, Tin, K, stride);
  else {
    set_error("convtr1d: unsupported dtype %d", dtype);
    return BVG_ERR_INVALID;
  }
  BVG_LAUNCHED();
  return BVG_OK;
}

int repack_conv_weight_launch(float* dst, const float* src, int64_t Cout, int64_t Cin, int K, int transposed,
                              cudaStream_t st) {
  const int64_t n = Cout * Cin * K;
  if (n == 0) return BVG_OK;
  int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  repack_kernel<<<blocks, 256, 0, st>>>(dst, src, Cout, Cin, K, transposed);
  BVG_LAUNCHED();
  return BVG_OK;
}

}  // namespace bvg
