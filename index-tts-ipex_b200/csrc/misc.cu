// Small reduction / elementwise kernels of the decode path (speaker encoder glue, conv_post).
// References: ECAPA_TDNN.py:228-242 (SEBlock), :282-338 (AttentiveStatisticsPooling),
// models.py:246-248 (activation_post -> conv_post -> tanh), infer.py:206-212 (int16 epilogue).
#include "bvg_common.cuh"

namespace bvg {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// one warp per row
__global__ void row_mean_kernel(float* __restrict__ out, const float* __restrict__ x, int64_t rows, int64_t T) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* p = x + r * T;
  float s = 0.f;
  for (int64_t t = lane; t < T; t += 32) s += p[t];
  s = warp_sum(s);
  if (lane == 0) out[r] = s / (float)T;
}

__global__ void row_stats_kernel(float* __restrict__ ms, const float* __restrict__ x, int64_t B, int64_t C, int64_t T) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= B * C) return;
  const int lane = threadIdx.x & 31;
  const float* p = x + r * T;
  float s = 0.f;
  for (int64_t t = lane; t < T; t += 32) s += p[t];
  const float mean = warp_sum(s) / (float)T;
  float v = 0.f;
  for (int64_t t = lane; t < T; t += 32) { const float d = p[t] - mean; v = fmaf(d, d, v); }
  v = warp_sum(v) / (float)T;
  if (lane == 0) {
    const int64_t b = r / C, c = r % C;
    ms[b * 2 * C + c] = mean;
    ms[b * 2 * C + C + c] = sqrtf(fmaxf(v, 1e-12f));
  }
}

__global__ void scale_residual_kernel(float* __restrict__ out, int64_t osb, const float* __restrict__ s,
                                      const float* __restrict__ y, const float* __restrict__ res, int64_t rsb,
                                      int64_t B, int64_t C, int64_t T) {
  const int64_t n = B * C * T;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i % T, c = (i / T) % C, b = i / (T * C);
    out[b * osb + c * T + t] = fmaf(s[b * C + c], y[i], res[b * rsb + c * T + t]);
  }
}

// one warp per (b,c) row
__global__ void attn_stats_kernel(float* __restrict__ pooled, const float* __restrict__ logits,
                                  const float* __restrict__ x, const float* __restrict__ bn_scale,
                                  const float* __restrict__ bn_shift, int64_t B, int64_t C, int64_t T) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= B * C) return;
  const int lane = threadIdx.x & 31;
  const float* l = logits + r * T;
  const float* p = x + r * T;
  float m = -INFINITY;
  for (int64_t t = lane; t < T; t += 32) m = fmaxf(m, l[t]);
  m = warp_max(m);
  float z = 0.f, sx = 0.f;
  for (int64_t t = lane; t < T; t += 32) { const float e = expf(l[t] - m); z += e; sx = fmaf(e, p[t], sx); }
  z = warp_sum(z);
  const float mean = warp_sum(sx) / z;
  float v = 0.f;
  for (int64_t t = lane; t < T; t += 32) { const float e = expf(l[t] - m); const float d = p[t] - mean; v = fmaf(e, d * d, v); }
  v = warp_sum(v) / z;
  if (lane == 0) {
    const int64_t b = r / C, c = r % C;
    pooled[b * 2 * C + c] = fmaf(mean, bn_scale[c], bn_shift[c]);
    pooled[b * 2 * C + C + c] = fmaf(sqrtf(fmaxf(v, 1e-12f)), bn_scale[C + c], bn_shift[C + c]);
  }
}

// y[b, co] = epilogue(sum_ci w[ci][co] * x[b, ci]): the T == 1 "convs" of the path (cond vectors, SE block,
// attentive-statistics context, final fc).  256 threads = 64 outputs x 4 input slices; weights are read
// coalesced along co from the [Cin][Cout] pack, partial sums meet in shared memory.
__global__ void __launch_bounds__(256)
matvec_kernel(float* __restrict__ y, const float* __restrict__ x, const float* __restrict__ w,
              const float* __restrict__ bias, int relu, const float* __restrict__ post_scale,
              const float* __restrict__ post_shift, int act, int Cin, int Cout) {
  __shared__ float part[4][64];
  const int co = blockIdx.x * 64 + (threadIdx.x & 63);
  const int sl = threadIdx.x >> 6;
  const int b = blockIdx.y;
  const float* xb = x + (int64_t)b * Cin;
  float acc = 0.f;
  if (co < Cout) {
    const int per = (Cin + 3) / 4;
    const int lo = sl * per, hi = min(Cin, lo + per);
#pragma unroll 4
    for (int ci = lo; ci < hi; ++ci) acc = fmaf(w[(int64_t)ci * Cout + co], xb[ci], acc);
  }
  part[sl][threadIdx.x & 63] = acc;
  __syncthreads();
  if (sl == 0 && co < Cout) {
    float v = part[0][threadIdx.x] + part[1][threadIdx.x] + part[2][threadIdx.x] + part[3][threadIdx.x];
    if (bias) v += bias[co];
    if (relu) v = fmaxf(v, 0.f);
    if (post_scale) v = fmaf(v, post_scale[co], post_shift[co]);
    if (act == 1) v = tanhf(v);
    if (act == 2) v = 1.f / (1.f + expf(-v));
    y[(int64_t)b * Cout + co] = v;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
conv_post_kernel(float* __restrict__ wav, int16_t* __restrict__ pcm, const T* __restrict__ x,
                 const float* __restrict__ w, const float* __restrict__ bias, int64_t Cin, int64_t Tlen, int K,
                 int64_t s_lo, int64_t s_hi) {
  extern __shared__ float wsm[];   // [Cin*K]
  for (int i = threadIdx.x; i < Cin * K; i += blockDim.x) wsm[i] = w[i];
  __syncthreads();
  const int64_t b = blockIdx.y;
  const int64_t Tout = Tlen - s_lo - s_hi;
  const int64_t to = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (to >= Tout) return;
  const int64_t t = to + s_lo;
  const int pad = (K - 1) / 2;
  const T* xb = x + b * Cin * Tlen;
  float acc = bias ? bias[0] : 0.f;
  for (int64_t ci = 0; ci < Cin; ++ci) {
    const T* xr = xb + ci * Tlen;
    for (int k = 0; k < K; ++k) {
      const int64_t tt = t + k - pad;
      if (tt >= 0 && tt < Tlen) acc = fmaf(wsm[ci * K + k], to_f<T>(xr[tt]), acc);
    }
  }
  const float y = tanhf(acc);
  if (wav) wav[b * Tout + to] = y;
  if (pcm) {
    // infer.py:207 torch.clamp(32767 * wav, -32767, 32767); :234 .type(torch.int16) truncates
    const float s = fminf(fmaxf(32767.f * y, -32767.f), 32767.f);
    pcm[b * Tout + to] = (int16_t)s;
  }
}

}  // namespace

int matvec_launch(float* y, const float* x, const float* w_ic, const ConvEpilogue& ep, int64_t B, int Cin, int Cout,
                  cudaStream_t st) {
  BVG_CHECK_ARG(y && x && w_ic && Cin > 0 && Cout > 0, "matvec: bad argument");
  BVG_CHECK_ARG(!ep.res1 && !ep.res2 && !ep.cond && ep.scale == 1.f, "matvec: unsupported epilogue");
  if (B == 0) return BVG_OK;
  BVG_CHECK_ARG(B <= 65535, "matvec: batch too large");
  dim3 grid((unsigned)((Cout + 63) / 64), (unsigned)B);
  ProfScope prof(st, KC_OTHER);
  matvec_kernel<<<grid, 256, 0, st>>>(y, x, w_ic, ep.bias, ep.relu, ep.post_scale, ep.post_shift, ep.act, Cin, Cout);
  BVG_LAUNCHED();
  return BVG_OK;
}

int row_mean_launch(float* out, const float* x, int64_t rows, int64_t T, cudaStream_t st) {
  if (rows == 0) return BVG_OK;
  ProfScope prof(st, KC_OTHER);
  row_mean_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, st>>>(out, x, rows, T);
  BVG_LAUNCHED();
  return BVG_OK;
}
int row_stats_launch(float* ms, const float* x, int64_t B, int64_t C, int64_t T, cudaStream_t st) {
  if (B * C == 0) return BVG_OK;
  ProfScope prof(st, KC_OTHER);
  row_stats_kernel<<<(unsigned)((B * C + 3) / 4), 128, 0, st>>>(ms, x, B, C, T);
  BVG_LAUNCHED();
  return BVG_OK;
}
int scale_residual_launch(float* out, int64_t osb, const float* s, const float* y, const float* res, int64_t rsb,
                          int64_t B, int64_t C, int64_t T, cudaStream_t st) {
  const int64_t n = B * C * T;
  if (n == 0) return BVG_OK;
  int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
  ProfScope prof(st, KC_OTHER);
  scale_residual_kernel<<<blocks, 256, 0, st>>>(out, osb, s, y, res, rsb, B, C, T);
  BVG_LAUNCHED();
  return BVG_OK;
}
int attn_stats_launch(float* pooled, const float* logits, const float* x, const float* bn_scale,
                      const float* bn_shift, int64_t B, int64_t C, int64_t T, cudaStream_t st) {
  if (B * C == 0) return BVG_OK;
  ProfScope prof(st, KC_OTHER);
  attn_stats_kernel<<<(unsigned)((B * C + 3) / 4), 128, 0, st>>>(pooled, logits, x, bn_scale, bn_shift, B, C, T);
  BVG_LAUNCHED();
  return BVG_OK;
}
int conv_post_launch(float* wav, int16_t* pcm, const void* x, const float* w, const float* bias, int64_t B,
                     int64_t Cin, int64_t T, int K, int64_t s_lo, int64_t s_hi, int dtype, cudaStream_t st) {
  BVG_CHECK_ARG(wav || pcm, "conv_post: no output buffer");
  BVG_CHECK_ARG(Cin * K * 4 <= 48 * 1024, "conv_post: Cin*K too large");
  const int64_t Tout = T - s_lo - s_hi;
  BVG_CHECK_ARG(Tout >= 0 && s_lo >= 0 && s_hi >= 0, "conv_post: bad crop");
  if (B == 0 || Tout == 0) return BVG_OK;
  BVG_CHECK_ARG(B <= 65535, "conv_post: batch too large");
  dim3 grid((unsigned)((Tout + 255) / 256), (unsigned)B);
  const size_t smem = (size_t)Cin * K * 4;
  ProfScope prof(st, KC_OTHER);
  if (dtype == BVG_F32)
    conv_post_kernel<float><<<grid, 256, smem, st>>>(wav, pcm, (const float*)x, w, bias, Cin, T, K, s_lo, s_hi);
  else if (dtype == BVG_BF16)
    conv_post_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(wav, pcm, (const __nv_bfloat16*)x, w, bias, Cin, T, K, s_lo, s_hi);
  else { set_error("conv_post: unsupported dtype %d", dtype); return BVG_ERR_INVALID; }
  BVG_LAUNCHED();
  return BVG_OK;
}

namespace {
template <typename TS>
__global__ void cast_to_f32_kernel(float* __restrict__ dst, const TS* __restrict__ src, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = to_f<TS>(src[i]);
}
}  // namespace

int cast_to_f32_launch(float* dst, const void* src, int src_dtype, int64_t n, cudaStream_t st) {
  if (n == 0) return BVG_OK;
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  ProfScope prof(st, KC_OTHER);
  if (src_dtype == BVG_BF16) cast_to_f32_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(dst, (const __nv_bfloat16*)src, n);
  else if (src_dtype == BVG_F16) cast_to_f32_kernel<__half><<<blocks, 256, 0, st>>>(dst, (const __half*)src, n);
  else { set_error("cast_to_f32: unsupported dtype"); return BVG_ERR_INVALID; }
  BVG_LAUNCHED();
  return BVG_OK;
}

}  // namespace bvg
