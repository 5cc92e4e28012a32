// Small reduction / elementwise kernels of the decode path (speaker encoder glue, conv_post).
// References: ECAPA_TDNN.py:228-242 (SEBlock), :282-338 (AttentiveStatisticsPooling),
// models.py:246-248 (activation_post -> conv_post -> tanh), infer.py:206-212 (int16 epilogue).
#include <string.h>

#include "bvg_common.cuh"
#include "umma.cuh"

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

// one thread = 8 channels (one c8t chunk) of one time step: fp32 out and, when yc is given, the same values as one 16-byte
// row of the c8t bf16 tensor the next 1x1 GEMM reads
__global__ void scale_residual_kernel(float* __restrict__ out, int64_t osb, const float* __restrict__ s,
                                      const float* __restrict__ y, const float* __restrict__ res, int64_t rsb,
                                      int64_t B, int64_t C, int64_t T, __nv_bfloat16* __restrict__ yc, int64_t yc_bstride,
                                      int yc_tp, int yc_pad) {
  const int64_t C8 = (C + 7) / 8, n = B * C8 * T;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i % T, c8 = (i / T) % C8, b = i / (T * C8);
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int64_t c = c8 * 8 + k;
      v[k] = 0.f;
      if (c < C) {
        v[k] = fmaf(s[b * C + c], y[(b * C + c) * T + t], res[b * rsb + c * T + t]);
        out[b * osb + c * T + t] = v[k];
      }
    }
    if (yc) {
      __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
      uint4 o;
      o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
      o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
      *reinterpret_cast<uint4*>(yc + b * yc_bstride + (c8 * yc_tp + yc_pad + t) * 8) = o;
    }
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
struct MatvecJobs {
  MatvecJob job[kMaxMatvecJobs];
  int blk0[kMaxMatvecJobs + 1];        // first blockIdx.x of every job
  int njobs;
};
// 256 (1024) threads = 32 outputs x 8 (32) input slices, sixteen independent loads in flight per thread (one accumulator chain with four
// loads per round trip made each of these launches ~25 us of L2 latency); several layers that read the same x share a launch.
__global__ void __launch_bounds__(1024)
matvec_kernel(const MatvecJobs J, const float* __restrict__ x, int Cin) {
  __shared__ float part[32][32];
  const int nsl = blockDim.x >> 5;                 // input slices: 8, or 32 for the 3072-input layers of the pooling head
  int j = 0;
  while (j + 1 < J.njobs && (int)blockIdx.x >= J.blk0[j + 1]) ++j;
  const MatvecJob& job = J.job[j];
  const int Cout = job.Cout;
  const int co = ((int)blockIdx.x - J.blk0[j]) * 32 + (threadIdx.x & 31);
  const int sl = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const float* xb = x + (int64_t)b * Cin;
  const float* w = job.w;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  if (co < Cout) {
    const int per = (Cin + nsl - 1) / nsl;
    const int lo = sl * per, hi = min(Cin, lo + per);
    int ci = lo;
    for (; ci + 16 <= hi; ci += 16) {
      float wv[16], xv[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) { wv[u] = __ldg(w + (int64_t)(ci + u) * Cout + co); xv[u] = __ldg(xb + ci + u); }
#pragma unroll
      for (int u = 0; u < 16; u += 4) {
        acc0 = fmaf(wv[u], xv[u], acc0); acc1 = fmaf(wv[u + 1], xv[u + 1], acc1);
        acc2 = fmaf(wv[u + 2], xv[u + 2], acc2); acc3 = fmaf(wv[u + 3], xv[u + 3], acc3);
      }
    }
    for (; ci < hi; ++ci) acc0 = fmaf(__ldg(w + (int64_t)ci * Cout + co), __ldg(xb + ci), acc0);
  }
  part[sl][threadIdx.x & 31] = (acc0 + acc1) + (acc2 + acc3);
  __syncthreads();
  if (sl == 0 && co < Cout) {
    float v = 0.f;
    for (int q = 0; q < nsl; ++q) v += part[q][threadIdx.x];
    if (job.bias) v += job.bias[co];
    if (job.relu) v = fmaxf(v, 0.f);
    if (job.post_scale) v = fmaf(v, job.post_scale[co], job.post_shift[co]);
    if (job.act == 1) v = tanhf(v);
    if (job.act == 2) v = 1.f / (1.f + expf(-v));
    job.y[(int64_t)b * Cout + co] = v;
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

int matvec_multi_launch(const MatvecJob* jobs, int njobs, const float* x, int64_t B, int Cin, cudaStream_t st) {
  BVG_CHECK_ARG(jobs && x && Cin > 0 && njobs >= 1 && njobs <= kMaxMatvecJobs, "matvec: bad argument");
  if (B == 0) return BVG_OK;
  BVG_CHECK_ARG(B <= 65535, "matvec: batch too large");
  MatvecJobs J;
  memset(&J, 0, sizeof J);
  J.njobs = njobs;
  int blk = 0;
  for (int j = 0; j < njobs; ++j) {
    BVG_CHECK_ARG(jobs[j].w && jobs[j].y && jobs[j].Cout > 0, "matvec: bad job");
    J.job[j] = jobs[j];
    J.blk0[j] = blk;
    blk += (jobs[j].Cout + 31) / 32;
  }
  J.blk0[njobs] = blk;
  dim3 grid((unsigned)blk, (unsigned)B);
  ProfScope prof(st, KC_OTHER);
  matvec_kernel<<<grid, Cin >= 2048 ? 1024 : 256, 0, st>>>(J, x, Cin);
  BVG_LAUNCHED();
  return BVG_OK;
}

int matvec_launch(float* y, const float* x, const float* w_ic, const ConvEpilogue& ep, int64_t B, int Cin, int Cout,
                  cudaStream_t st) {
  BVG_CHECK_ARG(y && x && w_ic && Cin > 0 && Cout > 0, "matvec: bad argument");
  BVG_CHECK_ARG(!ep.res1 && !ep.res2 && !ep.cond && ep.scale == 1.f, "matvec: unsupported epilogue");
  MatvecJob j;
  j.w = w_ic; j.bias = ep.bias; j.y = y; j.Cout = Cout; j.relu = ep.relu; j.post_scale = ep.post_scale;
  j.post_shift = ep.post_shift; j.act = ep.act;
  return matvec_multi_launch(&j, 1, x, B, Cin, st);
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
                          int64_t B, int64_t C, int64_t T, cudaStream_t st, const C8T* yc) {
  const int64_t n = B * ((C + 7) / 8) * T;
  if (n == 0) return BVG_OK;
  BVG_CHECK_ARG(!yc || (yc->T == (int)T && yc->chunks * 8 >= C), "scale_residual: c8t output geometry");
  int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
  ProfScope prof(st, KC_OTHER);
  scale_residual_kernel<<<blocks, 256, 0, st>>>(out, osb, s, y, res, rsb, B, C, T, yc ? yc->p : nullptr,
                                                yc ? yc->batch_stride() : 0, yc ? yc->Tp : 0, yc ? yc->pad : 0);
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
