// Shared helpers for libbigvgan_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <atomic>

#include "../../include/bigvgan_b200.h"

namespace bvg {

void set_error(const char* fmt, ...);
extern thread_local int64_t g_launches;
extern long long* g_dbg_buf;      // optional per-role cycle counters (bvg_debug_set_umma_counters), else null

// ---- optional per-kernel-class timing (bvg_profile_*): CUDA events around every launch ---------
enum KernelClass { KC_ACT1D = 0, KC_CONV = 1, KC_CONVTR = 2, KC_OTHER = 3, KC_ACTCONV = 4, KC_COUNT = 5 };
void prof_mark(cudaStream_t st, int kclass, bool begin);
struct ProfScope {
  cudaStream_t st;
  int kc;
  ProfScope(cudaStream_t s, int k) : st(s), kc(k) { prof_mark(st, kc, true); }
  ~ProfScope() { prof_mark(st, kc, false); }
};

#define BVG_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      bvg::set_error(__VA_ARGS__);               \
      return BVG_ERR_INVALID;                    \
    }                                            \
  } while (0)

#define BVG_CUDA(call)                                                               \
  do {                                                                               \
    cudaError_t e__ = (call);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      bvg::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return BVG_ERR_CUDA;                                                           \
    }                                                                                \
  } while (0)

// called after every kernel launch
#define BVG_LAUNCHED()                                                               \
  do {                                                                               \
    ++bvg::g_launches;                                                               \
    cudaError_t e__ = cudaPeekAtLastError();                                         \
    if (e__ != cudaSuccess) {                                                        \
      bvg::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return BVG_ERR_CUDA;                                                           \
    }                                                                                \
  } while (0)

#define BVG_TRY(call)            \
  do {                           \
    int rc__ = (call);           \
    if (rc__ != BVG_OK) return rc__; \
  } while (0)

// ---- per-device caches ----------------------------------------------------------------------------
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the CURRENT device only and one process may hold plans
// on several devices (the reference web UI is one process): remember the opt-in per device, not per process.
template <class Kern>
inline int smem_opt_in(Kern kern, std::atomic<uint64_t>& done_mask, int bytes) {
  int dev = 0;
  BVG_CUDA(cudaGetDevice(&dev));
  const uint64_t bit = 1ull << (dev & 63);
  if (!(done_mask.load(std::memory_order_acquire) & bit)) {
    BVG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done_mask.fetch_or(bit, std::memory_order_release);
  }
  return BVG_OK;
}
// size threshold (million elements) below which the c8t Activation1d / fused kernels stay on the CUDA-core stencil (plan.cu)
int tc_min_melems();
// SM count of the current device (cached per device; plan.cu)
int current_device_sms(int* sms);
// integer environment switch, read ONCE per process per call site (tuning knobs only)
int env_int_once(const char* name, int dflt);
#define BVG_ENV_ONCE(name, dflt) ([]() -> int { static const int v__ = bvg::env_int_once(name, dflt); return v__; }())

// rows of utterance b in a (possibly ragged) batch; the shuffle marks the loaded value warp-uniform for ptxas (the
// tcgen05.mma issue loops must stay on the uniform datapath: profiles/README.md, round 2)
#ifdef __CUDACC__
__device__ __forceinline__ int rows_of(const int* lens, int len_mul, int b, int T) {
  return lens ? __shfl_sync(0xffffffffu, __ldg(lens + b) * len_mul, 0) : T;
}
#endif

static inline size_t dtype_size(int dt) { return (dt == BVG_F32 || dt == BVG_F32X3) ? 4 : 2; }

// ---- element load/store as float -----------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// The anti-aliasing filter of every Activation1d in the generator: kaiser_sinc_filter1d(0.25,
// 0.3, 12) (alias_free_torch/filter.py:29-58), exact fp32 values of the reference's
// `upsample.filter` / `downsample.lowpass.filter` buffers.  Symmetric: f[k] == f[11-k].
#define BVG_F0 0x1.09f0c2p-9f
#define BVG_F1 0x1.33ac8cp-7f
#define BVG_F2 -0x1.a28108p-6f
#define BVG_F3 -0x1.d8544cp-5f
#define BVG_F4 0x1.075110p-3f
#define BVG_F5 0x1.c5d8cap-2f

// internal launchers (defined in the .cu files), all enqueue on `st`
// Activation1d (fast-math) on fp32 [B,C,T], written as the [hi | lo] bf16 c8t tensor the split convs read (fp32x3 path)
int act1d_split_launch(__nv_bfloat16* dst_c8t, int dst_chunks, int dst_Tp, int dst_pad, const float* src,
                       const float* alpha_log, const float* beta_log, int64_t B, int64_t C, int64_t T, cudaStream_t st);
int act1d_launch(void* dst, const void* src, const float* alpha_log, const float* beta_log,
                 int64_t B, int64_t C, int64_t T, int dtype, int precise, cudaStream_t st);

struct ConvEpilogue {
  const float* bias = nullptr;      // [Cout]
  const void* res1 = nullptr;       // [B,Cout,T] same dtype as out
  const void* res2 = nullptr;       // [B,Cout,T]
  float scale = 1.f;                // out = (acc + bias + res1 + res2) * scale
  const float* cond = nullptr;      // [Bc,Cout] per-(b,co) additive
  int64_t cond_B = 1;
  int relu = 0;                     // apply ReLU after bias (ECAPA TDNN)
  const float* post_scale = nullptr;  // [Cout] eval-BN folded affine applied after ReLU
  const float* post_shift = nullptr;
  int act = 0;                      // 0 none, 1 tanh, 2 sigmoid (applied last)
  int prof_other = 0;               // account this launch to the "other" profiling class (speaker encoder)
};

// src element (b, ci, t) lives at src[b*sb + ci*sc + t*st_] (src2, when non-null, has the same
// strides and is added to src while staging: conv(src + src2)).  dst element (b, co, t) lives at
// dst[b*dsb + co*T + t]; res1/res2 use dst's addressing.  weight_kic: [K][Cin][Cout] fp32
// (re-laid-out).  pad_mode 0 zero, 1 reflect.
int conv1d_simt_launch(void* dst, int64_t dsb, const void* src, const void* src2, int64_t sb, int64_t sc,
                       int64_t st_, const float* weight_kic, const ConvEpilogue& ep,
                       int64_t B, int64_t Cin, int64_t Cout, int64_t T, int K, int dil,
                       int pad_mode, int in_dtype, int out_dtype, cudaStream_t st);
// weight_kic: [K][Cin][Cout] fp32.
int convtr1d_simt_launch(void* dst, const void* src, const float* weight_kic, const ConvEpilogue& ep,
                         int64_t B, int64_t Cin, int64_t Cout, int64_t Tin, int K, int stride,
                         int dtype, cudaStream_t st);
// [Cout,Cin,K] -> [K][Cin][Cout]   /   [Cin,Cout,K] -> [K][Cin][Cout]   (device to device)
int repack_conv_weight_launch(float* dst, const float* src, int64_t Cout, int64_t Cin, int K,
                              int transposed, cudaStream_t st);


// ---- small kernels (misc.cu) ------------------------------------------------------------------
// dst[i] = float(src[i]) for bf16 / fp16 sources
int cast_to_f32_launch(float* dst, const void* src, int src_dtype, int64_t n, cudaStream_t st);
// y[b,co] = epilogue(sum_ci w[ci][co] x[b,ci]) for contiguous x [B,Cin], y [B,Cout]; w is the K=1 [Cin][Cout] pack
int matvec_launch(float* y, const float* x, const float* w_ic, const ConvEpilogue& ep, int64_t B, int Cin, int Cout,
                  cudaStream_t st);
// several such layers over the SAME x in one launch (the speaker-condition vectors of all stages: models.py:226,233-234)
constexpr int kMaxMatvecJobs = 8;
struct MatvecJob {
  const float* w = nullptr;           // [Cin][Cout]
  const float* bias = nullptr;
  float* y = nullptr;                 // [B, Cout]
  int Cout = 0, relu = 0, act = 0;
  const float* post_scale = nullptr;
  const float* post_shift = nullptr;
};
int matvec_multi_launch(const MatvecJob* jobs, int njobs, const float* x, int64_t B, int Cin, cudaStream_t st);
// mean over time of each row of x [rows, T] (row r at x + r*T) -> out[r]
int row_mean_launch(float* out, const float* x, int64_t rows, int64_t T, cudaStream_t st);
// mean/std over time: x [B,C,T] -> ms[b, c] = mean, ms[b, C + c] = sqrt(clamp(var, 1e-12))
int row_stats_launch(float* ms, const float* x, int64_t B, int64_t C, int64_t T, cudaStream_t st);
// out[b,c,t] = s[b,c] * y[b,c,t] + res[b,c,t]; out/res have batch strides osb/rsb, y is contiguous; yc (optional): the
// same values as a c8t bf16 tensor (or channel-slice view of one)
struct C8T;
int scale_residual_launch(float* out, int64_t osb, const float* s, const float* y, const float* res, int64_t rsb,
                          int64_t B, int64_t C, int64_t T, cudaStream_t st, const C8T* yc = nullptr);
// attentive statistics: softmax over time of logits[b,c,:], weighted mean/std of x[b,c,:], then the
// eval-BatchNorm affine (scale/shift [2C]):  pooled[b, c], pooled[b, C + c]
int attn_stats_launch(float* pooled, const float* logits, const float* x, const float* bn_scale,
                      const float* bn_shift, int64_t B, int64_t C, int64_t T, cudaStream_t st);
// activation_post output x [B,Cin,T] -> conv_post (Cin->1, K taps, zero pad) -> tanh -> wav fp32 and/or
// int16 pcm (clamp(32767*wav, +-32767)), keeping only samples [s_lo, T - s_hi) of each utterance.
int conv_post_launch(float* wav, int16_t* pcm, const void* x, const float* w /*[Cin][K]*/, const float* bias,
                     int64_t B, int64_t Cin, int64_t T, int K, int64_t s_lo, int64_t s_hi, int dtype,
                     cudaStream_t st);

}  // namespace bvg
