/*
 * bigvgan_b200.h -- C ABI of libbigvgan_b200.so: a B200 (sm_100a) implementation of the
 * IndexTTS BigVGAN vocoder decode path.
 *
 * This header is the drop-in boundary.  Every entry point takes plain pointers, sizes and a
 * cudaStream_t (passed as void*); there are no torch types.  All functions return 0 on
 * success and a non-zero code on failure; bvg_last_error() returns a thread-local message
 * (the Python wrapper raises RuntimeError with it, mirroring AT_ERROR -> RuntimeError in the
 * reference's dispatch macro, alias_free_activation/cuda/type_shim.h:41-42).
 *
 * Ownership: the caller owns every input/output/workspace buffer (in practice torch tensors,
 * passed by data_ptr()).  The library owns only what a bvg_plan holds (re-laid-out weights).
 * No allocation and no synchronisation happens inside a *_fwd / bvg_decode call; work is
 * enqueued on the caller's stream (the reference launches on the current torch stream,
 * anti_alias_activation_cuda.cu:209).  Calls are re-entrant: scratch lives in the
 * caller-provided workspace, so concurrent calls with distinct workspaces are safe
 * (webui.py:441-452 runs one thread per request against one shared engine).
 *
 * Reference paths are relative to /root/reference/indextts/BigVGAN/.
 */
#ifndef BIGVGAN_B200_H_
#define BIGVGAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define BVG_API __attribute__((visibility("default")))
#else
#define BVG_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* element types of activation tensors */
/* BVG_F32X3 (decode only): fp32 tensors like BVG_F32, but the generator's Conv1d layers run on the tensor cores
   with x and w split into two bf16 terms each and three of the four products accumulated in fp32 (relative error
   ~2^-17 per product; 2e-6 max-abs on the waveform against 5e-7 for BVG_F32) */
enum { BVG_F32 = 0, BVG_BF16 = 1, BVG_F16 = 2, BVG_F32X3 = 3 };

/* status codes */
enum {
  BVG_OK = 0,
  BVG_ERR_INVALID = 1,   /* bad argument / unsupported shape        */
  BVG_ERR_CUDA = 2,      /* a CUDA runtime/driver call failed       */
  BVG_ERR_STATE = 3,     /* plan not finalised / missing weight     */
  BVG_ERR_WORKSPACE = 4  /* workspace too small                     */
};

BVG_API const char* bvg_last_error(void);
/* library + build identification, e.g. "bigvgan_b200 0.1 sm_100a" */
BVG_API const char* bvg_version(void);
/* number of kernels this library launched from the calling thread since the last reset
   (bench.py reports it as gpu_launches) */
BVG_API int64_t bvg_launch_count(void);
BVG_API void bvg_launch_count_reset(void);
/* Per-kernel-class device timing for bench.py's roofline figure: between begin and end every
   kernel launched from the calling thread is bracketed by CUDA events on its stream.
   bvg_profile_end synchronises those events and returns, per class, the summed kernel time (ms)
   and the launch count.  Classes: 0 Activation1d, 1 dense Conv1d (generator), 2 ConvTranspose1d,
   3 everything else (speaker encoder, cond vectors, conv_post), 4 fused Activation1d->Conv1d
   (narrow generator stages).  Arrays have 5 entries. */
BVG_API void bvg_profile_begin(void);
BVG_API int bvg_profile_end(float* ms_per_class, int64_t* launches_per_class);

/* ------------------------------------------------------------------------------------------
 * Op boundary: the reference's one native op.
 *   replaces  anti_alias_activation_cuda.forward(inputs, up_ftr, down_ftr, alpha, beta)
 *             (alias_free_activation/cuda/anti_alias_activation.cpp:19-23,
 *              anti_alias_activation_cuda.cu:214-256)
 * Semantics follow the PyTorch module (alias_free_torch/act.py:24-29), including its edge
 * behaviour (replicate padding of the input for the up-FIR and of the ACTIVATED signal for the
 * down-FIR), which the reference CUDA kernel does not reproduce at the first/last 3 samples.
 *   src/dst  : [B, C, T] contiguous, dtype in {BVG_F32, BVG_BF16, BVG_F16}; dst != src
 *   alpha_log, beta_log : fp32 [C], log-scale (the kernel applies exp, like .cu:89-90)
 *   up_taps/down_taps   : fp32 [12] (device or host pointer is NOT accepted: pass HOST
 *                         pointers; they are validated against the compiled-in taps with
 *                         1e-6 tolerance; NULL skips validation)
 *   precise  : 1 = libdevice sinf/expf (fp32 parity path), 0 = range-reduced MUFU path
 * ------------------------------------------------------------------------------------------ */
BVG_API int bvg_act1d_fwd(void* dst, const void* src, const float* alpha_log, const float* beta_log,
                  const float* up_taps_host, const float* down_taps_host,
                  int64_t B, int64_t C, int64_t T, int dtype, int precise, void* stream);

/* ------------------------------------------------------------------------------------------
 * Layer-level entry points (parity tests call these through ctypes).
 *   conv1d : torch.nn.Conv1d forward, stride 1, "same" length, zero padding (generator convs,
 *            models.py:25-42,149,184) or reflect padding (ECAPA, nnet/CNN.py:458-488).
 *            weight [Cout, Cin, K] fp32 (torch layout, device), bias [Cout] fp32 or NULL.
 *            out = (conv + bias + res1 + res2) * scale   (res1/res2 nullable, same shape as out)
 *   convtr1d : torch.nn.ConvTranspose1d forward (models.py:155-161): weight [Cin, Cout, K]
 *            fp32, stride u, padding (K-u)/2, plus a per-(b,co) additive term `cond`
 *            ([Bc, Cout] fp32, Bc in {1,B}; nullable) -- the speaker conditioning add of
 *            models.py:232-234.
 * ------------------------------------------------------------------------------------------ */
BVG_API int bvg_conv1d_fwd(void* dst, const void* src, const float* weight, const float* bias,
                   const void* res1, const void* res2, float scale,
                   int64_t B, int64_t Cin, int64_t Cout, int64_t T, int K, int dilation,
                   int reflect_pad, int dtype, void* stream);
BVG_API int bvg_convtr1d_fwd(void* dst, const void* src, const float* weight, const float* bias,
                     const float* cond, int64_t Bc,
                     int64_t B, int64_t Cin, int64_t Cout, int64_t Tin, int K, int stride,
                     int dtype, void* stream);

/* Same two layers on the bf16 tensor-core path (tcgen05.mma, TMEM accumulators, TMA-staged
 * operands).  src/dst are plain [B,C,T] bf16 device tensors; the call converts to and from the
 * library's internal channel-chunked layout (layer-level test entry points: they allocate
 * temporaries with cudaMallocAsync; bvg_decode does not).  weight is fp32 in torch layout and is
 * rounded to bf16; accumulation is fp32; the epilogue adds bias (+cond) (+res1 +res2), scales. */
/* Activation1d through the channel-chunked kernel of the bf16 path (plain [B,C,T] bf16 in/out;
 * converts internally; test entry point). */
BVG_API int bvg_act1d_c8t_fwd(void* dst, const void* src, const float* alpha_log, const float* beta_log,
                      int64_t B, int64_t C, int64_t T, void* stream);
/* Same, choosing the implementation: 0 = what the decode path takes (tensor-core FIRs, csrc/act1d_tc.cu, for T >= 256;
 * the CUDA-core stencil otherwise), 1 = CUDA-core stencil (csrc/act1d_c8t.cu), 2 = tensor-core FIRs (BVG_ERR_INVALID if
 * the shape does not qualify).  Both follow alias_free_torch/act.py:24-29 of the reference, edges included. */
BVG_API int bvg_act1d_c8t_impl_fwd(void* dst, const void* src, const float* alpha_log, const float* beta_log,
                           int64_t B, int64_t C, int64_t T, int impl, void* stream);
/* Profiling aid: when non-NULL, the next bvg_conv*_umma_fwd launches write 8 int64 cycle counters per CTA
 * (148 CTAs max) into this device buffer: producer wait, MMA waits on input / weights / TMEM, MMA issue,
 * MMA-thread total, epilogue wait, epilogue busy. */
BVG_API void bvg_debug_set_umma_counters(long long* dev_buf);
/* Kernel selection of the bf16 path: c8t tensors with fewer than `melems` million elements (B * C * T) run Activation1d / the
 * fused Activation1d -> Conv1d on the CUDA-core stencil kernels (a single utterance: lower latency), larger ones on the
 * tensor-core FIR kernels.  Default 10 (env BVG_TC_MIN_MELEMS); 0 = tensor cores whenever the shape qualifies; < 0 restores
 * the default.  Process-wide. */
BVG_API void bvg_debug_set_tc_min_melems(int melems);
/* Measurement hook: 0 keeps the AMP blocks of a stage (models.py:238-245) on the caller's stream instead of spreading them
 * over three streams, so that CUDA events around a launch time that kernel alone (bench.py's per-class pass); 1 restores the
 * default.  A workspace sized while this is 1 serves both settings. */
BVG_API void bvg_debug_set_multi_stream(int on);
/* Activation1d(src) -> Conv1d (+bias, +res1, *scale) in ONE kernel (narrow layers, Cout <= 128): the form the
 * bf16 decode path uses for AMPBlock1's act->conv pairs (models.py:65-74).  Plain [B,C,T] bf16 in/out; status 3
 * when the shape does not qualify.  Test entry point (allocates temporaries). */
BVG_API int bvg_actconv_umma_fwd(void* dst, const void* src, const float* alpha_log, const float* beta_log,
                         const float* weight, const float* bias, const void* res1, float scale,
                         int64_t B, int64_t Cin, int64_t Cout, int64_t T, int K, int dilation, void* stream);
/* The same with a second residual and a choice of kernel: impl 0 = what the decode path runs (tensor-core FIR kernel
 * actconv_tc.cu when the layer qualifies -- C = 24 / 48 / 96, T >= 512 -- else the CUDA-core stencil kernel), 1 = the
 * CUDA-core stencil kernel (conv_umma_fused.cu), 2 = the tensor-core FIR kernel or status 3.  Also rewrites the output's
 * zero halo rows (the decode path's zero_pads).  Test entry point (allocates temporaries). */
BVG_API int bvg_actconv_impl_fwd(void* dst, const void* src, const float* alpha_log, const float* beta_log,
                         const float* weight, const float* bias, const void* res1, const void* res2, float scale,
                         int64_t B, int64_t Cin, int64_t Cout, int64_t T, int K, int dilation, int impl, void* stream);
BVG_API int bvg_conv1d_umma_fwd(void* dst, const void* src, const float* weight, const float* bias,
                        const void* res1, const void* res2, float scale,
                        int64_t B, int64_t Cin, int64_t Cout, int64_t T, int K, int dilation, void* stream);
BVG_API int bvg_convtr1d_umma_fwd(void* dst, const void* src, const float* weight, const float* bias,
                          const float* cond, int64_t Bc,
                          int64_t B, int64_t Cin, int64_t Cout, int64_t Tin, int K, int stride, void* stream);

/* ------------------------------------------------------------------------------------------
 * Whole-path plan.   replaces  indextts.BigVGAN.models.BigVGAN  (models.py:130-275) as built,
 * loaded and called by infer.py:61-67,204,498.
 * ------------------------------------------------------------------------------------------ */
typedef struct bvg_plan bvg_plan;

typedef struct bvg_config {
  int32_t gpt_dim;                  /* h.gpt_dim                      */
  int32_t upsample_initial_channel; /* h.upsample_initial_channel     */
  int32_t num_upsamples;            /* len(h.upsample_rates)  (<= 8)  */
  int32_t upsample_rates[8];
  int32_t upsample_kernel_sizes[8];
  int32_t num_kernels;              /* len(h.resblock_kernel_sizes) (<= 4) */
  int32_t resblock_kernel_sizes[4];
  int32_t resblock_dilation_sizes[4][3];
  int32_t speaker_embedding_dim;    /* h.speaker_embedding_dim        */
  int32_t num_mels;                 /* h.num_mels                     */
  int32_t cond_in_each_up_layer;    /* h.cond_d_vector_in_each_upsampling_layer */
  int32_t snake_logscale;           /* h.snake_logscale               */
  int32_t device;                   /* CUDA device ordinal            */
} bvg_config;

BVG_API int bvg_plan_create(bvg_plan** out, const bvg_config* cfg);
BVG_API void bvg_plan_destroy(bvg_plan* plan);

/* Upload one tensor of the (weight-norm-folded) state dict.  `key` is the reference's
 * state-dict key after remove_weight_norm (models.py:252-260), e.g. "conv_pre.weight",
 * "resblocks.3.convs1.0.bias", "resblocks.3.activations.2.act.alpha",
 * "speaker_encoder.blocks.0.norm.norm.running_var", "conds.2.weight".
 * `data` is a HOST fp32 pointer with `numel` elements in torch's contiguous layout. */
BVG_API int bvg_plan_set_tensor(bvg_plan* plan, const char* key, const float* data, int64_t numel);
/* Validates that every tensor the config requires was supplied, folds eval BatchNorm and
 * builds the device-side layouts (SIMT fp32 packs and, when `enable_bf16_umma` != 0, the bf16
 * per-tap K-major packs the tcgen05 path consumes). */
BVG_API int bvg_plan_finalize(bvg_plan* plan, int enable_bf16_umma);

/* workspace the decode call needs for (B utterances, T0 latent frames, Tm mel frames) in
 * precision `dtype` (BVG_F32: fp32 storage + fp32 CUDA-core math; BVG_BF16: bf16 storage,
 * tcgen05 bf16 convs with fp32 accumulate; BVG_F32X3: fp32 storage, tcgen05 convs on 3-term
 * bf16 splits, fast-math Activation1d) */
BVG_API size_t bvg_workspace_bytes(const bvg_plan* plan, int64_t B, int64_t T0, int64_t Tm, int dtype);

/* Reference-mel front end.   replaces  MelSpectrogramFeatures.forward  (indextts/utils/feature_extractors.py:24-50, called
 * from infer.py:82-93): torchaudio MelSpectrogram(n_fft, hop, periodic hann, center + reflect pad, power 1, htk mel scale)
 * + safe_log(clip 1e-7) (utils/common.py:110).
 *   audio [B, L] fp32 device (mono, already at the model's sample rate), fb [n_fft/2+1, n_mels] fp32 device = the mel
 *   filterbank (torchaudio.functional.melscale_fbanks; the host layer computes it), mel [B, bvg_mel_frames(L, hop), n_mels]
 *   fp32 device -- the layout bvg_speaker_embed / bvg_decode take (the reference's [B, n_mels, frames] transposed). */
BVG_API int64_t bvg_mel_frames(int64_t L, int hop);
BVG_API int bvg_mel_frontend(float* mel, const float* audio, const float* fb, int64_t B, int64_t L, int n_fft, int hop,
                             int n_mels, void* stream);

/* speaker encoder only:  mel [Bm, Tm, num_mels] fp32 device -> spk [Bm, emb] fp32 device
 * (ECAPA_TDNN.forward, ECAPA_TDNN.py:543-581, lengths=None) */
BVG_API int bvg_speaker_embed(const bvg_plan* plan, const float* mel, int64_t Bm, int64_t Tm,
                      float* spk, void* workspace, size_t workspace_bytes, void* stream);

/* BigVGAN.forward (models.py:201-250).
 *   latent [B, T0, gpt_dim] fp32 device; exactly one of {mel, spk} non-NULL:
 *   mel [Bm, Tm, num_mels] fp32 device (Bm in {1, B}) or a precomputed spk [Bm, emb] fp32.
 *   wav  [B, 1, T0*prod(rates)]  fp32 device  (tanh output), or
 *   when pcm16 != NULL: additionally the caller epilogue of infer.py:206-212,234 fused:
 *   pcm16[B, L] = int16(clamp(32767*wav, -32767, 32767)); wav may then be NULL.
 *   t_lo_pad/t_hi_pad: for chunked long-form decode -- latent frames at the start/end of this
 *   call that are halo (their output samples are computed but not stored; wav then has
 *   (T0 - t_lo_pad - t_hi_pad)*prod(rates) samples per utterance).  0,0 for a plain call. */
BVG_API int bvg_decode(const bvg_plan* plan, const float* latent, const float* mel, const float* spk,
               int64_t B, int64_t T0, int64_t Bm, int64_t Tm, int dtype,
               float* wav, int16_t* pcm16, int64_t t_lo_pad, int64_t t_hi_pad,
               void* workspace, size_t workspace_bytes, void* stream);

/* bvg_decode with the latent in the dtype the GPT hands it over in (gpt/model.py:462-477 under the autocast of infer.py:194):
 * latent [B, T0, gpt_dim] channels-last, latent_dtype in {BVG_F32, BVG_BF16, BVG_F16}.  The bf16 path ingests it directly
 * (no fp32 round trip); the fp32 paths widen it once in the workspace.  Everything else as bvg_decode. */
BVG_API int bvg_decode_lat(const bvg_plan* plan, const void* latent, int latent_dtype, const float* mel, const float* spk,
                   int64_t B, int64_t T0, int64_t Bm, int64_t Tm, int dtype, float* wav, int16_t* pcm16,
                   int64_t t_lo_pad, int64_t t_hi_pad, void* workspace, size_t workspace_bytes, void* stream);

/* Ragged batch: true batched vocoding of utterances of DIFFERENT lengths in one call (what infer_fast lacks,
 * infer.py:480-503: it concatenates sentences along time instead).  latent [B, T0_max, gpt_dim] (rows past an utterance's
 * length are ignored), lens_dev = device int32 [B] latent-frame counts (1 <= lens[b] <= T0_max).  bf16 path only.
 * Every kernel clips its tiles and applies its edge semantics (zero / replicate padding) at each utterance's own length, so
 * utterance b's samples [0, lens[b] * prod(rates)) equal those of decoding it alone through the same kernels; the rest of
 * its row in wav / pcm16 ([B, T0_max * prod(rates)]) is zero. */
BVG_API int bvg_decode_varlen(const bvg_plan* plan, const void* latent, int latent_dtype, const int32_t* lens_dev,
                      const float* mel, const float* spk, int64_t B, int64_t T0_max, int64_t Bm, int64_t Tm,
                      float* wav, int16_t* pcm16, void* workspace, size_t workspace_bytes, void* stream);

/* Same call with HOST buffers: copies latent/mel host->device, decodes, copies the waveform
 * back (the end-to-end figure bench.py reports as `e2e`).  Uses the plan's device and the
 * given stream; synchronises the stream before returning.  latent/mel/wav should be pinned. */
BVG_API int bvg_decode_host(const bvg_plan* plan, const float* latent_host, const float* mel_host,
                    int64_t B, int64_t T0, int64_t Bm, int64_t Tm, int dtype,
                    float* wav_host, int16_t* pcm16_host,
                    void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BIGVGAN_B200_H_ */
