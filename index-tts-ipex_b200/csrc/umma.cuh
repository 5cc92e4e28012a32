// Types shared by the tcgen05 conv path (conv_umma.cu), the c8t elementwise kernels and plan.cu.
#pragma once
#include "bvg_common.cuh"

namespace bvg {

constexpr int kC8tPad = 32;    // zero rows before t=0 and after t=T-1 (max conv padding is 25)

// "c8t" activation tensor: [B][chunks][Tp][8] bf16, element (b,c,t) at
// ((b*chunks + c/8)*Tp + pad + t)*8 + c%8.  chunks*8 >= C rounded up to a multiple of 16.
struct C8T {
  __nv_bfloat16* p = nullptr;
  int C = 0, chunks = 0, T = 0, Tp = 0, pad = kC8tPad;
  // ragged batches (bvg_decode_varlen): utterance b holds lens[b] * len_mul valid rows (device int32 array of latent-frame
  // counts, times the stage's cumulative upsampling factor); T is then the LONGEST utterance = the memory geometry.  Every
  // kernel clips its tiles, its edge semantics (zero / replicate padding) and its zero halo rows at the utterance's own
  // length, so each utterance of a ragged batch is computed exactly as if decoded alone.  nullptr: all utterances have T rows.
  const int* lens = nullptr;
  int len_mul = 1;
  // elements between utterances; non-zero `bstride` makes this a channel-slice view of a wider tensor (speaker encoder:
  // the three SERes2Net outputs live side by side in the 1536-channel tensor the MFA layer reads)
  int64_t bstride = 0;
  int64_t batch_stride() const { return bstride ? bstride : (int64_t)chunks * Tp * 8; }
};
inline C8T make_c8t(void* p, int C, int T) {
  C8T t;
  t.p = static_cast<__nv_bfloat16*>(p);
  t.C = C; t.chunks = ((C + 15) / 16) * 2; t.T = T; t.Tp = T + 2 * kC8tPad;
  return t;
}
// bytes of a c8t buffer for B batch elements, including the tail slack tiles may over-read
inline size_t c8t_bytes(int64_t B, int C, int64_t T) {
  return (size_t)B * (((C + 15) / 16) * 2) * (T + 2 * kC8tPad) * 16 + 1024 * 16;
}

struct UmmaLayer {
  const __nv_bfloat16* w = nullptr;   // packed [n_blk][ci_blk][tap][kchunk][NB][8]
  const float* bias = nullptr;
  int Cin = 0, Cout = 0, K = 1, dil = 1;
  int transposed = 0, stride = 1;
  // 3-term bf16 split ("fp32x3", the tensor-core form of the fp32 path): x = hi + lo and w = hi + lo in bf16,
  // y = x_hi*w_hi + x_hi*w_lo + x_lo*w_hi accumulated in fp32 (relative error ~2^-17 per product).  The input
  // tensor carries the halves as channels [hi (Cp) | lo (Cp)], Cp = Cin rounded up to 8, and the weights are packed
  // with 2K taps: taps [0,K) = (w_hi | w_hi), taps [K,2K) = (w_lo | 0) at the same shifts.
  int split = 0;
};

struct UmmaEpilogue {
  const float* bias = nullptr;
  const float* cond = nullptr;        // [cond_B][Cout]
  int64_t cond_B = 1;
  const __nv_bfloat16* res1 = nullptr;  // same c8t geometry as the output
  const __nv_bfloat16* res2 = nullptr;
  float scale = 1.f;
  int relu = 0;                       // ReLU after bias/cond (ECAPA TDNN: conv -> ReLU -> BN)
  const float* post_scale = nullptr;  // [Cout] folded eval-BN affine applied after the ReLU
  const float* post_shift = nullptr;
  int act = 0;                        // 1: tanh after the affine
  int prof_other = 0;                 // account this launch to the "other" profiling class (speaker encoder)
  int zero_pads = 0;                  // also (re)write the output's zero halo rows
  float* yf32 = nullptr;              // fp32 plain [B,Cout,T] output instead of the c8t tensor (split convs)
  const float* res1_f32 = nullptr;    // fp32 plain residuals for the fp32 output
  const float* res2_f32 = nullptr;
  int dry = 0;                        // debug: run only the MMA issue loop (no TMA, waits or epilogue)
  long long* dbg = nullptr;           // optional [grid][8] cycle counters (profiling builds of the tests)
};

struct UmmaConvParams {
  const __nv_bfloat16* x; int64_t x_bstride; int x_tp; int x_row0;
  __nv_bfloat16* y; int64_t y_bstride; int y_tp; int y_row0; int y_chunks;
  const __nv_bfloat16* w;
  const __nv_bfloat16* res1; const __nv_bfloat16* res2;
  const float* bias; const float* cond; int cond_B;
  float scale;
  int Tout, u, p;                     // output row of coarse row q, phase s: t = u*q + s - p
  int ntaps; int16_t tap_shift[16]; int16_t tap_acc[16];
  int lo, XR;                         // staged rows: [q0 - lo, q0 - lo + XR)
  int n_ci_blk, Cin_p, NB, Cout, NPH, MT, tiles_per_batch, zero_pads, tmem_cols;
  int acc_stages, n_nblk, B;
  int x_stages, w_stages, w_resident, kc_max;
  int a_stages, rows_out;             // fused kernel: A-operand ring depth, output rows per tile (XR - conv halo)
  long long* dbg;
  int dry;
  int n_issuers;
  const int* lens; int len_mul;       // ragged batch: valid OUTPUT rows of utterance b = lens[b] * len_mul (else Tout)
  int transposed, dil;
  int relu, act;
  int Cin;                            // true input channels (fused activation)
  const float* act_alpha; const float* act_beta;   // fused Activation1d parameters (log scale), or null
  const float* post_scale; const float* post_shift;
  float* yf; const float* r1f; const float* r2f;    // fp32 plain output / residuals (split convs), else null
  int tap_mod;                        // tap t reads the input at shift (t % tap_mod) * dil (== ntaps unless split)
  int split_hi_chunks;                // split convs: 8-channel chunks of the hi half (taps >= tap_mod have zero weights beyond), else 0
};

// Res2Net chain of one SERes2NetBlock in one launch (ecapa.cu); *taken = false when the shape does not fit (the caller
// then runs the seven convs separately).  y2 (fp32 [B,512,T]) and / or yc (c8t bf16) receive all eight groups.
int res2net_chain_launch(const float* y1, float* y2, const C8T* yc, const float* const* w, const float* const* bias,
                         const float* const* bn_scale, const float* const* bn_shift, int dil, int64_t B, int64_t T,
                         bool* taken, cudaStream_t st);
void umma_choose_nb(int Cout, int nph, int* NB, int* n_nblk);
int64_t umma_pack_elems(int Cout, int Cin, int K, int nph);
int umma_pack_launch(__nv_bfloat16* dst, const float* src_torch_layout, int Cout, int Cin, int K, int transposed,
                     int nph, cudaStream_t st);
// split-weight pack (see UmmaLayer::split) of a torch-layout [Cout][Cin][K] fp32 weight
int64_t umma_pack_split_elems(int Cout, int Cin, int K, int nph);
int umma_pack_split_launch(__nv_bfloat16* dst, const float* src_torch_layout, int Cout, int Cin, int K, int transposed,
                           int nph, cudaStream_t st);
// fp32 [B,C,T] (general strides) -> c8t bf16 with 2*roundup8(C) channels: [hi | lo] halves of every element
int split_to_c8t_launch(const C8T& dst, const float* src, int64_t sb, int64_t sc, int64_t st_, int C, int64_t B, cudaStream_t st);
int conv_umma_launch(const UmmaLayer& L, const C8T& x, const C8T& y, const UmmaEpilogue& ep, int64_t B, cudaStream_t st);
// Activation1d -> Conv1d in one kernel for narrow layers; BVG_ERR_STATE (nothing launched) if the layer does not qualify
// max_nb: widest output-channel block accepted (0: the decode path's default, 128 unless BVG_FUSE_MAX_NB says otherwise)
int conv_umma_fused_launch(const UmmaLayer& L, const C8T& x, const float* act_alpha, const float* act_beta,
                           const C8T& y, const UmmaEpilogue& ep, int64_t B, cudaStream_t st, int max_nb = 0);
// the same with both anti-alias FIRs on the tensor cores (actconv_tc.cu; C = 24 / 48 / 96).  `scratch` holds
// actconv_tc_scratch_bytes(B) bytes (the exact edge rows of the activation); BVG_ERR_STATE if the layer does not qualify
size_t actconv_tc_scratch_bytes(int64_t B);
int actconv_tc_launch(const UmmaLayer& L, const C8T& x, const float* act_alpha, const float* act_beta, const C8T& y,
                      const UmmaEpilogue& ep, int64_t B, void* scratch, cudaStream_t st);
int to_c8t_launch(const C8T& dst, const void* src, int64_t sb, int64_t sc, int64_t st_, int src_dtype, int64_t B, cudaStream_t st,
                  int reflect = 0);   // reflect: halo rows on each side that mirror the signal instead of being zero
int from_c8t_launch(void* dst, const C8T& src, int dst_dtype, int64_t B, cudaStream_t st);
// Activation1d on c8t tensors (writes the output's zero halo rows / padding channels too)
// impl: 0 = tensor-core FIRs when the tensor qualifies, else the CUDA-core stencil; 1 = stencil; 2 = tensor cores or error
int act1d_c8t_launch(const C8T& y, const C8T& x, const float* alpha_log, const float* beta_log, int64_t B,
                     cudaStream_t st, int impl = 0);
// the same on the tensor cores (act1d_tc.cu); BVG_ERR_STATE (nothing launched) if the tensor does not qualify
int act1d_tc_launch(const C8T& y, const C8T& x, const float* alpha_log, const float* beta_log, int64_t B, cudaStream_t st);
// conv_post (Cin->1) + tanh on a c8t tensor whose halo rows are zero; w is [Cin][K] fp32
int conv_post_c8t_launch(float* wav, int16_t* pcm, const C8T& x, const float* w, const float* bias, int K,
                         int64_t s_lo, int64_t s_hi, int64_t B, cudaStream_t st);

}  // namespace bvg
