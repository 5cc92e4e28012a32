// Plan (weights + layer graph) and the C ABI of libbigvgan_b200.
//
// A bvg_plan mirrors what `IndexTTS.__init__` builds at infer.py:61-67: the generator of
// BigVGAN/models.py:130-275 with weight norm already folded (models.py:252-260), every tensor
// supplied under its reference state-dict key.  bvg_decode() enqueues BigVGAN.forward
// (models.py:201-250) on the caller's stream: ECAPA speaker encoder -> conv_pre + cond ->
// 6 x (ConvTranspose1d + cond + mean of 3 AMPBlock1) -> Activation1d -> conv_post -> tanh.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "bvg_common.cuh"
#include "umma.cuh"

namespace bvg {

thread_local std::string g_error;
thread_local int64_t g_launches = 0;
long long* g_dbg_buf = nullptr;
// tensors below this many million elements stay on the CUDA-core stencil kernels (BVG_TC_MIN_MELEMS, default 10: a single 10 s
// utterance); bvg_debug_set_tc_min_melems overrides it (tests: the same kernels for single and batched decodes)
static std::atomic<int> g_tc_min_melems{-1};
int tc_min_melems() {
  int v = g_tc_min_melems.load(std::memory_order_relaxed);
  if (v < 0) { v = env_int_once("BVG_TC_MIN_MELEMS", 10); g_tc_min_melems.store(v, std::memory_order_relaxed); }
  return v;
}

// profiler state (per calling thread)
struct ProfRec { int kc; cudaEvent_t a, b; };
thread_local bool g_prof_on = false;
thread_local std::vector<ProfRec> g_prof_recs;
thread_local std::vector<cudaEvent_t> g_prof_pool;

void prof_mark(cudaStream_t st, int kclass, bool begin) {
  if (!g_prof_on) return;
  if (begin) {
    ProfRec r;
    r.kc = kclass;
    for (cudaEvent_t* e : {&r.a, &r.b}) {
      if (!g_prof_pool.empty()) { *e = g_prof_pool.back(); g_prof_pool.pop_back(); }
      else cudaEventCreate(e);
    }
    cudaEventRecord(r.a, st);
    g_prof_recs.push_back(r);
  } else if (!g_prof_recs.empty()) {
    cudaEventRecord(g_prof_recs.back().b, st);
  }
}

int current_device_sms(int* sms) {
  static std::atomic<int> cache[64];
  int dev = 0;
  BVG_CUDA(cudaGetDevice(&dev));
  int v = cache[dev & 63].load(std::memory_order_relaxed);
  if (v == 0) {
    BVG_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    cache[dev & 63].store(v, std::memory_order_relaxed);
  }
  *sms = v;
  return BVG_OK;
}

int env_int_once(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_error = buf;
}

struct ConvLayer {
  float* w = nullptr;     // [K][Cin][Cout] fp32
  float* bias = nullptr;  // [Cout] or null
  int Cin = 0, Cout = 0, K = 1;
  __nv_bfloat16* wu = nullptr;   // tcgen05 pack (conv_umma.cu), generator convs only
  __nv_bfloat16* wx3 = nullptr;  // 3-term split pack for the fp32x3 path (generator convs and transposed convs)
};
struct Tdnn {             // conv -> ReLU -> eval-BN (ECAPA_TDNN.py:126-128)
  ConvLayer conv;
  float* bn_scale = nullptr;
  float* bn_shift = nullptr;
  int dil = 1;
};
struct ResBlock {         // AMPBlock1 (models.py:20-80)
  ConvLayer c1[3], c2[3];
  float* alpha[6] = {};
  float* beta[6] = {};
  int K = 3;
  int dil[3] = {1, 3, 5};
};
struct SERes2 {           // SERes2NetBlock (ECAPA_TDNN.py:341-426)
  Tdnn tdnn1, r2n[7], tdnn2;
  ConvLayer se1, se2;
};

}  // namespace bvg

using namespace bvg;

struct bvg_plan {
  bvg_config cfg{};
  std::map<std::string, std::vector<float>> host;
  bool finalized = false;
  bool umma = false;        // bf16 tcgen05 packs built
  std::vector<void*> allocs;
  int n_stage = 0;
  int C[9] = {};          // C[0] = upsample_initial_channel, C[i+1] = channels after ups[i]
  int64_t total_up = 1;
  // generator
  ConvLayer conv_pre, cond_layer;
  std::vector<ConvLayer> ups, conds;
  std::vector<ResBlock> res;
  float *post_alpha = nullptr, *post_beta = nullptr, *post_w = nullptr, *post_bias = nullptr;
  // speaker encoder (ECAPA_TDNN.py:429-541, default channels [512,512,512,512,1536])
  Tdnn e_block0, e_mfa, e_asp_tdnn;
  SERes2 e_blk[3];
  ConvLayer e_asp_ctx, e_asp_conv, e_fc;
  float *asp_bn_scale = nullptr, *asp_bn_shift = nullptr;
  // small batches: the (up to three) AMP blocks of a stage read the same input and only meet in the final sum
  // (models.py:238-245), so they run on three streams -- one utterance does not fill 148 SMs with one block's kernels
  cudaStream_t side[2] = {};
  cudaEvent_t ev_fork = nullptr, ev_chain[3] = {};
  cudaEvent_t ev_h2d[2] = {};            // bvg_decode_host: the latent upload runs on side[0] beside the speaker encoder
  mutable std::atomic<int> h2d_busy{0};
  mutable std::atomic<int> ms_busy{0};   // held while one decode enqueues on the side streams (a second host thread decoding
                                         // on the same plan at that moment simply stays on its one stream)
};

namespace {

constexpr int kEC = 512;      // ECAPA trunk channels
constexpr int kEM = 1536;     // MFA channels
constexpr int kEA = 128;      // attention / SE bottleneck channels

int dev_alloc(bvg_plan* P, void** out, size_t bytes) {
  BVG_CUDA(cudaMalloc(out, std::max<size_t>(bytes, 16)));
  P->allocs.push_back(*out);
  return BVG_OK;
}

int upload(bvg_plan* P, const float* h, int64_t n, float** out) {
  void* d = nullptr;
  BVG_TRY(dev_alloc(P, &d, n * sizeof(float)));
  BVG_CUDA(cudaMemcpy(d, h, n * sizeof(float), cudaMemcpyHostToDevice));
  *out = static_cast<float*>(d);
  return BVG_OK;
}

int get_host(bvg_plan* P, const std::string& key, int64_t numel, const std::vector<float>** out) {
  auto it = P->host.find(key);
  if (it == P->host.end()) {
    set_error("plan: missing state-dict tensor '%s'", key.c_str());
    return BVG_ERR_STATE;
  }
  if ((int64_t)it->second.size() != numel) {
    set_error("plan: tensor '%s' has %lld elements, expected %lld", key.c_str(), (long long)it->second.size(),
              (long long)numel);
    return BVG_ERR_STATE;
  }
  *out = &it->second;
  return BVG_OK;
}

int upload_key(bvg_plan* P, const std::string& key, int64_t numel, float** out) {
  const std::vector<float>* v;
  BVG_TRY(get_host(P, key, numel, &v));
  return upload(P, v->data(), numel, out);
}

// weight key holds torch layout [Cout,Cin,K] (or [Cin,Cout,K] when transposed); optional input-channel slice
int make_conv(bvg_plan* P, const std::string& wkey, const std::string& bkey, int Cout, int Cin_total, int K,
              bool transposed, ConvLayer* L, int ci_lo = 0, int ci_hi = -1, int umma_nph = 0) {
  if (ci_hi < 0) ci_hi = Cin_total;
  const int Cin = ci_hi - ci_lo;
  const std::vector<float>* w;
  BVG_TRY(get_host(P, wkey, (int64_t)Cout * Cin_total * K, &w));
  float* raw = nullptr;
  const int64_t n = (int64_t)Cout * Cin * K;
  if (ci_lo == 0 && ci_hi == Cin_total) {
    BVG_CUDA(cudaMalloc(&raw, n * sizeof(float)));
    BVG_CUDA(cudaMemcpy(raw, w->data(), n * sizeof(float), cudaMemcpyHostToDevice));
  } else {
    if (transposed) { set_error("plan: sliced transposed conv unsupported"); return BVG_ERR_INVALID; }
    std::vector<float> tmp(n);
    for (int co = 0; co < Cout; ++co)
      memcpy(&tmp[(size_t)co * Cin * K], &(*w)[((size_t)co * Cin_total + ci_lo) * K], (size_t)Cin * K * sizeof(float));
    BVG_CUDA(cudaMalloc(&raw, n * sizeof(float)));
    BVG_CUDA(cudaMemcpy(raw, tmp.data(), n * sizeof(float), cudaMemcpyHostToDevice));
  }
  void* d = nullptr;
  int rc = dev_alloc(P, &d, n * sizeof(float));
  if (rc == BVG_OK) rc = repack_conv_weight_launch((float*)d, raw, Cout, Cin, K, transposed ? 1 : 0, 0);
  void* du = nullptr;
  if (rc == BVG_OK && umma_nph > 0) {
    rc = dev_alloc(P, &du, (size_t)umma_pack_elems(Cout, Cin, K, umma_nph) * 2);
    if (rc == BVG_OK) rc = umma_pack_launch((__nv_bfloat16*)du, raw, Cout, Cin, K, transposed ? 1 : 0, umma_nph, 0);
  }
  L->wu = (__nv_bfloat16*)du;
  void* dx = nullptr;
  if (rc == BVG_OK && umma_nph > 0 && 2 * K <= (transposed ? 16 : 32)) {
    rc = dev_alloc(P, &dx, (size_t)umma_pack_split_elems(Cout, Cin, K, umma_nph) * 2);
    if (rc == BVG_OK) rc = umma_pack_split_launch((__nv_bfloat16*)dx, raw, Cout, Cin, K, transposed ? 1 : 0, umma_nph, 0);
  }
  L->wx3 = (__nv_bfloat16*)dx;
  cudaError_t e = cudaStreamSynchronize(0);
  cudaFree(raw);
  if (rc != BVG_OK) return rc;
  BVG_CUDA(e);
  L->w = (float*)d;
  L->Cin = Cin; L->Cout = Cout; L->K = K;
  L->bias = nullptr;
  if (!bkey.empty()) BVG_TRY(upload_key(P, bkey, Cout, &L->bias));
  return BVG_OK;
}

// eval BatchNorm1d folded to y = x*scale + shift (nnet/normalization.py:75-108, eps 1e-5)
int make_bn(bvg_plan* P, const std::string& p, int Cn, float** scale, float** shift) {
  const std::vector<float>*w, *b, *m, *v;
  BVG_TRY(get_host(P, p + ".weight", Cn, &w));
  BVG_TRY(get_host(P, p + ".bias", Cn, &b));
  BVG_TRY(get_host(P, p + ".running_mean", Cn, &m));
  BVG_TRY(get_host(P, p + ".running_var", Cn, &v));
  std::vector<float> sc(Cn), sh(Cn);
  for (int i = 0; i < Cn; ++i) {
    sc[i] = (*w)[i] / sqrtf((*v)[i] + 1e-5f);
    sh[i] = (*b)[i] - (*m)[i] * sc[i];
  }
  BVG_TRY(upload(P, sc.data(), Cn, scale));
  BVG_TRY(upload(P, sh.data(), Cn, shift));
  return BVG_OK;
}

int make_tdnn(bvg_plan* P, const std::string& p, int Cin, int Cout, int K, int dil, Tdnn* t, int ci_lo = 0,
              int ci_hi = -1, int umma_nph = 0) {
  BVG_TRY(make_conv(P, p + ".conv.conv.weight", p + ".conv.conv.bias", Cout, Cin, K, false, &t->conv, ci_lo, ci_hi, umma_nph));
  BVG_TRY(make_bn(P, p + ".norm.norm", Cout, &t->bn_scale, &t->bn_shift));
  t->dil = dil;
  return BVG_OK;
}

// Activation1d buffers must be the compiled-in kaiser-sinc taps (they are persistent buffers in the
// reference state dict: alias_free_torch/resample.py:22, filter.py:82)
int check_taps(bvg_plan* P, const std::string& key) {
  static const float ref[12] = {BVG_F0, BVG_F1, BVG_F2, BVG_F3, BVG_F4, BVG_F5,
                                BVG_F5, BVG_F4, BVG_F3, BVG_F2, BVG_F1, BVG_F0};
  auto it = P->host.find(key);
  if (it == P->host.end()) return BVG_OK;  // buffers are optional (older checkpoints omit them)
  if (it->second.size() != 12) { set_error("plan: '%s' must have 12 taps", key.c_str()); return BVG_ERR_STATE; }
  for (int i = 0; i < 12; ++i)
    if (fabsf(it->second[i] - ref[i]) > 1e-6f) {
      set_error("plan: '%s' differs from kaiser_sinc_filter1d(0.25,0.3,12) at tap %d (%g vs %g); only the "
                "reference's default Activation1d filter is supported", key.c_str(), i, it->second[i], ref[i]);
      return BVG_ERR_STATE;
    }
  return BVG_OK;
}

int make_act(bvg_plan* P, const std::string& p, int Cn, float** alpha, float** beta) {
  const std::vector<float>*a, *b;
  BVG_TRY(get_host(P, p + ".act.alpha", Cn, &a));
  BVG_TRY(get_host(P, p + ".act.beta", Cn, &b));
  BVG_TRY(check_taps(P, p + ".upsample.filter"));
  BVG_TRY(check_taps(P, p + ".downsample.lowpass.filter"));
  std::vector<float> la(*a), lb(*b);
  if (!P->cfg.snake_logscale) {  // kernel applies exp: cancel it (cuda/activation1d.py:67-71)
    for (auto& x : la) x = logf(x);
    for (auto& x : lb) x = logf(x);
  }
  BVG_TRY(upload(P, la.data(), Cn, alpha));
  BVG_TRY(upload(P, lb.data(), Cn, beta));
  return BVG_OK;
}

struct Bump {
  char* base;
  size_t cap, off = 0;
  bool overflow = false;
  Bump(void* p, size_t c) : base(static_cast<char*>(p)), cap(c) {}
  void* take(size_t bytes) {
    off = (off + 255) & ~size_t(255);
    void* r = base ? base + off : nullptr;
    off += bytes;
    if (off > cap) overflow = true;
    return r;
  }
  float* takef(int64_t n) { return static_cast<float*>(take((size_t)n * 4)); }
};

struct EcapaWs {
  float *X0, *Y1, *Y2, *Y3, *XL, *M, *A1, *A2, *sem, *se1, *se2, *ms, *actx, *pooled;
  void *cx, *cy;           // c8t bf16 staging for the GEMM-shaped layers on the tensor-core path
};
void carve_ecapa(Bump& b, int64_t Bm, int64_t Tm, EcapaWs* w) {
  w->X0 = b.takef(Bm * kEC * Tm); w->Y1 = b.takef(Bm * kEC * Tm); w->Y2 = b.takef(Bm * kEC * Tm);
  w->Y3 = b.takef(Bm * kEC * Tm); w->XL = b.takef(Bm * kEM * Tm); w->M = b.takef(Bm * kEM * Tm);
  w->A1 = b.takef(Bm * kEA * Tm); w->A2 = b.takef(Bm * kEM * Tm);
  w->sem = b.takef(Bm * kEC); w->se1 = b.takef(Bm * kEA); w->se2 = b.takef(Bm * kEC);
  w->ms = b.takef(Bm * 2 * kEM); w->actx = b.takef(Bm * kEA); w->pooled = b.takef(Bm * 2 * kEM);
  w->cx = b.take(c8t_bytes(Bm, kEM, Tm)); w->cy = b.take(c8t_bytes(Bm, kEM, Tm));
}

// conv on [B,C,T]-contiguous fp32 tensors viewed through explicit strides
int conv_f32(float* dst, int64_t dsb, const float* src, const float* src2, int64_t sb, int64_t sc, int64_t st_,
             const ConvLayer& L, ConvEpilogue ep, int64_t B, int64_t T, int dil, int pad_mode, cudaStream_t st) {
  ep.bias = L.bias;
  // the T == 1 layers (cond vectors, SE, attentive-statistics context, fc) are matrix-vector products
  if (T == 1 && L.K == 1 && !src2 && sc == 1 && sb == L.Cin && dsb == L.Cout && !ep.cond && !ep.res1 && !ep.res2)
    return matvec_launch(dst, src, L.w, ep, B, L.Cin, L.Cout, st);
  return conv1d_simt_launch(dst, dsb, src, src2, sb, sc, st_, L.w, ep, B, L.Cin, L.Cout, T, L.K, dil, pad_mode,
                            BVG_F32, BVG_F32, st);
}
int tdnn_f32(float* dst, int64_t dsb, const float* src, const float* src2, int64_t sb, int64_t sc, int64_t st_,
             const Tdnn& t, int64_t B, int64_t T, cudaStream_t st, const float* cond = nullptr, int act = 0) {
  ConvEpilogue ep;
  ep.relu = 1; ep.post_scale = t.bn_scale; ep.post_shift = t.bn_shift; ep.act = act;
  ep.cond = cond; ep.cond_B = B; ep.prof_other = 1;
  return conv_f32(dst, dsb, src, src2, sb, sc, st_, t.conv, ep, B, T, t.dil, 1, st);
}

// A 1x1 TDNN layer (a plain GEMM over B*T columns) on the tcgen05 path: c8t bf16 input, bias / cond / ReLU / folded-BN /
// tanh epilogue, output as fp32 [B,C,T] (dst) and / or c8t bf16 (dstc).  Used by the bf16 decode path only.
int tdnn_umma(float* dst, const C8T* dstc, const C8T& x, const ConvLayer& L, const float* bn_scale, const float* bn_shift,
              int relu, int act, const float* cond, int64_t B, int64_t T, cudaStream_t st) {
  UmmaLayer u;
  u.w = L.wu; u.Cin = L.Cin; u.Cout = L.Cout; u.K = 1; u.dil = 1;
  UmmaEpilogue ep;
  ep.bias = L.bias; ep.cond = cond; ep.cond_B = B; ep.relu = relu; ep.post_scale = bn_scale; ep.post_shift = bn_shift;
  ep.act = act; ep.prof_other = 1; ep.yf32 = dst;
  return conv_umma_launch(u, x, dstc ? *dstc : make_c8t(nullptr, L.Cout, (int)T), ep, B, st);
}

// Res2NetBlock (ECAPA_TDNN.py:179-191) through ecapa.cu's chain kernel when the block has its standard geometry
int res2net_chain(const SERes2& S, const float* y1, float* y2, const C8T* yc, int64_t B, int64_t T, bool* taken,
                  cudaStream_t st) {
  *taken = false;
  if (BVG_ENV_ONCE("BVG_ECAPA_CHAIN", 1) == 0) return BVG_OK;
  const float *w[7], *bias[7], *sc[7], *sh[7];
  for (int j = 0; j < 7; ++j) {
    const Tdnn& t = S.r2n[j];
    if (t.conv.K != 3 || t.conv.Cin != kEC / 8 || t.conv.Cout != kEC / 8 || t.dil != S.r2n[0].dil || !t.conv.bias ||
        !t.bn_scale || !t.bn_shift)
      return BVG_OK;
    w[j] = t.conv.w; bias[j] = t.conv.bias; sc[j] = t.bn_scale; sh[j] = t.bn_shift;
  }
  return res2net_chain_launch(y1, y2, yc, w, bias, sc, sh, S.r2n[0].dil, B, T, taken, st);
}

// ECAPA_TDNN.forward (ECAPA_TDNN.py:543-581), lengths=None.  mel [Bm,Tm,num_mels] -> spk [Bm,E]
int ecapa_forward(const bvg_plan* P, const float* mel, int64_t Bm, int64_t Tm, float* spk, const EcapaWs& w,
                  cudaStream_t st, bool tc = false) {
  const int NM = P->cfg.num_mels;
  BVG_CHECK_ARG(Tm >= 5, "speaker encoder: reference mel needs >= 5 frames for reflect padding (got %lld)", (long long)Tm);
  // x.transpose(1,2): read channels-last mel through strides
  const bool b0_tc = tc && P->e_block0.conv.wu && BVG_ENV_ONCE("BVG_ECAPA_B0_TC", 1) != 0;
  if (!b0_tc) BVG_TRY(tdnn_f32(w.X0, kEC * Tm, mel, nullptr, Tm * NM, 1, NM, P->e_block0, Bm, Tm, st));
  const float* X = w.X0;
  int64_t Xsb = kEC * Tm;
  // tensor-core path: activations meet the 1x1 GEMMs as c8t bf16 tensors written by their producers (no fp32 -> c8t ->
  // fp32 round trips): XLc = the three block outputs side by side (the MFA input), Mc = the MFA output; the block-0
  // output, the Res2Net chain output and the attention hidden layer borrow whichever of the two is idle.
  C8T XLc = make_c8t(w.cx, kEM, (int)Tm), Mc = make_c8t(w.cy, kEM, (int)Tm);
  C8T X0c = make_c8t(w.cy, kEC, (int)Tm);
  C8T Y2c = make_c8t(static_cast<char*>(w.cy) + c8t_bytes(Bm, kEC, Tm), kEC, (int)Tm);
  C8T A1c = make_c8t(w.cx, kEA, (int)Tm);
  auto xl_slice = [&](int i) {                                   // channels [512 i, 512 (i + 1)) of XLc
    C8T v = make_c8t(XLc.p + (int64_t)i * (kEC / 8) * XLc.Tp * 8, kEC, (int)Tm);
    v.bstride = XLc.batch_stride();
    return v;
  };
  if (b0_tc) {
    // blocks[0] (k = 5 over the mel bins, reflect padding) on the tensor cores as well: the mel goes to c8t bf16 with its
    // two reflected rows per side in the halo; the output is written in both forms (fp32 residual, c8t GEMM input)
    const Tdnn& t0 = P->e_block0;
    const int rp = t0.dil * (t0.conv.K - 1) / 2;
    C8T melc = make_c8t(w.cx, NM, (int)Tm);
    BVG_TRY(to_c8t_launch(melc, mel, Tm * NM, 1, NM, BVG_F32, Bm, st, rp));
    UmmaLayer u;
    u.w = t0.conv.wu; u.Cin = t0.conv.Cin; u.Cout = t0.conv.Cout; u.K = t0.conv.K; u.dil = t0.dil;
    UmmaEpilogue ep;
    ep.bias = t0.conv.bias; ep.relu = 1; ep.post_scale = t0.bn_scale; ep.post_shift = t0.bn_shift; ep.prof_other = 1;
    ep.yf32 = w.X0;
    BVG_TRY(conv_umma_launch(u, melc, X0c, ep, Bm, st));
  } else if (tc) {
    BVG_TRY(to_c8t_launch(X0c, w.X0, kEC * Tm, Tm, 1, BVG_F32, Bm, st));
  }
  for (int i = 0; i < 3; ++i) {
    const SERes2& S = P->e_blk[i];
    if (tc) BVG_TRY(tdnn_umma(w.Y1, nullptr, i == 0 ? X0c : xl_slice(i - 1), S.tdnn1.conv, S.tdnn1.bn_scale, S.tdnn1.bn_shift, 1, 0, nullptr, Bm, Tm, st));
    else BVG_TRY(tdnn_f32(w.Y1, kEC * Tm, X, nullptr, Xsb, Tm, 1, S.tdnn1, Bm, Tm, st));
    // Res2NetBlock :179-191: 8 chunks of 64 channels, y_i = f(x_i + y_{i-1})
    const int CH = kEC / 8;
    bool chained = false;
    BVG_TRY(res2net_chain(S, w.Y1, tc ? nullptr : w.Y2, tc ? &Y2c : nullptr, Bm, Tm, &chained, st));   // one cluster launch
    if (!chained) {
      BVG_CUDA(cudaMemcpy2DAsync(w.Y2, kEC * Tm * 4, w.Y1, kEC * Tm * 4, (size_t)CH * Tm * 4, Bm,
                                 cudaMemcpyDeviceToDevice, st));
      for (int j = 1; j < 8; ++j) {
        const float* s2 = (j >= 2) ? w.Y2 + (int64_t)(j - 1) * CH * Tm : nullptr;
        BVG_TRY(tdnn_f32(w.Y2 + (int64_t)j * CH * Tm, kEC * Tm, w.Y1 + (int64_t)j * CH * Tm, s2, kEC * Tm, Tm, 1,
                         S.r2n[j - 1], Bm, Tm, st));
      }
      if (tc) BVG_TRY(to_c8t_launch(Y2c, w.Y2, kEC * Tm, Tm, 1, BVG_F32, Bm, st));
    }
    if (tc) BVG_TRY(tdnn_umma(w.Y3, nullptr, Y2c, S.tdnn2.conv, S.tdnn2.bn_scale, S.tdnn2.bn_shift, 1, 0, nullptr, Bm, Tm, st));
    else BVG_TRY(tdnn_f32(w.Y3, kEC * Tm, w.Y2, nullptr, kEC * Tm, Tm, 1, S.tdnn2, Bm, Tm, st));
    // SEBlock :228-242
    BVG_TRY(row_mean_launch(w.sem, w.Y3, Bm * kEC, Tm, st));
    ConvEpilogue e1; e1.relu = 1;
    BVG_TRY(conv_f32(w.se1, kEA, w.sem, nullptr, kEC, 1, 1, S.se1, e1, Bm, 1, 1, 0, st));
    ConvEpilogue e2; e2.act = 2;
    BVG_TRY(conv_f32(w.se2, kEC, w.se1, nullptr, kEA, 1, 1, S.se2, e2, Bm, 1, 1, 0, st));
    float* out = w.XL + (int64_t)i * kEC * Tm;
    C8T outc = xl_slice(i);
    BVG_TRY(scale_residual_launch(out, kEM * Tm, w.se2, w.Y3, X, Xsb, Bm, kEC, Tm, st, tc ? &outc : nullptr));
    X = out; Xsb = kEM * Tm;
  }
  if (tc) BVG_TRY(tdnn_umma(w.M, &Mc, XLc, P->e_mfa.conv, P->e_mfa.bn_scale, P->e_mfa.bn_shift, 1, 0, nullptr, Bm, Tm, st));
  else BVG_TRY(tdnn_f32(w.M, kEM * Tm, w.XL, nullptr, kEM * Tm, Tm, 1, P->e_mfa, Bm, Tm, st));
  // AttentiveStatisticsPooling :282-338.  The tdnn over cat([x, mean, std]) splits into a conv over x
  // plus a per-utterance term from (mean, std).
  BVG_TRY(row_stats_launch(w.ms, w.M, Bm, kEM, Tm, st));
  ConvEpilogue ec;
  BVG_TRY(conv_f32(w.actx, kEA, w.ms, nullptr, 2 * kEM, 1, 1, P->e_asp_ctx, ec, Bm, 1, 1, 0, st));
  if (tc) {
    BVG_TRY(tdnn_umma(nullptr, &A1c, Mc, P->e_asp_tdnn.conv, P->e_asp_tdnn.bn_scale, P->e_asp_tdnn.bn_shift, 1,
                      /*tanh*/ 1, w.actx, Bm, Tm, st));
    BVG_TRY(tdnn_umma(w.A2, nullptr, A1c, P->e_asp_conv, nullptr, nullptr, 0, 0, nullptr, Bm, Tm, st));
  } else {
    BVG_TRY(tdnn_f32(w.A1, kEA * Tm, w.M, nullptr, kEM * Tm, Tm, 1, P->e_asp_tdnn, Bm, Tm, st, w.actx, /*tanh*/ 1));
    ConvEpilogue ea;
    BVG_TRY(conv_f32(w.A2, kEM * Tm, w.A1, nullptr, kEA * Tm, Tm, 1, P->e_asp_conv, ea, Bm, Tm, 1, 0, st));
  }
  BVG_TRY(attn_stats_launch(w.pooled, w.A2, w.M, P->asp_bn_scale, P->asp_bn_shift, Bm, kEM, Tm, st));
  ConvEpilogue ef;
  BVG_TRY(conv_f32(spk, P->cfg.speaker_embedding_dim, w.pooled, nullptr, 2 * kEM, 1, 1, P->e_fc, ef, Bm, 1, 1, 0, st));
  return BVG_OK;
}

struct GenWs {
  void *A, *Y, *T1, *T2, *XS;
  void* edge;               // actconv_tc_launch scratch (exact edge rows of the fused activation)
  void *Yx[2], *T1x[2], *T2x[2], *edgex[2], *X3x[2];   // private buffers of AMP blocks 1 and 2 when they run on their own streams
  void* X3;                 // [hi | lo] c8t staging of one conv input (fp32x3 path)
  float* cond[9];
  float* spk;
  EcapaWs e;
  float *latent_dev, *mel_dev, *wav_dev;
  int16_t* pcm_dev;
};

// size of each of the generator's rotating activation buffers: the largest c8t tensor on the bf16 tensor-core path, the
// largest plain [B, C, T] tensor otherwise
size_t gen_buf_bytes(const bvg_plan* P, int64_t B, int64_t T0, int dtype) {
  int64_t maxel = (int64_t)P->C[0] * T0;
  size_t c8 = std::max(c8t_bytes(B, P->C[0], T0), c8t_bytes(B, P->cfg.gpt_dim, T0));
  int64_t T = T0;
  for (int i = 0; i < P->n_stage; ++i) {
    T *= P->cfg.upsample_rates[i];
    maxel = std::max<int64_t>(maxel, (int64_t)P->C[i + 1] * T);
    c8 = std::max(c8, c8t_bytes(B, P->C[i + 1], T));
  }
  return (dtype == BVG_BF16 && P->umma) ? c8 : (size_t)B * maxel * dtype_size(dtype);
}
// [hi | lo] staging buffer of the fp32x3 path
size_t gen_x3_bytes(const bvg_plan* P, int64_t B, int64_t T0) {
  size_t x3 = c8t_bytes(B, 2 * ((P->cfg.gpt_dim + 7) / 8 * 8), T0);
  int64_t Tx = T0;
  for (int i = 0; i < P->n_stage; ++i) {
    Tx *= P->cfg.upsample_rates[i];
    x3 = std::max(x3, c8t_bytes(B, 2 * ((P->C[i] + 7) / 8 * 8), Tx / P->cfg.upsample_rates[i]));   // ups[i] input
    x3 = std::max(x3, c8t_bytes(B, 2 * ((P->C[i + 1] + 7) / 8 * 8), Tx));
  }
  return x3;
}

// AMP blocks of a stage on separate streams (all three precisions), at most three blocks.  Measured (B200, 10 s utterances):
// B = 1 4.9 -> 3.6 ms, B = 4 +24 %, B = 8 +18 %, B = 16 +11 %, B = 32 +5 % audio-s/s -- while one block's persistent kernel
// ramps up, drains or leaves SMs idle in its last round of tiles, the other blocks' kernels take the free SMs.  The price is
// six more rotating buffers; BVG_MS_MAX_EXTRA_GB (default 32) keeps huge batches on one stream, BVG_MS_MAX_FRAMES (latent
// frames in the batch) and BVG_MULTI_STREAM=0 exist for A/B runs.
std::atomic<int> g_ms_enable{1};     // bvg_debug_set_multi_stream
bool multi_stream(const bvg_plan* P, int64_t B, int64_t T0, int dtype, bool for_workspace = false) {
  if (!(P->side[0] && P->cfg.num_kernels >= 2 && P->cfg.num_kernels <= 3)) return false;
  if (!for_workspace && !g_ms_enable.load()) return false;     // (workspaces are always sized for the multi-stream form)
  if (B * T0 > BVG_ENV_ONCE("BVG_MS_MAX_FRAMES", 1 << 30)) return false;
  const double extra = 6.0 * (double)gen_buf_bytes(P, B, T0, dtype) + (dtype == BVG_F32X3 ? 2.0 * (double)gen_x3_bytes(P, B, T0) : 0.0);
  return extra <= (double)BVG_ENV_ONCE("BVG_MS_MAX_EXTRA_GB", 32) * 1073741824.0;
}

void carve_gen(const bvg_plan* P, Bump& b, int64_t B, int64_t T0, int64_t Bm, int64_t Tm, int dtype, GenWs* g) {
  const size_t bytes = gen_buf_bytes(P, B, T0, dtype);
  const bool ms = multi_stream(P, B, T0, dtype, /*for_workspace=*/true);
  g->A = b.take(bytes); g->Y = b.take(bytes); g->T1 = b.take(bytes); g->T2 = b.take(bytes); g->XS = b.take(bytes);
  g->edge = b.take(actconv_tc_scratch_bytes(B));
  for (int j = 0; j < 2; ++j) {
    g->Yx[j] = ms ? b.take(bytes) : nullptr; g->T1x[j] = ms ? b.take(bytes) : nullptr; g->T2x[j] = ms ? b.take(bytes) : nullptr;
    g->edgex[j] = ms ? b.take(actconv_tc_scratch_bytes(B)) : nullptr;
    g->X3x[j] = nullptr;
  }
  g->X3 = nullptr;
  if (dtype == BVG_F32X3) {
    const size_t x3 = gen_x3_bytes(P, B, T0);
    g->X3 = b.take(x3);
    if (ms) for (int j = 0; j < 2; ++j) g->X3x[j] = b.take(x3);
  }
  for (int i = 0; i <= P->n_stage; ++i) g->cond[i] = b.takef(Bm * P->C[i]);
  g->spk = b.takef(Bm * P->cfg.speaker_embedding_dim);
  carve_ecapa(b, Bm, Tm, &g->e);
  // staging for the host-buffer entry point
  g->latent_dev = b.takef(B * T0 * P->cfg.gpt_dim);
  g->mel_dev = b.takef(Bm * Tm * P->cfg.num_mels);
  g->wav_dev = b.takef(B * T0 * P->total_up);
  g->pcm_dev = static_cast<int16_t*>(b.take((size_t)B * T0 * P->total_up * 2));
}

// precise: libdevice sinf (the fp32 parity path); otherwise the packed fast-math stencil (MUFU cosine), which the
// fp32x3 path can afford inside its 1e-4 budget (BVG_X3_PRECISE_ACT=1 keeps sinf there as well)
int act_launch(void* dst, const void* src, const float* a, const float* b_, int64_t B, int64_t Cn, int64_t T,
               int dtype, cudaStream_t st, bool fast_f32 = false) {
  return act1d_launch(dst, src, a, b_, B, Cn, T, dtype, /*precise=*/(dtype == BVG_F32 && !fast_f32) ? 1 : 0, st);
}

// fp32 conv on the tensor cores: split the fp32 input into [hi | lo] bf16 halves, run the tcgen05 conv on the 3-term
// split weights, fp32 output with the bias / cond / residual / scale epilogue (UmmaLayer::split)
int gen_conv_x3(float* dst, const float* src, int64_t sb, int64_t sc, int64_t st_, const ConvLayer& L, const ConvEpilogue& ep,
                int64_t B, int64_t T, int dil, void* x3buf, cudaStream_t st) {
  C8T xs = make_c8t(x3buf, 2 * ((L.Cin + 7) / 8 * 8), (int)T);
  if (src) BVG_TRY(split_to_c8t_launch(xs, src, sb, sc, st_, L.Cin, B, st));   // (null: x3buf already holds the split input)
  UmmaLayer u;
  u.w = L.wx3; u.Cin = L.Cin; u.Cout = L.Cout; u.K = L.K; u.dil = dil; u.split = 1;
  UmmaEpilogue e;
  e.bias = L.bias; e.cond = ep.cond; e.cond_B = ep.cond_B; e.scale = ep.scale;
  e.yf32 = dst; e.res1_f32 = static_cast<const float*>(ep.res1); e.res2_f32 = static_cast<const float*>(ep.res2);
  return conv_umma_launch(u, xs, make_c8t(nullptr, L.Cout, (int)T), e, B, st);
}

// fp32x3: Activation1d written straight into the split c8t tensor, then the split conv (no fp32 round trip through HBM)
int act_conv_x3(float* dst, const float* src, const float* alpha, const float* beta, const ConvLayer& L, ConvEpilogue ep,
                int64_t B, int64_t T, int dil, void* x3buf, cudaStream_t st) {
  C8T xs = make_c8t(x3buf, 2 * ((L.Cin + 7) / 8 * 8), (int)T);
  const int rc = act1d_split_launch(xs.p, xs.chunks, xs.Tp, xs.pad, src, alpha, beta, B, L.Cin, T, st);
  if (rc != BVG_OK) return rc;                      // BVG_ERR_STATE: shape does not qualify, the caller takes the 2-step route
  ep.bias = L.bias;
  return gen_conv_x3(dst, nullptr, 0, 0, 0, L, ep, B, T, dil, x3buf, st);
}

int gen_conv(void* dst, const void* src, const ConvLayer& L, ConvEpilogue ep, int64_t B, int64_t T, int dil,
             int dtype, cudaStream_t st, void* x3buf = nullptr) {
  ep.bias = L.bias;
  if (x3buf && L.wx3)
    return gen_conv_x3((float*)dst, (const float*)src, (int64_t)L.Cin * T, T, 1, L, ep, B, T, dil, x3buf, st);
  return conv1d_simt_launch(dst, (int64_t)L.Cout * T, src, nullptr, (int64_t)L.Cin * T, T, 1, L.w, ep, B, L.Cin,
                            L.Cout, T, L.K, dil, 0, dtype, dtype, st);
}

// AMPBlock1's act->conv pairs of the narrow stages run as one kernel (conv_umma_fused_kernel: the activated tensor
// never exists in HBM); layers that do not qualify fall back to the Activation1d kernel + conv kernel.
// BVG_FUSE=0 forces the two-kernel path everywhere (A/B measurements, tests).
const bool g_fuse_act = [] { const char* e = getenv("BVG_FUSE"); return !(e && e[0] == '0'); }();
// BVG_FUSE_TC=0 keeps the round-1 fused kernel (FIRs on the FP32 pipe) for every fused layer (A/B measurements)
const bool g_fuse_tc = [] { const char* e = getenv("BVG_FUSE_TC"); return !(e && e[0] == '0'); }();

// Activation1d -> Conv1d as one kernel: tensor-core FIRs when the layer qualifies, else the CUDA-core stencil version
int fused_act_conv(const UmmaLayer& L, const C8T& x, const float* alpha, const float* beta, const C8T& y, const UmmaEpilogue& ep,
                   int64_t B, void* scratch, cudaStream_t st) {
  if (!g_fuse_act) return BVG_ERR_STATE;
  // (BVG_FUSE_TC_MINC: narrowest layer that takes the tensor-core FIR kernel; A/B measurements)
  const bool big = B * (int64_t)L.Cin * x.T >= (int64_t)tc_min_melems() * 1000000;   // (as act1d_c8t_launch)
  if (g_fuse_tc && big && L.Cin >= BVG_ENV_ONCE("BVG_FUSE_TC_MINC", 24)) {
    const int rc = actconv_tc_launch(L, x, alpha, beta, y, ep, B, scratch, st);
    if (rc != BVG_ERR_STATE) return rc;
  }
  return conv_umma_fused_launch(L, x, alpha, beta, y, ep, B, st);
}

UmmaLayer ulayer(const ConvLayer& L, int dil, int transposed = 0, int stride = 1) {
  UmmaLayer u;
  u.w = L.wu; u.bias = L.bias; u.Cin = L.Cin; u.Cout = L.Cout; u.K = L.K; u.dil = dil;
  u.transposed = transposed; u.stride = stride;
  return u;
}

// The bf16 throughput path: c8t activations, tcgen05 convs (models.py:220-248).
int decode_bf16_umma(const bvg_plan* P, const void* latent, int latent_dtype, const GenWs& g, int64_t B, int64_t T0, int64_t Bm,
                     float* wav, int16_t* pcm16, int64_t t_lo_pad, int64_t t_hi_pad, cudaStream_t st0, const int* lens = nullptr) {
  const bvg_config& c = P->cfg;
  const bool ms = multi_stream(P, B, T0, BVG_BF16) && P->ms_busy.exchange(1) == 0;
  struct Release { const bvg_plan* p; bool on; ~Release() { if (on) p->ms_busy.store(0); } } release{P, ms};
  // ragged batches (bvg_decode_varlen): `lens` = device int32 [B] latent-frame counts; every tensor of a stage carries them
  // with the stage's cumulative upsampling factor, the kernels clip at each utterance's own length (umma.cuh, C8T::lens)
  auto make_c8t = [&](void* p, int C, int T) {
    C8T t = bvg::make_c8t(p, C, T);
    t.lens = lens; t.len_mul = (int)(T / T0);
    return t;
  };
  // latent [B,T0,gpt_dim] fp32 (channels-last) -> c8t bf16
  C8T lat = make_c8t(g.T1, c.gpt_dim, (int)T0);
  // (the GPT hands its latent over in its autocast dtype, gpt/model.py:462-477 under infer.py:194: fp16 / bf16 / fp32 are
  // all ingested directly, [B, T, C] channels-last, no fp32 round trip)
  BVG_TRY(to_c8t_launch(lat, latent, T0 * c.gpt_dim, 1, c.gpt_dim, latent_dtype, B, st0));
  int64_t T = T0;
  C8T xs = make_c8t(g.XS, P->C[0], (int)T);
  {
    UmmaEpilogue ep;
    ep.bias = P->conv_pre.bias; ep.cond = g.cond[0]; ep.cond_B = Bm; ep.zero_pads = 1;
    BVG_TRY(conv_umma_launch(ulayer(P->conv_pre, 1), lat, xs, ep, B, st0));
  }
  const float inv_nk = 1.0f / (float)c.num_kernels;
  for (int i = 0; i < P->n_stage; ++i) {
    const int u = c.upsample_rates[i];
    const int ch = P->C[i + 1];
    const int64_t Tn = T * u;
    C8T a = make_c8t(g.A, ch, (int)Tn);
    {
      UmmaEpilogue ep;
      ep.bias = P->ups[i].bias;
      if (c.cond_in_each_up_layer) { ep.cond = g.cond[i + 1]; ep.cond_B = Bm; }
      BVG_TRY(conv_umma_launch(ulayer(P->ups[i], 1, 1, u), xs, a, ep, B, st0));
    }
    T = Tn;
    xs = make_c8t(g.XS, ch, (int)T);
    if (ms) {                                              // fork: the side streams start once `a` is complete
      BVG_CUDA(cudaEventRecord(P->ev_fork, st0));
      for (int j = 1; j < c.num_kernels; ++j) BVG_CUDA(cudaStreamWaitEvent(P->side[j - 1], P->ev_fork, 0));
    }
    for (int j = 0; j < c.num_kernels; ++j) {
      const ResBlock& R = P->res[(size_t)i * c.num_kernels + j];
      const bool own = ms && j > 0;                        // block j on its own stream with its own temporaries
      cudaStream_t st = own ? P->side[j - 1] : st0;
      C8T y = make_c8t(own ? g.Yx[j - 1] : g.Y, ch, (int)T), t1 = make_c8t(own ? g.T1x[j - 1] : g.T1, ch, (int)T),
          t2 = make_c8t(own ? g.T2x[j - 1] : g.T2, ch, (int)T);
      void* edge = own ? g.edgex[j - 1] : g.edge;
      const C8T* cur = &a;
      for (int m = 0; m < 3; ++m) {
        // xt = c1(a1(x)): one fused kernel for narrow layers, Activation1d kernel + conv kernel otherwise
        UmmaEpilogue e1;
        e1.bias = R.c1[m].bias;
        int rc = fused_act_conv(ulayer(R.c1[m], R.dil[m]), *cur, R.alpha[2 * m], R.beta[2 * m], t2, e1, B, edge, st);
        if (rc == BVG_ERR_STATE) {
          BVG_TRY(act1d_c8t_launch(t1, *cur, R.alpha[2 * m], R.beta[2 * m], B, st));
          rc = conv_umma_launch(ulayer(R.c1[m], R.dil[m]), t1, t2, e1, B, st);
        }
        BVG_TRY(rc);
        // x = c2(a2(xt)) + x   (and the 3-block sum / mean on the last pair)
        UmmaEpilogue e2;
        e2.bias = R.c2[m].bias;
        e2.res1 = cur->p;
        const C8T* out = &y;
        if (m == 2) {
          if (j > 0) e2.res2 = xs.p;
          if (j == c.num_kernels - 1) e2.scale = inv_nk;
          e2.zero_pads = 1;
          out = &xs;
          // the running sum lives in xs: block j's last conv reads what block j - 1's last conv wrote (same order of
          // additions as on one stream, so the result is bit-identical)
          if (ms && j > 0) BVG_CUDA(cudaStreamWaitEvent(st, P->ev_chain[j - 1], 0));
        }
        rc = fused_act_conv(ulayer(R.c2[m], 1), t2, R.alpha[2 * m + 1], R.beta[2 * m + 1], *out, e2, B, edge, st);
        if (rc == BVG_ERR_STATE) {
          BVG_TRY(act1d_c8t_launch(t1, t2, R.alpha[2 * m + 1], R.beta[2 * m + 1], B, st));
          rc = conv_umma_launch(ulayer(R.c2[m], 1), t1, *out, e2, B, st);
        }
        BVG_TRY(rc);
        if (m < 2) cur = &y;
      }
      if (ms) BVG_CUDA(cudaEventRecord(P->ev_chain[j], st));
    }
    if (ms) BVG_CUDA(cudaStreamWaitEvent(st0, P->ev_chain[c.num_kernels - 1], 0));   // join (the chain events are transitive)
  }
  cudaStream_t st = st0;
  const int chp = P->C[P->n_stage];
  C8T t1 = make_c8t(g.T1, chp, (int)T);
  BVG_TRY(act1d_c8t_launch(t1, xs, P->post_alpha, P->post_beta, B, st));
  BVG_TRY(conv_post_c8t_launch(wav, pcm16, t1, P->post_w, P->post_bias, 7, t_lo_pad * P->total_up,
                               t_hi_pad * P->total_up, B, st));
  return BVG_OK;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char* bvg_last_error(void) { return g_error.c_str(); }
const char* bvg_version(void) { return "bigvgan_b200 0.1 sm_100a"; }
int64_t bvg_launch_count(void) { return g_launches; }
void bvg_launch_count_reset(void) { g_launches = 0; }

void bvg_debug_set_tc_min_melems(int melems) { bvg::g_tc_min_melems.store(melems < 0 ? -1 : melems); }
void bvg_debug_set_multi_stream(int on) { g_ms_enable.store(on ? 1 : 0); }

void bvg_profile_begin(void) {
  g_prof_on = true;
  for (auto& r : g_prof_recs) { g_prof_pool.push_back(r.a); g_prof_pool.push_back(r.b); }
  g_prof_recs.clear();
}

int bvg_profile_end(float* ms_per_class, int64_t* launches_per_class) {
  g_prof_on = false;
  for (int i = 0; i < KC_COUNT; ++i) { if (ms_per_class) ms_per_class[i] = 0.f; if (launches_per_class) launches_per_class[i] = 0; }
  for (auto& r : g_prof_recs) {
    BVG_CUDA(cudaEventSynchronize(r.b));
    float ms = 0.f;
    BVG_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    if (ms_per_class) ms_per_class[r.kc] += ms;
    if (launches_per_class) launches_per_class[r.kc] += 1;
    g_prof_pool.push_back(r.a); g_prof_pool.push_back(r.b);
  }
  g_prof_recs.clear();
  return BVG_OK;
}

int bvg_act1d_fwd(void* dst, const void* src, const float* alpha_log, const float* beta_log,
                  const float* up_taps_host, const float* down_taps_host, int64_t B, int64_t C, int64_t T, int dtype,
                  int precise, void* stream) {
  static const float ref[12] = {BVG_F0, BVG_F1, BVG_F2, BVG_F3, BVG_F4, BVG_F5,
                                BVG_F5, BVG_F4, BVG_F3, BVG_F2, BVG_F1, BVG_F0};
  for (const float* taps : {up_taps_host, down_taps_host})
    if (taps)
      for (int i = 0; i < 12; ++i)
        BVG_CHECK_ARG(fabsf(taps[i] - ref[i]) <= 1e-6f,
                      "act1d: filter tap %d = %g differs from the built-in kaiser-sinc tap %g (the fused kernel "
                      "hard-codes filter 12 / ratio 2, like the reference's)", i, taps[i], ref[i]);
  return act1d_launch(dst, src, alpha_log, beta_log, B, C, T, dtype, precise, (cudaStream_t)stream);
}

int bvg_conv1d_fwd(void* dst, const void* src, const float* weight, const float* bias, const void* res1,
                   const void* res2, float scale, int64_t B, int64_t Cin, int64_t Cout, int64_t T, int K, int dilation,
                   int reflect_pad, int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BVG_CHECK_ARG(weight && Cin > 0 && Cout > 0 && K > 0, "conv1d: bad weight");
  float* wk = nullptr;   // layer-level test entry point: a temporary repack buffer is acceptable here
  BVG_CUDA(cudaMallocAsync((void**)&wk, (size_t)Cin * Cout * K * 4, st));
  int rc = repack_conv_weight_launch(wk, weight, Cout, Cin, K, 0, st);
  if (rc == BVG_OK) {
    ConvEpilogue ep;
    ep.bias = bias; ep.res1 = res1; ep.res2 = res2; ep.scale = scale;
    rc = conv1d_simt_launch(dst, Cout * T, src, nullptr, Cin * T, T, 1, wk, ep, B, Cin, Cout, T, K, dilation,
                            reflect_pad ? 1 : 0, dtype, dtype, st);
  }
  cudaFreeAsync(wk, st);
  return rc;
}

int bvg_convtr1d_fwd(void* dst, const void* src, const float* weight, const float* bias, const float* cond,
                     int64_t Bc, int64_t B, int64_t Cin, int64_t Cout, int64_t Tin, int K, int stride, int dtype,
                     void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BVG_CHECK_ARG(weight && Cin > 0 && Cout > 0 && K > 0, "convtr1d: bad weight");
  BVG_CHECK_ARG(!cond || Bc == 1 || Bc == B, "convtr1d: cond batch must be 1 or B");
  float* wk = nullptr;
  BVG_CUDA(cudaMallocAsync((void**)&wk, (size_t)Cin * Cout * K * 4, st));
  int rc = repack_conv_weight_launch(wk, weight, Cout, Cin, K, 1, st);
  if (rc == BVG_OK) {
    ConvEpilogue ep;
    ep.bias = bias; ep.cond = cond; ep.cond_B = Bc;
    rc = convtr1d_simt_launch(dst, src, wk, ep, B, Cin, Cout, Tin, K, stride, dtype, st);
  }
  cudaFreeAsync(wk, st);
  return rc;
}


static long long* g_umma_dbg = nullptr;   // set by bvg_debug_set_umma_counters (profiling only)
static int g_umma_dry = 0;
void bvg_debug_set_umma_counters(long long* dev_buf) {
  g_umma_dbg = dev_buf;
  bvg::g_dbg_buf = dev_buf;
#ifdef BVG_DEBUG
  const char* e = getenv("BVG_UMMA_DRY");          // debug builds only: issue-loop-only dry run (results are garbage)
  g_umma_dry = (dev_buf && e && e[0] == '1') ? 1 : 0;
#else
  g_umma_dry = 0;
#endif
}

static int umma_layer_test(void* dst, const void* src, const float* weight, const float* bias, const void* res1,
                           const void* res2, float scale, const float* cond, int64_t Bc, int64_t B, int64_t Cin,
                           int64_t Cout, int64_t Tin, int K, int dil, int transposed, int stride, cudaStream_t st) {
  BVG_CHECK_ARG(dst && src && weight && B >= 1 && Cin >= 1 && Cout >= 1 && Tin >= 1, "conv_umma: bad argument");
  const int64_t Tout = transposed ? Tin * stride : Tin;
  const int nph = transposed ? stride : 1;
  const size_t xb = c8t_bytes(B, (int)Cin, Tin), yb = c8t_bytes(B, (int)Cout, Tout);
  const size_t wb = (size_t)umma_pack_elems((int)Cout, (int)Cin, K, nph) * 2;
  char* tmp = nullptr;
  BVG_CUDA(cudaMallocAsync((void**)&tmp, xb + 3 * yb + wb + 1024, st));
  auto al = [](size_t v) { return (v + 255) & ~size_t(255); };
  C8T x = make_c8t(tmp, (int)Cin, (int)Tin);
  C8T y = make_c8t(tmp + al(xb), (int)Cout, (int)Tout);
  C8T r1 = make_c8t(tmp + al(xb) + al(yb), (int)Cout, (int)Tout);
  C8T r2 = make_c8t(tmp + al(xb) + 2 * al(yb), (int)Cout, (int)Tout);
  __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(tmp + al(xb) + 3 * al(yb));
  int rc = to_c8t_launch(x, src, Cin * Tin, Tin, 1, BVG_BF16, B, st);
  if (rc == BVG_OK && res1) rc = to_c8t_launch(r1, res1, Cout * Tout, Tout, 1, BVG_BF16, B, st);
  if (rc == BVG_OK && res2) rc = to_c8t_launch(r2, res2, Cout * Tout, Tout, 1, BVG_BF16, B, st);
  if (rc == BVG_OK) rc = umma_pack_launch(wp, weight, (int)Cout, (int)Cin, K, transposed, nph, st);
  if (rc == BVG_OK) {
    UmmaLayer L;
    L.w = wp; L.Cin = (int)Cin; L.Cout = (int)Cout; L.K = K; L.dil = dil; L.transposed = transposed; L.stride = stride;
    UmmaEpilogue ep;
    ep.bias = bias; ep.cond = cond; ep.cond_B = Bc; ep.scale = scale;
    ep.res1 = res1 ? r1.p : nullptr; ep.res2 = res2 ? r2.p : nullptr;
    ep.dbg = g_umma_dbg;
    ep.dry = g_umma_dry;
    rc = conv_umma_launch(L, x, y, ep, B, st);
  }
  if (rc == BVG_OK) rc = from_c8t_launch(dst, y, BVG_BF16, B, st);
  cudaFreeAsync(tmp, st);
  return rc;
}

int bvg_act1d_c8t_fwd(void* dst, const void* src, const float* alpha_log, const float* beta_log, int64_t B,
                      int64_t Cn, int64_t T, void* stream) {
  return bvg_act1d_c8t_impl_fwd(dst, src, alpha_log, beta_log, B, Cn, T, 0, stream);
}

int bvg_act1d_c8t_impl_fwd(void* dst, const void* src, const float* alpha_log, const float* beta_log, int64_t B,
                           int64_t Cn, int64_t T, int impl, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BVG_CHECK_ARG(impl >= 0 && impl <= 2, "act1d_c8t: impl must be 0, 1 or 2");
  BVG_CHECK_ARG(dst && src && alpha_log && beta_log && B >= 1 && Cn >= 1 && T >= 1, "act1d_c8t: bad argument");
  const size_t nb = c8t_bytes(B, (int)Cn, T);
  char* tmp = nullptr;
  BVG_CUDA(cudaMallocAsync((void**)&tmp, 2 * nb + 512, st));
  C8T x = make_c8t(tmp, (int)Cn, (int)T), y = make_c8t(tmp + ((nb + 255) & ~size_t(255)), (int)Cn, (int)T);
  // poison the halo rows of x: the kernel must not depend on them (conv outputs leave them undefined)
  int rc = BVG_OK;
  if (cudaMemsetAsync(tmp, 0x7f, 2 * nb + 512, st) != cudaSuccess) rc = BVG_ERR_CUDA;
  if (rc == BVG_OK) rc = to_c8t_launch(x, src, Cn * T, T, 1, BVG_BF16, B, st);
  if (rc == BVG_OK && cudaMemsetAsync(x.p, 0x7f, (size_t)x.pad * 16, st) != cudaSuccess) rc = BVG_ERR_CUDA;
  if (rc == BVG_OK) rc = act1d_c8t_launch(y, x, alpha_log, beta_log, B, st, impl);
  if (rc == BVG_OK) rc = from_c8t_launch(dst, y, BVG_BF16, B, st);
  cudaFreeAsync(tmp, st);
  return rc;
}

int bvg_actconv_impl_fwd(void* dst, const void* src, const float* alpha_log, const float* beta_log,
                         const float* weight, const float* bias, const void* res1, const void* res2, float scale,
                         int64_t B, int64_t Cin, int64_t Cout, int64_t T, int K, int dilation, int impl, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BVG_CHECK_ARG(dst && src && weight && alpha_log && beta_log && B >= 1 && Cin >= 1 && Cout >= 1 && T >= 1, "actconv: bad argument");
  BVG_CHECK_ARG(impl >= 0 && impl <= 2, "actconv: impl must be 0, 1 or 2");
  const size_t xb = c8t_bytes(B, (int)Cin, T), yb = c8t_bytes(B, (int)Cout, T);
  const size_t wb = (size_t)umma_pack_elems((int)Cout, (int)Cin, K, 1) * 2;
  const size_t sb = actconv_tc_scratch_bytes(B);
  char* tmp = nullptr;
  auto al = [](size_t v) { return (v + 255) & ~size_t(255); };
  BVG_CUDA(cudaMallocAsync((void**)&tmp, al(xb) + 3 * al(yb) + al(wb) + sb + 1024, st));
  C8T x = make_c8t(tmp, (int)Cin, (int)T);
  C8T y = make_c8t(tmp + al(xb), (int)Cout, (int)T);
  C8T r1 = make_c8t(tmp + al(xb) + al(yb), (int)Cout, (int)T);
  C8T r2 = make_c8t(tmp + al(xb) + 2 * al(yb), (int)Cout, (int)T);
  __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(tmp + al(xb) + 3 * al(yb));
  void* scratch = tmp + al(xb) + 3 * al(yb) + al(wb);
  int rc = BVG_OK;
  if (cudaMemsetAsync(tmp, 0x7f, xb, st) != cudaSuccess) rc = BVG_ERR_CUDA;      // poison: halo rows must not matter
  if (rc == BVG_OK) rc = to_c8t_launch(x, src, Cin * T, T, 1, BVG_BF16, B, st);
  if (rc == BVG_OK && cudaMemsetAsync(x.p, 0x7f, (size_t)x.pad * 16, st) != cudaSuccess) rc = BVG_ERR_CUDA;
  if (rc == BVG_OK && res1) rc = to_c8t_launch(r1, res1, Cout * T, T, 1, BVG_BF16, B, st);
  if (rc == BVG_OK && res2) rc = to_c8t_launch(r2, res2, Cout * T, T, 1, BVG_BF16, B, st);
  if (rc == BVG_OK) rc = umma_pack_launch(wp, weight, (int)Cout, (int)Cin, K, 0, 1, st);
  if (rc == BVG_OK) {
    UmmaLayer L;
    L.w = wp; L.Cin = (int)Cin; L.Cout = (int)Cout; L.K = K; L.dil = dilation;
    UmmaEpilogue ep;
    ep.bias = bias; ep.scale = scale; ep.res1 = res1 ? r1.p : nullptr; ep.res2 = res2 ? r2.p : nullptr;
    ep.zero_pads = 1;
    ep.dbg = g_umma_dbg;
    rc = BVG_ERR_STATE;
    if (impl != 1) rc = actconv_tc_launch(L, x, alpha_log, beta_log, y, ep, B, scratch, st);
    if (rc == BVG_ERR_STATE && impl != 2)
      rc = conv_umma_fused_launch(L, x, alpha_log, beta_log, y, ep, B, st, /*max_nb=*/256);   // (the decode path stops at 128)
    if (rc == BVG_ERR_STATE) set_error("actconv: this layer shape does not qualify for the fused kernel");
  }
  if (rc == BVG_OK) rc = from_c8t_launch(dst, y, BVG_BF16, B, st);
  cudaFreeAsync(tmp, st);
  return rc;
}

int bvg_actconv_umma_fwd(void* dst, const void* src, const float* alpha_log, const float* beta_log,
                         const float* weight, const float* bias, const void* res1, float scale, int64_t B, int64_t Cin,
                         int64_t Cout, int64_t T, int K, int dilation, void* stream) {
  return bvg_actconv_impl_fwd(dst, src, alpha_log, beta_log, weight, bias, res1, nullptr, scale, B, Cin, Cout, T, K, dilation, 1,
                              stream);
}

int bvg_conv1d_umma_fwd(void* dst, const void* src, const float* weight, const float* bias, const void* res1,
                        const void* res2, float scale, int64_t B, int64_t Cin, int64_t Cout, int64_t T, int K,
                        int dilation, void* stream) {
  return umma_layer_test(dst, src, weight, bias, res1, res2, scale, nullptr, 1, B, Cin, Cout, T, K, dilation, 0, 1,
                         (cudaStream_t)stream);
}

int bvg_convtr1d_umma_fwd(void* dst, const void* src, const float* weight, const float* bias, const float* cond,
                          int64_t Bc, int64_t B, int64_t Cin, int64_t Cout, int64_t Tin, int K, int stride,
                          void* stream) {
  return umma_layer_test(dst, src, weight, bias, nullptr, nullptr, 1.f, cond, Bc, B, Cin, Cout, Tin, K, 1, 1, stride,
                         (cudaStream_t)stream);
}

int bvg_plan_create(bvg_plan** out, const bvg_config* cfg) {
  BVG_CHECK_ARG(out && cfg, "plan_create: null argument");
  BVG_CHECK_ARG(cfg->num_upsamples >= 1 && cfg->num_upsamples <= 8, "plan_create: num_upsamples out of range");
  BVG_CHECK_ARG(cfg->num_kernels >= 1 && cfg->num_kernels <= 4, "plan_create: num_kernels out of range");
  BVG_CHECK_ARG(cfg->gpt_dim > 0 && cfg->upsample_initial_channel > 0 && cfg->speaker_embedding_dim > 0 &&
                    cfg->num_mels > 0, "plan_create: non-positive dimension");
  BVG_CHECK_ARG((cfg->upsample_initial_channel >> cfg->num_upsamples) >= 1,
                "plan_create: upsample_initial_channel too small for %d stages", cfg->num_upsamples);
  int ndev = 0;
  BVG_CUDA(cudaGetDeviceCount(&ndev));
  BVG_CHECK_ARG(cfg->device >= 0 && cfg->device < ndev, "plan_create: no CUDA device %d", cfg->device);
  cudaDeviceProp prop;
  BVG_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
  BVG_CHECK_ARG(prop.major == 10, "plan_create: this library is built for sm_100a only (device is sm_%d%d)",
                prop.major, prop.minor);
  bvg_plan* P = new bvg_plan();
  P->cfg = *cfg;
  P->n_stage = cfg->num_upsamples;
  P->C[0] = cfg->upsample_initial_channel;
  for (int i = 0; i < P->n_stage; ++i) {
    P->C[i + 1] = cfg->upsample_initial_channel >> (i + 1);
    P->total_up *= cfg->upsample_rates[i];
  }
  *out = P;
  return BVG_OK;
}

void bvg_plan_destroy(bvg_plan* P) {
  if (!P) return;
  cudaSetDevice(P->cfg.device);
  for (cudaStream_t s : P->side) if (s) cudaStreamDestroy(s);
  if (P->ev_fork) cudaEventDestroy(P->ev_fork);
  for (cudaEvent_t e : P->ev_chain) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : P->ev_h2d) if (e) cudaEventDestroy(e);
  for (void* p : P->allocs) cudaFree(p);
  delete P;
}

int bvg_plan_set_tensor(bvg_plan* P, const char* key, const float* data, int64_t numel) {
  BVG_CHECK_ARG(P && key && (data || numel == 0) && numel >= 0, "plan_set_tensor: bad argument");
  if (P->finalized) { set_error("plan_set_tensor: plan already finalised"); return BVG_ERR_STATE; }
  P->host[key].assign(data, data + numel);
  return BVG_OK;
}

int bvg_plan_finalize(bvg_plan* P, int enable_bf16_umma) {
  BVG_CHECK_ARG(P, "plan_finalize: null plan");
  if (P->finalized) return BVG_OK;
  const int um = enable_bf16_umma ? 1 : 0;
  P->umma = um != 0;
  BVG_CUDA(cudaSetDevice(P->cfg.device));
  const bvg_config& c = P->cfg;
  const int E = c.speaker_embedding_dim;
  BVG_TRY(make_conv(P, "conv_pre.weight", "conv_pre.bias", P->C[0], c.gpt_dim, 7, false, &P->conv_pre, 0, -1, um));
  BVG_TRY(make_conv(P, "cond_layer.weight", "cond_layer.bias", P->C[0], E, 1, false, &P->cond_layer));
  P->ups.resize(P->n_stage);
  P->conds.resize(P->n_stage);
  for (int i = 0; i < P->n_stage; ++i) {
    const std::string u = "ups." + std::to_string(i) + ".0";
    BVG_TRY(make_conv(P, u + ".weight", u + ".bias", P->C[i + 1], P->C[i], c.upsample_kernel_sizes[i], true, &P->ups[i],
                      0, -1, um * c.upsample_rates[i]));
    if (c.cond_in_each_up_layer) {
      const std::string k = "conds." + std::to_string(i);
      BVG_TRY(make_conv(P, k + ".weight", k + ".bias", P->C[i + 1], E, 1, false, &P->conds[i]));
    }
  }
  P->res.resize((size_t)P->n_stage * c.num_kernels);
  for (int i = 0; i < P->n_stage; ++i)
    for (int j = 0; j < c.num_kernels; ++j) {
      const int n = i * c.num_kernels + j;
      ResBlock& R = P->res[n];
      R.K = c.resblock_kernel_sizes[j];
      const int ch = P->C[i + 1];
      const std::string p = "resblocks." + std::to_string(n);
      for (int m = 0; m < 3; ++m) {
        R.dil[m] = c.resblock_dilation_sizes[j][m];
        const std::string a = p + ".convs1." + std::to_string(m), b = p + ".convs2." + std::to_string(m);
        BVG_TRY(make_conv(P, a + ".weight", a + ".bias", ch, ch, R.K, false, &R.c1[m], 0, -1, um));
        BVG_TRY(make_conv(P, b + ".weight", b + ".bias", ch, ch, R.K, false, &R.c2[m], 0, -1, um));
      }
      for (int m = 0; m < 6; ++m)
        BVG_TRY(make_act(P, p + ".activations." + std::to_string(m), ch, &R.alpha[m], &R.beta[m]));
    }
  const int chp = P->C[P->n_stage];
  BVG_TRY(make_act(P, "activation_post", chp, &P->post_alpha, &P->post_beta));
  BVG_TRY(upload_key(P, "conv_post.weight", (int64_t)chp * 7, &P->post_w));   // [1,Cin,7] == [Cin][7]
  BVG_TRY(upload_key(P, "conv_post.bias", 1, &P->post_bias));
  // speaker encoder
  const std::string S = "speaker_encoder.";
  BVG_TRY(make_tdnn(P, S + "blocks.0", c.num_mels, kEC, 5, 1, &P->e_block0, 0, -1, um));
  for (int i = 0; i < 3; ++i) {
    const std::string b = S + "blocks." + std::to_string(i + 1);
    SERes2& R = P->e_blk[i];
    BVG_TRY(make_tdnn(P, b + ".tdnn1", kEC, kEC, 1, 1, &R.tdnn1, 0, -1, um));
    for (int j = 0; j < 7; ++j)
      BVG_TRY(make_tdnn(P, b + ".res2net_block.blocks." + std::to_string(j), kEC / 8, kEC / 8, 3, i + 2, &R.r2n[j]));
    BVG_TRY(make_tdnn(P, b + ".tdnn2", kEC, kEC, 1, 1, &R.tdnn2, 0, -1, um));
    BVG_TRY(make_conv(P, b + ".se_block.conv1.conv.weight", b + ".se_block.conv1.conv.bias", kEA, kEC, 1, false, &R.se1));
    BVG_TRY(make_conv(P, b + ".se_block.conv2.conv.weight", b + ".se_block.conv2.conv.bias", kEC, kEA, 1, false, &R.se2));
  }
  BVG_TRY(make_tdnn(P, S + "mfa", kEM, kEM, 1, 1, &P->e_mfa, 0, -1, um));
  // asp.tdnn over cat([x, mean, std]) (4608 ch): x part keeps ReLU+BN, the (mean,std) part becomes a
  // per-utterance additive term that also carries the conv bias
  BVG_TRY(make_tdnn(P, S + "asp.tdnn", 3 * kEM, kEA, 1, 1, &P->e_asp_tdnn, 0, kEM, um));
  BVG_TRY(make_conv(P, S + "asp.tdnn.conv.conv.weight", "", kEA, 3 * kEM, 1, false, &P->e_asp_ctx, kEM, 3 * kEM));
  P->e_asp_ctx.bias = P->e_asp_tdnn.conv.bias;
  P->e_asp_tdnn.conv.bias = nullptr;
  BVG_TRY(make_conv(P, S + "asp.conv.conv.weight", S + "asp.conv.conv.bias", kEM, kEA, 1, false, &P->e_asp_conv, 0, -1, um));
  BVG_TRY(make_bn(P, S + "asp_bn.norm", 2 * kEM, &P->asp_bn_scale, &P->asp_bn_shift));
  BVG_TRY(make_conv(P, S + "fc.conv.weight", S + "fc.conv.bias", E, 2 * kEM, 1, false, &P->e_fc));
  if (BVG_ENV_ONCE("BVG_MULTI_STREAM", 1) != 0) {
    // the later blocks have the larger kernels (k = 7, 11 against 3): BVG_MS_PRIO=1 gives their streams scheduling priority
    int prio_lo = 0, prio_hi = 0;
    BVG_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    const int use_prio = BVG_ENV_ONCE("BVG_MS_PRIO", 0);
    for (int j = 0; j < 2; ++j)
      BVG_CUDA(cudaStreamCreateWithPriority(&P->side[j], cudaStreamNonBlocking, use_prio ? std::max(prio_hi, prio_lo - 1 - j) : prio_lo));
    BVG_CUDA(cudaEventCreateWithFlags(&P->ev_fork, cudaEventDisableTiming));
    for (cudaEvent_t& e : P->ev_chain) BVG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (cudaEvent_t& e : P->ev_h2d) BVG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  BVG_CUDA(cudaDeviceSynchronize());
  P->host.clear();
  P->finalized = true;
  return BVG_OK;
}

size_t bvg_workspace_bytes(const bvg_plan* P, int64_t B, int64_t T0, int64_t Tm, int dtype) {
  if (!P || B <= 0 || T0 <= 0) return 0;
  Bump b(nullptr, ~size_t(0));
  GenWs g;
  // sized for per-utterance reference mels (Bm == B); a broadcast mel needs less
  carve_gen(P, b, B, T0, B, std::max<int64_t>(Tm, 1), dtype, &g);
  return b.off + 256;
}

int bvg_speaker_embed(const bvg_plan* P, const float* mel, int64_t Bm, int64_t Tm, float* spk, void* workspace,
                      size_t workspace_bytes, void* stream) {
  BVG_CHECK_ARG(P && mel && spk && workspace, "speaker_embed: null argument");
  if (!P->finalized) { set_error("speaker_embed: plan not finalised"); return BVG_ERR_STATE; }
  Bump b(workspace, workspace_bytes);
  EcapaWs w;
  carve_ecapa(b, Bm, Tm, &w);
  if (b.overflow) { set_error("speaker_embed: workspace too small (%zu < %zu)", workspace_bytes, b.off); return BVG_ERR_WORKSPACE; }
  return ecapa_forward(P, mel, Bm, Tm, spk, w, (cudaStream_t)stream);
}

int bvg_decode(const bvg_plan* P, const float* latent, const float* mel, const float* spk_in, int64_t B, int64_t T0,
               int64_t Bm, int64_t Tm, int dtype, float* wav, int16_t* pcm16, int64_t t_lo_pad, int64_t t_hi_pad,
               void* workspace, size_t workspace_bytes, void* stream) {
  return bvg_decode_lat(P, latent, BVG_F32, mel, spk_in, B, T0, Bm, Tm, dtype, wav, pcm16, t_lo_pad, t_hi_pad, workspace,
                        workspace_bytes, stream);
}

static int decode_any(const bvg_plan* P, const void* latent_any, int latent_dtype, const int* lens, const float* mel,
                      const float* spk_in, int64_t B, int64_t T0, int64_t Bm, int64_t Tm, int dtype, float* wav, int16_t* pcm16,
                      int64_t t_lo_pad, int64_t t_hi_pad, void* workspace, size_t workspace_bytes, void* stream);
// set by bvg_decode_host around its bvg_decode call: the latent is still being uploaded on another stream; decode_any waits
// for this event on its own stream right before the first kernel that reads the latent (i.e. after the speaker encoder)
static thread_local cudaEvent_t tl_latent_ready = nullptr;

int bvg_decode_lat(const bvg_plan* P, const void* latent_any, int latent_dtype, const float* mel, const float* spk_in, int64_t B,
                   int64_t T0, int64_t Bm, int64_t Tm, int dtype, float* wav, int16_t* pcm16, int64_t t_lo_pad,
                   int64_t t_hi_pad, void* workspace, size_t workspace_bytes, void* stream) {
  return decode_any(P, latent_any, latent_dtype, nullptr, mel, spk_in, B, T0, Bm, Tm, dtype, wav, pcm16, t_lo_pad, t_hi_pad,
                    workspace, workspace_bytes, stream);
}

int bvg_decode_varlen(const bvg_plan* P, const void* latent, int latent_dtype, const int32_t* lens_dev, const float* mel,
                      const float* spk, int64_t B, int64_t T0_max, int64_t Bm, int64_t Tm, float* wav, int16_t* pcm16,
                      void* workspace, size_t workspace_bytes, void* stream) {
  BVG_CHECK_ARG(lens_dev, "decode_varlen: null length array");
  BVG_CHECK_ARG(P && P->umma, "decode_varlen: ragged batches run on the bf16 tensor-core path (plan finalised without its packs)");
  return decode_any(P, latent, latent_dtype, lens_dev, mel, spk, B, T0_max, Bm, Tm, BVG_BF16, wav, pcm16, 0, 0, workspace,
                    workspace_bytes, stream);
}

static int decode_any(const bvg_plan* P, const void* latent_any, int latent_dtype, const int* lens, const float* mel,
                      const float* spk_in, int64_t B, int64_t T0, int64_t Bm, int64_t Tm, int dtype, float* wav, int16_t* pcm16,
                      int64_t t_lo_pad, int64_t t_hi_pad, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BVG_CHECK_ARG(latent_dtype == BVG_F32 || latent_dtype == BVG_BF16 || latent_dtype == BVG_F16,
                "decode: latent dtype must be BVG_F32, BVG_BF16 or BVG_F16");
  const float* latent = static_cast<const float*>(latent_any);
  BVG_CHECK_ARG(P && latent && workspace, "decode: null argument");
  if (!P->finalized) { set_error("decode: plan not finalised"); return BVG_ERR_STATE; }
  BVG_CHECK_ARG((mel != nullptr) != (spk_in != nullptr), "decode: pass exactly one of mel / spk");
  BVG_CHECK_ARG(dtype == BVG_F32 || dtype == BVG_BF16 || dtype == BVG_F32X3, "decode: dtype must be BVG_F32, BVG_BF16 or BVG_F32X3");
  BVG_CHECK_ARG(dtype != BVG_F32X3 || P->umma, "decode: BVG_F32X3 needs the plan finalised with tensor-core packs");
  BVG_CHECK_ARG(B >= 1 && T0 >= 1, "decode: empty batch or zero latent frames (B=%lld T0=%lld)", (long long)B, (long long)T0);
  // models.py:205-209: Bm == 2B enters a training-only branch that references an undefined attribute
  BVG_CHECK_ARG(Bm == 1 || Bm == B, "decode: reference mel batch must be 1 or B (got %lld for B=%lld)", (long long)Bm, (long long)B);
  BVG_CHECK_ARG(t_lo_pad >= 0 && t_hi_pad >= 0 && t_lo_pad + t_hi_pad < T0, "decode: bad halo");
  BVG_CHECK_ARG(wav || pcm16, "decode: no output buffer");
  const bvg_config& c = P->cfg;
  Bump bump(workspace, workspace_bytes);
  GenWs g;
  carve_gen(P, bump, B, T0, Bm, mel ? Tm : 1, dtype, &g);
  if (bump.overflow) {
    set_error("decode: workspace too small (%zu bytes given, %zu needed)", workspace_bytes, bump.off);
    return BVG_ERR_WORKSPACE;
  }
  const int E = c.speaker_embedding_dim;
  const float* spk = spk_in;
  if (mel) {
    BVG_TRY(ecapa_forward(P, mel, Bm, Tm, g.spk, g.e, st, /*tensor cores*/ dtype == BVG_BF16 && P->umma));
    spk = g.spk;
  }
  // speaker conditioning vectors cond_layer(spk), conds[i](spk)  (models.py:226, :233-234): [Bm, C]
  {
    // all of them read the same [Bm, E] embedding: one launch per kMaxMatvecJobs layers
    MatvecJob jobs[kMaxMatvecJobs];
    int nj = 0;
    auto add = [&](const ConvLayer& L, float* y) -> int {
      MatvecJob j;
      j.w = L.w; j.bias = L.bias; j.y = y; j.Cout = L.Cout;
      jobs[nj++] = j;
      if (nj == kMaxMatvecJobs) { BVG_TRY(matvec_multi_launch(jobs, nj, spk, Bm, E, st)); nj = 0; }
      return BVG_OK;
    };
    BVG_CHECK_ARG(P->cond_layer.K == 1 && P->cond_layer.Cin == E, "decode: cond_layer is a 1x1 conv of the embedding");
    BVG_TRY(add(P->cond_layer, g.cond[0]));
    if (c.cond_in_each_up_layer)
      for (int i = 0; i < P->n_stage; ++i) {
        BVG_CHECK_ARG(P->conds[i].K == 1 && P->conds[i].Cin == E, "decode: conds[%d] is a 1x1 conv of the embedding", i);
        BVG_TRY(add(P->conds[i], g.cond[i + 1]));
      }
    if (nj) BVG_TRY(matvec_multi_launch(jobs, nj, spk, Bm, E, st));
  }
  if (tl_latent_ready) {
    BVG_CUDA(cudaStreamWaitEvent(st, tl_latent_ready, 0));
    tl_latent_ready = nullptr;
  }
  if (dtype == BVG_BF16 && P->umma)
    return decode_bf16_umma(P, latent_any, latent_dtype, g, B, T0, Bm, wav, pcm16, t_lo_pad, t_hi_pad, st, lens);
  BVG_CHECK_ARG(!lens, "decode_varlen: ragged batches need the bf16 path");
  if (latent_dtype != BVG_F32) {
    // the fp32 paths read the latent through the CUDA-core conv: widen it once into the staging buffer
    BVG_TRY(cast_to_f32_launch(g.latent_dev, latent_any, latent_dtype, B * T0 * c.gpt_dim, st));
    latent = g.latent_dev;
  }
  // BVG_F32X3: the fp32 path below with the Conv1d layers on the tensor cores (3-term bf16 split); everything else
  // (Activation1d with libdevice sinf, ConvTranspose1d, speaker encoder, conv_post) is the fp32 CUDA-core code
  void* const x3_main = g.X3;
  void* x3 = g.X3;
  cudaStream_t const st0 = st;
  const bool ms = multi_stream(P, B, T0, dtype) && P->ms_busy.exchange(1) == 0;
  struct Release { const bvg_plan* p; bool on; ~Release() { if (on) p->ms_busy.store(0); } } release{P, ms};
  static const bool x3_precise = [] { const char* e = getenv("BVG_X3_PRECISE_ACT"); return e && e[0] == '1'; }();
  const bool fast_act = dtype == BVG_F32X3 && !x3_precise;
  if (dtype == BVG_F32X3) dtype = BVG_F32;
  // conv_pre on latent^T (models.py:220-226): read [B,T0,gpt_dim] through strides
  {
    ConvEpilogue ep;
    ep.bias = P->conv_pre.bias; ep.cond = g.cond[0]; ep.cond_B = Bm;
    if (x3 && P->conv_pre.wx3)
      BVG_TRY(gen_conv_x3((float*)g.XS, latent, T0 * c.gpt_dim, 1, c.gpt_dim, P->conv_pre, ep, B, T0, 1, x3, st));
    else
      BVG_TRY(conv1d_simt_launch(g.XS, (int64_t)P->C[0] * T0, latent, nullptr, T0 * c.gpt_dim, 1, c.gpt_dim,
                                 P->conv_pre.w, ep, B, c.gpt_dim, P->C[0], T0, 7, 1, 0, BVG_F32, dtype, st));
  }
  int64_t T = T0;
  const float inv_nk = 1.0f / (float)c.num_kernels;
  for (int i = 0; i < P->n_stage; ++i) {
    const int u = c.upsample_rates[i];
    const int ch = P->C[i + 1];
    {
      ConvEpilogue ep;
      ep.bias = P->ups[i].bias;
      if (c.cond_in_each_up_layer) { ep.cond = g.cond[i + 1]; ep.cond_B = Bm; }
      if (x3 && P->ups[i].wx3) {
        // ConvTranspose1d on the tensor cores as well (phase mode of conv_umma_kernel, split weights)
        const ConvLayer& L = P->ups[i];
        C8T xsp = make_c8t(x3, 2 * ((L.Cin + 7) / 8 * 8), (int)T);
        BVG_TRY(split_to_c8t_launch(xsp, (const float*)g.XS, (int64_t)L.Cin * T, T, 1, L.Cin, B, st));
        UmmaLayer ul;
        ul.w = L.wx3; ul.Cin = L.Cin; ul.Cout = L.Cout; ul.K = L.K; ul.dil = 1; ul.transposed = 1; ul.stride = u; ul.split = 1;
        UmmaEpilogue ue;
        ue.bias = L.bias; ue.cond = ep.cond; ue.cond_B = ep.cond_B; ue.yf32 = (float*)g.A;
        BVG_TRY(conv_umma_launch(ul, xsp, make_c8t(nullptr, L.Cout, (int)(T * u)), ue, B, st));
      } else {
        BVG_TRY(convtr1d_simt_launch(g.A, g.XS, P->ups[i].w, ep, B, P->C[i], ch, T, P->ups[i].K, u, dtype, st));
      }
    }
    T *= u;
    if (ms) {                                     // fork: the AMP blocks of the stage on three streams (see decode_bf16_umma)
      BVG_CUDA(cudaEventRecord(P->ev_fork, st0));
      for (int j = 1; j < c.num_kernels; ++j) BVG_CUDA(cudaStreamWaitEvent(P->side[j - 1], P->ev_fork, 0));
    }
    for (int j = 0; j < c.num_kernels; ++j) {
      const ResBlock& R = P->res[(size_t)i * c.num_kernels + j];
      const bool own = ms && j > 0;
      st = own ? P->side[j - 1] : st0;
      x3 = own && x3_main ? g.X3x[j - 1] : x3_main;
      void* const bT1 = own ? g.T1x[j - 1] : g.T1;
      void* const bT2 = own ? g.T2x[j - 1] : g.T2;
      void* const bY = own ? g.Yx[j - 1] : g.Y;
      const void* y = g.A;                       // AMPBlock1.forward models.py:65-74
      for (int m = 0; m < 3; ++m) {
        // act -> conv: one pass through the split tensor on the fp32x3 path when the shape qualifies
        auto act_conv = [&](void* dst, const void* src, const float* al, const float* be, const ConvLayer& L, const ConvEpilogue& e,
                            int dil) -> int {
          if (x3 && fast_act && L.wx3) {
            const int rc = act_conv_x3((float*)dst, (const float*)src, al, be, L, e, B, T, dil, x3, st);
            if (rc != BVG_ERR_STATE) return rc;
          }
          BVG_TRY(act_launch(bT1, src, al, be, B, ch, T, dtype, st, fast_act));
          return gen_conv(dst, bT1, L, e, B, T, dil, dtype, st, x3);
        };
        ConvEpilogue e1;
        BVG_TRY(act_conv(bT2, y, R.alpha[2 * m], R.beta[2 * m], R.c1[m], e1, R.dil[m]));
        ConvEpilogue e2;
        e2.res1 = y;                              // x = xt + x
        if (m < 2) {
          BVG_TRY(act_conv(bY, bT2, R.alpha[2 * m + 1], R.beta[2 * m + 1], R.c2[m], e2, 1));
          y = bY;
        } else {                                  // xs += block(x); x = xs / num_kernels (models.py:237-243)
          if (j > 0) e2.res2 = g.XS;
          if (j == c.num_kernels - 1) e2.scale = inv_nk;
          if (ms && j > 0) BVG_CUDA(cudaStreamWaitEvent(st, P->ev_chain[j - 1], 0));   // the running sum in XS, in block order
          BVG_TRY(act_conv(g.XS, bT2, R.alpha[2 * m + 1], R.beta[2 * m + 1], R.c2[m], e2, 1));
        }
      }
      if (ms) BVG_CUDA(cudaEventRecord(P->ev_chain[j], st));
    }
    st = st0; x3 = x3_main;
    if (ms) BVG_CUDA(cudaStreamWaitEvent(st0, P->ev_chain[c.num_kernels - 1], 0));   // join
  }
  const int chp = P->C[P->n_stage];
  BVG_TRY(act_launch(g.T1, g.XS, P->post_alpha, P->post_beta, B, chp, T, dtype, st, fast_act));
  BVG_TRY(conv_post_launch(wav, pcm16, g.T1, P->post_w, P->post_bias, B, chp, T, 7, t_lo_pad * P->total_up,
                           t_hi_pad * P->total_up, dtype, st));
  return BVG_OK;
}

int bvg_decode_host(const bvg_plan* P, const float* latent_host, const float* mel_host, int64_t B, int64_t T0,
                    int64_t Bm, int64_t Tm, int dtype, float* wav_host, int16_t* pcm16_host, void* workspace,
                    size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BVG_CHECK_ARG(P && latent_host && mel_host && workspace && (wav_host || pcm16_host), "decode_host: null argument");
  BVG_CHECK_ARG(B >= 1 && T0 >= 1 && Tm >= 1 && (Bm == 1 || Bm == B), "decode_host: bad shape");
  Bump bump(workspace, workspace_bytes);
  GenWs g;
  carve_gen(P, bump, B, T0, Bm, Tm, dtype, &g);
  if (bump.overflow) { set_error("decode_host: workspace too small (%zu < %zu)", workspace_bytes, bump.off); return BVG_ERR_WORKSPACE; }
  const int64_t L = T0 * P->total_up;
  // the (small) reference mel first; the latent (38 MB for 32 x 10 s) goes up on a side stream while the speaker encoder runs
  BVG_CUDA(cudaMemcpyAsync(g.mel_dev, mel_host, (size_t)Bm * Tm * P->cfg.num_mels * 4, cudaMemcpyHostToDevice, st));
  const bool split = P->side[0] && P->ev_h2d[0] && g_ms_enable.load() && P->h2d_busy.exchange(1) == 0;
  struct ReleaseH { const bvg_plan* p; bool on; ~ReleaseH() { if (on) p->h2d_busy.store(0); tl_latent_ready = nullptr; } } release_h{P, split};
  if (split) {
    BVG_CUDA(cudaEventRecord(P->ev_h2d[0], st));                    // (the staging buffer may still be read by earlier work on st)
    BVG_CUDA(cudaStreamWaitEvent(P->side[0], P->ev_h2d[0], 0));
    BVG_CUDA(cudaMemcpyAsync(g.latent_dev, latent_host, (size_t)B * T0 * P->cfg.gpt_dim * 4, cudaMemcpyHostToDevice, P->side[0]));
    BVG_CUDA(cudaEventRecord(P->ev_h2d[1], P->side[0]));
    tl_latent_ready = P->ev_h2d[1];
  } else {
    BVG_CUDA(cudaMemcpyAsync(g.latent_dev, latent_host, (size_t)B * T0 * P->cfg.gpt_dim * 4, cudaMemcpyHostToDevice, st));
  }
  BVG_TRY(bvg_decode(P, g.latent_dev, g.mel_dev, nullptr, B, T0, Bm, Tm, dtype, wav_host ? g.wav_dev : nullptr,
                     pcm16_host ? g.pcm_dev : nullptr, 0, 0, workspace, workspace_bytes, stream));
  if (wav_host) BVG_CUDA(cudaMemcpyAsync(wav_host, g.wav_dev, (size_t)B * L * 4, cudaMemcpyDeviceToHost, st));
  if (pcm16_host) BVG_CUDA(cudaMemcpyAsync(pcm16_host, g.pcm_dev, (size_t)B * L * 2, cudaMemcpyDeviceToHost, st));
  BVG_CUDA(cudaStreamSynchronize(st));
  return BVG_OK;
}

}  // extern "C"
