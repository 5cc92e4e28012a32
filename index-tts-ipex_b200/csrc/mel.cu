// Reference-mel front end: MelSpectrogramFeatures.forward of the reference (indextts/utils/feature_extractors.py:24-50, called from
// infer.py:82-93 on the 24 kHz mono prompt): torchaudio MelSpectrogram(n_fft 1024, hop 256, hann window (periodic), center = True
// with reflect padding, power 1, htk mel scale, no norm, 100 bands 0..12 kHz) followed by safe_log = log(clip(x, 1e-7))
// (utils/common.py:110).  Output is the layout the speaker encoder and bvg_decode take: mel [B, frames, n_mels] fp32.
//
// One CTA per (frame, utterance): the windowed frame goes to shared memory, 513 + 512 threads-worth of bins are computed as a
// direct DFT with an exact integer twiddle index (f * n mod n_fft into a sincospi table: no FFT butterflies to round
// differently from one length to the next; 1 M FMA per frame, a 3 s prompt is 0.3 GFLOP), magnitudes stay in shared memory,
// then the mel filterbank rows and the clipped log.  A per-voice constant: throughput is irrelevant, accuracy is not.
#include <math.h>

#include "bvg_common.cuh"

namespace bvg {
namespace {

constexpr int kMelThreads = 256;

__global__ void __launch_bounds__(kMelThreads) mel_frontend_kernel(float* __restrict__ mel, const float* __restrict__ audio,
                                                                   const float* __restrict__ fb /*[n_freq][n_mels]*/, int64_t L,
                                                                   int n_fft, int hop, int n_mels, int frames, float clip) {
  extern __shared__ float sm[];
  float* xw = sm;                       // [n_fft] windowed frame
  float* cs = xw + n_fft;               // [n_fft] cos(2 pi k / n_fft)
  float* sn = cs + n_fft;               // [n_fft] sin(2 pi k / n_fft)
  float* mag = sn + n_fft;              // [n_fft / 2 + 1]
  const int t = blockIdx.x, b = blockIdx.y;
  const int n_freq = n_fft / 2 + 1;
  const float* a = audio + (int64_t)b * L;
  const int64_t start = (int64_t)t * hop - n_fft / 2;           // center = True
  for (int n = threadIdx.x; n < n_fft; n += kMelThreads) {
    int64_t i = start + n;                                      // reflect padding (no edge repeat), as torch.stft
    if (i < 0) i = -i;
    if (i >= L) i = 2 * (L - 1) - i;
    i = min(max(i, (int64_t)0), L - 1);
    float s, c;
    sincospif(2.0f * (float)n / (float)n_fft, &s, &c);
    cs[n] = c; sn[n] = s;
    xw[n] = a[i] * (0.5f - 0.5f * c);                           // periodic hann window
  }
  __syncthreads();
  for (int f = threadIdx.x; f < n_freq; f += kMelThreads) {
    float re = 0.f, im = 0.f;
    int k = 0;                                                  // f * n mod n_fft
    for (int n = 0; n < n_fft; ++n) {
      const float x = xw[n];
      re = fmaf(x, cs[k], re);
      im = fmaf(x, sn[k], im);
      k = (k + f) & (n_fft - 1);
    }
    mag[f] = sqrtf(re * re + im * im);                          // power = 1
  }
  __syncthreads();
  for (int m = threadIdx.x; m < n_mels; m += kMelThreads) {
    float acc = 0.f;
    for (int f = 0; f < n_freq; ++f) acc = fmaf(mag[f], fb[(int64_t)f * n_mels + m], acc);
    mel[((int64_t)b * frames + t) * n_mels + m] = logf(fmaxf(acc, clip));
  }
}

}  // namespace
}  // namespace bvg

extern "C" int64_t bvg_mel_frames(int64_t L, int hop) { return L / hop + 1; }

extern "C" int bvg_mel_frontend(float* mel, const float* audio, const float* fb, int64_t B, int64_t L, int n_fft, int hop,
                                int n_mels, void* stream) {
  using namespace bvg;
  BVG_CHECK_ARG(mel && audio && fb, "mel_frontend: null pointer");
  BVG_CHECK_ARG(B >= 1 && B <= 65535 && n_mels >= 1 && hop >= 1, "mel_frontend: bad sizes");
  BVG_CHECK_ARG(n_fft >= 16 && n_fft <= 4096 && (n_fft & (n_fft - 1)) == 0, "mel_frontend: n_fft must be a power of two in [16, 4096]");
  BVG_CHECK_ARG(L > n_fft / 2, "mel_frontend: the prompt (%lld samples) is shorter than the reflect padding (%d)", (long long)L, n_fft / 2);
  const int64_t frames = bvg_mel_frames(L, hop);
  BVG_CHECK_ARG(frames < (1ll << 31), "mel_frontend: too many frames");
  const size_t smem = (size_t)(3 * n_fft + n_fft / 2 + 1) * sizeof(float);
  static std::atomic<uint64_t> opted{0};
  BVG_TRY(smem_opt_in(mel_frontend_kernel, opted, 64 * 1024));
  ProfScope prof((cudaStream_t)stream, KC_OTHER);
  mel_frontend_kernel<<<dim3((unsigned)frames, (unsigned)B), kMelThreads, smem, (cudaStream_t)stream>>>(
      mel, audio, fb, L, n_fft, hop, n_mels, (int)frames, 1e-7f);
  BVG_LAUNCHED();
  return BVG_OK;
}
