/* Plain-C restatement of the reference's Activation1d (TEST INFRASTRUCTURE ONLY; see bigvgan_oracle.py).
 *
 * Follows alias_free_torch/act.py:24-29 = UpSample1d.forward (resample.py:25-33: replicate pad 5/5,
 * conv_transpose1d stride 2 with the 12-tap kaiser-sinc filter, x ratio, crop 15/15) ->
 * SnakeBeta.forward (activations.py:109-122, log-scale alpha/beta) -> DownSample1d.forward
 * (resample.py:46-49 -> filter.py:87-96: replicate pad 5 left / 6 right, depthwise conv stride 2).
 * Unlike bigvgan_oracle.act1d (a polyphase closed form) this version materialises the padded and
 * 2x-upsampled signals exactly as the PyTorch ops do, in double precision, so the two restatements
 * check each other.  Built by __graft_entry__.build() into oracle/_build/libact1d_oracle.so. */
#include <math.h>
#include <stdlib.h>

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* x, y: [C][T] row-major; alpha_log, beta_log: [C]; taps: [12].  Returns 0, or 1 on allocation failure. */
int act1d_oracle(const double* x, double* y, const double* alpha_log, const double* beta_log, const double* taps,
                 int C, int T) {
  const int K = 12, ratio = 2, pad = K / ratio - 1;                 /* 5 */
  const int pad_left = pad * ratio + (K - ratio) / 2;               /* 15 */
  const int Tp = T + 2 * pad;                                       /* padded length */
  const int Tu = (Tp - 1) * ratio + K;                              /* conv_transpose1d output length */
  double* xp = (double*)malloc(sizeof(double) * Tp);
  double* up = (double*)malloc(sizeof(double) * Tu);
  double* a = (double*)malloc(sizeof(double) * (2 * T + 11));
  if (!xp || !up || !a) { free(xp); free(up); free(a); return 1; }
  for (int c = 0; c < C; ++c) {
    const double* xc = x + (size_t)c * T;
    for (int i = 0; i < Tp; ++i) xp[i] = xc[clampi(i - pad, 0, T - 1)];           /* F.pad replicate */
    for (int i = 0; i < Tu; ++i) up[i] = 0.0;
    for (int i = 0; i < Tp; ++i)                                                  /* conv_transpose1d */
      for (int k = 0; k < K; ++k) up[i * ratio + k] += xp[i] * taps[k];
    const double ea = exp(alpha_log[c]), ib = 1.0 / (exp(beta_log[c]) + 1e-9);
    /* crop [pad_left : -pad_right], x ratio, SnakeBeta, then replicate pad 5 / 6 */
    for (int m = -5; m < 2 * T + 6; ++m) {
      const int mm = clampi(m, 0, 2 * T - 1);
      const double u = ratio * up[mm + pad_left];
      const double s = sin(u * ea);
      a[m + 5] = u + ib * s * s;
    }
    for (int t = 0; t < T; ++t) {                                                 /* lowpass, stride 2 */
      double acc = 0.0;
      for (int k = 0; k < K; ++k) acc += taps[k] * a[2 * t + k];
      y[(size_t)c * T + t] = acc;
    }
  }
  free(xp); free(up); free(a);
  return 0;
}
