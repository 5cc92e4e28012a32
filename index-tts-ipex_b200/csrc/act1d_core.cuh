// Per-thread Activation1d stencil shared by the plain [B,C,T] kernel (act1d.cu) and the c8t kernel
// (act1d_c8t.cu).  See act1d.cu for the closed form and reference citations.
#pragma once
#include "bvg_common.cuh"

namespace bvg {

// SnakeBeta on one upsampled sample (activations.py:109-122):  u + 1/(e^beta + 1e-9) * sin(e^alpha u)^2
//   PRECISE: libdevice sinf (fp32 parity path).  sc0 = e^alpha, sc1 = 1/(e^beta+1e-9)
//   fast:    sin^2 z = (1 - cos 2z)/2  ->  u + hb - hb*cos(2 e^alpha u), one FMUL + MUFU.COS + FADD + FFMA.
//            sc0 = 2 e^alpha, sc1 = hb = 0.5/(e^beta+1e-9)
template <bool PRECISE>
__device__ __forceinline__ float snake(float u, float sc0, float sc1) {
  if (PRECISE) {
    float s = sinf(u * sc0);
    return fmaf(sc1 * s, s, u);
  } else {
    return fmaf(-sc1, __cosf(u * sc0), u + sc1);
  }
}
template <bool PRECISE>
__device__ __forceinline__ void snake_params(float alpha_log, float beta_log, float& sc0, float& sc1) {
  if (PRECISE) {
    sc0 = expf(alpha_log);
    sc1 = 1.0f / (expf(beta_log) + 1e-9f);
  } else {
    sc0 = 2.0f * __expf(alpha_log);
    sc1 = __fdividef(0.5f, __expf(beta_log) + 1e-9f);
  }
}

// The per-thread stencil: (V+16)-sample input window (xw[i] = x[tg-8+i]) -> V outputs y[tg..tg+V-1].
template <int V, bool PRECISE>
__device__ __forceinline__ void act1d_window(float (&xw)[V + 16], float (&y)[V], float sc0, float sc1,
                                             int64_t tg, int64_t T) {
  constexpr int W = V + 16;
  constexpr int NA = 2 * V + 10;
  // replicate padding of the input (F.pad(x,(5,5),'replicate'), resample.py:28)
  if (tg - 5 < 0 || tg + V + 4 > T - 1) {
    float xl = 0.f, xr = 0.f;
#pragma unroll
    for (int i = 0; i < W; ++i) {
      int64_t t = tg - 8 + i;
      if (t == 0) xl = xw[i];
      if (t == T - 1) xr = xw[i];
    }
#pragma unroll
    for (int i = 0; i < W; ++i) {
      int64_t t = tg - 8 + i;
      if (t < 0) xw[i] = xl;
      if (t > T - 1) xw[i] = xr;
    }
  }
  const float g0 = 2.f * BVG_F0, g1 = 2.f * BVG_F1, g2 = 2.f * BVG_F2, g3 = 2.f * BVG_F3,
              g4 = 2.f * BVG_F4, g5 = 2.f * BVG_F5;
  // a[i] <-> upsampled index m = 2*tg - 5 + i, i = 0..2V+9
  float a[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    float u;
    if ((i & 1) == 0) {
      // m odd = 2j+1, j = tg-3+i/2 -> xw index of x[j] is j-tg+8 = 5+i/2
      const int c = 5 + i / 2;
      u = g1 * xw[c - 2];
      u = fmaf(g3, xw[c - 1], u);
      u = fmaf(g5, xw[c], u);
      u = fmaf(g4, xw[c + 1], u);
      u = fmaf(g2, xw[c + 2], u);
      u = fmaf(g0, xw[c + 3], u);
    } else {
      // m even = 2j, j = tg-2+(i-1)/2 -> xw index 6+(i-1)/2
      const int c = 6 + (i - 1) / 2;
      u = g0 * xw[c - 3];
      u = fmaf(g2, xw[c - 2], u);
      u = fmaf(g4, xw[c - 1], u);
      u = fmaf(g5, xw[c], u);
      u = fmaf(g3, xw[c + 1], u);
      u = fmaf(g1, xw[c + 2], u);
    }
    a[i] = snake<PRECISE>(u, sc0, sc1);
  }
  // replicate padding of the ACTIVATED signal (F.pad(x,(5,6),'replicate'), filter.py:90-92)
  const int64_t m0 = 2 * tg - 5;
  if (m0 < 0 || m0 + NA - 1 > 2 * T - 1) {
    float al = 0.f, ar = 0.f;
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      if (m0 + i == 0) al = a[i];
      if (m0 + i == 2 * T - 1) ar = a[i];
    }
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      if (m0 + i < 0) a[i] = al;
      if (m0 + i > 2 * T - 1) a[i] = ar;
    }
  }
#pragma unroll
  for (int q = 0; q < V; ++q) {
    const int i = 2 * q;
    // symmetric taps: f[k] == f[11-k]
    float s = BVG_F0 * (a[i] + a[i + 11]);
    s = fmaf(BVG_F1, a[i + 1] + a[i + 10], s);
    s = fmaf(BVG_F2, a[i + 2] + a[i + 9], s);
    s = fmaf(BVG_F3, a[i + 3] + a[i + 8], s);
    s = fmaf(BVG_F4, a[i + 4] + a[i + 7], s);
    s = fmaf(BVG_F5, a[i + 5] + a[i + 6], s);
    y[q] = s;
  }
}

}  // namespace bvg
