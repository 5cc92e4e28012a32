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

// ---- two channels at once on packed fp32x2 math (FFMA2 / FADD2 / FMUL2, sm_100+) ----------------------------
// The stencil is issue-bound on scalar fp32 (about 58 issue slots per sample); a thread of the c8t kernel owns
// two channels with identical instruction streams, so every FIR multiply-add pairs up across the channels and
// the slot count halves.  Interior threads only (no replicate padding in the window); results are bit-identical
// to the scalar path (same operation order; the down-FIR uses the x2 taps and a final exact *0.5).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float x, float y) { f32x2 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(x), "f"(y)); return d; }
__device__ __forceinline__ void unpk2(f32x2 v, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// bf16x2 word (lo = channel A, hi = channel B) -> (float A, float B)
// (the low half goes through PRMT so that the shift does not become an IMAD on the FMA pipe, which the stencil saturates)
__device__ __forceinline__ f32x2 unpack_bf16x2(uint32_t w) {
  uint32_t lo;
  asm("prmt.b32 %0, %1, 0, 0x1044;" : "=r"(lo) : "r"(w));
  return pk2(__uint_as_float(lo), __uint_as_float(w & 0xffff0000u));
}

// Two signals (channels) at once.  load(j) returns the pair of samples at window position j (= time tg-8+j),
// j in [3, V+13); store(q, ya, yb) receives output tg+q of both signals.
// sc0 = (2 e^alphaA, 2 e^alphaB), sc1 = (hbA, hbB)  (fast snake, see snake<false>).
// The intermediates are kept at half scale (a' = a/2, an exact power-of-two scaling folded into the snake's
// constants), so the down-FIR can reuse the doubled taps g = 2f of the up-FIR without a final multiply:
// sum g a' == sum f a, rounding for rounding.
template <int V, class Load, class Store>
__device__ __forceinline__ void act1d_window2(Load load, Store store, f32x2 sc0, f32x2 sc1) {
  const f32x2 g0 = pk2(2.f * BVG_F0, 2.f * BVG_F0), g1 = pk2(2.f * BVG_F1, 2.f * BVG_F1),
              g2 = pk2(2.f * BVG_F2, 2.f * BVG_F2), g3 = pk2(2.f * BVG_F3, 2.f * BVG_F3),
              g4 = pk2(2.f * BVG_F4, 2.f * BVG_F4), g5 = pk2(2.f * BVG_F5, 2.f * BVG_F5);
  const f32x2 half = pk2(0.5f, 0.5f);
  const f32x2 sc1h = mul2(sc1, half), nsc1h = mul2(sc1, pk2(-0.5f, -0.5f));
  f32x2 X[V + 16];      // unpacked rows (only p+3..p+8 are live around intermediate pair p)
  f32x2 a[2 * V + 10];  // activated intermediates (half scale), a[i] <-> m = 2*tg - 5 + i (12 live at a time)
#pragma unroll
  for (int j = 3; j < 8; ++j) X[j] = load(j);
#pragma unroll
  for (int p = 0; p < V + 5; ++p) {
    X[p + 8] = load(p + 8);
    // i = 2p (m odd):  g1 x[c-2] + g3 x[c-1] + g5 x[c] + g4 x[c+1] + g2 x[c+2] + g0 x[c+3], c = 5+p
    f32x2 u = mul2(g1, X[p + 3]);
    u = fma2(g3, X[p + 4], u);
    u = fma2(g5, X[p + 5], u);
    u = fma2(g4, X[p + 6], u);
    u = fma2(g2, X[p + 7], u);
    u = fma2(g0, X[p + 8], u);
    // i = 2p+1 (m even): g0 x[c-3] + g2 x[c-2] + g4 x[c-1] + g5 x[c] + g3 x[c+1] + g1 x[c+2], c = 6+p
    f32x2 w = mul2(g0, X[p + 3]);
    w = fma2(g2, X[p + 4], w);
    w = fma2(g4, X[p + 5], w);
    w = fma2(g5, X[p + 6], w);
    w = fma2(g3, X[p + 7], w);
    w = fma2(g1, X[p + 8], w);
    {
      float zx, zy;
      unpk2(mul2(u, sc0), zx, zy);
      a[2 * p] = fma2(nsc1h, pk2(__cosf(zx), __cosf(zy)), fma2(u, half, sc1h));
      unpk2(mul2(w, sc0), zx, zy);
      a[2 * p + 1] = fma2(nsc1h, pk2(__cosf(zx), __cosf(zy)), fma2(w, half, sc1h));
    }
    if (p >= 5) {
      const int i = 2 * (p - 5);
      f32x2 s = mul2(g0, add2(a[i], a[i + 11]));
      s = fma2(g1, add2(a[i + 1], a[i + 10]), s);
      s = fma2(g2, add2(a[i + 2], a[i + 9]), s);
      s = fma2(g3, add2(a[i + 3], a[i + 8]), s);
      s = fma2(g4, add2(a[i + 4], a[i + 7]), s);
      s = fma2(g5, add2(a[i + 5], a[i + 6]), s);
      float yx, yy;
      unpk2(s, yx, yy);
      store(p - 5, yx, yy);
    }
  }
}

}  // namespace bvg
