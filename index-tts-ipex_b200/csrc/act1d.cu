// Activation1d for sm_100a:  2x kaiser-sinc upsample -> SnakeBeta -> 2x low-pass downsample,
// fused into one stencil pass that never materialises the 2x intermediate.
//
// Restates (does not port) alias_free_torch/act.py:24-29 = resample.py:25-33 ->
// activations.py:109-122 -> resample.py:46-49 / filter.py:87-96 of the reference, with the
// PyTorch path's edge semantics (replicate padding of x before the up-FIR, replicate padding
// of the ACTIVATED signal before the down-FIR).  The reference's own CUDA kernel
// (anti_alias_activation_cuda.cu:43-181) deviates from that at the first/last 3 samples.
//
// Closed form, f = 12 symmetric taps (bvg_common.cuh), per row x[0..T):
//   u[2j]   = 2*sum_{d=-3..2} f[5-2d] * x[clamp(j+d)]
//   u[2j+1] = 2*sum_{d=-2..3} f[6-2d] * x[clamp(j+d)]
//   a[m]    = u[m] + 1/(exp(beta)+1e-9) * sin(exp(alpha)*u[m])^2
//   y[t]    = sum_{k=0..11} f[k] * a[clamp(2t+k-5, 0, 2T-1)]
//
// Data movement: one CTA = 8 rows x 256 samples.  The tile plus an 8-sample halo per side is
// staged in shared memory by the TMA unit (cp.async.bulk, one bulk copy per row, completion
// on an mbarrier) when rows are 16-byte aligned; each thread then owns 8 consecutive outputs:
// three 128-bit conflict-free LDS for its 24-sample window (bf16), all FIR/activation math in
// registers, one 128-bit coalesced store.  HBM sees each element once in, once out.
#include "bvg_common.cuh"
#include "act1d_core.cuh"
#include "umma_ptx.cuh"

namespace bvg {

namespace {

constexpr int kRows = 8;      // rows per CTA
constexpr int kW = 256;       // outputs per row per CTA
constexpr int kHalo = 8;      // staged halo per side (5 needed; 8 keeps 16B alignment)
constexpr int kPitch = kW + 2 * kHalo;   // 272 elements
constexpr int kThreads = 256;

template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); };

// load 24 consecutive elements starting at a 16B-aligned smem address
template <typename T> __device__ __forceinline__ void load_window(const T* p, float (&xw)[24]);
template <> __device__ __forceinline__ void load_window<float>(const float* p, float (&xw)[24]) {
  const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    float4 v = q[i];
    xw[4 * i + 0] = v.x; xw[4 * i + 1] = v.y; xw[4 * i + 2] = v.z; xw[4 * i + 3] = v.w;
  }
}
template <> __device__ __forceinline__ void load_window<__nv_bfloat16>(const __nv_bfloat16* p, float (&xw)[24]) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    uint4 v = q[i];
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      xw[8 * i + 2 * j + 0] = __uint_as_float(w[j] << 16);
      xw[8 * i + 2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
    }
  }
}
template <> __device__ __forceinline__ void load_window<__half>(const __half* p, float (&xw)[24]) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    uint4 v = q[i];
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
      xw[8 * i + 2 * j + 0] = f2.x;
      xw[8 * i + 2 * j + 1] = f2.y;
    }
  }
}

template <typename T> __device__ __forceinline__ void store8_vec(T* p, const float (&y)[8]);
template <> __device__ __forceinline__ void store8_vec<float>(float* p, const float (&y)[8]) {
  float4* q = reinterpret_cast<float4*>(p);
  q[0] = make_float4(y[0], y[1], y[2], y[3]);
  q[1] = make_float4(y[4], y[5], y[6], y[7]);
}
template <> __device__ __forceinline__ void store8_vec<__nv_bfloat16>(__nv_bfloat16* p, const float (&y)[8]) {
  uint4 v;
  __nv_bfloat162 b0 = __floats2bfloat162_rn(y[0], y[1]);
  __nv_bfloat162 b1 = __floats2bfloat162_rn(y[2], y[3]);
  __nv_bfloat162 b2 = __floats2bfloat162_rn(y[4], y[5]);
  __nv_bfloat162 b3 = __floats2bfloat162_rn(y[6], y[7]);
  v.x = *reinterpret_cast<uint32_t*>(&b0); v.y = *reinterpret_cast<uint32_t*>(&b1);
  v.z = *reinterpret_cast<uint32_t*>(&b2); v.w = *reinterpret_cast<uint32_t*>(&b3);
  *reinterpret_cast<uint4*>(p) = v;
}
template <> __device__ __forceinline__ void store8_vec<__half>(__half* p, const float (&y)[8]) {
  uint4 v;
  __half2 b0 = __floats2half2_rn(y[0], y[1]);
  __half2 b1 = __floats2half2_rn(y[2], y[3]);
  __half2 b2 = __floats2half2_rn(y[4], y[5]);
  __half2 b3 = __floats2half2_rn(y[6], y[7]);
  v.x = *reinterpret_cast<uint32_t*>(&b0); v.y = *reinterpret_cast<uint32_t*>(&b1);
  v.z = *reinterpret_cast<uint32_t*>(&b2); v.w = *reinterpret_cast<uint32_t*>(&b3);
  *reinterpret_cast<uint4*>(p) = v;
}

template <typename T, bool PRECISE, bool ALIGNED>
__global__ void __launch_bounds__(kThreads)
act1d_kernel(T* __restrict__ dst, const T* __restrict__ src, const float* __restrict__ alpha_log,
             const float* __restrict__ beta_log, int64_t rows, int C, int64_t Tlen, int col_tiles) {
  __shared__ __align__(128) T tile[kRows][kPitch];
  __shared__ __align__(8) uint64_t bar;

  const int tid = threadIdx.x;
  const int64_t blk = blockIdx.x;
  const int64_t row0 = (blk / col_tiles) * kRows;
  const int64_t t0 = (blk % col_tiles) * (int64_t)kW;
  const int nrows = (int)min((int64_t)kRows, rows - row0);

  if (ALIGNED) {
    // rows are 16B aligned and T is a multiple of the 16B vector width: TMA bulk copies.
    const int64_t lo = max(t0 - kHalo, (int64_t)0);
    const int64_t hi = min(t0 + kW + kHalo, Tlen);
    const uint32_t bytes = (uint32_t)((hi - lo) * sizeof(T));
    if (tid == 0) {
      mbar_init(&bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(&bar, bytes * nrows);
      for (int r = 0; r < nrows; ++r)
        bulk_g2s(&tile[r][lo - (t0 - kHalo)], src + (row0 + r) * Tlen + lo, bytes, &bar);
    }
    mbar_wait(&bar, 0);
  } else {
    for (int idx = tid; idx < kRows * kPitch; idx += kThreads) {
      const int r = idx / kPitch, c = idx % kPitch;
      const int64_t t = t0 - kHalo + c;
      T v = from_f<T>(0.f);
      if (r < nrows && t >= 0 && t < Tlen) v = src[(row0 + r) * Tlen + t];
      tile[r][c] = v;
    }
    __syncthreads();
  }

  const int r = tid >> 5;
  const int c0 = (tid & 31) * 8;
  const int64_t tg = t0 + c0;
  if (r >= nrows || tg >= Tlen) return;
  const int64_t row = row0 + r;
  const int ch = (int)(row % C);
  float sc0, sc1;
  snake_params<PRECISE>(alpha_log[ch], beta_log[ch], sc0, sc1);
  float xw[24], y[8];
  load_window<T>(&tile[r][c0], xw);
  act1d_window<8, PRECISE>(xw, y, sc0, sc1, tg, Tlen);
  T* out = dst + row * Tlen + tg;
  if (ALIGNED) {
    store8_vec<T>(out, y);       // T % (16/sizeof(T)) == 0 and tg % 8 == 0 -> all 8 in range
  } else {
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (tg + q < Tlen) out[q] = from_f<T>(y[q]);
  }
}

template <typename T>
int launch_typed(void* dst, const void* src, const float* a, const float* b, int64_t rows, int64_t C,
                 int64_t Tlen, int precise, cudaStream_t st) {
  const int64_t col_tiles = (Tlen + kW - 1) / kW;
  const int64_t row_groups = (rows + kRows - 1) / kRows;
  const int64_t nblk = col_tiles * row_groups;
  BVG_CHECK_ARG(nblk < (1ll << 31) && col_tiles < (1ll << 31), "act1d: problem too large (%lld CTAs)", (long long)nblk);
  // vector path needs 8-element granularity on the store (8 outputs/thread) and 16B-aligned rows
  const bool aligned = (Tlen % 8 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                       ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
  dim3 grid((unsigned)nblk), block(kThreads);
  ProfScope prof(st, KC_ACT1D);
  T* d = static_cast<T*>(dst);
  const T* s = static_cast<const T*>(src);
#define BVG_ACT_LAUNCH(P, A) act1d_kernel<T, P, A><<<grid, block, 0, st>>>(d, s, a, b, rows, (int)C, Tlen, (int)col_tiles)
  if (precise) { if (aligned) BVG_ACT_LAUNCH(true, true); else BVG_ACT_LAUNCH(true, false); }
  else         { if (aligned) BVG_ACT_LAUNCH(false, true); else BVG_ACT_LAUNCH(false, false); }
#undef BVG_ACT_LAUNCH
  BVG_LAUNCHED();
  return BVG_OK;
}


// ---------------------------------------------------------------------------------------------------------
// Fast-math variant on packed fp32x2 (FFMA2): a thread owns TWO rows (channels) x 16 consecutive samples, so
// every FIR multiply-add pairs up across the two rows and the issue-slot count per sample halves (see
// act1d_core.cuh::act1d_window2).  One CTA = 16 rows x 512 samples (+8-sample halo per side), staged by TMA
// bulk copies like the kernel above.  Used when precise == 0; the libdevice-sinf parity path keeps the scalar kernel.
// ---------------------------------------------------------------------------------------------------------
constexpr int kPRows = 16;
constexpr int kPW = 512;
constexpr int kPPitch = kPW + 2 * kHalo;       // 528 staged elements per row
// Row pitches in shared memory are == 16 (mod 128) bytes and the 8 lanes of a quarter-warp own 8 different rows
// (same 16-sample segment): their 128-bit loads / stores then fall into 8 different 16-byte bank groups.  (With lanes
// spread along a row instead, every segment starts at a multiple of 64 bytes and the accesses are 4-way conflicted.)
template <typename T> struct PairSmem {
  static constexpr int kInPitch = (kPPitch * (int)sizeof(T) + 127) / 128 * 128 + 16;   // bytes
  static constexpr int kOutPitch = kPW * (int)sizeof(T) + 16;                           // bytes
  static constexpr int kBytes = kPRows * (kInPitch + kOutPitch) + 16;
};

// Split output (fp32x3 path, fp32 input only): instead of fp32 rows the tile leaves as the [hi | lo] bf16 halves of the
// c8t tensor the split tcgen05 conv reads (UmmaLayer::split), so the activated fp32 tensor never goes through HBM.
// Staging (in place of the fp32 result rows): [half][8-row group][512 time rows][8 channels] bf16.
struct SplitOut {
  __nv_bfloat16* p;     // c8t base, null = plain output
  int chunks, Tp, pad, Cp8;
};

template <bool SPLIT, typename T>
__device__ __forceinline__ void put16(uint8_t* otile_b, int row, int c0, const float (&y)[16]) {
  if (!SPLIT) {
    T* o = reinterpret_cast<T*>(otile_b + row * PairSmem<T>::kOutPitch) + c0;
    store8_vec<T>(o, *reinterpret_cast<const float(*)[8]>(&y[0]));
    store8_vec<T>(o + 8, *reinterpret_cast<const float(*)[8]>(&y[8]));
  } else {
    __nv_bfloat16* st = reinterpret_cast<__nv_bfloat16*>(otile_b);
    const int grp = row >> 3, pch = row & 7;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const __nv_bfloat16 hi = __float2bfloat16_rn(y[q]);
      st[((size_t)(0 * 2 + grp) * kPW + c0 + q) * 8 + pch] = hi;
      st[((size_t)(1 * 2 + grp) * kPW + c0 + q) * 8 + pch] = __float2bfloat16_rn(y[q] - __bfloat162float(hi));
    }
  }
}

template <typename T, bool ALIGNED, bool SPLIT = false>
__global__ void __launch_bounds__(256, 3)
act1d_pair_kernel(T* __restrict__ dst, const T* __restrict__ src, const float* __restrict__ alpha_log,
                  const float* __restrict__ beta_log, int64_t rows, int C, int64_t Tlen, int col_tiles,
                  SplitOut sp = SplitOut{nullptr, 0, 0, 0, 0}) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int IP = PairSmem<T>::kInPitch, OP = PairSmem<T>::kOutPitch;
  uint8_t* tile_b = smem_raw;                                     // [kPRows] rows of IP bytes
  uint8_t* otile_b = smem_raw + kPRows * IP;                      // [kPRows] rows of OP bytes (results, row-major)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + kPRows * (IP + OP));
  auto tile = [&](int rr) { return reinterpret_cast<T*>(tile_b + rr * IP); };
  auto otile = [&](int rr) { return reinterpret_cast<T*>(otile_b + rr * OP); };

  const int tid = threadIdx.x;
  const int64_t blk = blockIdx.x;
  const int64_t row0 = (blk / col_tiles) * kPRows;
  const int64_t t0 = (blk % col_tiles) * (int64_t)kPW;
  const int nrows = (int)min((int64_t)kPRows, rows - row0);

  // thread = rows (p, p + 8) x 16-sample group: p = tid & 7, group = tid >> 3  (a tile with <= 8 rows pairs (p, p + 4)
  // on half of the threads instead of falling back to the scalar single-row path)
  const int half = nrows > 8 ? 8 : 4;
  const int r = tid & 7, r2 = r + half;
  const int c0 = (tid >> 3) * 16;
  const int64_t tg = t0 + c0;
  float a0 = 0.f, b0 = 0.f, a1 = 0.f, b1 = 0.f;     // snake parameters first: the loads overlap the TMA wait
  if (r < nrows) snake_params<false>(alpha_log[(int)((row0 + r) % C)], beta_log[(int)((row0 + r) % C)], a0, b0);
  if (r2 < nrows) snake_params<false>(alpha_log[(int)((row0 + r2) % C)], beta_log[(int)((row0 + r2) % C)], a1, b1);

  if (ALIGNED) {
    const int64_t lo = max(t0 - kHalo, (int64_t)0);
    const int64_t hi = min(t0 + kPW + kHalo, Tlen);
    const uint32_t bytes = (uint32_t)((hi - lo) * sizeof(T));
    if (tid == 0) {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(bar, bytes * nrows);
      for (int rr = 0; rr < nrows; ++rr)
        bulk_g2s(tile(rr) + (lo - (t0 - kHalo)), src + (row0 + rr) * Tlen + lo, bytes, bar);
    }
    mbar_wait(bar, 0);
  } else {
    for (int idx = tid; idx < kPRows * kPPitch; idx += 256) {
      const int rr = idx / kPPitch, c = idx % kPPitch;
      const int64_t t = t0 - kHalo + c;
      T v = from_f<T>(0.f);
      if (rr < nrows && t >= 0 && t < Tlen) v = src[(row0 + rr) * Tlen + t];
      tile(rr)[c] = v;
    }
    __syncthreads();
  }
  constexpr int V = 16;
  if (r < half && r < nrows && tg < Tlen) {
  const int nr = r2 < nrows ? 2 : 1;
  const bool interior = (tg - 5 >= 0) && (tg + V + 4 <= Tlen - 1) && nr == 2;
  if (interior) {
    // window position j <-> tile(row)[c0 + j] (time tg - 8 + j); rows are fetched 16 bytes at a time as the stencil
    // walks down the window (everything is unrolled, so `j` is a compile-time constant in each call)
    constexpr int VE = 16 / sizeof(T);                     // elements per 128-bit shared-memory load
    const uint4* va = reinterpret_cast<const uint4*>(tile(r) + c0);
    const uint4* vb = reinterpret_cast<const uint4*>(tile(r2) + c0);
    uint4 ca = va[0], cb = vb[0];
    float ya[V], yb[V];
    act1d_window2<V>(
        [&](int j) {
          if (j % VE == 0) { ca = va[j / VE]; cb = vb[j / VE]; }
          const int e = j % VE;
          float fa, fb;
          if (sizeof(T) == 4) {
            const uint32_t wa = e == 0 ? ca.x : e == 1 ? ca.y : e == 2 ? ca.z : ca.w;
            const uint32_t wb = e == 0 ? cb.x : e == 1 ? cb.y : e == 2 ? cb.z : cb.w;
            fa = __uint_as_float(wa); fb = __uint_as_float(wb);
          } else {
            const int w = e >> 1;
            const uint32_t wa = w == 0 ? ca.x : w == 1 ? ca.y : w == 2 ? ca.z : ca.w;
            const uint32_t wb = w == 0 ? cb.x : w == 1 ? cb.y : w == 2 ? cb.z : cb.w;
            const uint16_t ha = (e & 1) ? (uint16_t)(wa >> 16) : (uint16_t)(wa & 0xffffu);
            const uint16_t hb = (e & 1) ? (uint16_t)(wb >> 16) : (uint16_t)(wb & 0xffffu);
            fa = to_f<T>(*reinterpret_cast<const T*>(&ha)); fb = to_f<T>(*reinterpret_cast<const T*>(&hb));
          }
          return pk2(fa, fb);
        },
        [&](int q, float va_, float vb_) { ya[q] = va_; yb[q] = vb_; },
        pk2(a0, a1), pk2(b0, b1));
    put16<SPLIT, T>(otile_b, r, c0, ya);
    put16<SPLIT, T>(otile_b, r2, c0, yb);
  } else {
    for (int h = 0; h < nr; ++h) {          // sequence edges / odd last row: scalar stencil with replicate padding
      const int rr = h ? r2 : r;
      float xw[V + 16], y[V];
#pragma unroll
      for (int j = 0; j < V + 16; ++j) xw[j] = to_f<T>(tile(rr)[c0 + j]);
      act1d_window<V, false>(xw, y, h ? a1 : a0, h ? b1 : b0, tg, Tlen);
      if (SPLIT) {
        put16<SPLIT, T>(otile_b, rr, c0, y);
      } else {
#pragma unroll
        for (int q = 0; q < V; ++q) otile(rr)[c0 + q] = from_f<T>(y[q]);
      }
    }
  }
  }
  // results leave row by row: TMA bulk stores when rows are 16-byte aligned, coalesced stores otherwise
  const int ncols = (int)min((int64_t)kPW, Tlen - t0);
  if (SPLIT) {
    // four chunk tiles (half x 8-row group), each a contiguous run of ncols 16-byte rows in the c8t tensor
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    auto chunk_base = [&](int ct) -> __nv_bfloat16* {               // ct = half * 2 + group; row t = 0 of that chunk
      const int64_t R = row0 + (ct & 1) * 8;
      const int64_t b = R / C, ch = (R % C) >> 3;
      return sp.p + (((int64_t)b * sp.chunks + (ct >> 1) * sp.Cp8 + ch) * sp.Tp + sp.pad) * 8;
    };
    if (tid < 4 && (tid & 1) * 8 < nrows) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(chunk_base(tid) + t0 * 8),
                   "r"(smem_u32(otile_b + (size_t)tid * kPW * 16)), "r"((uint32_t)(ncols * 16)) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    // the c8t zero halo rows in front of t = 0 and behind t = T-1 (the conv's zero padding) come from the edge tiles
    const uint4 z = make_uint4(0, 0, 0, 0);
    if (tid >= 128 && tid < 128 + 4 * 32) {
      const int ct = (tid - 128) >> 5, rowi = (tid - 128) & 31;
      if ((ct & 1) * 8 < nrows && rowi < sp.pad) {
        if (t0 == 0) *reinterpret_cast<uint4*>(chunk_base(ct) + ((int64_t)rowi - sp.pad) * 8) = z;
        if (t0 + kPW >= Tlen) *reinterpret_cast<uint4*>(chunk_base(ct) + (Tlen + rowi) * 8) = z;
      }
    }
    if (tid < 4) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    return;
  }
  if (ALIGNED) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid < nrows) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (row0 + tid) * Tlen + t0),
                   "r"(smem_u32(otile(tid))), "r"((uint32_t)(ncols * sizeof(T))) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  } else {
    __syncthreads();
    for (int idx = tid; idx < nrows * kPW; idx += 256) {
      const int rr = idx / kPW, c = idx % kPW;
      if (c < ncols) dst[(row0 + rr) * Tlen + t0 + c] = otile(rr)[c];
    }
  }
}

template <typename T>
int launch_pair(void* dst, const void* src, const float* a, const float* b, int64_t rows, int64_t C, int64_t Tlen,
                cudaStream_t st) {
  const int64_t col_tiles = (Tlen + kPW - 1) / kPW;
  const int64_t row_groups = (rows + kPRows - 1) / kPRows;
  const int64_t nblk = col_tiles * row_groups;
  BVG_CHECK_ARG(nblk < (1ll << 31), "act1d: problem too large (%lld CTAs)", (long long)nblk);
  const bool aligned = (Tlen % 16 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                       ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
  const size_t smem = PairSmem<T>::kBytes;
  static std::atomic<uint64_t> opted_a{0}, opted_u{0};
  BVG_TRY(smem_opt_in(act1d_pair_kernel<T, true>, opted_a, (int)smem));
  BVG_TRY(smem_opt_in(act1d_pair_kernel<T, false>, opted_u, (int)smem));
  ProfScope prof(st, KC_ACT1D);
  T* d = static_cast<T*>(dst);
  const T* s = static_cast<const T*>(src);
  if (aligned) act1d_pair_kernel<T, true><<<(unsigned)nblk, 256, smem, st>>>(d, s, a, b, rows, (int)C, Tlen, (int)col_tiles);
  else act1d_pair_kernel<T, false><<<(unsigned)nblk, 256, smem, st>>>(d, s, a, b, rows, (int)C, Tlen, (int)col_tiles);
  BVG_LAUNCHED();
  return BVG_OK;
}

}  // namespace

// Activation1d on fp32 [B,C,T] with the result written as the [hi | lo] bf16 c8t tensor of the split convs.
// BVG_ERR_STATE (nothing launched) when the shape does not qualify (C % 8, T % 16, 16-byte alignment).
int act1d_split_launch(__nv_bfloat16* dst_c8t, int dst_chunks, int dst_Tp, int dst_pad, const float* src,
                       const float* alpha_log, const float* beta_log, int64_t B, int64_t C, int64_t T, cudaStream_t st) {
  if (B == 0 || T == 0) return BVG_OK;
  if (C % 8 != 0 || T % 16 != 0 || (reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst_c8t) & 15))
    return BVG_ERR_STATE;
  BVG_CHECK_ARG(dst_chunks == 2 * (int)(C / 8) && dst_Tp == T + 2 * dst_pad, "act1d_split: destination geometry");
  const int64_t rows = B * C;
  const int64_t col_tiles = (T + kPW - 1) / kPW, row_groups = (rows + kPRows - 1) / kPRows;
  const int64_t nblk = col_tiles * row_groups;
  BVG_CHECK_ARG(nblk < (1ll << 31), "act1d_split: problem too large");
  const size_t smem = PairSmem<float>::kBytes;
  static std::atomic<uint64_t> opted{0};
  BVG_TRY(smem_opt_in(act1d_pair_kernel<float, true, true>, opted, (int)smem));
  SplitOut sp{dst_c8t, dst_chunks, dst_Tp, dst_pad, (int)(C / 8)};
  ProfScope prof(st, KC_ACT1D);
  act1d_pair_kernel<float, true, true><<<(unsigned)nblk, 256, smem, st>>>(nullptr, src, alpha_log, beta_log, rows, (int)C, T,
                                                                         (int)col_tiles, sp);
  BVG_LAUNCHED();
  return BVG_OK;
}

int act1d_launch(void* dst, const void* src, const float* alpha_log, const float* beta_log,
                 int64_t B, int64_t C, int64_t T, int dtype, int precise, cudaStream_t st) {
  BVG_CHECK_ARG(dst && src && alpha_log && beta_log, "act1d: null pointer");
  BVG_CHECK_ARG(dst != src, "act1d: in-place is not supported (halo reads would race)");
  BVG_CHECK_ARG(B >= 0 && C > 0 && T >= 0 && C < (1ll << 31), "act1d: bad shape B=%lld C=%lld T=%lld", (long long)B, (long long)C, (long long)T);
  if (B == 0 || T == 0) return BVG_OK;
  const int64_t rows = B * C;
  if (!precise) {
    switch (dtype) {
      case BVG_F32: return launch_pair<float>(dst, src, alpha_log, beta_log, rows, C, T, st);
      case BVG_BF16: return launch_pair<__nv_bfloat16>(dst, src, alpha_log, beta_log, rows, C, T, st);
      case BVG_F16: return launch_pair<__half>(dst, src, alpha_log, beta_log, rows, C, T, st);
      default: set_error("act1d: unsupported dtype %d", dtype); return BVG_ERR_INVALID;
    }
  }
  switch (dtype) {
    case BVG_F32: return launch_typed<float>(dst, src, alpha_log, beta_log, rows, C, T, precise, st);
    case BVG_BF16: return launch_typed<__nv_bfloat16>(dst, src, alpha_log, beta_log, rows, C, T, precise, st);
    case BVG_F16: return launch_typed<__half>(dst, src, alpha_log, beta_log, rows, C, T, precise, st);
    default: set_error("act1d: unsupported dtype %d", dtype); return BVG_ERR_INVALID;
  }
}

}  // namespace bvg
