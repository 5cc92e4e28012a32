// Speaker-encoder kernels that replace chains of small launches (ECAPA_TDNN.py of the reference's BigVGAN package).
//
// res2net_chain_kernel -- Res2NetBlock.forward (ECAPA_TDNN.py:179-191): the 512 channels of a SERes2NetBlock are eight
//   groups of 64; group 0 passes through, y_1 = f_1(x_1), y_j = f_j(x_j + y_{j-1}) with f = Conv1d(64 -> 64, k = 3,
//   dilation d, reflect "same" padding) -> ReLU -> eval-BatchNorm (:126-128).  The seven convs depend on each other
//   over ALL of time, so as separate launches they are seven latency-bound kernels per block (21 per decode, 0.7 ms).
//   Here one 4-CTA thread-block cluster owns one utterance for the whole chain: every CTA keeps the current conv
//   input (64 channels x T, fp32, reflect halo included) in shared memory, computes 16 of the 64 output channels, and
//   stores its slice of y_j into the input buffers of all four CTAs over distributed shared memory; two cluster
//   barriers per conv.  Exact fp32 FMAs (the parity path and the bf16 decode path share it); the output goes to
//   fp32 [B,512,T] and / or straight into the c8t bf16 tensor the following 1x1 GEMM (tdnn2) reads.
//
// se_gate_kernel -- SEBlock (ECAPA_TDNN.py:228-242, lengths = None): mean over time -> 1x1 conv -> ReLU -> 1x1 conv ->
//   sigmoid, one CTA per utterance (three launches before).
//
// matvec_multi_kernel -- the per-stage speaker-condition vectors (models.py:184,229,236: cond_layer, conds[i], all
//   1x1 convs of the same [B,512,1] embedding) in one launch.
#include <cooperative_groups.h>

#include "bvg_common.cuh"
#include "umma.cuh"

namespace cg = cooperative_groups;

namespace bvg {
namespace {

constexpr int kCH = 64;            // channels per Res2Net group
// RANKS CTAs per cluster (4, or 8 for small batches: twice the SMs per utterance), each owning 64 / RANKS output channels;
// 256 threads = (64 / RANKS / 4) channel quads x LANES time lanes

struct ChainParams {
  const float* y1;                 // [B, 8*64, T] fp32 (tdnn1 output)
  float* y2;                       // [B, 8*64, T] fp32 or null
  __nv_bfloat16* yc;               // c8t bf16 [B][64 chunks][Tp][8] or null
  int yc_tp, yc_pad;
  const float* w[7];               // [3][64][64] (k, ci, co)
  const float* bias[7];
  const float* bn_scale[7];
  const float* bn_shift[7];
  int T, dil, row;                 // row = floats per shared-memory row
  int npass;
  int staged;                      // shared memory holds a [64][T] staging buffer for the next conv's x_j
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }
// cluster barrier without memory ordering: "I have finished READING my input buffer" (the values are consumed)
__device__ __forceinline__ void cluster_sync_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}

template <int MP, int RANKS>
__global__ void __launch_bounds__(256) res2net_chain_kernel(const ChainParams P) {
  constexpr int kRanks = RANKS, kCo = kCH / RANKS, kNQ = kCo / 4, kLanes = 256 / kNQ;
  extern __shared__ __align__(16) float smem_f[];
  float* in = smem_f;                                   // [64][row]: position dil + t holds sample t
  float* ws = smem_f + (size_t)kCH * P.row;             // [2][3][64][16]: this conv's weight slice, the next one's in flight
  float* stage = ws + 2 * 3 * kCH * kCo;                // [64][T]: x_{j+1} lands here (cp.async) while conv j runs
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  const int cgq = tid % kNQ, tt = tid / kNQ;
  const int T = P.T, dil = P.dil, row = P.row;
  const float* y1b = P.y1 + (int64_t)b * 8 * kCH * T;
  float* y2b = P.y2 ? P.y2 + (int64_t)b * 8 * kCH * T : nullptr;
  __nv_bfloat16* ycb = P.yc ? P.yc + (int64_t)b * 64 * P.yc_tp * 8 : nullptr;
  float* rin[kRanks];
#pragma unroll
  for (int r = 0; r < kRanks; ++r) rin[r] = cluster.map_shared_rank(in, r);
  const int n4 = kCH * T / 4;                           // float4s of one group (64 * T * 4 bytes: always 16-byte granular)

  // x_j (64 x T contiguous floats) and conv j's weight slice, asynchronously into shared memory
  auto prefetch = [&](int j) {
    if (P.staged) {
      const float4* xj4 = reinterpret_cast<const float4*>(y1b + (int64_t)j * kCH * T);
      for (int i = tid; i < n4; i += 256) cp_async16(reinterpret_cast<float4*>(stage) + i, xj4 + i);
    }
    const float* wj = P.w[j - 1];
    float* wd = ws + (j & 1) * (3 * kCH * kCo);
    for (int i = tid; i < 3 * kCH * kNQ; i += 256) {
      const int q = i % kNQ, kc = i / kNQ;              // kc = k * 64 + ci
      cp_async16(wd + kc * kCo + q * 4, wj + (int64_t)kc * kCH + rank * kCo + q * 4);
    }
  };
  prefetch(1);
  // rows are zero beyond the halo: the strided time lanes of the last pass read (and discard) those positions
  for (int i = tid; i < kCH * (row - T); i += 256) { const int c = i / (row - T); in[c * row + T + (i - c * (row - T))] = 0.f; }
  // group 0 passes through (this CTA copies its 16 channels), eight independent loads at a time
  for (int i0 = tid; i0 < kCo * T; i0 += 256 * 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int i = i0 + u * 256; v[u] = i < kCo * T ? __ldg(y1b + (int64_t)rank * kCo * T + i) : 0.f; }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * 256;
      if (i >= kCo * T) continue;
      const int c = rank * kCo + i / T, t = i - (i / T) * T;
      if (y2b) y2b[(int64_t)c * T + t] = v[u];
      if (ycb) ycb[((int64_t)(c >> 3) * P.yc_tp + P.yc_pad + t) * 8 + (c & 7)] = __float2bfloat16_rn(v[u]);
    }
  }

  for (int j = 1; j < 8; ++j) {
    // (a) conv input x_j (+ y_{j-1}, already in the buffer), reflect halo
    cp_async_wait_all();
    __syncthreads();
    {
      const float4* xj4 = P.staged ? reinterpret_cast<const float4*>(stage)
                                   : reinterpret_cast<const float4*>(y1b + (int64_t)j * kCH * T);
      for (int i0 = tid; i0 < n4; i0 += 256 * 6) {
        float4 v[6];
#pragma unroll
        for (int u = 0; u < 6; ++u) { const int i = i0 + u * 256; v[u] = i < n4 ? xj4[i] : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
        for (int u = 0; u < 6; ++u) {
          const int i = i0 + u * 256;
          if (i >= n4) continue;
          int c = (i * 4) / T, t = i * 4 - c * T;
          const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float* d = in + c * row + dil + t;
            *d = (j == 1) ? e[q] : *d + e[q];
            if (++t == T) { t = 0; ++c; }
          }
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < kCH * dil; i += 256) {
      const int c = i / dil, o = 1 + (i - c * dil);
      float* rp = in + c * row + dil;
      rp[-o] = rp[o];
      rp[T - 1 + o] = rp[T - 1 - o];
    }
    __syncthreads();
    if (j < 7) prefetch(j + 1);                         // lands while this conv computes
    const float* wsj = ws + (j & 1) * (3 * kCH * kCo);
    const int co = rank * kCo + cgq * 4;                // first of this thread's four output channels
    const float4 bi = *reinterpret_cast<const float4*>(P.bias[j - 1] + co);
    const float4 sc = *reinterpret_cast<const float4*>(P.bn_scale[j - 1] + co);
    const float4 sh = *reinterpret_cast<const float4*>(P.bn_shift[j - 1] + co);
    float outv[2][4][MP];                               // (at most two passes keep their results in registers)
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      if (p >= P.npass) break;
      float acc[4][MP];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int m = 0; m < MP; ++m) acc[i][m] = 0.f;
      const float* xin = in + p * (kLanes * MP) + tt;
#pragma unroll 2
      for (int ci = 0; ci < kCH; ++ci) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float4 w4 = *reinterpret_cast<const float4*>(wsj + (k * kCH + ci) * kCo + cgq * 4);
          const float* xr = xin + ci * row + k * dil;
#pragma unroll
          for (int m = 0; m < MP; ++m) {
            const float xv = xr[m * kLanes];
            acc[0][m] = fmaf(w4.x, xv, acc[0][m]);
            acc[1][m] = fmaf(w4.y, xv, acc[1][m]);
            acc[2][m] = fmaf(w4.z, xv, acc[2][m]);
            acc[3][m] = fmaf(w4.w, xv, acc[3][m]);
          }
        }
      }
#pragma unroll
      for (int m = 0; m < MP; ++m) {
        outv[p][0][m] = fmaf(fmaxf(acc[0][m] + bi.x, 0.f), sc.x, sh.x);
        outv[p][1][m] = fmaf(fmaxf(acc[1][m] + bi.y, 0.f), sc.y, sh.y);
        outv[p][2][m] = fmaf(fmaxf(acc[2][m] + bi.z, 0.f), sc.z, sh.z);
        outv[p][3][m] = fmaf(fmaxf(acc[3][m] + bi.w, 0.f), sc.w, sh.w);
      }
    }
    if (j < 7) {
      cluster_sync_relaxed();                           // every CTA has finished reading its input buffer
      // (d) y_j into the input buffers of the whole cluster (the next conv adds x_{j+1})
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        if (p >= P.npass) break;
#pragma unroll
        for (int m = 0; m < MP; ++m) {
          const int t = p * (kLanes * MP) + m * kLanes + tt;
          if (t >= T) continue;
#pragma unroll
          for (int r = 0; r < kRanks; ++r) {
            float* d = rin[r] + co * row + dil + t;
            d[0] = outv[p][0][m]; d[row] = outv[p][1][m]; d[2 * row] = outv[p][2][m]; d[3 * row] = outv[p][3][m];
          }
        }
      }
      cluster.sync();                                   // y_j is in place everywhere
    }
    // y_j to global memory, after the barrier: a release barrier would otherwise wait for these stores to drain
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      if (p >= P.npass) break;
#pragma unroll
      for (int m = 0; m < MP; ++m) {
        const int t = p * (kLanes * MP) + m * kLanes + tt;
        if (t >= T) continue;
        const float v0 = outv[p][0][m], v1 = outv[p][1][m], v2 = outv[p][2][m], v3 = outv[p][3][m];
        const int c = j * kCH + co;
        if (y2b) {
          float* o = y2b + (int64_t)c * T + t;
          o[0] = v0; o[T] = v1; o[2 * (int64_t)T] = v2; o[3 * (int64_t)T] = v3;
        }
        if (ycb) {
          const __nv_bfloat162 lo = __floats2bfloat162_rn(v0, v1), hi = __floats2bfloat162_rn(v2, v3);
          uint2 pk;
          pk.x = *reinterpret_cast<const uint32_t*>(&lo); pk.y = *reinterpret_cast<const uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(ycb + ((int64_t)(c >> 3) * P.yc_tp + P.yc_pad + t) * 8 + (c & 7)) = pk;
        }
      }
    }
  }
  cluster.sync();                                       // no CTA leaves while a peer may still address its shared memory
}

}  // namespace

// Whether the chain kernel takes this shape (else the caller runs the seven convs as separate launches).
static bool chain_geometry(int T, int dil, int ranks, int* mp, int* npass, int* row, size_t* smem, int* staged) {
  if (dil < 1 || dil > 8 || T <= dil) return false;
  const int lanes = 256 / (kCH / ranks / 4);
  const int m_lo = ranks == 8 ? 2 : 3, m_hi = ranks == 8 ? 3 : 5;
  int best = 0, best_cover = 1 << 30;
  for (int m = m_lo; m <= m_hi; ++m) {
    const int np = (T + lanes * m - 1) / (lanes * m);
    if (np > 2) continue;
    if (np * lanes * m < best_cover) { best_cover = np * lanes * m; best = m; }
  }
  if (!best) return false;
  *mp = best;
  *npass = best_cover / (lanes * best);
  *row = ((best_cover + 2 * dil + 3) / 4) * 4 + 4;
  const size_t base = ((size_t)kCH * *row + 2 * 3 * kCH * (kCH / ranks)) * sizeof(float);
  const size_t with_stage = base + (size_t)kCH * T * sizeof(float);
  *staged = with_stage <= 227 * 1024;
  *smem = *staged ? with_stage : base;
  return *smem <= 227 * 1024;
}

template <int MP, int RANKS>
static int chain_launch_t(const ChainParams& P, int64_t B, size_t smem, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(RANKS, (unsigned)B);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = RANKS; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  static std::atomic<uint64_t> opted{0};
  BVG_TRY(smem_opt_in(res2net_chain_kernel<MP, RANKS>, opted, 227 * 1024));
  BVG_CUDA(cudaLaunchKernelEx(&cfg, res2net_chain_kernel<MP, RANKS>, P));
  return BVG_OK;
}

int res2net_chain_launch(const float* y1, float* y2, const C8T* yc, const float* const* w, const float* const* bias,
                         const float* const* bn_scale, const float* const* bn_shift, int dil, int64_t B, int64_t T,
                         bool* taken, cudaStream_t st) {
  int mp, npass, row, staged;
  size_t smem;
  *taken = false;
  if (B == 0 || B > 65535 || T > (1 << 20)) return BVG_OK;
  int num_sms = 0;
  BVG_TRY(current_device_sms(&num_sms));
  // 8-CTA clusters while they all fit on the GPU at once (one utterance: 8 SMs instead of 4), else 4-CTA clusters
  int ranks = (B * 8 <= num_sms && BVG_ENV_ONCE("BVG_ECAPA_RANKS8", 1)) ? 8 : 4;
  if (!chain_geometry((int)T, dil, ranks, &mp, &npass, &row, &smem, &staged)) {
    if (ranks == 4) return BVG_OK;
    ranks = 4;
    if (!chain_geometry((int)T, dil, ranks, &mp, &npass, &row, &smem, &staged)) return BVG_OK;
  }
  ChainParams P;
  P.y1 = y1; P.y2 = y2;
  P.yc = yc ? yc->p : nullptr; P.yc_tp = yc ? yc->Tp : 0; P.yc_pad = yc ? yc->pad : 0;
  BVG_CHECK_ARG(!yc || (yc->chunks == 8 * kCH / 8 && yc->T == (int)T), "res2net_chain: c8t output geometry");
  for (int j = 0; j < 7; ++j) { P.w[j] = w[j]; P.bias[j] = bias[j]; P.bn_scale[j] = bn_scale[j]; P.bn_shift[j] = bn_shift[j]; }
  P.T = (int)T; P.dil = dil; P.row = row; P.npass = npass; P.staged = staged;
  ProfScope prof(st, KC_OTHER);
  if (ranks == 8) {
    if (mp == 2) BVG_TRY((chain_launch_t<2, 8>(P, B, smem, st)));
    else BVG_TRY((chain_launch_t<3, 8>(P, B, smem, st)));
  } else {
    if (mp == 3) BVG_TRY((chain_launch_t<3, 4>(P, B, smem, st)));
    else if (mp == 4) BVG_TRY((chain_launch_t<4, 4>(P, B, smem, st)));
    else BVG_TRY((chain_launch_t<5, 4>(P, B, smem, st)));
  }
  BVG_LAUNCHED();
  *taken = true;
  return BVG_OK;
}

}  // namespace bvg
