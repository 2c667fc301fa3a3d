// backward.cu -- bandwidth-bound kernels of the DSM training step (reference: loss.backward() through
// sbgm/score_unet.py:936-985 and the modules it calls; torch autograd supplies these in the reference).
//
//   * train-mode normalisation forward from a statistics table (BatchNorm batch statistics / GroupNorm), and the
//     running-statistics update of nn.BatchNorm2d;
//   * one unified normalisation backward for BatchNorm (statistics over n,h,w per channel) and GroupNorm /
//     InstanceNorm (statistics over h,w,channels-of-group per sample) with the fused epilogue of the forward
//     (+skip / +residual, +time projection before or after the activation, activation) differentiated in place;
//   * LayerNorm backward, activation forward/backward, bilinear x2 upsample backward (the adjoint stencil),
//     gradient accumulation, per-channel sums (bias gradients), time-embedding/projection backward, DSM loss backward.
//
// Everything is deterministic: two-stage reductions with fixed partition and order, no floating-point atomics.
#include "common.cuh"

namespace sbgm {

__device__ __forceinline__ float act_grad(float u, int act) {
  switch (act) {
    case SBGM_ACT_RELU: return u > 0.0f ? 1.0f : 0.0f;
    case SBGM_ACT_SILU: {
      const float s = 1.0f / (1.0f + expf(-u));
      return s * (1.0f + u * (1.0f - s));
    }
    case SBGM_ACT_GELU: {
      const float cdf = 0.5f * (1.0f + erff(u * 0.70710678118654752f));
      return cdf + u * 0.3989422804014327f * expf(-0.5f * u * u);
    }
    default: return 1.0f;
  }
}

static int bgrid_for(size_t items, int block, int max_blocks = 148 * 16) {
  size_t g = (items + block - 1) / block;
  if (g < 1) g = 1;
  if (g > static_cast<size_t>(max_blocks)) g = max_blocks;
  return static_cast<int>(g);
}

constexpr int kNormChunks = 32;   // must equal kGnChunks of elementwise.cu (the forward partial statistics)

// ---- statistics tables ------------------------------------------------------------------------
// GroupNorm: finish [n][chunks][pgroups][2] partial (sum, sumsq) into stats[n][groups][2] = (mean, rstd).
__global__ void gn_stats_finalize_kernel(const float* __restrict__ partials, int chunks, int pgroups, int groups, double cnt,
                                         float eps, float* __restrict__ stats) {
  pdl_grid_sync();
  const int n = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int sub = pgroups / groups;
  for (int g = warp; g < groups; g += nwarp) {
    double s = 0.0, q = 0.0;
    for (int k = lane; k < chunks * sub; k += 32) {
      const int ck = k / sub, sg = k - ck * sub;
      const float2 p = __ldg(reinterpret_cast<const float2*>(partials + ((static_cast<size_t>(n) * chunks + ck) * pgroups + g * sub + sg) * 2));
      s += p.x;
      q += p.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) {
      const double mean = s / cnt;
      const double var = fmax(q / cnt - mean * mean, 0.0);
      stats[(static_cast<size_t>(n) * groups + g) * 2] = static_cast<float>(mean);
      stats[(static_cast<size_t>(n) * groups + g) * 2 + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
  }
}

// BatchNorm (training): finish per-channel partials [n][chunks][c][2] over the whole batch; update the running
// statistics the way nn.BatchNorm2d does (momentum, unbiased variance).  One warp per channel.
__global__ void bn_stats_finalize_kernel(const float* __restrict__ partials, int n, int chunks, int c, double cnt, float eps,
                                         float momentum, float* __restrict__ stats, float* __restrict__ running_mean,
                                         float* __restrict__ running_var) {
  pdl_grid_sync();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= c) return;
  double s = 0.0, q = 0.0;
  for (int k = lane; k < n * chunks; k += 32) {
    const float2 p = __ldg(reinterpret_cast<const float2*>(partials + (static_cast<size_t>(k) * c + warp) * 2));
    s += p.x;
    q += p.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane == 0) {
    const double mean = s / cnt;
    const double var = fmax(q / cnt - mean * mean, 0.0);
    stats[2 * warp] = static_cast<float>(mean);
    stats[2 * warp + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    if (running_mean) running_mean[warp] = (1.0f - momentum) * running_mean[warp] + momentum * static_cast<float>(mean);
    if (running_var) {
      const double unbiased = cnt > 1.0 ? var * cnt / (cnt - 1.0) : var;
      running_var[warp] = (1.0f - momentum) * running_var[warp] + momentum * static_cast<float>(unbiased);
    }
  }
}

// ---- normalisation forward from a statistics table ------------------------------------------------
//   u = (x - mean) * rstd * gamma + beta + add + (tproj_pre ? tproj : 0);  y = act(u) + (tproj_pre ? 0 : tproj)
// stats index = n * n_stride + ch / cpg  (BatchNorm: n_stride = 0, cpg = 1; GroupNorm: n_stride = groups)
struct NormArgs {
  const void* x; size_t x_plane;
  const float* stats; int n_stride; int cpg;
  const float* gamma; const float* beta;
  const void* add; size_t add_plane;
  const float* tproj; int tproj_stride; int tproj_pre;
  int act, hw, c;
};

__device__ __forceinline__ void norm_coefs(const NormArgs& a, int n, int ch0, float (&mean)[8], float (&rstd)[8], float (&ga)[8],
                                           float (&sh)[8], float (&post)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = ch0 + j;
    const float* st = a.stats + (static_cast<size_t>(n) * a.n_stride + ch / a.cpg) * 2;
    mean[j] = st[0];
    rstd[j] = st[1];
    ga[j] = a.gamma ? a.gamma[ch] : 1.0f;
    sh[j] = a.beta ? a.beta[ch] : 0.0f;
    const float tp = a.tproj ? a.tproj[static_cast<size_t>(n) * a.tproj_stride + ch] : 0.0f;
    if (a.tproj_pre) { sh[j] += tp; post[j] = 0.0f; } else { post[j] = tp; }
  }
}

template <int FMT>
__global__ void norm_apply_kernel(const NormArgs a, void* __restrict__ y, size_t y_plane) {
  pdl_grid_sync();
  const int n = blockIdx.y;
  const int vecs = a.c >> 3;
  const size_t total = static_cast<size_t>(a.hw) * vecs;
  // a thread's channel vector is fixed when the grid stride is a multiple of vecs
  const size_t stride = (static_cast<size_t>(gridDim.x) * blockDim.x / vecs) * vecs;
  size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= stride) return;
  const int vec = static_cast<int>(i % vecs);
  float mean[8], rstd[8], ga[8], sh[8], post[8];
  norm_coefs(a, n, vec * 8, mean, rstd, ga, sh, post);
  for (; i < total; i += stride) {
    const size_t idx = (static_cast<size_t>(n) * a.hw) * a.c + i * 8;
    float v[8];
    Act<FMT>::load8(a.x, a.x_plane, idx, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf((v[j] - mean[j]) * rstd[j], ga[j], sh[j]);
    if (a.add) {
      float r[8];
      Act<FMT>::load8(a.add, a.add_plane, idx, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += r[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = apply_act(v[j], a.act) + post[j];
    Act<FMT>::store8(y, y_plane, idx, v);
  }
}

// ---- normalisation backward ---------------------------------------------------------------------
// Two launches per layer.  The layers are small (a few MB each, 28 of them per step), so both are built for LATENCY:
// Launch 1 (norm_bwd_sums_kernel): a block owns ALL the pixels of one sample for a column of `vpb` channel vectors
//   (threads = [pixel lanes][vpb]), so its block reduction already yields the finished per-sample sums
//     sums[n][ch][4] = sum over the pixels of  dY,  dU = dY * act'(u),  dU * xhat,  xhat
//   -- no per-chunk partials, no second pass over them.  The block also writes dtproj[n][ch] and, for GroupNorm (a column
//   always holds whole groups), the sample's projection coefficients
//     A = sum(gamma * dU) / cnt,  B = sum(gamma * dU * xhat) / cnt      (coef[stat index][2]).
//   The LAST sample of a column to finish (integer ticket per column; no floating-point atomics, every sum has a fixed
//   order) adds its channels over the samples: dgamma / dbeta, BatchNorm's per-channel A / B, and -- closed form, no pass
//   over dx -- the bias gradient of the convolution in front of the norm:
//     sum_p dx = rstd * (gamma * sum dU - hw * A - B * sum xhat)
// Launch 2 (norm_bwd_apply_kernel): dx = rstd * (gamma * dU - A - xhat * B);  dadd = dU.
// Tickets live in the first kNormTicketWords words of the scratch buffer: zero before the first use, left zero by every launch.
constexpr int kNormTicketWords = 4096;

struct NormBwdTail {
  unsigned int* tickets;      // [column]: samples of the column finished
  float* sums;                // [n][c][4]
  const float* sums_all;      // BatchNorm over the global batch (synchronised): [n_all][c][4]; else == sums
  float* coef;
  float* dgamma; float* dbeta; float* dbias_prev;
  float* dtproj; int dtproj_stride;
  int n, n_all, fixed_stats, with_coef;
  int vpb, pl, shuffle;       // block geometry: vectors per column, pixel lanes, warp-butterfly reduction (vpb a power of two <= 32)
  float inv_cnt;
};

// true in every thread of the block that takes the last of `total` tickets; the caller's global writes are published first
__device__ __forceinline__ bool last_ticket(unsigned int* ticket, unsigned int total) {
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == total - 1);
    if (is_last) *ticket = 0;          // self-cleaning: the next launch finds zero
  }
  __syncthreads();
  const bool last = is_last != 0;
  if (last) __threadfence();
  return last;
}

// Sums over the samples for the channels [ch_begin, ch_begin + ch_count), by one block: dgamma, dbeta, BatchNorm's
// coefficients, the bias gradient in front.  Latency-bound (n x channels float4 from L2): every thread takes a slice of the
// samples of one channel with eight loads in flight; the slices are added in a fixed order through `sm` (>= 4 * blockDim.x floats).
__device__ void norm_bwd_tail_cols(const NormArgs& a, const NormBwdTail& t, float* sm, int ch_begin, int ch_count) {
  const int c = a.c, nt = blockDim.x;
  const int cw = min(ch_count, nt), parts = nt / cw;          // threads = [parts][cw]
  const int part = threadIdx.x / cw, ch0 = threadIdx.x % cw;
  const bool bn = a.n_stride == 0;
  const bool gathered = bn && !t.fixed_stats && t.sums_all != t.sums;
  const float4* sums4 = reinterpret_cast<const float4*>(t.sums);
  for (int base = 0; base < ch_count; base += cw) {           // (one pass unless the column is wider than the block)
    const int ch = ch_begin + base + ch0;
    const bool live = part < parts && base + ch0 < ch_count;
    float s1 = 0.0f, s2 = 0.0f, s3 = 0.0f, g1 = 0.0f, g2 = 0.0f;
    if (live) {
      const float ga = a.gamma ? a.gamma[ch] : 1.0f;
      if (bn) {
#pragma unroll 8
        for (int k = part; k < t.n; k += parts) {
          const float4 s = __ldcg(sums4 + static_cast<size_t>(k) * c + ch);
          s1 += s.y; s2 += s.z; s3 += s.w;
        }
        if (gathered) {
#pragma unroll 8
          for (int k = part; k < t.n_all; k += parts) {
            const float4 s = __ldcg(reinterpret_cast<const float4*>(t.sums_all) + static_cast<size_t>(k) * c + ch);
            g1 += s.y; g2 += s.z;
          }
        }
      } else {
        // GroupNorm: s3 collects the bias gradient itself, its coefficients being per (sample, group)
#pragma unroll 8
        for (int k = part; k < t.n; k += parts) {
          const float4 s = __ldcg(sums4 + static_cast<size_t>(k) * c + ch);
          s1 += s.y; s2 += s.z;
          if (t.dbias_prev) {
            const size_t si = static_cast<size_t>(k) * a.n_stride + ch / a.cpg;
            const float2 cf = __ldcg(reinterpret_cast<const float2*>(t.coef) + si);
            s3 += a.stats[2 * si + 1] * (ga * s.y - a.hw * cf.x - s.w * cf.y);
          }
        }
      }
    }
    __syncthreads();
    if (live) *reinterpret_cast<float4*>(sm + (static_cast<size_t>(part) * cw + ch0) * 4) = make_float4(s1, s2, s3, g1);
    __syncthreads();
    float g2s = g2;                                     // the gathered second moment travels through a second round
    if (live && part == 0) {
      for (int q = 1; q < parts; ++q) {
        const float4 o = *reinterpret_cast<const float4*>(sm + (static_cast<size_t>(q) * cw + ch0) * 4);
        s1 += o.x; s2 += o.y; s3 += o.z; g1 += o.w;
      }
    }
    if (gathered) {
      __syncthreads();
      if (live) sm[static_cast<size_t>(part) * cw + ch0] = g2;
      __syncthreads();
      if (live && part == 0) for (int q = 1; q < parts; ++q) g2s += sm[static_cast<size_t>(q) * cw + ch0];
    }
    if (live && part == 0) {
      const float ga = a.gamma ? a.gamma[ch] : 1.0f;
      if (t.dgamma) { t.dgamma[ch] = s2; t.dbeta[ch] = s1; }
      if (bn) {                         // one statistics group per channel, spanning the (global) batch
        float ca = 0.0f, cb = 0.0f;
        if (!t.fixed_stats) {           // eval mode: the statistics are constants, no projection terms
          ca = ga * (gathered ? g1 : s1) * t.inv_cnt;
          cb = ga * (gathered ? g2s : s2) * t.inv_cnt;
        }
        t.coef[2 * ch] = ca;
        t.coef[2 * ch + 1] = cb;
        if (t.dbias_prev) t.dbias_prev[ch] = a.stats[2 * ch + 1] * (ga * s1 - static_cast<float>(t.n) * a.hw * ca - s3 * cb);
      } else if (t.dbias_prev) {
        t.dbias_prev[ch] = s3;
      }
    }
  }
}

template <int FMT>
__global__ void __launch_bounds__(256, 2) norm_bwd_sums_kernel(const NormArgs a, const void* __restrict__ dy, size_t dy_plane, const NormBwdTail t) {
  pdl_grid_sync();
  extern __shared__ float red[];  // [rows][vpb * 32] block reduction, then 4 * blockDim.x floats for the column tail
  const int n = blockIdx.y, c = a.c;
  const int vpb = t.vpb, pl = t.pl;
  const int vl = threadIdx.x % vpb, lane = threadIdx.x / vpb;
  const int vec = blockIdx.x * vpb + vl;
  const int width = vpb * 32;                        // floats per reduction row: [vpb][8 channels][4 sums]
  const bool active = lane < pl;
  float s0[8], s1[8], s2[8], s3[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s0[j] = s1[j] = s2[j] = s3[j] = 0.0f;
  if (active) {
    float mean[8], rstd[8], ga[8], sh[8], post[8];
    norm_coefs(a, n, vec * 8, mean, rstd, ga, sh, post);
    // latency-bound: the loads of kUn pixels are issued before any is used (kUn = 4 needs 254 registers: one block per SM)
    constexpr int kUn = 2;
    for (int p = lane; p < a.hw; p += pl * kUn) {
      float v[kUn][8], g[kUn][8], r[kUn][8];
#pragma unroll
      for (int k = 0; k < kUn; ++k) {
        if (p + k * pl < a.hw) {
          const size_t idx = (static_cast<size_t>(n) * a.hw + p + k * pl) * c + vec * 8;
          Act<FMT>::load8(a.x, a.x_plane, idx, v[k]);
          Act<FMT>::load8(dy, dy_plane, idx, g[k]);
          if (a.add) Act<FMT>::load8(a.add, a.add_plane, idx, r[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < kUn; ++k) {
        if (p + k * pl >= a.hw) break;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (v[k][j] - mean[j]) * rstd[j];
          float u = fmaf(xh, ga[j], sh[j]);
          if (a.add) u += r[k][j];
          const float du = g[k][j] * act_grad(u, a.act);
          s0[j] += g[k][j];
          s1[j] += du;
          s2[j] = fmaf(du, xh, s2[j]);
          s3[j] += xh;
        }
      }
    }
  }
  // ---- block reduction over the pixel lanes -> row 0 of `red` = the sample's sums for this column ----
  int rows;
  if (t.shuffle) {                  // threads `vpb` apart in a warp share a channel vector: butterfly, then one row per warp
    for (int off = vpb; off < 32; off <<= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s0[j] += __shfl_xor_sync(0xffffffffu, s0[j], off);
        s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], off);
        s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], off);
        s3[j] += __shfl_xor_sync(0xffffffffu, s3[j], off);
      }
    }
    rows = blockDim.x >> 5;
    if ((threadIdx.x & 31) < vpb) {
      float* o = red + static_cast<size_t>(threadIdx.x >> 5) * width + vl * 32;
#pragma unroll
      for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(o + j * 4) = make_float4(s0[j], s1[j], s2[j], s3[j]);
    }
  } else {                          // one row per pixel lane
    rows = pl;
    if (active) {
      float* o = red + static_cast<size_t>(lane) * width + vl * 32;
#pragma unroll
      for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(o + j * 4) = make_float4(s0[j], s1[j], s2[j], s3[j]);
    }
  }
  __syncthreads();
  const int ch_begin = blockIdx.x * vpb * 8, ch_count = vpb * 8;
  for (int i = threadIdx.x; i < width; i += blockDim.x) {
    float s = red[i];
    for (int r = 1; r < rows; ++r) s += red[static_cast<size_t>(r) * width + i];
    red[i] = s;                     // (row 0, column i: read and written by this thread only)
    const int ch = ch_begin + (i >> 2), which = i & 3;
    t.sums[(static_cast<size_t>(n) * c + ch) * 4 + which] = s;
    if (t.dtproj && which == (a.tproj_pre ? 1 : 0)) t.dtproj[static_cast<size_t>(n) * t.dtproj_stride + ch] = s;
  }
  if (!t.with_coef) return;         // synchronised BatchNorm, stage 1: the sums are gathered over the ranks first
  __syncthreads();
  if (a.n_stride != 0) {            // GroupNorm: the column holds whole groups (host-checked)
    const int g_begin = ch_begin / a.cpg, g_count = ch_count / a.cpg;
    for (int gl = threadIdx.x; gl < g_count; gl += blockDim.x) {
      float ca = 0.0f, cb = 0.0f;
      for (int j = 0; j < a.cpg; ++j) {
        const int ch = (g_begin + gl) * a.cpg + j;
        const float ga = a.gamma ? a.gamma[ch] : 1.0f;
        ca = fmaf(ga, red[(ch - ch_begin) * 4 + 1], ca);
        cb = fmaf(ga, red[(ch - ch_begin) * 4 + 2], cb);
      }
      const size_t si = static_cast<size_t>(n) * a.n_stride + g_begin + gl;
      t.coef[2 * si] = ca * t.inv_cnt;
      t.coef[2 * si + 1] = cb * t.inv_cnt;
    }
  }
  // ---- last sample of this column: sums over the samples ----
  if (!last_ticket(t.tickets + blockIdx.x, gridDim.y)) return;
  norm_bwd_tail_cols(a, t, red + width, ch_begin, ch_count);
}

// synchronised BatchNorm, stage 2: the per-channel tail alone (a block per column), over the gathered sums of every rank
__global__ void norm_bwd_tail_kernel(const NormArgs a, const NormBwdTail t) {
  pdl_grid_sync();
  extern __shared__ float red[];
  norm_bwd_tail_cols(a, t, red, blockIdx.x * t.vpb * 8, t.vpb * 8);
}

// Stage 3: dx = rstd * (gamma * dU - A - xhat * B);  dadd = dU.
template <int FMT>
__global__ void norm_bwd_apply_kernel(const NormArgs a, const void* __restrict__ dy, size_t dy_plane, const float* __restrict__ coef,
                                      void* __restrict__ dx, size_t dx_plane, void* __restrict__ dadd, size_t dadd_plane) {
  pdl_grid_sync();
  const int n = blockIdx.y;
  const int vecs = a.c >> 3;
  const size_t total = static_cast<size_t>(a.hw) * vecs;
  const size_t stride = (static_cast<size_t>(gridDim.x) * blockDim.x / vecs) * vecs;
  size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= stride) return;
  const int vec = static_cast<int>(i % vecs);
  float mean[8], rstd[8], ga[8], sh[8], post[8], ca[8], cb[8];
  norm_coefs(a, n, vec * 8, mean, rstd, ga, sh, post);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float* cf = coef + (static_cast<size_t>(n) * a.n_stride + (vec * 8 + j) / a.cpg) * 2;
    ca[j] = cf[0];
    cb[j] = cf[1];
  }
  for (; i < total; i += stride) {
    const size_t idx = (static_cast<size_t>(n) * a.hw) * a.c + i * 8;
    float v[8], g[8], r[8], du[8], o[8];
    Act<FMT>::load8(a.x, a.x_plane, idx, v);
    Act<FMT>::load8(dy, dy_plane, idx, g);
    if (a.add) Act<FMT>::load8(a.add, a.add_plane, idx, r);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (v[j] - mean[j]) * rstd[j];
      float u = fmaf(xh, ga[j], sh[j]);
      if (a.add) u += r[j];
      du[j] = g[j] * act_grad(u, a.act);
      o[j] = rstd[j] * (ga[j] * du[j] - ca[j] - xh * cb[j]);
    }
    Act<FMT>::store8(dx, dx_plane, idx, o);
    if (dadd) Act<FMT>::store8(dadd, dadd_plane, idx, du);
  }
}

// ---- LayerNorm backward: one warp per token row, grid-stride; dgamma / dbeta via per-block partials -------
constexpr int kLnBwdBlocks = 296;
template <int FMT>
__global__ void layernorm_bwd_kernel(const void* __restrict__ dy, size_t dy_plane, const void* __restrict__ x, size_t x_plane,
                                     const float* __restrict__ gamma, float eps, void* __restrict__ dx, size_t dx_plane,
                                     int rows, int c, float* __restrict__ partials) {
  pdl_grid_sync();
  extern __shared__ float red[];   // [warps][c][2]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int vecs = c >> 3;
  float dg[2][8], db[2][8];
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) dg[k][j] = db[k][j] = 0.0f;
  for (int row = blockIdx.x * nwarp + warp; row < rows; row += gridDim.x * nwarp) {
    float v[2][8], g[2][8];
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int vec = lane + 32 * k;
      if (vec < vecs) {
        Act<FMT>::load8(x, x_plane, static_cast<size_t>(row) * c + vec * 8, v[k]);
        Act<FMT>::load8(dy, dy_plane, static_cast<size_t>(row) * c + vec * 8, g[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[k][j];
      }
    }
    const float mean = warp_sum(s) / c;
    float q = 0.0f;
#pragma unroll
    for (int k = 0; k < 2; ++k)
      if (lane + 32 * k < vecs) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[k][j] - mean;
          q = fmaf(d, d, q);
        }
      }
    const float rstd = rsqrtf(warp_sum(q) / c + eps);
    float a = 0.0f, b = 0.0f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int vec = lane + 32 * k;
      if (vec < vecs) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (v[k][j] - mean) * rstd;
          const float dxh = g[k][j] * gamma[vec * 8 + j];
          dg[k][j] = fmaf(g[k][j], xh, dg[k][j]);
          db[k][j] += g[k][j];
          a += dxh;
          b = fmaf(dxh, xh, b);
          v[k][j] = xh;
          g[k][j] = dxh;
        }
      }
    }
    a = warp_sum(a) / c;
    b = warp_sum(b) / c;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int vec = lane + 32 * k;
      if (vec < vecs) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (g[k][j] - a - v[k][j] * b);
        Act<FMT>::store8(dx, dx_plane, static_cast<size_t>(row) * c + vec * 8, o);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int vec = lane + 32 * k;
    if (vec < vecs) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        red[(static_cast<size_t>(warp) * c + vec * 8 + j) * 2] = dg[k][j];
        red[(static_cast<size_t>(warp) * c + vec * 8 + j) * 2 + 1] = db[k][j];
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c * 2; i += blockDim.x) {
    float acc = 0.0f;
    for (int w = 0; w < nwarp; ++w) acc += red[static_cast<size_t>(w) * c * 2 + i];
    partials[static_cast<size_t>(blockIdx.x) * c * 2 + i] = acc;
  }
}
__global__ void layernorm_bwd_finish_kernel(const float* __restrict__ partials, int blocks, int c, float* __restrict__ dgamma,
                                            float* __restrict__ dbeta) {
  pdl_grid_sync();
  const int ch = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (ch >= c) return;
  float a = 0.0f, b = 0.0f;
  for (int k = lane; k < blocks; k += 32) {
    a += partials[(static_cast<size_t>(k) * c + ch) * 2];
    b += partials[(static_cast<size_t>(k) * c + ch) * 2 + 1];
  }
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) {
    dgamma[ch] = a;
    dbeta[ch] = b;
  }
}

// ---- elementwise --------------------------------------------------------------------------------
template <int FMT>
__global__ void act_fwd_kernel(const void* __restrict__ x, size_t xp, void* __restrict__ y, size_t yp, size_t nvec, int act) {
  pdl_grid_sync();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < nvec; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float v[8];
    Act<FMT>::load8(x, xp, i * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = apply_act(v[j], act);
    Act<FMT>::store8(y, yp, i * 8, v);
  }
}
template <int FMT>
__global__ void act_bwd_kernel(const void* __restrict__ dy, size_t dyp, const void* __restrict__ x, size_t xp, void* __restrict__ dx,
                               size_t dxp, size_t nvec, int act) {
  pdl_grid_sync();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < nvec; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float v[8], g[8];
    Act<FMT>::load8(x, xp, i * 8, v);
    Act<FMT>::load8(dy, dyp, i * 8, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= act_grad(v[j], act);
    Act<FMT>::store8(dx, dxp, i * 8, g);
  }
}
template <int FMT>
__global__ void add_inplace_kernel(void* __restrict__ dst, size_t dp, const void* __restrict__ src, size_t sp, size_t nvec) {
  pdl_grid_sync();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < nvec; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float a[8], b[8];
    Act<FMT>::load8(dst, dp, i * 8, a);
    Act<FMT>::load8(src, sp, i * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += b[j];
    Act<FMT>::store8(dst, dp, i * 8, a);
  }
}

// per-sample channel sums: out[n][c] = sum over the hw pixels of sample n (two-stage via [n][chunks][c] partials)
template <int FMT>
__global__ void chansum_partial_kernel(const void* __restrict__ x, size_t plane, int hw, int c, float* __restrict__ partials, int chunks) {
  pdl_grid_sync();
  extern __shared__ float red[];  // [lanes][c]
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int vecs = c >> 3, lanes = blockDim.x / vecs;
  const int vec = threadIdx.x % vecs, lane = threadIdx.x / vecs;
  const int per_chunk = (hw + chunks - 1) / chunks;
  const int p_begin = chunk * per_chunk, p_end = min(hw, p_begin + per_chunk);
  if (lane < lanes) {
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = p_begin + lane; p < p_end; p += lanes) {
      float v[8];
      Act<FMT>::load8(x, plane, (static_cast<size_t>(n) * hw + p) * c + vec * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[static_cast<size_t>(lane) * c + vec * 8 + j] = s[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    float acc = 0.0f;
    for (int l = 0; l < lanes; ++l) acc += red[static_cast<size_t>(l) * c + i];
    partials[(static_cast<size_t>(n) * chunks + chunk) * c + i] = acc;
  }
}
// out_n[n][c] (nullable) and out_total[c] (nullable).  One warp per channel: lane <-> chunk (kNormChunks == 32).
__global__ void chansum_finish_kernel(const float* __restrict__ partials, int n, int c, int chunks, float* __restrict__ out_n,
                                      int out_n_stride, float* __restrict__ out_total) {
  pdl_grid_sync();
  const int ch = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (ch >= c) return;
  float tot = 0.0f;
  if (out_n) {
    // per-sample sums: lanes over the samples, each adds its sample's chunks (a loop of 64 warp reductions in sequence was a
    // 30 us kernel for 16 KB of partials); the total adds the per-sample sums in lane order
    for (int k0 = 0; k0 < n; k0 += 32) {
      const int k = k0 + lane;
      float s = 0.0f;
      if (k < n) {
        for (int q = 0; q < chunks; ++q) s += partials[(static_cast<size_t>(k) * chunks + q) * c + ch];
        out_n[static_cast<size_t>(k) * out_n_stride + ch] = s;
      }
      tot += warp_sum(s);
    }
  } else {
    for (int k = lane; k < n * chunks; k += 32) tot += partials[static_cast<size_t>(k) * c + ch];
    tot = warp_sum(tot);
  }
  if (out_total && lane == 0) out_total[ch] = tot;
}

// Total only, few pixels (bias gradients of the attention blocks' Linear layers, x = [tokens][c], a few MB): a block owns ONE
// 8-channel column and all the pixels, so there is no second reduction level at all (no partials, no ticket, no tail
// through L2): eight loads in flight per thread, warp butterfly, one shared-memory step over the warps.
template <int FMT>
__global__ void chansum_cols_kernel(const void* __restrict__ x, size_t plane, int pixels, int c, float* __restrict__ out_total) {
  pdl_grid_sync();
  __shared__ float red[32][8];
  const int vec = blockIdx.x;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 8
  for (int p = threadIdx.x; p < pixels; p += blockDim.x) {
    float v[8];
    Act<FMT>::load8(x, plane, static_cast<size_t>(p) * c + vec * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += v[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = warp_sum(s[j]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][j] = s[j];
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float acc = 0.0f;
    for (int w = 0; w < nwarps; ++w) acc += red[w][threadIdx.x];
    out_total[vec * 8 + threadIdx.x] = acc;
  }
}

// Total only (bias gradients of the Linear layers, x = [n * hw][c]): ONE launch.  The pixels of all samples are dealt to up to
// kChansumBlocks blocks; the last block to finish (integer ticket; summation order fixed) adds the per-block rows.
constexpr int kChansumBlocks = 296;
constexpr int kChansumTicketWords = 4096;     // == kNormTicketWords: one ticketed-scratch convention (train_engine.TrainKernels.scratch)
template <int FMT>
__global__ void chansum_total_kernel(const void* __restrict__ x, size_t plane, int pixels, int c, float* __restrict__ partials,
                                     unsigned int* __restrict__ ticket, float* __restrict__ out_total) {
  pdl_grid_sync();
  extern __shared__ float red[];  // [max(lanes, parts)][c]
  const int vecs = c >> 3, lanes = blockDim.x / vecs;
  const int vec = threadIdx.x % vecs, lane = threadIdx.x / vecs;
  if (lane < lanes) {
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 4
    for (int p = blockIdx.x * lanes + lane; p < pixels; p += gridDim.x * lanes) {
      float v[8];
      Act<FMT>::load8(x, plane, static_cast<size_t>(p) * c + vec * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[static_cast<size_t>(lane) * c + vec * 8 + j] = s[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    float acc = 0.0f;
    for (int l = 0; l < lanes; ++l) acc += red[static_cast<size_t>(l) * c + i];
    partials[static_cast<size_t>(blockIdx.x) * c + i] = acc;
  }
  if (!last_ticket(ticket, gridDim.x)) return;
  const int cw = min(c, static_cast<int>(blockDim.x)), parts = blockDim.x / cw;
  const int part = threadIdx.x / cw, ch0 = threadIdx.x % cw;
  for (int base = 0; base < c; base += cw) {
    const int ch = base + ch0;
    float acc = 0.0f;
    if (part < parts && ch < c) {
#pragma unroll 8
      for (int k = part; k < static_cast<int>(gridDim.x); k += parts) acc += __ldcg(partials + static_cast<size_t>(k) * c + ch);
    }
    __syncthreads();
    if (part < parts && ch < c) red[static_cast<size_t>(part) * cw + ch0] = acc;
    __syncthreads();
    if (part == 0 && ch < c) {
      for (int q = 1; q < parts; ++q) acc += red[static_cast<size_t>(q) * cw + ch0];
      out_total[ch] = acc;
    }
  }
}

// ---- bilinear x2 upsample backward (adjoint of upsample2x_kernel) ----------------------------------
// Input pixel i receives from output rows 2i-1, 2i, 2i+1, 2i+2 the weights
//   0.25 (i >= 1) | 0.75 (+0.25 if i == 0) | 0.75 (+0.25 if i == h-1) | 0.25 (i <= h-2)        (same along w)
template <int FMT>
__global__ void upsample2x_bwd_kernel(const void* __restrict__ dy, size_t dy_plane, void* __restrict__ dx, size_t dx_plane,
                                      int n, int h, int w, int c) {
  pdl_grid_sync();
  const uint32_t vecs = c >> 3;
  const uint32_t total = static_cast<uint32_t>(n) * h * w * vecs;     // < 2^32 (host-checked): 32-bit divisions only
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int vec = static_cast<int>(i % vecs);
    uint32_t r = i / vecs;
    const int ix = static_cast<int>(r % w);
    r /= w;
    const int iy = static_cast<int>(r % h);
    const int b = static_cast<int>(r / h);
    const float wy[4] = {iy >= 1 ? 0.25f : 0.0f, iy == 0 ? 1.0f : 0.75f, iy == h - 1 ? 1.0f : 0.75f, iy <= h - 2 ? 0.25f : 0.0f};
    const float wx[4] = {ix >= 1 ? 0.25f : 0.0f, ix == 0 ? 1.0f : 0.75f, ix == w - 1 ? 1.0f : 0.75f, ix <= w - 2 ? 0.25f : 0.0f};
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (wy[a] == 0.0f) continue;
      const int oy = 2 * iy - 1 + a;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (wx[e] == 0.0f) continue;
        const int ox = 2 * ix - 1 + e;
        float v[8];
        Act<FMT>::load8(dy, dy_plane, ((static_cast<size_t>(b) * 2 * h + oy) * 2 * w + ox) * c + vec * 8, v);
        const float wgt = wy[a] * wx[e];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(wgt, v[j], acc[j]);
      }
    }
    Act<FMT>::store8(dx, dx_plane, static_cast<size_t>(i) * 8, acc);
  }
}

// ---- time embedding + projections backward ---------------------------------------------------------
// forward (elementwise.cu: time_embed_project_kernel):  e_s[n][k] = fourier_s(t[n])[k] (+ label_emb[y[n]][k] on set 0),
// h = silu(e),  out[n][col] = b[col] + sum_k h_{set(col)}[n][k] * W[col][k].
// (1) e -> workspace;  (2) dW, db per column;  (3) d(label_emb) through set 0.
__global__ void time_embed_bwd_embed_kernel(const float* __restrict__ t, const int64_t* __restrict__ y, const float* __restrict__ fw,
                                            int n_sets, int te, const float* __restrict__ label_emb, float* __restrict__ e_ws,
                                            float* __restrict__ h_ws) {
  pdl_grid_sync();
  const int row = blockIdx.x, half = te / 2, rows = gridDim.x;
  const float tv = t[row];
  const int64_t lab = y ? y[row] : -1;
  for (int i = threadIdx.x; i < n_sets * half; i += blockDim.x) {
    const int s = i / half, j = i - s * half;
    float sv, cv;
    sincosf((tv * fw[i]) * 6.283185307179586f, &sv, &cv);
    if (s == 0 && lab >= 0 && label_emb) {
      sv += label_emb[lab * te + j];
      cv += label_emb[lab * te + half + j];
    }
    float* e = e_ws + (static_cast<size_t>(s) * rows + row) * te;
    e[j] = sv;
    e[half + j] = cv;
    float* hh = h_ws + (static_cast<size_t>(s) * rows + row) * te;     // h = silu(e): every output column re-used it 1 700 times
    hh[j] = silu(sv);
    hh[half + j] = silu(cv);
  }
}
// one block per output column; threads over k
__global__ void time_embed_bwd_weight_kernel(const float* __restrict__ dout, int c_total, const float* __restrict__ h_ws, int rows, int te,
                                             const int32_t* __restrict__ pset, float* __restrict__ dW, float* __restrict__ db) {
  pdl_grid_sync();
  const int col = blockIdx.x;
  const float* hh = h_ws + static_cast<size_t>(pset[col]) * rows * te;
  for (int k = threadIdx.x; k < te; k += blockDim.x) {
    float acc = 0.0f;
#pragma unroll 8
    for (int r = 0; r < rows; ++r) acc = fmaf(dout[static_cast<size_t>(r) * c_total + col], hh[static_cast<size_t>(r) * te + k], acc);
    dW[static_cast<size_t>(col) * te + k] = acc;
  }
  if (threadIdx.x == 0) {
    float acc = 0.0f;
    for (int r = 0; r < rows; ++r) acc += dout[static_cast<size_t>(r) * c_total + col];
    db[col] = acc;
  }
}
// de0[row][k] = silu'(e0[row][k]) * sum_{col in set 0} dout[row][col] * W[col][k]
// A 64 x 1024 x 256 product (33 MFLOP) that is pure latency: block = (row, 16 values of k) x 16 column groups, so a thread's
// chain is c_total / 16 columns with eight loads in flight; the groups' partial sums meet in shared memory in a fixed order.
__global__ void __launch_bounds__(256) time_embed_bwd_input_kernel(const float* __restrict__ dout, int c_total, const float* __restrict__ e_ws, int te,
                                                                  const int32_t* __restrict__ pset, const float* __restrict__ pw, float* __restrict__ de0) {
  pdl_grid_sync();
  __shared__ float part[16][17];
  const int row = blockIdx.x, k = blockIdx.y * 16 + (threadIdx.x & 15), grp = threadIdx.x >> 4;
  float acc = 0.0f;
  if (k < te) {
#pragma unroll 8
    for (int col = grp; col < c_total; col += 16) {
      const float g = pset[col] == 0 ? dout[static_cast<size_t>(row) * c_total + col] : 0.0f;
      acc = fmaf(g, __ldg(pw + static_cast<size_t>(col) * te + k), acc);
    }
  }
  part[grp][threadIdx.x & 15] = acc;
  __syncthreads();
  if (grp == 0 && k < te) {
    float s = 0.0f;
#pragma unroll
    for (int q = 0; q < 16; ++q) s += part[q][threadIdx.x];
    de0[static_cast<size_t>(row) * te + k] = s * act_grad(e_ws[static_cast<size_t>(row) * te + k], SBGM_ACT_SILU);
  }
}
__global__ void label_emb_bwd_kernel(const float* __restrict__ de0, const int64_t* __restrict__ y, int rows, int te,
                                     float* __restrict__ dlabel) {
  pdl_grid_sync();
  const int cls = blockIdx.x;
  for (int k = threadIdx.x; k < te; k += blockDim.x) {
    float acc = 0.0f;
    for (int r = 0; r < rows; ++r)
      if (y[r] == cls) acc += de0[static_cast<size_t>(r) * te + k];
    dlabel[static_cast<size_t>(cls) * te + k] = acc;
  }
}

// ---- DSM loss backward:  d loss / d score = grad * 2 w (score * std + z) * std / n -----------------------
__global__ void dsm_loss_bwd_kernel(const float* __restrict__ score, const float* __restrict__ std, const float* __restrict__ z,
                                    const float* __restrict__ sdf, const float* __restrict__ grad_loss, size_t nq, int per_member_q,
                                    float inv_n, float* __restrict__ dscore) {
  pdl_grid_sync();
  const float gl = grad_loss ? *grad_loss : 1.0f;
  for (size_t q = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; q < nq; q += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float sd = std[q / per_member_q];
    const float4 s = __ldg(reinterpret_cast<const float4*>(score) + q);
    const float4 nz = __ldg(reinterpret_cast<const float4*>(z) + q);
    float w[4] = {1.0f, 1.0f, 1.0f, 1.0f};
    if (sdf) {
      const float4 d = __ldg(reinterpret_cast<const float4*>(sdf) + q);
      w[0] = 0.5f / (1.0f + expf(-d.x)) + 0.5f; w[1] = 0.5f / (1.0f + expf(-d.y)) + 0.5f;
      w[2] = 0.5f / (1.0f + expf(-d.z)) + 0.5f; w[3] = 0.5f / (1.0f + expf(-d.w)) + 0.5f;
    }
    const float k = 2.0f * sd * inv_n * gl;
    reinterpret_cast<float4*>(dscore)[q] = make_float4(k * w[0] * fmaf(s.x, sd, nz.x), k * w[1] * fmaf(s.y, sd, nz.y),
                                                       k * w[2] * fmaf(s.z, sd, nz.z), k * w[3] * fmaf(s.w, sd, nz.w));
  }
}

}  // namespace sbgm

using namespace sbgm;

static NormArgs make_norm_args(const void* x, size_t x_plane, const float* stats, int per_sample, int groups, const float* gamma,
                               const float* beta, const void* add, size_t add_plane, const float* tproj, int tproj_stride,
                               int tproj_pre, int act, int hw, int c) {
  NormArgs a;
  a.x = x; a.x_plane = x_plane; a.stats = stats; a.n_stride = (per_sample == 1) ? groups : 0; a.cpg = c / groups;
  a.gamma = gamma; a.beta = beta; a.add = add; a.add_plane = add_plane; a.tproj = tproj; a.tproj_stride = tproj_stride;
  a.tproj_pre = tproj_pre; a.act = act; a.hw = hw; a.c = c;
  return a;
}

extern "C" {

int sbgm_gn_stats_finalize(const float* partials, int chunks, int pgroups, int groups, int n, int hw, int c, float eps,
                           float* stats, void* stream) {
  SBGM_REQUIRE(groups >= 1 && pgroups >= groups && pgroups % groups == 0 && c % groups == 0, "gn_stats_finalize: bad groups=%d pgroups=%d c=%d", groups, pgroups, c);
  launch_k((gn_stats_finalize_kernel), n, 256, 0, as_stream(stream), partials, chunks, pgroups, groups, static_cast<double>(hw) * (c / groups), eps, stats);
  return check_launch("gn_stats_finalize");
}

int sbgm_bn_stats_finalize(const float* partials, int chunks, int n, int hw, int c, float eps, float momentum, float* stats,
                           float* running_mean, float* running_var, void* stream) {
  launch_k((bn_stats_finalize_kernel), ceil_div(static_cast<long long>(c) * 32, 256), 256, 0, as_stream(stream), 
      partials, n, chunks, c, static_cast<double>(n) * hw, eps, momentum, stats, running_mean, running_var);
  return check_launch("bn_stats_finalize");
}

int sbgm_norm_apply(const void* x, size_t x_plane, const float* stats, int per_sample_stats, int groups, const float* gamma,
                    const float* beta, const void* add, size_t add_plane, const float* tproj, int tproj_stride,
                    int tproj_pre_act, int act, void* y, size_t y_plane, int fmt, int n, int hw, int c, void* stream) {
  SBGM_REQUIRE(c % 8 == 0 && groups >= 1 && c % groups == 0, "norm_apply: bad c=%d groups=%d", c, groups);
  const NormArgs a = make_norm_args(x, x_plane, stats, per_sample_stats, groups, gamma, beta, add, add_plane, tproj, tproj_stride,
                                    tproj_pre_act, act, hw, c);
  const int vecs = c / 8;
  int slots = 148 * 8;     // one wave of resident blocks (see resident_blocks)
  SBGM_DISPATCH_FMT(fmt, (slots = resident_blocks(reinterpret_cast<const void*>(norm_apply_kernel<FMT>), 256, 0)));
  const int per_n_blocks = max(1, min(ceil_div(static_cast<long long>(hw) * vecs, 256 * 4), slots / max(n, 1)));
  SBGM_REQUIRE(vecs <= 256, "norm_apply: c too large");
  dim3 grid(per_n_blocks, n);
  SBGM_DISPATCH_FMT(fmt, (launch_k((norm_apply_kernel<FMT>), grid, 256, 0, as_stream(stream), a, y, y_plane)));
  return check_launch("norm_apply");
}

size_t sbgm_norm_backward_sums_offset(int n, int c) { (void)n; (void)c; return kNormTicketWords; }
size_t sbgm_norm_backward_sums_floats(int n, int c) { return static_cast<size_t>(n) * c * 4; }

size_t sbgm_norm_backward_scratch_floats(int n, int c) {
  return kNormTicketWords + static_cast<size_t>(n) * c * 4 + static_cast<size_t>(n) * c * 2 + 64;
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// Block geometry of norm_bwd_sums_kernel: `vpb` channel vectors per column (a divisor of c / 8) and `pl` pixel lanes.
static void norm_bwd_geometry(int hw, int c, int per_sample_stats, int groups, int* vpb, int* pl, int* shuffle) {
  const int vecs = c / 8, cpg = c / groups;
  bool whole = !is_pow2(vecs);                       // odd channel counts: one column with every vector, no butterfly
  if (per_sample_stats == 1 && !(cpg % 8 == 0 && is_pow2(cpg / 8)) && 8 % cpg != 0) whole = true;   // groups straddling vectors
  if (whole) {
    *vpb = vecs;
    *pl = max(1, 256 / vecs);
  } else {
    int lanes = 32;                                  // enough lanes for the pixels, at most 128 (two vectors = one 32-byte sector)
    while (lanes < 128 && lanes < hw) lanes <<= 1;
    int v = max(256 / lanes, 1);
    if (per_sample_stats == 1 && cpg > 8) v = max(v, cpg / 8);       // a column holds whole groups
    v = min(v, vecs);
    *vpb = v;
    *pl = max(1, 256 / v);
  }
  *shuffle = is_pow2(*vpb) && *vpb <= 32 && is_pow2(*pl);
}

int sbgm_norm_backward(const void* dy, size_t dy_plane, const void* x, size_t x_plane, const float* stats, int per_sample_stats,
                       int groups, const float* gamma, const float* beta, const void* add, size_t add_plane,
                       const float* tproj, int tproj_stride, int tproj_pre_act, int act, void* dx, size_t dx_plane,
                       void* dadd, size_t dadd_plane, float* dgamma, float* dbeta, float* dbias_prev, float* dtproj, int dtproj_stride,
                       int fmt, int n, int hw, int c, float* scratch, int stage, const float* sums_all, int n_all, void* stream) {
  SBGM_REQUIRE(c % 8 == 0 && c <= 2048 && groups >= 1 && c % groups == 0, "norm_backward: bad c=%d groups=%d", c, groups);
  SBGM_REQUIRE(stage >= 0 && stage <= 2, "norm_backward: stage %d", stage);
  SBGM_REQUIRE(stage == 0 || per_sample_stats == 0, "norm_backward: stages 1 / 2 (gathered sums) are for batch statistics only");
  SBGM_REQUIRE(sums_all == nullptr || (stage == 2 && n_all >= n), "norm_backward: gathered sums belong to stage 2");
  const int vecs = c / 8;
  SBGM_REQUIRE(vecs <= 256, "norm_backward: c too large");
  const NormArgs a = make_norm_args(x, x_plane, stats, per_sample_stats, groups, gamma, beta, add, add_plane, tproj, tproj_stride,
                                    tproj_pre_act, act, hw, c);
  float* sums = scratch + kNormTicketWords;
  float* coef = sums + static_cast<size_t>(n) * c * 4;
  cudaStream_t st = as_stream(stream);
  if (sums_all == nullptr) { sums_all = sums; n_all = n; }
  const double cnt = per_sample_stats == 1 ? static_cast<double>(hw) * (c / groups) : static_cast<double>(n_all) * hw;
  NormBwdTail t;
  t.tickets = reinterpret_cast<unsigned int*>(scratch);
  t.sums = sums; t.sums_all = sums_all; t.coef = coef; t.dgamma = dgamma; t.dbeta = dbeta; t.dbias_prev = dbias_prev;
  t.dtproj = dtproj; t.dtproj_stride = dtproj_stride; t.n = n; t.n_all = n_all; t.fixed_stats = per_sample_stats == 2;
  t.with_coef = stage == 0; t.inv_cnt = static_cast<float>(1.0 / cnt);
  norm_bwd_geometry(hw, c, per_sample_stats, groups, &t.vpb, &t.pl, &t.shuffle);
  const int columns = vecs / t.vpb;
  SBGM_REQUIRE(columns <= kNormTicketWords, "norm_backward: too many channel columns");
  const int threads = max(32, min(256, ((t.pl * t.vpb + 31) / 32) * 32));
  const int rows = t.shuffle ? threads / 32 : t.pl;
  const size_t smem1 = (static_cast<size_t>(rows) * t.vpb * 32 + 4 * 256) * sizeof(float);
  int slots = 148 * 8;
  SBGM_DISPATCH_FMT(fmt, (slots = resident_blocks(reinterpret_cast<const void*>(norm_bwd_apply_kernel<FMT>), 256, 0)));
  const int per_n_blocks = max(1, min(ceil_div(static_cast<long long>(hw) * vecs, 256 * 4), slots / max(n, 1)));
  dim3 g1(columns, n), g3(per_n_blocks, n);
  SBGM_DISPATCH_FMT(fmt, {
    if (stage != 2) {
      auto k1 = norm_bwd_sums_kernel<FMT>;
      if (smem1 > 48 * 1024 && cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem1)) != cudaSuccess) {
        set_error("norm_backward: cannot reserve %zu bytes of shared memory", smem1);
        return 1;
      }
      launch_k((k1), g1, threads, smem1, st, a, dy, dy_plane, t);
    }
    if (stage == 2) launch_k((norm_bwd_tail_kernel), columns, 256, static_cast<size_t>(4 * 256) * sizeof(float), st, a, t);
    if (stage != 1) launch_k((norm_bwd_apply_kernel<FMT>), g3, 256, 0, st, a, dy, dy_plane, coef, dx, dx_plane, dadd, dadd_plane);
  });
  return check_launch("norm_backward");
}

size_t sbgm_layernorm_backward_scratch_floats(int c) { return static_cast<size_t>(kLnBwdBlocks) * c * 2; }

int sbgm_layernorm_backward(const void* dy, size_t dy_plane, const void* x, size_t x_plane, const float* gamma, float eps,
                            void* dx, size_t dx_plane, float* dgamma, float* dbeta, int fmt, int rows, int c, float* scratch,
                            void* stream) {
  SBGM_REQUIRE(c % 8 == 0 && c <= 512, "layernorm_backward: c=%d must be a multiple of 8 and <= 512", c);
  cudaStream_t st = as_stream(stream);
  const int blocks = min(kLnBwdBlocks, ceil_div(rows, 8));
  const size_t smem = static_cast<size_t>(8) * c * 2 * sizeof(float);
  SBGM_DISPATCH_FMT(fmt, (launch_k((layernorm_bwd_kernel<FMT>), blocks, 256, smem, st, dy, dy_plane, x, x_plane, gamma, eps, dx, dx_plane,
                                                                                rows, c, scratch)));
  launch_k((layernorm_bwd_finish_kernel), ceil_div(c, 8), 256, 0, st, scratch, blocks, c, dgamma, dbeta);
  return check_launch("layernorm_backward");
}

int sbgm_act_forward(const void* x, size_t x_plane, void* y, size_t y_plane, int fmt, size_t count, int act, void* stream) {
  SBGM_REQUIRE(count % 8 == 0, "act_forward: count must be a multiple of 8");
  SBGM_DISPATCH_FMT(fmt, (launch_k((act_fwd_kernel<FMT>), bgrid_for(count / 8, 256), 256, 0, as_stream(stream), x, x_plane, y, y_plane, count / 8, act)));
  return check_launch("act_forward");
}
int sbgm_act_backward(const void* dy, size_t dy_plane, const void* x, size_t x_plane, void* dx, size_t dx_plane, int fmt,
                      size_t count, int act, void* stream) {
  SBGM_REQUIRE(count % 8 == 0, "act_backward: count must be a multiple of 8");
  SBGM_DISPATCH_FMT(fmt, (launch_k((act_bwd_kernel<FMT>), bgrid_for(count / 8, 256), 256, 0, as_stream(stream), dy, dy_plane, x, x_plane, dx, dx_plane, count / 8, act)));
  return check_launch("act_backward");
}
int sbgm_add_inplace(void* dst, size_t dst_plane, const void* src, size_t src_plane, int fmt, size_t count, void* stream) {
  SBGM_REQUIRE(count % 8 == 0, "add_inplace: count must be a multiple of 8");
  SBGM_DISPATCH_FMT(fmt, (launch_k((add_inplace_kernel<FMT>), bgrid_for(count / 8, 256), 256, 0, as_stream(stream), dst, dst_plane, src, src_plane, count / 8)));
  return check_launch("add_inplace");
}

size_t sbgm_channel_sums_scratch_floats(int n, int c) {
  const size_t per_sample = static_cast<size_t>(n) * kNormChunks * c, total = static_cast<size_t>(kChansumBlocks) * c;
  return kChansumTicketWords + (per_sample > total ? per_sample : total);
}

int sbgm_channel_sums(const void* x, size_t x_plane, int fmt, int n, int hw, int c, float* out_per_sample, int out_stride,
                      float* out_total, float* scratch, void* stream) {
  SBGM_REQUIRE(c % 8 == 0 && c / 8 <= 256, "channel_sums: bad c=%d", c);
  const int vecs = c / 8, lanes = 256 / vecs;
  cudaStream_t st = as_stream(stream);
  float* partials = scratch + kChansumTicketWords;
  if (out_per_sample == nullptr) {
    SBGM_REQUIRE(out_total != nullptr, "channel_sums: no output");
    const long long pixels = static_cast<long long>(n) * hw;
    SBGM_REQUIRE(pixels < (1ll << 31), "channel_sums: tensor too large");
    if (pixels <= 32768) {        // one block per channel vector, <= 32 pixels per thread
      const int threads = pixels >= 8192 ? 1024 : (pixels >= 2048 ? 512 : 256);
      SBGM_DISPATCH_FMT(fmt, (launch_k((chansum_cols_kernel<FMT>), vecs, threads, 0, st, x, x_plane, static_cast<int>(pixels), c, out_total)));
      return check_launch("channel_sums");
    }
    // the main pass shrinks with the block count, the last block's tail (c x blocks partials through one SM) grows with it:
    // both are latency-bound and balance near blocks = sqrt(pixels)
    const int blocks = max(1, min(min(kChansumBlocks, ceil_div(pixels, static_cast<long long>(lanes) * 8)),
                                  static_cast<int>(sqrt(static_cast<double>(pixels)))));
    const size_t smem = static_cast<size_t>(max(lanes, 256 / min(c, 256))) * c * sizeof(float);
    SBGM_DISPATCH_FMT(fmt, (launch_k((chansum_total_kernel<FMT>), blocks, 256, smem, st, x, x_plane, static_cast<int>(pixels), c, partials,
                                                                            reinterpret_cast<unsigned int*>(scratch), out_total)));
    return check_launch("channel_sums");
  }
  const int chunks = max(1, min(kNormChunks, hw / (lanes * 8)));
  dim3 g1(chunks, n);
  SBGM_DISPATCH_FMT(fmt, (launch_k((chansum_partial_kernel<FMT>), g1, 256, static_cast<size_t>(lanes) * c * sizeof(float), st, x, x_plane, hw, c, partials, chunks)));
  launch_k((chansum_finish_kernel), ceil_div(c, 8), 256, 0, st, partials, n, c, chunks, out_per_sample, out_stride, out_total);
  return check_launch("channel_sums");
}

int sbgm_upsample2x_backward(const void* dy, size_t dy_plane, void* dx, size_t dx_plane, int fmt, int n, int h, int w, int c,
                             void* stream) {
  SBGM_REQUIRE(c % 8 == 0, "upsample2x_backward: c=%d must be a multiple of 8", c);
  const size_t total = static_cast<size_t>(n) * h * w * (c / 8);
  SBGM_REQUIRE(total < (1ull << 32), "upsample2x_backward: tensor too large for 32-bit indexing");
  SBGM_DISPATCH_FMT(fmt, (launch_k((upsample2x_bwd_kernel<FMT>), bgrid_for(total, 256), 256, 0, as_stream(stream), dy, dy_plane, dx, dx_plane, n, h, w, c)));
  return check_launch("upsample2x_backward");
}

size_t sbgm_time_embed_backward_scratch_floats(int n_sets, int te, int rows) {
  return static_cast<size_t>(2) * n_sets * rows * te + static_cast<size_t>(rows) * te;
}

int sbgm_time_embed_backward(const float* dout, const float* t, const int64_t* y, const float* fourier_w, int n_sets, int te,
                             const float* label_emb, int n_classes, const float* proj_w, const int32_t* proj_set, int c_total,
                             int rows, float* d_proj_w, float* d_proj_b, float* d_label_emb, float* scratch, void* stream) {
  SBGM_REQUIRE(te % 2 == 0 && n_sets >= 1 && rows >= 1, "time_embed_backward: bad sizes");
  cudaStream_t st = as_stream(stream);
  float* e_ws = scratch;
  float* h_ws = scratch + static_cast<size_t>(n_sets) * rows * te;
  float* de0 = h_ws + static_cast<size_t>(n_sets) * rows * te;
  launch_k((time_embed_bwd_embed_kernel), rows, 256, 0, st, t, y, fourier_w, n_sets, te, label_emb, e_ws, h_ws);
  launch_k((time_embed_bwd_weight_kernel), c_total, 256, 0, st, dout, c_total, h_ws, rows, te, proj_set, d_proj_w, d_proj_b);
  if (d_label_emb != nullptr && y != nullptr) {
    launch_k((time_embed_bwd_input_kernel), dim3(rows, ceil_div(te, 16)), 256, 0, st, dout, c_total, e_ws, te, proj_set, proj_w, de0);
    launch_k((label_emb_bwd_kernel), n_classes, 256, 0, st, de0, y, rows, te, d_label_emb);
  }
  return check_launch("time_embed_backward");
}

int sbgm_dsm_loss_backward(const float* score, const float* std, const float* z, const float* sdf, const float* grad_loss, int n,
                           int per_member, float* dscore, void* stream) {
  SBGM_REQUIRE(per_member % 4 == 0, "dsm_loss_backward: per_member must be a multiple of 4");
  const size_t nq = static_cast<size_t>(n) * per_member / 4;
  launch_k((dsm_loss_bwd_kernel), bgrid_for(nq, 256), 256, 0, as_stream(stream), score, std, z, sdf, grad_loss, nq, per_member / 4,
                                                                         1.0f / static_cast<float>(n), dscore);
  return check_launch("dsm_loss_backward");
}

}  // extern "C"
