// post_sampler.cu -- (i) the stage arithmetic of the device-resident Dormand-Prince integrator behind ode_sampler
// (sbgm/score_sampling.py:239-300 calls scipy.integrate.solve_ivp(RK45) on the host in float64) and (ii) the
// extreme-value sentinel of the generation path (sbgm/utils.py:1642-1671 report_precip_extremes, driven by
// sbgm/training.py:700-755) fused with the back-transform to physical units.
//
// (i)  State, stages and error estimate stay on the device in float64; a stage combination  y + h * sum_s a_s K_s  is one
//      pass that also emits the float32 copy the score network reads; the scaled error norm is a deterministic two-stage
//      reduction whose single double is the only thing the host controller reads per attempted step.
// (ii) One block per sample: back-transform every value (affine -> clamp -> exp, the arithmetic of back_transform_kernel),
//      store it, and find the two order statistics around the 0.999 quantile with an exact 4-pass radix select over
//      order-preserving 32-bit keys held in shared-memory histograms (torch.quantile sorts the whole sample: here no
//      value leaves the SM except the three results), plus the sample maximum.
#include "common.cuh"

namespace sbgm {

struct Rk45Coef {
  double c[7];
};

// out = y + h * sum_{s < ns} c[s] * K[s][.]   (float64), out32 = float(out) (either may be null)
__global__ void rk45_combine_kernel(const double* __restrict__ y, const double* __restrict__ K, size_t n, int ns, Rk45Coef c, double h,
                                    double* __restrict__ out, float* __restrict__ out32) {
  pdl_grid_sync();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    double acc = 0.0;
    for (int s = 0; s < ns; ++s) acc += K[static_cast<size_t>(s) * n + i] * c.c[s];     // torch.mv order: ascending s
    const double v = y[i] + acc * h;
    if (out) out[i] = v;
    if (out32) out32[i] = static_cast<float>(v);
  }
}

// K_s = scale * score  (the probability-flow right-hand side  -1/2 g(t)^2 score, score_sampling.py:287-291)
__global__ void rk45_rhs_kernel(const float* __restrict__ score, double scale, double* __restrict__ k_out, size_t n) {
  pdl_grid_sync();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    k_out[i] = scale * static_cast<double>(score[i]);
}

// partial[b] = sum over the block's elements of ( h * sum_s e[s] K[s][i] / (atol + max(|y|, |y_new|) * rtol) )^2
// mode 1: ( v[i] / (atol + |y[i]| * rtol) )^2 with v = K (the norms of select_initial_step)
__global__ void rk45_errnorm_partial_kernel(const double* __restrict__ K, size_t n, int ns, Rk45Coef e, double h, const double* __restrict__ y,
                                            const double* __restrict__ y_new, double atol, double rtol, double* __restrict__ partial) {
  pdl_grid_sync();
  __shared__ double red[256];
  double acc = 0.0;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    double v = 0.0;
    for (int s = 0; s < ns; ++s) v += K[static_cast<size_t>(s) * n + i] * e.c[s];
    v *= h;
    const double a = fabs(y[i]), b = y_new ? fabs(y_new[i]) : a;
    const double r = v / (atol + fmax(a, b) * rtol);
    acc += r * r;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}
__global__ void rk45_errnorm_finish_kernel(const double* __restrict__ partial, int blocks, double* __restrict__ out) {
  pdl_grid_sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int b = 0; b < blocks; ++b) s += partial[b];
    *out = s;
  }
}

// ---- extremes -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t order_key(float v) {          // monotone float -> uint (NaN sorts last, as torch.sort does)
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_value(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// k-th smallest key (0-based) of the block's `per` values in `vals` (global, just written by this block): MSB-first radix select.
__device__ uint32_t radix_select(const float* __restrict__ vals, int per, uint32_t k, uint32_t* hist /* smem[256] */) {
  uint32_t prefix = 0, mask = 0;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < per; i += blockDim.x) {
      const uint32_t key = order_key(vals[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    // every thread walks the 256 bins (cheap, and keeps `k` / `prefix` uniform without another broadcast)
    uint32_t acc = 0, bin = 0;
    for (uint32_t b = 0; b < 256; ++b) {
      const uint32_t c = hist[b];
      if (acc + c > k) { bin = b; break; }
      acc += c;
    }
    k -= acc;
    prefix |= bin << shift;
    mask |= 255u << shift;
    __syncthreads();
  }
  return prefix;
}

// y = back_transform(x) stored; out[n][4] = {quantile q of the sample (torch.quantile 'linear'), max, min, NaN count}
__global__ void back_transform_extremes_kernel(const float* __restrict__ x, float* __restrict__ y, int per, float pre, float a, float b,
                                               float lo, float hi, int do_clamp, int do_exp, float q, float* __restrict__ out) {
  pdl_grid_sync();
  __shared__ uint32_t hist[256];
  __shared__ float red_mx[32], red_mn[32];
  const float* xs = x + static_cast<size_t>(blockIdx.x) * per;
  float* ys = y + static_cast<size_t>(blockIdx.x) * per;
  float mx = -INFINITY, mn = INFINITY;
  for (int i = threadIdx.x; i < per; i += blockDim.x) {
    float v = __fadd_rn(__fmul_rn(__fadd_rn(xs[i], pre), a), b);
    if (do_clamp) v = fminf(fmaxf(v, lo), hi);
    if (do_exp) v = expf(v);
    ys[i] = v;
    mx = fmaxf(mx, v);
    mn = fminf(mn, v);
  }
  mx = warp_max(mx);
  mn = -warp_max(-mn);
  if ((threadIdx.x & 31) == 0) { red_mx[threadIdx.x >> 5] = mx; red_mn[threadIdx.x >> 5] = mn; }
  __syncthreads();       // also makes this block's writes to ys visible to its own threads below
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    float m1 = threadIdx.x < nw ? red_mx[threadIdx.x] : -INFINITY, m2 = threadIdx.x < nw ? red_mn[threadIdx.x] : INFINITY;
    m1 = warp_max(m1);
    m2 = -warp_max(-m2);
    if (threadIdx.x == 0) { red_mx[0] = m1; red_mn[0] = m2; }
  }
  __syncthreads();
  // torch.quantile(interpolation='linear'): rank = q * (n - 1); lerp(sorted[floor], sorted[ceil], rank - floor)
  const float rank = q * static_cast<float>(per - 1);
  const uint32_t k_lo = static_cast<uint32_t>(floorf(rank));
  const uint32_t k_hi = min(k_lo + 1u, static_cast<uint32_t>(per - 1));
  const float v_lo = key_value(radix_select(ys, per, k_lo, hist));
  const float v_hi = key_value(radix_select(ys, per, k_hi, hist));
  if (threadIdx.x == 0) {
    const float w = rank - static_cast<float>(k_lo);
    const float d = v_hi - v_lo;
    const float qv = (w < 0.5f) ? v_lo + w * d : v_hi - d * (1.0f - w);    // at::lerp
    float* o = out + static_cast<size_t>(blockIdx.x) * 4;
    o[0] = qv; o[1] = red_mx[0]; o[2] = red_mn[0]; o[3] = 0.0f;
  }
}

}  // namespace sbgm

using namespace sbgm;

static int rk_grid(size_t n) {
  size_t g = (n + 255) / 256;
  if (g < 1) g = 1;
  if (g > 148 * 8) g = 148 * 8;
  return static_cast<int>(g);
}

extern "C" {

int sbgm_rk45_combine(const double* y, const double* k_stages, size_t n, int n_stages, const double* coef_host, double h,
                      double* out, float* out_f32, void* stream) {
  SBGM_REQUIRE(n_stages >= 0 && n_stages <= 7, "rk45_combine: n_stages=%d out of range", n_stages);
  SBGM_REQUIRE(out != nullptr || out_f32 != nullptr, "rk45_combine: no output");
  Rk45Coef c = {};
  for (int s = 0; s < n_stages; ++s) c.c[s] = coef_host[s];
  launch_k((rk45_combine_kernel), rk_grid(n), 256, 0, as_stream(stream), y, k_stages, n, n_stages, c, h, out, out_f32);
  return check_launch("rk45_combine");
}

int sbgm_rk45_rhs(const float* score, double scale, double* k_out, size_t n, void* stream) {
  launch_k((rk45_rhs_kernel), rk_grid(n), 256, 0, as_stream(stream), score, scale, k_out, n);
  return check_launch("rk45_rhs");
}

size_t sbgm_rk45_scratch_doubles(size_t n) { return static_cast<size_t>(rk_grid(n)); }

int sbgm_rk45_error_norm(const double* k_stages, size_t n, int n_stages, const double* coef_host, double h, const double* y,
                         const double* y_new, double atol, double rtol, double* scratch, double* out_sumsq, void* stream) {
  SBGM_REQUIRE(n_stages >= 1 && n_stages <= 7, "rk45_error_norm: n_stages=%d out of range", n_stages);
  Rk45Coef c = {};
  for (int s = 0; s < n_stages; ++s) c.c[s] = coef_host[s];
  const int g = rk_grid(n);
  cudaStream_t st = as_stream(stream);
  launch_k((rk45_errnorm_partial_kernel), g, 256, 0, st, k_stages, n, n_stages, c, h, y, y_new, atol, rtol, scratch);
  launch_k((rk45_errnorm_finish_kernel), 1, 32, 0, st, static_cast<const double*>(scratch), g, out_sumsq);
  return check_launch("rk45_error_norm");
}

int sbgm_back_transform_extremes(const float* x, float* y, int n, int per, float pre_shift, float scale, float shift, float lo,
                                 float hi, int do_clamp, int do_exp, float quantile, float* out, void* stream) {
  SBGM_REQUIRE(n >= 1 && per >= 1, "back_transform_extremes: empty input");
  SBGM_REQUIRE(quantile >= 0.0f && quantile <= 1.0f, "back_transform_extremes: quantile %f outside [0, 1]", quantile);
  launch_k((back_transform_extremes_kernel), n, 512, 0, as_stream(stream), x, y, per, pre_shift, scale, shift, lo, hi, do_clamp, do_exp,
           quantile, out);
  return check_launch("back_transform_extremes");
}

}  // extern "C"
