// sampler.cu -- fused VE-SDE sampler updates with in-kernel counter-based noise (Philox4x32-10),
// and the DSM perturbation / loss reduction.  One float4 (4 pixels) per thread per iteration:
// 12 B/pixel of HBM traffic for a predictor step (read x, read score, write x) + 4 B/pixel for
// the noise-free mean; the noise never touches memory.
//
// Reference: sbgm/score_sampling.py:93-127 (Euler-Maruyama), :167-230 (predictor-corrector),
//            sbgm/score_unet.py:936-985 (loss_fn).
#include "common.cuh"

namespace sbgm {

constexpr int kBlock = 256;

static int grid_for(size_t items) {
  size_t g = (items + kBlock - 1) / kBlock;
  if (g < 1) g = 1;
  if (g > 148 * 16) g = 148 * 16;
  return static_cast<int>(g);
}

template <bool NORMAL>
__global__ void philox_fill_kernel(float* __restrict__ out, size_t count, uint64_t seed, uint32_t draw, uint64_t first_q) {
  pdl_grid_sync();
  const size_t nq = (count + 3) / 4;
  for (size_t q = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; q < nq;
       q += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 z = NORMAL ? Philox::normal4(first_q + q, draw, seed) : Philox::uniform4(first_q + q, draw, seed);
    const float v[4] = {z.x, z.y, z.z, z.w};
    if (4 * q + 3 < count) {
      *reinterpret_cast<float4*>(out + 4 * q) = z;
    } else {
      for (int i = 0; i < 4 && 4 * q + i < count; ++i) out[4 * q + i] = v[i];
    }
  }
}

// sampler state (device, int32[4]): [0] step, [1] block-arrival scratch, [2..3] Philox seed (lo, hi)
__device__ __forceinline__ uint64_t state_seed(const int32_t* state) {
  return static_cast<uint64_t>(static_cast<uint32_t>(state[2])) | (static_cast<uint64_t>(static_cast<uint32_t>(state[3])) << 32);
}

__global__ void sampler_init_kernel(float* __restrict__ x, size_t nq, float std1, uint64_t seed, uint64_t first_q) {
  pdl_grid_sync();
  for (size_t q = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; q < nq;
       q += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float4 z = Philox::normal4(first_q + q, 0u, seed);
    z.x *= std1; z.y *= std1; z.z *= std1; z.w *= std1;
    reinterpret_cast<float4*>(x)[q] = z;
  }
}

// mean = x + (g^2 dt) score ; x = mean + noise_scale z ; last block to finish bumps the step counter.
__global__ void predictor_kernel(float* __restrict__ x, const float* __restrict__ score, float* __restrict__ mean_out,
                                 size_t nq, const float* __restrict__ table, int32_t* __restrict__ step_counter,
                                 uint32_t draw_base, uint32_t draw_stride, uint64_t first_q) {
  pdl_grid_sync();
  const int step = *step_counter;
  const uint64_t seed = state_seed(step_counter);
  const float* row = table + static_cast<size_t>(step) * SBGM_STEP_COLS;
  const float drift = row[4], nscale = row[5];
  const uint32_t draw = draw_base + draw_stride * static_cast<uint32_t>(step);
  for (size_t q = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; q < nq;
       q += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 xv = reinterpret_cast<const float4*>(x)[q];
    const float4 sv = __ldg(reinterpret_cast<const float4*>(score) + q);
    const float4 z = Philox::normal4(first_q + q, draw, seed);
    // same association as the reference: x + ((g*g) * s) * dt is folded to x + s * (g*g*dt)
    float4 m;
    m.x = fmaf(sv.x, drift, xv.x); m.y = fmaf(sv.y, drift, xv.y);
    m.z = fmaf(sv.z, drift, xv.z); m.w = fmaf(sv.w, drift, xv.w);
    reinterpret_cast<float4*>(mean_out)[q] = m;
    float4 o;
    o.x = fmaf(nscale, z.x, m.x); o.y = fmaf(nscale, z.y, m.y);
    o.z = fmaf(nscale, z.z, m.z); o.w = fmaf(nscale, z.w, m.w);
    reinterpret_cast<float4*>(x)[q] = o;
  }
  // every block has read *step_counter before any block can be the last one to arrive here
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int* done_blocks = reinterpret_cast<unsigned int*>(step_counter + 1);
    const unsigned int prev = atomicAdd(done_blocks, 1u);
    if (prev == gridDim.x - 1) {
      *done_blocks = 0u;
      *step_counter = step + 1;
    }
  }
}

// per-member sum of squares: one 1024-thread block per member, fixed reduction order (deterministic)
__global__ void sumsq_kernel(const float* __restrict__ score, float* __restrict__ sumsq, int per_member) {
  pdl_grid_sync();
  __shared__ float red[32];
  const int member = blockIdx.x;
  const size_t nq = per_member / 4;
  const float4* p = reinterpret_cast<const float4*>(score + static_cast<size_t>(member) * per_member);
  float acc = 0.0f;
  for (size_t q = threadIdx.x; q < nq; q += blockDim.x) {
    const float4 v = __ldg(p + q);
    acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0f;
    s = warp_sum(s);
    if (threadIdx.x == 0) sumsq[member] = s;
  }
}

__global__ void corrector_kernel(float* __restrict__ x, const float* __restrict__ score, const float* __restrict__ sumsq,
                                 int members_total, float noise_norm, float snr, size_t nq,
                                 const int32_t* __restrict__ step_counter, uint32_t draw_base,
                                 uint32_t draw_stride, uint64_t first_q) {
  pdl_grid_sync();
  const uint64_t seed = state_seed(step_counter);
  __shared__ float s_eps;
  if (threadIdx.x == 0) {
    float gn = 0.0f;
    for (int m = 0; m < members_total; ++m) gn += sqrtf(sumsq[m]);
    gn /= members_total;
    const float r = snr * noise_norm / gn;
    s_eps = 2.0f * (r * r);
  }
  __syncthreads();
  const float eps = s_eps, nscale = sqrtf(2.0f * eps);
  const uint32_t draw = draw_base + draw_stride * static_cast<uint32_t>(*step_counter);
  for (size_t q = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; q < nq;
       q += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 xv = reinterpret_cast<const float4*>(x)[q];
    const float4 sv = __ldg(reinterpret_cast<const float4*>(score) + q);
    const float4 z = Philox::normal4(first_q + q, draw, seed);
    float4 o;
    o.x = fmaf(nscale, z.x, fmaf(eps, sv.x, xv.x)); o.y = fmaf(nscale, z.y, fmaf(eps, sv.y, xv.y));
    o.z = fmaf(nscale, z.z, fmaf(eps, sv.z, xv.z)); o.w = fmaf(nscale, z.w, fmaf(eps, sv.w, xv.w));
    reinterpret_cast<float4*>(x)[q] = o;
  }
}

__global__ void dsm_perturb_kernel(const float* __restrict__ x, const float* __restrict__ std, float* __restrict__ xt,
                                   float* __restrict__ zout, size_t nq, int per_member_q, uint64_t seed, uint32_t draw,
                                   uint64_t first_q) {
  pdl_grid_sync();
  for (size_t q = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; q < nq;
       q += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float sd = std[q / per_member_q];
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + q);
    const float4 z = Philox::normal4(first_q + q, draw, seed);
    reinterpret_cast<float4*>(zout)[q] = z;
    reinterpret_cast<float4*>(xt)[q] = make_float4(fmaf(sd, z.x, xv.x), fmaf(sd, z.y, xv.y), fmaf(sd, z.z, xv.z), fmaf(sd, z.w, xv.w));
  }
}

constexpr int kLossBlocks = 592;  // 148 SMs x 4
__global__ void dsm_loss_partial_kernel(const float* __restrict__ score, const float* __restrict__ std,
                                        const float* __restrict__ z, const float* __restrict__ sdf, size_t nq,
                                        int per_member_q, float* __restrict__ partials) {
  pdl_grid_sync();
  __shared__ float red[kBlock / 32];
  float acc = 0.0f;
  for (size_t q = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; q < nq;
       q += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float sd = std[q / per_member_q];
    const float4 s = __ldg(reinterpret_cast<const float4*>(score) + q);
    const float4 n = __ldg(reinterpret_cast<const float4*>(z) + q);
    float w[4] = {1.0f, 1.0f, 1.0f, 1.0f};
    if (sdf) {
      const float4 d = __ldg(reinterpret_cast<const float4*>(sdf) + q);
      w[0] = 0.5f / (1.0f + expf(-d.x)) + 0.5f; w[1] = 0.5f / (1.0f + expf(-d.y)) + 0.5f;
      w[2] = 0.5f / (1.0f + expf(-d.z)) + 0.5f; w[3] = 0.5f / (1.0f + expf(-d.w)) + 0.5f;
    }
    const float e0 = fmaf(s.x, sd, n.x), e1 = fmaf(s.y, sd, n.y), e2 = fmaf(s.z, sd, n.z), e3 = fmaf(s.w, sd, n.w);
    acc += w[0] * e0 * e0 + w[1] * e1 * e1 + w[2] * e2 * e2 + w[3] * e3 * e3;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.0f;
    for (int i = 0; i < kBlock / 32; ++i) s += red[i];
    partials[blockIdx.x] = s;
  }
}
__global__ void dsm_loss_finish_kernel(const float* __restrict__ partials, int nblocks, float inv_n, float* __restrict__ out) {
  pdl_grid_sync();
  __shared__ double red[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) acc += partials[i];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) s += red[i];
    *out = static_cast<float>(s * inv_n);
  }
}

// ---- ensemble statistics: per-pixel mean, std (Bessel), CRPS = E|X - y| - 1/2 E|X - X'| over the M members ------------
// One thread per pixel; member m of pixel p is members[m * pixels + p], so every load is coalesced across the warp.
// The pair term is the O(M^2) double loop (M = 64: 2 016 pairs) -- the members of a pixel stay in L1.
__global__ void ensemble_stats_kernel(const float* __restrict__ members, const float* __restrict__ truth, int m, size_t pixels,
                                      float* __restrict__ mean, float* __restrict__ stdev, float* __restrict__ crps) {
  pdl_grid_sync();
  for (size_t p = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; p < pixels; p += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float s = 0.0f;
    for (int i = 0; i < m; ++i) s += members[static_cast<size_t>(i) * pixels + p];
    const float mu = s / m;
    float q = 0.0f, a = 0.0f, pair = 0.0f;
    const float y = truth ? truth[p] : 0.0f;
    for (int i = 0; i < m; ++i) {
      const float xi = members[static_cast<size_t>(i) * pixels + p];
      q = fmaf(xi - mu, xi - mu, q);
      if (crps) {
        a += fabsf(xi - y);
        for (int j = i + 1; j < m; ++j) pair += fabsf(xi - members[static_cast<size_t>(j) * pixels + p]);
      }
    }
    mean[p] = mu;
    stdev[p] = m > 1 ? sqrtf(q / (m - 1)) : 0.0f;
    if (crps) crps[p] = a / m - pair / (static_cast<float>(m) * m);
  }
}

// ---- back-transforms of sampled fields: y = a * (x + pre) + b, optionally clamp to [lo, hi], optionally exp --------
// (ZScoreBackTransform / ScaleBackTransform / PrcpLogBackTransform of sbgm/special_transforms.py:103-138, 187-237, 360-462)
__global__ void back_transform_kernel(const float* __restrict__ x, float* __restrict__ y, size_t count, float pre, float a, float b, float lo,
                                      float hi, int do_clamp, int do_exp) {
  pdl_grid_sync();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < count; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float v = __fadd_rn(__fmul_rn(__fadd_rn(x[i], pre), a), b);     // unfused: the reference (torch) rounds the product before the sum
    if (do_clamp) v = fminf(fmaxf(v, lo), hi);
    y[i] = do_exp ? expf(v) : v;
  }
}

}  // namespace sbgm

using namespace sbgm;

extern "C" {

int sbgm_philox_normal(float* out, size_t count, uint64_t seed, uint32_t draw, uint64_t first_elem, void* stream) {
  SBGM_REQUIRE(first_elem % 4 == 0, "philox_normal: first_elem must be a multiple of 4");
  launch_k((philox_fill_kernel<true>), grid_for((count + 3) / 4), kBlock, 0, as_stream(stream), out, count, seed, draw, first_elem / 4);
  return check_launch("philox_normal");
}
int sbgm_philox_uniform(float* out, size_t count, uint64_t seed, uint32_t draw, uint64_t first_elem, void* stream) {
  SBGM_REQUIRE(first_elem % 4 == 0, "philox_uniform: first_elem must be a multiple of 4");
  launch_k((philox_fill_kernel<false>), grid_for((count + 3) / 4), kBlock, 0, as_stream(stream), out, count, seed, draw, first_elem / 4);
  return check_launch("philox_uniform");
}

int sbgm_sampler_init(float* x, size_t count, float std1, uint64_t seed, uint64_t first_elem, void* stream) {
  SBGM_REQUIRE(count % 4 == 0 && first_elem % 4 == 0, "sampler_init: count and first_elem must be multiples of 4");
  launch_k((sampler_init_kernel), grid_for(count / 4), kBlock, 0, as_stream(stream), x, count / 4, std1, seed, first_elem / 4);
  return check_launch("sampler_init");
}

int sbgm_sampler_predictor(float* x, const float* score, float* mean_out, size_t count, const float* step_table,
                           int32_t* step_counter, uint32_t draw_base, uint32_t draw_stride,
                           uint64_t first_elem, void* stream) {
  SBGM_REQUIRE(count % 4 == 0 && first_elem % 4 == 0, "sampler_predictor: count and first_elem must be multiples of 4");
  launch_k((predictor_kernel), grid_for(count / 4), kBlock, 0, as_stream(stream), x, score, mean_out, count / 4, step_table,
                                                                          step_counter, draw_base, draw_stride,
                                                                          first_elem / 4);
  return check_launch("sampler_predictor");
}

int sbgm_sampler_sumsq(const float* score, float* sumsq, int members, int per_member, void* stream) {
  SBGM_REQUIRE(per_member % 4 == 0, "sampler_sumsq: per_member must be a multiple of 4");
  launch_k((sumsq_kernel), members, 1024, 0, as_stream(stream), score, sumsq, per_member);
  return check_launch("sampler_sumsq");
}

int sbgm_sampler_corrector(float* x, const float* score, const float* sumsq, int members_total, int per_member,
                           float snr, size_t count, const int32_t* step_counter,
                           uint32_t draw_base, uint32_t draw_stride, uint64_t first_elem, void* stream) {
  SBGM_REQUIRE(count % 4 == 0 && first_elem % 4 == 0, "sampler_corrector: count and first_elem must be multiples of 4");
  launch_k((corrector_kernel), grid_for(count / 4), kBlock, 0, as_stream(stream), 
      x, score, sumsq, members_total, sqrtf(static_cast<float>(per_member)), snr, count / 4, step_counter,
      draw_base, draw_stride, first_elem / 4);
  return check_launch("sampler_corrector");
}

int sbgm_dsm_perturb(const float* x, const float* std, float* xt, float* z, int n, int per_member, uint64_t seed,
                     uint32_t draw, uint64_t first_elem, void* stream) {
  SBGM_REQUIRE(per_member % 4 == 0 && first_elem % 4 == 0, "dsm_perturb: per_member and first_elem must be multiples of 4");
  const size_t nq = static_cast<size_t>(n) * per_member / 4;
  launch_k((dsm_perturb_kernel), grid_for(nq), kBlock, 0, as_stream(stream), x, std, xt, z, nq, per_member / 4, seed, draw, first_elem / 4);
  return check_launch("dsm_perturb");
}

size_t sbgm_dsm_scratch_floats(size_t) { return kLossBlocks; }

int sbgm_dsm_loss(const float* score, const float* std, const float* z, const float* sdf, int n, int per_member,
                  float* partials, float* loss_out, void* stream) {
  SBGM_REQUIRE(per_member % 4 == 0, "dsm_loss: per_member must be a multiple of 4");
  const size_t nq = static_cast<size_t>(n) * per_member / 4;
  const int blocks = static_cast<int>(nq < static_cast<size_t>(kLossBlocks) * kBlock ? (nq + kBlock - 1) / kBlock : kLossBlocks);
  launch_k((dsm_loss_partial_kernel), blocks, kBlock, 0, as_stream(stream), score, std, z, sdf, nq, per_member / 4, partials);
  launch_k((dsm_loss_finish_kernel), 1, 256, 0, as_stream(stream), partials, blocks, 1.0f / n, loss_out);
  return check_launch("dsm_loss");
}

int sbgm_ensemble_stats(const float* members, const float* truth, int m, size_t pixels, float* mean, float* stdev, float* crps,
                        void* stream) {
  SBGM_REQUIRE(m >= 1 && pixels >= 1, "ensemble_stats: empty ensemble");
  SBGM_REQUIRE(crps == nullptr || truth != nullptr, "ensemble_stats: CRPS needs the verifying field");
  launch_k((ensemble_stats_kernel), grid_for(pixels), kBlock, 0, as_stream(stream), members, truth, m, pixels, mean, stdev, crps);
  return check_launch("ensemble_stats");
}

int sbgm_back_transform(const float* x, float* y, size_t count, float pre_shift, float scale, float shift, float lo, float hi,
                        int do_clamp, int do_exp, void* stream) {
  launch_k((back_transform_kernel), grid_for(count), kBlock, 0, as_stream(stream), x, y, count, pre_shift, scale, shift, lo, hi, do_clamp, do_exp);
  return check_launch("back_transform");
}

}  // extern "C"
