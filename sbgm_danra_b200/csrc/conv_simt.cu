// conv_simt.cu -- CUDA-core convolutions.
//  * sbgm_stem_conv   : Encoder.conv1 (8x8 stride 2 pad 3 over 2..16 input channels, score_unet.py:206-211).
//                       K per input channel is only 64 and the input is NCHW fp32 planes, so this
//                       is a direct convolution with the weights of one channel staged in smem.
//  * sbgm_conv2d_simt : generic fp32 NHWC implicit GEMM (exact-arithmetic mode; also the reference
//                       the tcgen05 kernel is unit-tested against on the GPU).
#include "common.cuh"

namespace sbgm {

// ---- stem ------------------------------------------------------------------------------------
// Block = 256 threads = 16 rows x 16 column-pairs of output pixels (a 16 x 32 tile); every thread keeps
// 2 pixels x 64 output channels in registers so each weight vector read from smem feeds two FMAs.
constexpr int kStemTH = 16, kStemTW = 32;
constexpr int kStemInH = 2 * kStemTH + 6;        // 38 input rows
constexpr int kStemInW = 2 * kStemTW + 6;        // 70 input cols

template <int FMT>
__global__ void __launch_bounds__(256)
stem_conv_kernel(const float* __restrict__ x, const float* __restrict__ planes, int np, int cc, int c_begin, int c_end,
                 const float* __restrict__ wp, const float* __restrict__ addend, int na, const float* __restrict__ tproj,
                 int tproj_stride, void* __restrict__ out, size_t out_plane, int h, int w) {
  pdl_grid_sync();
  __shared__ float s_in[kStemInH][kStemInW + 1];
  __shared__ __align__(16) float s_w[64][64];    // [tap][co]
  const int n = blockIdx.z, ho = h / 2, wo = w / 2;
  const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
  const int oy = blockIdx.y * kStemTH + ty, ox = blockIdx.x * kStemTW + 2 * tx;
  const int iy0 = blockIdx.y * kStemTH * 2 - 3, ix0 = blockIdx.x * kStemTW * 2 - 3;
  float acc0[64], acc1[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) { acc0[i] = 0.0f; acc1[i] = 0.0f; }

  for (int c = c_begin; c < c_end; ++c) {
    const float* src = (c == 0) ? x + static_cast<size_t>(n) * h * w
                                : planes + (static_cast<size_t>(np == 1 ? 0 : n) * cc + (c - 1)) * h * w;
    __syncthreads();
    for (int i = threadIdx.x; i < kStemInH * kStemInW; i += 256) {
      const int yy = i / kStemInW, xx = i % kStemInW;
      const int iy = iy0 + yy, ix = ix0 + xx;
      s_in[yy][xx] = (iy >= 0 && iy < h && ix >= 0 && ix < w) ? __ldg(src + static_cast<size_t>(iy) * w + ix) : 0.0f;
    }
    const float4* wsrc = reinterpret_cast<const float4*>(wp + static_cast<size_t>(c) * 64 * 64);
    for (int i = threadIdx.x; i < 64 * 16; i += 256) reinterpret_cast<float4*>(&s_w[0][0])[i] = __ldg(wsrc + i);
    __syncthreads();
#pragma unroll 1
    for (int r = 0; r < 8; ++r) {
#pragma unroll 2
      for (int s = 0; s < 8; ++s) {
        const float v0 = s_in[2 * ty + r][4 * tx + s];
        const float v1 = s_in[2 * ty + r][4 * tx + 2 + s];
        const float4* wrow = reinterpret_cast<const float4*>(&s_w[r * 8 + s][0]);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 wv = wrow[j];
          acc0[4 * j + 0] = fmaf(v0, wv.x, acc0[4 * j + 0]); acc1[4 * j + 0] = fmaf(v1, wv.x, acc1[4 * j + 0]);
          acc0[4 * j + 1] = fmaf(v0, wv.y, acc0[4 * j + 1]); acc1[4 * j + 1] = fmaf(v1, wv.y, acc1[4 * j + 1]);
          acc0[4 * j + 2] = fmaf(v0, wv.z, acc0[4 * j + 2]); acc1[4 * j + 2] = fmaf(v1, wv.z, acc1[4 * j + 2]);
          acc0[4 * j + 3] = fmaf(v0, wv.w, acc0[4 * j + 3]); acc1[4 * j + 3] = fmaf(v1, wv.w, acc1[4 * j + 3]);
        }
      }
    }
  }
  if (oy >= ho) return;
  const float* tp = tproj ? tproj + static_cast<size_t>(n) * tproj_stride : nullptr;
#pragma unroll
  for (int px = 0; px < 2; ++px) {
    if (ox + px >= wo) continue;
    const size_t pix = (static_cast<size_t>(n) * ho + oy) * wo + ox + px;
    const float* ad = addend ? addend + ((static_cast<size_t>(na == 1 ? 0 : n) * ho + oy) * wo + ox + px) * 64 : nullptr;
#pragma unroll
    for (int v8 = 0; v8 < 8; ++v8) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = px == 0 ? acc0[v8 * 8 + j] : acc1[v8 * 8 + j];
        if (ad) t += __ldg(ad + v8 * 8 + j);
        if (tp) t += __ldg(tp + v8 * 8 + j);
        o[j] = t;
      }
      Act<FMT>::store8(out, out_plane, pix * 64 + v8 * 8, o);
    }
  }
}

// ---- generic fp32 implicit GEMM ------------------------------------------------------------------
constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256)
conv_simt_kernel(const float* __restrict__ in, const float* __restrict__ wgt, const float* __restrict__ bias,
                 const float* __restrict__ residual, const float* __restrict__ tproj, int tproj_stride,
                 float* __restrict__ out, int n, int h, int w, int cin, int cout, int kh, int kw, int stride, int pad,
                 int ho, int wo, int act) {
  pdl_grid_sync();
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int M = n * ho * wo;
  const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
  // A-load role: row = tid / 4 (0..63), k-quad = tid % 4
  const int arow = threadIdx.x >> 2, aq = threadIdx.x & 3;
  const int am = m0 + arow;
  int a_n = 0, a_oy = 0, a_ox = 0;
  const bool a_valid = am < M;
  if (a_valid) {
    a_ox = am % wo;
    a_oy = (am / wo) % ho;
    a_n = am / (wo * ho);
  }
  // B-load role: k = tid / 16, co-quad = tid % 16
  const int bk = threadIdx.x >> 4, bq = threadIdx.x & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  const int kchunks = cin / BK;
  for (int tap = 0; tap < kh * kw; ++tap) {
    const int r = tap / kw, s = tap % kw;
    const int iy = a_oy * stride + r - pad, ix = a_ox * stride + s - pad;
    const bool ok = a_valid && iy >= 0 && iy < h && ix >= 0 && ix < w;
    const float* arow_ptr = in + ((static_cast<size_t>(a_n) * h + iy) * w + ix) * cin;
    for (int kc = 0; kc < kchunks; ++kc) {
      float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) av = __ldg(reinterpret_cast<const float4*>(arow_ptr + kc * BK + aq * 4));
      const size_t krow = static_cast<size_t>(tap) * cin + kc * BK + bk;
      const float4 bv = __ldg(reinterpret_cast<const float4*>(wgt + krow * cout + n0 + bq * 4));
      __syncthreads();
      As[aq * 4 + 0][arow] = av.x; As[aq * 4 + 1][arow] = av.y; As[aq * 4 + 2][arow] = av.z; As[aq * 4 + 3][arow] = av.w;
      *reinterpret_cast<float4*>(&Bs[bk][bq * 4]) = bv;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int co = n0 + tx * 4;
    const int bn = m / (wo * ho);
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + (bias ? __ldg(bias + co + j) : 0.0f);
    if (residual) {
      const float4 rv = __ldg(reinterpret_cast<const float4*>(residual + static_cast<size_t>(m) * cout + co));
      v[0] += rv.x; v[1] += rv.y; v[2] += rv.z; v[3] += rv.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = apply_act(v[j], act);
    if (tproj) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += __ldg(tproj + static_cast<size_t>(bn) * tproj_stride + co + j);
    }
    *reinterpret_cast<float4*>(out + static_cast<size_t>(m) * cout + co) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

}  // namespace sbgm

using namespace sbgm;

extern "C" {

int sbgm_stem_conv(const float* x, const float* planes, int np, int cc, int c_begin, int c_end,
                   const float* w_packed, const float* addend, int na, const float* tproj, int tproj_stride,
                   void* out, size_t out_plane, int fmt, int n, int h, int w, void* stream) {
  SBGM_REQUIRE(h % 2 == 0 && w % 2 == 0, "stem_conv: h=%d w=%d must be even", h, w);
  SBGM_REQUIRE(c_begin >= 0 && c_end <= cc + 1 && c_begin <= c_end, "stem_conv: bad channel range [%d,%d) of %d", c_begin, c_end, cc + 1);
  SBGM_REQUIRE(c_begin > 0 || x != nullptr || c_end == 0, "stem_conv: x is NULL but channel 0 requested");
  dim3 grid(ceil_div(w / 2, kStemTW), ceil_div(h / 2, kStemTH), n);
  SBGM_DISPATCH_FMT(fmt, (launch_k((stem_conv_kernel<FMT>), grid, 256, 0, as_stream(stream), 
                             x, planes, np, cc, c_begin, c_end, w_packed, addend, na, tproj, tproj_stride, out,
                             out_plane, h, w)));
  return check_launch("stem_conv");
}

int sbgm_conv2d_simt(const float* in, const float* weight, const float* bias, const float* residual,
                     const float* tproj, int tproj_stride, float* out, int n, int h, int w, int cin, int cout,
                     int kh, int kw, int stride, int pad, int act, void* stream) {
  SBGM_REQUIRE(cin % BK == 0 && cout % BN == 0, "conv2d_simt: cin=%d must be a multiple of %d and cout=%d of %d", cin, BK, cout, BN);
  const int ho = (h + 2 * pad - kh) / stride + 1, wo = (w + 2 * pad - kw) / stride + 1;
  SBGM_REQUIRE(ho > 0 && wo > 0, "conv2d_simt: empty output");
  dim3 grid(ceil_div(static_cast<long long>(n) * ho * wo, BM), cout / BN);
  launch_k((conv_simt_kernel), grid, 256, 0, as_stream(stream), in, weight, bias, residual, tproj, tproj_stride, out, n, h, w,
                                                        cin, cout, kh, kw, stride, pad, ho, wo, act);
  return check_launch("conv2d_simt");
}

}  // extern "C"
