// attention.cu -- softmax(Q K^T / sqrt(d)) V for the low-resolution self-attention blocks
// (ImageSelfAttention, sbgm/score_unet.py:112-148; S <= 1024 tokens, 1..16 heads, d = 8..512).
//
// fp32 online-softmax kernel on CUDA cores: a block owns 32 queries of one (batch, head); each warp
// owns 4 queries.  Keys/values stream through shared memory in chunks of 32 (lane j <-> key j for
// the scores, lane <-> output dimension for P V).  The attention core is ~3% of the network FLOPs;
// the four Linear layers around it run on the tensor cores through sbgm_conv2d_tc.
#include "common.cuh"

namespace sbgm {

constexpr int kAttnWarps = 8;
constexpr int kKeyChunk = 32;

// QPW queries per warp (8 when the per-lane output slice is small, else 4).  Per 32-key chunk a warp does
//   scores: lane <-> key, d x (1 LDS k + QPW/4 LDS.128 q-broadcast + QPW FMA)
//   online softmax per query (warp max / sum)
//   P V:    probabilities staged in smem, lane <-> output dim, 32 x (QPW/4 LDS.128 p-broadcast + DPL LDS v + QPW*DPL FMA)
template <int FMT, int DPL, int QPW>
__global__ void __launch_bounds__(kAttnWarps * 32)
attention_kernel(const void* __restrict__ qkv, size_t plane, void* __restrict__ out, size_t out_plane, int s, int c,
                 int heads, float scale) {
  pdl_grid_sync();
  extern __shared__ __align__(16) float smem[];
  const int d = c / heads;
  float* ks = smem;                                              // [32][d + 1]
  float* vs = ks + kKeyChunk * (d + 1) + ((kKeyChunk * (d + 1)) & 3 ? 4 - ((kKeyChunk * (d + 1)) & 3) : 0);  // [32][d], 16-B aligned
  float* qs = vs + kKeyChunk * d;                                // [warps][d][QPW]
  float* ps = qs + kAttnWarps * d * QPW;                         // [warps][32][QPW]
  const int b = blockIdx.z, head = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (kAttnWarps * QPW) + warp * QPW;
  const size_t row_stride = static_cast<size_t>(3) * c;
  const size_t base = static_cast<size_t>(b) * s * row_stride;
  const int dvec = d >> 3;

  float* myq = qs + static_cast<size_t>(warp) * d * QPW;
  float* myp = ps + static_cast<size_t>(warp) * kKeyChunk * QPW;
  for (int item = lane; item < QPW * dvec; item += 32) {
    const int qi = item / dvec, vec = item % dvec;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (q0 + qi < s) Act<FMT>::load8(qkv, plane, base + static_cast<size_t>(q0 + qi) * row_stride + head * d + vec * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) myq[(vec * 8 + j) * QPW + qi] = v[j] * scale;
  }

  float m[QPW], l[QPW], acc[QPW][DPL];
#pragma unroll
  for (int qi = 0; qi < QPW; ++qi) {
    m[qi] = -INFINITY;
    l[qi] = 0.0f;
#pragma unroll
    for (int dd = 0; dd < DPL; ++dd) acc[qi][dd] = 0.0f;
  }

  for (int k0 = 0; k0 < s; k0 += kKeyChunk) {
    __syncthreads();
    for (int item = threadIdx.x; item < kKeyChunk * dvec; item += blockDim.x) {
      const int kj = item / dvec, vec = item % dvec;
      float kv[8] = {0, 0, 0, 0, 0, 0, 0, 0}, vv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (k0 + kj < s) {
        const size_t rowp = base + static_cast<size_t>(k0 + kj) * row_stride + head * d + vec * 8;
        Act<FMT>::load8(qkv, plane, rowp + c, kv);
        Act<FMT>::load8(qkv, plane, rowp + 2 * c, vv);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ks[kj * (d + 1) + vec * 8 + j] = kv[j];
        vs[kj * d + vec * 8 + j] = vv[j];
      }
    }
    __syncthreads();
    float sc[QPW];
#pragma unroll
    for (int qi = 0; qi < QPW; ++qi) sc[qi] = 0.0f;
    const float* krow = ks + lane * (d + 1);
    for (int i = 0; i < d; ++i) {
      const float kvv = krow[i];
#pragma unroll
      for (int g = 0; g < QPW / 4; ++g) {
        const float4 qv = *reinterpret_cast<const float4*>(myq + i * QPW + 4 * g);
        sc[4 * g + 0] = fmaf(kvv, qv.x, sc[4 * g + 0]); sc[4 * g + 1] = fmaf(kvv, qv.y, sc[4 * g + 1]);
        sc[4 * g + 2] = fmaf(kvv, qv.z, sc[4 * g + 2]); sc[4 * g + 3] = fmaf(kvv, qv.w, sc[4 * g + 3]);
      }
    }
    const bool key_ok = (k0 + lane) < s;
    float corr[QPW];
#pragma unroll
    for (int qi = 0; qi < QPW; ++qi) {
      const float sv = key_ok ? sc[qi] : -INFINITY;
      const float mnew = fmaxf(m[qi], warp_max(sv));
      const float pv = key_ok ? expf(sv - mnew) : 0.0f;
      corr[qi] = expf(m[qi] - mnew);          // m = -inf on the first chunk -> 0
      l[qi] = l[qi] * corr[qi] + warp_sum(pv);
      m[qi] = mnew;
      sc[qi] = pv;
    }
    __syncwarp();
#pragma unroll
    for (int g = 0; g < QPW / 4; ++g)
      *reinterpret_cast<float4*>(myp + lane * QPW + 4 * g) = make_float4(sc[4 * g], sc[4 * g + 1], sc[4 * g + 2], sc[4 * g + 3]);
    __syncwarp();
#pragma unroll
    for (int qi = 0; qi < QPW; ++qi)
#pragma unroll
      for (int dd = 0; dd < DPL; ++dd) acc[qi][dd] *= corr[qi];
    for (int j = 0; j < kKeyChunk; ++j) {
      float pj[QPW];
#pragma unroll
      for (int g = 0; g < QPW / 4; ++g) {
        const float4 t = *reinterpret_cast<const float4*>(myp + j * QPW + 4 * g);
        pj[4 * g] = t.x; pj[4 * g + 1] = t.y; pj[4 * g + 2] = t.z; pj[4 * g + 3] = t.w;
      }
#pragma unroll
      for (int dd = 0; dd < DPL; ++dd) {
        const int dim = lane + 32 * dd;
        if (dim < d) {
          const float vv = vs[j * d + dim];
#pragma unroll
          for (int qi = 0; qi < QPW; ++qi) acc[qi][dd] = fmaf(pj[qi], vv, acc[qi][dd]);
        }
      }
    }
  }
  // write: stage through smem (this warp's q slot) so stores are 8-channel vectors
  __syncwarp();
#pragma unroll
  for (int qi = 0; qi < QPW; ++qi) {
    const float inv = 1.0f / l[qi];
#pragma unroll
    for (int dd = 0; dd < DPL; ++dd) {
      const int dim = lane + 32 * dd;
      if (dim < d) myq[dim * QPW + qi] = acc[qi][dd] * inv;
    }
  }
  __syncwarp();
  for (int item = lane; item < QPW * dvec; item += 32) {
    const int qi = item / dvec, vec = item % dvec;
    if (q0 + qi >= s) continue;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = myq[(vec * 8 + j) * QPW + qi];
    Act<FMT>::store8(out, out_plane, (static_cast<size_t>(b) * s + q0 + qi) * c + head * d + vec * 8, v);
  }
}

template <int FMT, int DPL>
static int launch_attention(const void* qkv, size_t plane, void* out, size_t out_plane, int b, int s, int c, int heads,
                            cudaStream_t st) {
  constexpr int QPW = (DPL <= 4) ? 8 : 4;
  const int d = c / heads;
  const size_t ks_floats = (static_cast<size_t>(kKeyChunk) * (d + 1) + 3) & ~static_cast<size_t>(3);
  const size_t smem = (ks_floats + kKeyChunk * d + kAttnWarps * d * QPW + kAttnWarps * kKeyChunk * QPW) * sizeof(float);
  auto kern = attention_kernel<FMT, DPL, QPW>;
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) {
      set_error("attention: cannot reserve %zu bytes of shared memory", smem);
      return 1;
    }
  }
  dim3 grid(ceil_div(s, kAttnWarps * QPW), heads, b);
  launch_k((kern), grid, kAttnWarps * 32, smem, st, qkv, plane, out, out_plane, s, c, heads, 1.0f / sqrtf(static_cast<float>(d)));
  return check_launch("attention");
}

}  // namespace sbgm

namespace sbgm {
int attention_mma_dispatch(const void* qkv, size_t plane, void* out, size_t out_plane, int fmt, int b, int s, int c, int heads,
                           cudaStream_t st);
}
using namespace sbgm;

extern "C" int sbgm_attention(const void* qkv, size_t qkv_plane, void* out, size_t out_plane, int fmt,
                              int b, int s, int c, int heads, void* stream) {
  SBGM_REQUIRE(heads >= 1 && c % heads == 0, "attention: c=%d not divisible by heads=%d", c, heads);
  const int d = c / heads;
  SBGM_REQUIRE(d % 8 == 0 && d <= 512, "attention: head dim %d must be a multiple of 8 and <= 512", d);
  cudaStream_t st = as_stream(stream);
  {   // tensor-core formats: warp-level MMA flash attention (attention_mma.cu) for head dims 32 / 64 / 128
    const int rc = attention_mma_dispatch(qkv, qkv_plane, out, out_plane, fmt, b, s, c, heads, st);
    if (rc >= 0) return rc;
  }
  const int dpl = (d + 31) / 32;
#define SBGM_ATTN(D) SBGM_DISPATCH_FMT(fmt, return (launch_attention<FMT, D>(qkv, qkv_plane, out, out_plane, b, s, c, heads, st)))
  if (dpl <= 1) { SBGM_ATTN(1); }
  else if (dpl <= 2) { SBGM_ATTN(2); }
  else if (dpl <= 4) { SBGM_ATTN(4); }
  else if (dpl <= 8) { SBGM_ATTN(8); }
  else { SBGM_ATTN(16); }
#undef SBGM_ATTN
  return 0;
}
