// attention_mma.cu -- softmax(Q K^T / sqrt(d)) V on the warp-level tensor-core path (mma.sync m16n8k16 bf16, fp32
// accumulate) for the tensor-core storage formats; the fp32 strict mode keeps the CUDA-core kernel of attention.cu.
// (ImageSelfAttention, sbgm/score_unet.py:112-148: nn.MultiheadAttention core over S <= 1024 pixel tokens.)
//
// The attention blocks run on 4x4 .. 16x16 maps with head dims 32 .. 128: per (image, head) the whole problem is a few
// hundred KFLOP -- far below one tcgen05 tile -- so the register-resident flash-attention form is the right shape:
//   * a CTA of 4 warps owns 64 queries of one (image, head); a warp owns 16 of them (one m16 row block);
//   * keys / values stream through shared memory in chunks of 64 as bf16 hi|lo planes (V transposed so that both
//     products read their B fragments as conflict-free 32-bit words);
//   * scores S = Q K^T accumulate in registers, the online softmax runs on the accumulator fragment (a row lives in one
//     lane quad), and the probability fragment is re-used in place as the A operand of P V;
//   * split-bf16 (hi*hi + hi*lo + lo*hi) on both products keeps fp32-class accuracy in the bf16x3 mode.
#include "common.cuh"

namespace sbgm {

constexpr int kMmaQ = 64;       // queries per CTA
constexpr int kMmaKeys = 64;    // keys per chunk

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void mma_f16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// operand element type of the format: bfloat16 (one plane or hi|lo) or float16 (one plane, 11 significant bits)
template <int FMT>
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (FMT == SBGM_FMT_F16) mma_f16_16816(c, a, b0, b1); else mma_bf16_16816(c, a, b0, b1);
}
template <int FMT>
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  if (FMT == SBGM_FMT_F16) {
    hi = pack_f16x2(x, y);
    lo = 0u;
    return;
  }
  const float xh = bf16_round(x), yh = bf16_round(y);
  hi = pack_bf16x2(xh, yh);
  lo = pack_bf16x2(x - xh, y - yh);
}

template <int FMT, int D>
__global__ void __launch_bounds__(128)
attention_mma_kernel(const void* __restrict__ qkv, size_t plane, void* __restrict__ out, size_t out_plane, int s, int c, int heads,
                     float scale) {
  pdl_grid_sync();
  constexpr bool kLo = (FMT == SBGM_FMT_BF16X2);    // second (lo) operand planes
  using TF = TcFmt<FMT>;
  constexpr int kPlanes = kLo ? 2 : 1;
  constexpr int KP = D + 8;                         // row pitch of Q / K tiles (bf16 elements)
  constexpr int VP = kMmaKeys + 8;                  // row pitch of the transposed V tile
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* qs = reinterpret_cast<__nv_bfloat16*>(smem_raw);          // [planes][64][KP]
  __nv_bfloat16* ks = qs + kPlanes * kMmaQ * KP;                            // [planes][64][KP]
  __nv_bfloat16* vt = ks + kPlanes * kMmaKeys * KP;                         // [planes][D][VP]
  const int b = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * kMmaQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const size_t row_stride = static_cast<size_t>(3) * c;
  const size_t base = static_cast<size_t>(b) * s * row_stride + static_cast<size_t>(head) * D;
  constexpr int dvec = D / 8;

  // ---- Q tile -> smem (scaled, split), then A fragments ----
  for (int item = threadIdx.x; item < kMmaQ * dvec; item += blockDim.x) {
    const int qi = item / dvec, vec = item - qi * dvec;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (q0 + qi < s) Act<FMT>::load8(qkv, plane, base + static_cast<size_t>(q0 + qi) * row_stride + vec * 8, v);
    float hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float x = v[j] * scale;
      hi[j] = TF::round(x);
      lo[j] = x - hi[j];
    }
    *reinterpret_cast<uint4*>(qs + qi * KP + vec * 8) = TF::pack8(hi);
    if (kLo) *reinterpret_cast<uint4*>(qs + (kMmaQ + qi) * KP + vec * 8) = pack_bf16x8(lo);
  }
  __syncthreads();
  uint32_t qa[kPlanes][D / 16][4];
#pragma unroll
  for (int pl = 0; pl < kPlanes; ++pl)
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
      const __nv_bfloat16* p0 = qs + (pl * kMmaQ + warp * 16 + g) * KP + kk * 16 + t4 * 2;
      qa[pl][kk][0] = *reinterpret_cast<const uint32_t*>(p0);
      qa[pl][kk][1] = *reinterpret_cast<const uint32_t*>(p0 + 8 * KP);
      qa[pl][kk][2] = *reinterpret_cast<const uint32_t*>(p0 + 8);
      qa[pl][kk][3] = *reinterpret_cast<const uint32_t*>(p0 + 8 * KP + 8);
    }

  float o[D / 8][4];
#pragma unroll
  for (int dt = 0; dt < D / 8; ++dt)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[dt][e] = 0.0f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;     // rows g and g + 8 of this warp's block

  for (int k0 = 0; k0 < s; k0 += kMmaKeys) {
    __syncthreads();
    for (int item = threadIdx.x; item < kMmaKeys * dvec; item += blockDim.x) {
      const int kj = item / dvec, vec = item - kj * dvec;
      float kv[8] = {0, 0, 0, 0, 0, 0, 0, 0}, vv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (k0 + kj < s) {
        const size_t rowp = base + static_cast<size_t>(k0 + kj) * row_stride + vec * 8;
        Act<FMT>::load8(qkv, plane, rowp + c, kv);
        Act<FMT>::load8(qkv, plane, rowp + 2 * c, vv);
      }
      float hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        hi[j] = TF::round(kv[j]);
        lo[j] = kv[j] - hi[j];
      }
      *reinterpret_cast<uint4*>(ks + kj * KP + vec * 8) = TF::pack8(hi);
      if (kLo) *reinterpret_cast<uint4*>(ks + (kMmaKeys + kj) * KP + vec * 8) = pack_bf16x8(lo);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (FMT == SBGM_FMT_F16) {      // same 16-bit slots, float16 bit patterns
          reinterpret_cast<__half*>(vt)[(vec * 8 + j) * VP + kj] = __float2half_rn(vv[j]);
          continue;
        }
        const float vh = bf16_round(vv[j]);
        vt[(vec * 8 + j) * VP + kj] = __float2bfloat16_rn(vh);
        if (kLo) vt[(D + vec * 8 + j) * VP + kj] = __float2bfloat16_rn(vv[j] - vh);
      }
    }
    __syncthreads();

    // ---- scores: 16 queries x 64 keys ----
    float sc[kMmaKeys / 8][4];
#pragma unroll
    for (int nt = 0; nt < kMmaKeys / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) sc[nt][e] = 0.0f;
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        const __nv_bfloat16* kp = ks + (nt * 8 + g) * KP + kk * 16 + t4 * 2;
        const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(kp), bh1 = *reinterpret_cast<const uint32_t*>(kp + 8);
        mma_16816<FMT>(sc[nt], qa[0][kk], bh0, bh1);
        if (kLo) {
          const __nv_bfloat16* kl = kp + kMmaKeys * KP;
          mma_bf16_16816(sc[nt], qa[0][kk], *reinterpret_cast<const uint32_t*>(kl), *reinterpret_cast<const uint32_t*>(kl + 8));
          mma_bf16_16816(sc[nt], qa[kPlanes - 1][kk], bh0, bh1);
        }
      }
    }
    // ---- online softmax on the accumulator fragment (row g: e = 0, 1; row g + 8: e = 2, 3) ----
    float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < kMmaKeys / 8; ++nt) {
      const int key = k0 + nt * 8 + t4 * 2;
      if (key >= s) sc[nt][0] = sc[nt][2] = -INFINITY;
      if (key + 1 >= s) sc[nt][1] = sc[nt][3] = -INFINITY;
      cm0 = fmaxf(cm0, fmaxf(sc[nt][0], sc[nt][1]));
      cm1 = fmaxf(cm1, fmaxf(sc[nt][2], sc[nt][3]));
    }
    cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1)); cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
    cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1)); cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
    const float mn0 = fmaxf(m0, cm0), mn1 = fmaxf(m1, cm1);
    const float corr0 = __expf(m0 - mn0), corr1 = __expf(m1 - mn1);      // m = -inf on the first chunk -> 0
    m0 = mn0; m1 = mn1;
    l0 *= corr0; l1 *= corr1;
#pragma unroll
    for (int dt = 0; dt < D / 8; ++dt) {
      o[dt][0] *= corr0; o[dt][1] *= corr0; o[dt][2] *= corr1; o[dt][3] *= corr1;
    }
#pragma unroll
    for (int nt = 0; nt < kMmaKeys / 8; ++nt) {
      sc[nt][0] = __expf(sc[nt][0] - mn0); sc[nt][1] = __expf(sc[nt][1] - mn0);
      sc[nt][2] = __expf(sc[nt][2] - mn1); sc[nt][3] = __expf(sc[nt][3] - mn1);
      l0 += sc[nt][0] + sc[nt][1];
      l1 += sc[nt][2] + sc[nt][3];
    }
    // ---- O += P V: the probability fragment of key tiles (2 kk, 2 kk + 1) IS the A fragment of k-step kk ----
#pragma unroll
    for (int kk = 0; kk < kMmaKeys / 16; ++kk) {
      uint32_t ph[4], pl[4];
      split2<FMT>(sc[2 * kk][0], sc[2 * kk][1], ph[0], pl[0]);
      split2<FMT>(sc[2 * kk][2], sc[2 * kk][3], ph[1], pl[1]);
      split2<FMT>(sc[2 * kk + 1][0], sc[2 * kk + 1][1], ph[2], pl[2]);
      split2<FMT>(sc[2 * kk + 1][2], sc[2 * kk + 1][3], ph[3], pl[3]);
#pragma unroll
      for (int dt = 0; dt < D / 8; ++dt) {
        const __nv_bfloat16* vp = vt + (dt * 8 + g) * VP + kk * 16 + t4 * 2;
        const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(vp), bh1 = *reinterpret_cast<const uint32_t*>(vp + 8);
        mma_16816<FMT>(o[dt], ph, bh0, bh1);
        if (kLo) {
          const __nv_bfloat16* vl = vp + D * VP;
          mma_bf16_16816(o[dt], ph, *reinterpret_cast<const uint32_t*>(vl), *reinterpret_cast<const uint32_t*>(vl + 8));
          mma_bf16_16816(o[dt], pl, bh0, bh1);
        }
      }
    }
  }
  // ---- normalise, stage through smem (fp32 [64][D + 4]) and store 8-channel vectors ----
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
  __syncthreads();
  float* os = reinterpret_cast<float*>(smem_raw);
  constexpr int OP = D + 4;
#pragma unroll
  for (int dt = 0; dt < D / 8; ++dt) {
    float* r0 = os + (warp * 16 + g) * OP + dt * 8 + t4 * 2;
    r0[0] = o[dt][0] * inv0; r0[1] = o[dt][1] * inv0;
    r0[8 * OP] = o[dt][2] * inv1; r0[8 * OP + 1] = o[dt][3] * inv1;
  }
  __syncthreads();
  for (int item = threadIdx.x; item < kMmaQ * dvec; item += blockDim.x) {
    const int qi = item / dvec, vec = item - qi * dvec;
    if (q0 + qi >= s) continue;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = os[qi * OP + vec * 8 + j];
    Act<FMT>::store8(out, out_plane, (static_cast<size_t>(b) * s + q0 + qi) * c + head * D + vec * 8, v);
  }
}

template <int FMT, int D>
static int launch_attention_mma(const void* qkv, size_t plane, void* out, size_t out_plane, int b, int s, int c, int heads,
                                cudaStream_t st) {
  constexpr int kPlanes = (FMT == SBGM_FMT_BF16X2) ? 2 : 1;
  const size_t tiles = static_cast<size_t>(kPlanes) * (2 * kMmaQ * (D + 8) + D * (kMmaKeys + 8)) * sizeof(__nv_bfloat16);
  const size_t stage = static_cast<size_t>(kMmaQ) * (D + 4) * sizeof(float);
  const size_t smem = tiles > stage ? tiles : stage;
  auto kern = attention_mma_kernel<FMT, D>;
  static bool configured = false;
  if (!configured && smem > 48 * 1024) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) {
      set_error("attention_mma: cannot reserve %zu bytes of shared memory", smem);
      return 1;
    }
    configured = true;
  }
  dim3 grid(ceil_div(s, kMmaQ), heads, b);
  launch_k(kern, grid, 128, smem, st, qkv, plane, out, out_plane, s, c, heads, 1.0f / sqrtf(static_cast<float>(D)));
  return check_launch("attention_mma");
}

// returns -1 if this (format, head dim) is not covered (the caller falls back to the CUDA-core kernel)
int attention_mma_dispatch(const void* qkv, size_t plane, void* out, size_t out_plane, int fmt, int b, int s, int c, int heads,
                           cudaStream_t st) {
  const int d = c / heads;
#define SBGM_AM(F, DD) return launch_attention_mma<F, DD>(qkv, plane, out, out_plane, b, s, c, heads, st)
  if (fmt == SBGM_FMT_BF16X2) {
    if (d == 32) SBGM_AM(SBGM_FMT_BF16X2, 32);
    if (d == 64) SBGM_AM(SBGM_FMT_BF16X2, 64);
    if (d == 128) SBGM_AM(SBGM_FMT_BF16X2, 128);
  } else if (fmt == SBGM_FMT_BF16) {
    if (d == 32) SBGM_AM(SBGM_FMT_BF16, 32);
    if (d == 64) SBGM_AM(SBGM_FMT_BF16, 64);
    if (d == 128) SBGM_AM(SBGM_FMT_BF16, 128);
  } else if (fmt == SBGM_FMT_F16) {
    if (d == 32) SBGM_AM(SBGM_FMT_F16, 32);
    if (d == 64) SBGM_AM(SBGM_FMT_F16, 64);
    if (d == 128) SBGM_AM(SBGM_FMT_F16, 128);
  }
#undef SBGM_AM
  return -1;
}

}  // namespace sbgm
