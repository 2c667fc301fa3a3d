// attention_bwd_mma.cu -- backward of the attention core softmax(Q K^T / sqrt(d)) V as ONE warp-level tensor-core kernel
// (mma.sync m16n8k16 bf16, fp32 accumulate) for the bf16 training format; the fp32 / split-bf16 modes and odd shapes keep
// the batched CUDA-core path of attention_bwd.cu.  (ImageSelfAttention, sbgm/score_unet.py:112-148; torch autograd through
// nn.MultiheadAttention in the reference.)
//
// A CTA owns one (image, head): Q, K, V, dO of that head (S <= 256 tokens, head dim 32..128) sit in shared memory as bf16,
// row-major AND -- for the operands whose reduction index is the token -- transposed, so that every mma B fragment is a
// conflict-free 32-bit word.  No probability matrix ever leaves the registers:
//   phase A, a warp per 16-QUERY slab:  pass 1  S = Q K^T -> log-sum-exp per query (online over key chunks);
//                                       pass 2  P = exp(scale S - lse), dP = dO V^T, dS = P o (dP - D), dQ += dS K
//   phase B, a warp per 16-KEY slab:    S^T = K Q^T, dP^T = V dO^T (recomputed: no transposes of register fragments),
//                                       dV += P^T dO, dK += dS^T Q
// with D_i = sum_c dO_ic O_ic from the forward's output.  The accumulator fragment of S / dS is re-used in place as the
// A operand of the following product (as in attention_mma.cu).  7 small GEMMs instead of 5, but one launch instead of
// eight, no fp32 workspace round trips (7 x tokens x c + 2 x S^2 floats before), and the arithmetic is on the tensor cores.
#include "common.cuh"

namespace sbgm {

namespace {
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t ld32(const __nv_bfloat16* p) { return *reinterpret_cast<const uint32_t*>(p); }
// A fragments (k-steps of 16 along the row) of the 16-row block starting at `rows` in a row-major [.. ][pitch] tile
template <int KSTEPS>
__device__ __forceinline__ void load_a(uint32_t (&a)[KSTEPS][4], const __nv_bfloat16* rows, int pitch, int g, int t4) {
#pragma unroll
  for (int kk = 0; kk < KSTEPS; ++kk) {
    const __nv_bfloat16* p0 = rows + g * pitch + kk * 16 + t4 * 2;
    a[kk][0] = ld32(p0);
    a[kk][1] = ld32(p0 + 8 * pitch);
    a[kk][2] = ld32(p0 + 8);
    a[kk][3] = ld32(p0 + 8 * pitch + 8);
  }
}
// acc[nt] (16 x 8 tiles over CH columns) = A(16 x D) * B^T, B rows = `cols` .. in a row-major [.. ][pitch] tile
template <int KSTEPS, int NT>
__device__ __forceinline__ void gemm_rows(float (&acc)[NT][4], const uint32_t (&a)[KSTEPS][4], const __nv_bfloat16* cols, int pitch, int g, int t4) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[nt][e] = 0.0f;
#pragma unroll
    for (int kk = 0; kk < KSTEPS; ++kk) {
      const __nv_bfloat16* p = cols + (nt * 8 + g) * pitch + kk * 16 + t4 * 2;
      mma16816(acc[nt], a[kk], ld32(p), ld32(p + 8));
    }
  }
}
// out[dt] (16 x D) += F(16 x CH, accumulator fragments re-used as A) * B, B given transposed: bt[d][token], tokens from `tok0`
template <int NT, int DT>
__device__ __forceinline__ void gemm_frag(float (&out)[DT][4], const float (&f)[NT][4], const __nv_bfloat16* bt, int pitch, int tok0, int g, int t4) {
#pragma unroll
  for (int kk = 0; kk < NT / 2; ++kk) {
    uint32_t a[4];
    a[0] = pack_bf16x2(f[2 * kk][0], f[2 * kk][1]);
    a[1] = pack_bf16x2(f[2 * kk][2], f[2 * kk][3]);
    a[2] = pack_bf16x2(f[2 * kk + 1][0], f[2 * kk + 1][1]);
    a[3] = pack_bf16x2(f[2 * kk + 1][2], f[2 * kk + 1][3]);
#pragma unroll
    for (int dt = 0; dt < DT; ++dt) {
      const __nv_bfloat16* p = bt + (dt * 8 + g) * pitch + tok0 + kk * 16 + t4 * 2;
      mma16816(out[dt], a, ld32(p), ld32(p + 8));
    }
  }
}
__device__ __forceinline__ float dot8_bf16(const uint4& a, const uint4& b) {
  float x[8], y[8];
  unpack_bf16x8(a, x);
  unpack_bf16x8(b, y);
  float s = 0.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s = fmaf(x[j], y[j], s);
  return s;
}
}  // namespace

template <int D, int CH>
__global__ void __launch_bounds__(256)
attention_bwd_mma_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                         __nv_bfloat16* __restrict__ dqkv, int s, int c, float scale) {
  pdl_grid_sync();
  constexpr int RP = D + 8, KS = D / 16, NT = CH / 8, DT = D / 8, dvec = D / 8;
  const int TP = s + 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_raw);      // [s][RP] row-major
  __nv_bfloat16* Ks = Qs + s * RP;
  __nv_bfloat16* Vs = Ks + s * RP;
  __nv_bfloat16* Gs = Vs + s * RP;                                     // dO
  __nv_bfloat16* Qt = Gs + s * RP;                                     // [D][TP] transposed
  __nv_bfloat16* Kt = Qt + D * TP;
  __nv_bfloat16* Gt = Kt + D * TP;
  float* lse = reinterpret_cast<float*>(Gt + D * TP);                  // [s]
  float* dsum = lse + s;                                               // [s]  D_i = sum_c dO_ic O_ic
  const int head = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const size_t row_stride = static_cast<size_t>(3) * c;
  const size_t base = static_cast<size_t>(b) * s * row_stride + static_cast<size_t>(head) * D;
  const size_t obase = static_cast<size_t>(b) * s * c + static_cast<size_t>(head) * D;

  // ---- load: one (token, 8-channel vector) per item; s * dvec is a multiple of 64, so warps stay whole ----
  for (int item = threadIdx.x; item < s * dvec; item += blockDim.x) {
    const int row = item / dvec, vec = item - row * dvec;
    const __nv_bfloat16* p = qkv + base + static_cast<size_t>(row) * row_stride + vec * 8;
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
    const uint4 k = __ldg(reinterpret_cast<const uint4*>(p + c));
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p + 2 * c));
    const uint4 gd = __ldg(reinterpret_cast<const uint4*>(dout + obase + static_cast<size_t>(row) * c + vec * 8));
    const uint4 oo = __ldg(reinterpret_cast<const uint4*>(o + obase + static_cast<size_t>(row) * c + vec * 8));
    *reinterpret_cast<uint4*>(Qs + row * RP + vec * 8) = q;
    *reinterpret_cast<uint4*>(Ks + row * RP + vec * 8) = k;
    *reinterpret_cast<uint4*>(Vs + row * RP + vec * 8) = v;
    *reinterpret_cast<uint4*>(Gs + row * RP + vec * 8) = gd;
    const __nv_bfloat16* qe = reinterpret_cast<const __nv_bfloat16*>(&q);
    const __nv_bfloat16* ke = reinterpret_cast<const __nv_bfloat16*>(&k);
    const __nv_bfloat16* ge = reinterpret_cast<const __nv_bfloat16*>(&gd);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      Qt[(vec * 8 + j) * TP + row] = qe[j];
      Kt[(vec * 8 + j) * TP + row] = ke[j];
      Gt[(vec * 8 + j) * TP + row] = ge[j];
    }
    float part = dot8_bf16(gd, oo);            // the dvec lanes of a token are adjacent: fixed-order butterfly
#pragma unroll
    for (int off = dvec / 2; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if (vec == 0) dsum[row] = part;
  }
  __syncthreads();

  // ---- phase A: query slabs -> lse, dQ ----
  for (int slab = warp; slab < s / 16; slab += nwarps) {
    const int q0 = slab * 16;
    uint32_t qa[KS][4], ga[KS][4];
    load_a<KS>(qa, Qs + q0 * RP, RP, g, t4);
    load_a<KS>(ga, Gs + q0 * RP, RP, g, t4);
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;        // rows g and g + 8
    for (int k0 = 0; k0 < s; k0 += CH) {
      float sc[NT][4];
      gemm_rows<KS, NT>(sc, qa, Ks + k0 * RP, RP, g, t4);
      float c0 = -INFINITY, c1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) sc[nt][e] *= scale;
        c0 = fmaxf(c0, fmaxf(sc[nt][0], sc[nt][1]));
        c1 = fmaxf(c1, fmaxf(sc[nt][2], sc[nt][3]));
      }
      c0 = fmaxf(c0, __shfl_xor_sync(0xffffffffu, c0, 1)); c0 = fmaxf(c0, __shfl_xor_sync(0xffffffffu, c0, 2));
      c1 = fmaxf(c1, __shfl_xor_sync(0xffffffffu, c1, 1)); c1 = fmaxf(c1, __shfl_xor_sync(0xffffffffu, c1, 2));
      const float n0 = fmaxf(m0, c0), n1 = fmaxf(m1, c1);
      l0 *= __expf(m0 - n0); l1 *= __expf(m1 - n1);                     // m = -inf on the first chunk -> 0
      m0 = n0; m1 = n1;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        l0 += __expf(sc[nt][0] - n0) + __expf(sc[nt][1] - n0);
        l1 += __expf(sc[nt][2] - n1) + __expf(sc[nt][3] - n1);
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float lse0 = m0 + __logf(l0), lse1 = m1 + __logf(l1);
    if (t4 == 0) { lse[q0 + g] = lse0; lse[q0 + g + 8] = lse1; }
    const float d0 = dsum[q0 + g], d1 = dsum[q0 + g + 8];
    float dq[DT][4];
#pragma unroll
    for (int dt = 0; dt < DT; ++dt)
#pragma unroll
      for (int e = 0; e < 4; ++e) dq[dt][e] = 0.0f;
    for (int k0 = 0; k0 < s; k0 += CH) {
      float sc[NT][4], dp[NT][4];
      gemm_rows<KS, NT>(sc, qa, Ks + k0 * RP, RP, g, t4);
      gemm_rows<KS, NT>(dp, ga, Vs + k0 * RP, RP, g, t4);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        sc[nt][0] = __expf(sc[nt][0] * scale - lse0) * (dp[nt][0] - d0);
        sc[nt][1] = __expf(sc[nt][1] * scale - lse0) * (dp[nt][1] - d0);
        sc[nt][2] = __expf(sc[nt][2] * scale - lse1) * (dp[nt][2] - d1);
        sc[nt][3] = __expf(sc[nt][3] * scale - lse1) * (dp[nt][3] - d1);
      }
      gemm_frag<NT, DT>(dq, sc, Kt, TP, k0, g, t4);                     // dQ += dS K
    }
    __nv_bfloat16* dst = dqkv + base + static_cast<size_t>(q0 + g) * row_stride + t4 * 2;
#pragma unroll
    for (int dt = 0; dt < DT; ++dt) {
      *reinterpret_cast<uint32_t*>(dst + dt * 8) = pack_bf16x2(dq[dt][0] * scale, dq[dt][1] * scale);
      *reinterpret_cast<uint32_t*>(dst + 8 * row_stride + dt * 8) = pack_bf16x2(dq[dt][2] * scale, dq[dt][3] * scale);
    }
  }
  __syncthreads();

  // ---- phase B: key slabs -> dK, dV ----
  for (int slab = warp; slab < s / 16; slab += nwarps) {
    const int kb = slab * 16;
    uint32_t ka[KS][4], va[KS][4];
    load_a<KS>(ka, Ks + kb * RP, RP, g, t4);
    load_a<KS>(va, Vs + kb * RP, RP, g, t4);
    float dk[DT][4], dv[DT][4];
#pragma unroll
    for (int dt = 0; dt < DT; ++dt)
#pragma unroll
      for (int e = 0; e < 4; ++e) dk[dt][e] = dv[dt][e] = 0.0f;
    for (int q0 = 0; q0 < s; q0 += CH) {
      float pt[NT][4], dp[NT][4];
      gemm_rows<KS, NT>(pt, ka, Qs + q0 * RP, RP, g, t4);               // S^T: rows = keys, columns = queries
      gemm_rows<KS, NT>(dp, va, Gs + q0 * RP, RP, g, t4);               // dP^T
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int col = q0 + nt * 8 + t4 * 2;
        const float la = lse[col], lb = lse[col + 1], da = dsum[col], db = dsum[col + 1];
        pt[nt][0] = __expf(pt[nt][0] * scale - la); pt[nt][1] = __expf(pt[nt][1] * scale - lb);
        pt[nt][2] = __expf(pt[nt][2] * scale - la); pt[nt][3] = __expf(pt[nt][3] * scale - lb);
        dp[nt][0] = pt[nt][0] * (dp[nt][0] - da); dp[nt][1] = pt[nt][1] * (dp[nt][1] - db);
        dp[nt][2] = pt[nt][2] * (dp[nt][2] - da); dp[nt][3] = pt[nt][3] * (dp[nt][3] - db);
      }
      gemm_frag<NT, DT>(dv, pt, Gt, TP, q0, g, t4);                     // dV += P^T dO
      gemm_frag<NT, DT>(dk, dp, Qt, TP, q0, g, t4);                     // dK += dS^T Q
    }
    __nv_bfloat16* dst = dqkv + base + static_cast<size_t>(kb + g) * row_stride + t4 * 2;
#pragma unroll
    for (int dt = 0; dt < DT; ++dt) {
      *reinterpret_cast<uint32_t*>(dst + c + dt * 8) = pack_bf16x2(dk[dt][0] * scale, dk[dt][1] * scale);
      *reinterpret_cast<uint32_t*>(dst + c + 8 * row_stride + dt * 8) = pack_bf16x2(dk[dt][2] * scale, dk[dt][3] * scale);
      *reinterpret_cast<uint32_t*>(dst + 2 * c + dt * 8) = pack_bf16x2(dv[dt][0], dv[dt][1]);
      *reinterpret_cast<uint32_t*>(dst + 2 * c + 8 * row_stride + dt * 8) = pack_bf16x2(dv[dt][2], dv[dt][3]);
    }
  }
}

static size_t attn_bwd_smem(int s, int d) {
  return (static_cast<size_t>(4) * s * (d + 8) + static_cast<size_t>(3) * d * (s + 8)) * sizeof(__nv_bfloat16) + static_cast<size_t>(2) * s * sizeof(float);
}

template <int D, int CH>
static int launch_attn_bwd(const void* qkv, const void* out, const void* dout, void* dqkv, int b, int s, int c, int heads, cudaStream_t st) {
  const size_t smem = attn_bwd_smem(s, D);
  auto kern = attention_bwd_mma_kernel<D, CH>;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) {
      set_error("attention_backward: cannot reserve %zu bytes of shared memory", smem);
      return 1;
    }
    configured = smem;
  }
  const int warps = min(8, s / 16);
  launch_k(kern, dim3(heads, b), warps * 32, smem, st, static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(out),
           static_cast<const __nv_bfloat16*>(dout), static_cast<__nv_bfloat16*>(dqkv), s, c, 1.0f / sqrtf(static_cast<float>(D)));
  return check_launch("attention_backward");
}

// returns -1 if the shape / format is not covered (the caller falls back to the batched CUDA-core path)
int attention_bwd_mma_dispatch(const void* qkv, const void* out, const void* dout, void* dqkv, int fmt, int b, int s, int c, int heads,
                               cudaStream_t st) {
  if (fmt != SBGM_FMT_BF16 || out == nullptr || heads < 1 || c % heads != 0) return -1;
  const int d = c / heads;
  if (s % 16 != 0 || s < 16 || (d != 32 && d != 64 && d != 128) || attn_bwd_smem(s, d) > 200 * 1024) return -1;
  const bool wide = d <= 64 && s % 64 == 0;      // 64-token chunks where the registers allow, else 16
  if (d == 32) return wide ? launch_attn_bwd<32, 64>(qkv, out, dout, dqkv, b, s, c, heads, st) : launch_attn_bwd<32, 16>(qkv, out, dout, dqkv, b, s, c, heads, st);
  if (d == 64) return wide ? launch_attn_bwd<64, 64>(qkv, out, dout, dqkv, b, s, c, heads, st) : launch_attn_bwd<64, 16>(qkv, out, dout, dqkv, b, s, c, heads, st);
  return launch_attn_bwd<128, 16>(qkv, out, dout, dqkv, b, s, c, heads, st);
}

}  // namespace sbgm
