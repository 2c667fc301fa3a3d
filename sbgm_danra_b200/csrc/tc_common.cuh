// tc_common.cuh -- PTX wrappers (mbarrier, TMA, tcgen05, TMEM) and the shared epilogue of the
// tensor-core convolution kernels (conv_tc.cu, conv3x3_c64.cu).  sm_100a only.
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace sbgm {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a TMA fault or a descriptor bug would otherwise hang the GPU until the watchdog.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp):
//   [0,14) start >> 4 | [16,30) LBO >> 4 (unused with swizzle: 1) | [32,46) SBO >> 4 (8 rows x 128 B = 1024)
//   [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr) {
  return static_cast<uint64_t>((addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128, N = n.
__device__ __forceinline__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem_addr, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem_addr), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- epilogue ----------------------------------------------------------------------------------
// v = acc + bias[co];  v += residual[pix][co];  v = act(v);  v += tproj[n][co];  store (NHWC, fmt)
// Projection mode (cout == 64 tiles, one thread holds a whole pixel): instead of storing the 64
// channels, accumulate kProjN dot products  proj[q] += sum_c v[c] * c_proj_w[q][c]  -- the nine taps of
// the final 64 -> 1 convolution (Decoder.final_layer.conv, score_unet.py:713-730) -- and store them
// as proj_out[pix][SBGM_PROJ_STRIDE] fp32.  The 9 x 64 weights live in constant memory so every FFMA
// takes its weight as a constant-bank operand (no loads, warp-uniform address).
struct EpilogueParams {
  const float* bias;
  const void* residual;
  size_t res_plane;
  const float* tproj;
  int tproj_stride;
  int act;
  int cout;
  void* out;
  size_t out_plane;
  const float* proj_w;   // [kProjN][64] fp32 (device) or nullptr; copied to c_proj_w before the launch
  float* proj_out;       // [pix][SBGM_PROJ_STRIDE] fp32
  int n_proj;
};
constexpr int kProjN = 9;
constexpr int kProjMax = SBGM_PROJ_STRIDE;
static __constant__ float c_proj_w[kProjN * 64];

template <int ACT>
__device__ __forceinline__ float act_ct(float x) {
  if (ACT == SBGM_ACT_RELU) return fmaxf(x, 0.0f);
  if (ACT == SBGM_ACT_SILU) return silu(x);
  if (ACT == SBGM_ACT_GELU) return gelu_erf(x);
  return x;
}

// One 32-column chunk of one accumulator row.  ACT and PROJ are compile-time; COL0 is the chunk's first
// channel inside a 64-wide tile (only used by the projection).
template <int FMT, int ACT, bool PROJ, int COL0>
__device__ __forceinline__ void epilogue_chunk(const EpilogueParams& ep, const uint32_t (&r)[32], int co_base, int n,
                                               size_t pix, float (&proj_acc)[kProjMax]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int co = co_base + g * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
    if (ep.bias) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + co)), b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + co + 4));
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if (!PROJ && ep.residual) {
      float rv[8];
      Act<FMT>::load8(ep.residual, ep.res_plane, pix * ep.cout + co, rv);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += rv[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = act_ct<ACT>(v[j]);
    if (!PROJ && ep.tproj) {
      const float* tp = ep.tproj + static_cast<size_t>(n) * ep.tproj_stride + co;
      const float4 t0 = __ldg(reinterpret_cast<const float4*>(tp)), t1 = __ldg(reinterpret_cast<const float4*>(tp + 4));
      v[0] += t0.x; v[1] += t0.y; v[2] += t0.z; v[3] += t0.w; v[4] += t1.x; v[5] += t1.y; v[6] += t1.z; v[7] += t1.w;
    }
    if (PROJ) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int q = 0; q < kProjN; ++q) proj_acc[q] = fmaf(v[j], c_proj_w[q * 64 + COL0 + g * 8 + j], proj_acc[q]);
    } else {
      Act<FMT>::store8(ep.out, ep.out_plane, pix * ep.cout + co, v);
    }
  }
}

__device__ __forceinline__ void epilogue_store_proj(const EpilogueParams& ep, size_t pix, const float (&proj_acc)[kProjMax]) {
  float4* dst = reinterpret_cast<float4*>(ep.proj_out + pix * kProjMax);
  dst[0] = make_float4(proj_acc[0], proj_acc[1], proj_acc[2], proj_acc[3]);
  dst[1] = make_float4(proj_acc[4], proj_acc[5], proj_acc[6], proj_acc[7]);
  dst[2] = make_float4(proj_acc[8], proj_acc[9], proj_acc[10], proj_acc[11]);
}

// Runtime activation id -> compile-time template argument.
#define SBGM_DISPATCH_ACT(act, ...)                                                              \
  switch (act) {                                                                                  \
    case SBGM_ACT_NONE: { constexpr int ACT = SBGM_ACT_NONE; __VA_ARGS__; break; }                \
    case SBGM_ACT_RELU: { constexpr int ACT = SBGM_ACT_RELU; __VA_ARGS__; break; }                \
    case SBGM_ACT_GELU: { constexpr int ACT = SBGM_ACT_GELU; __VA_ARGS__; break; }                \
    case SBGM_ACT_SILU: { constexpr int ACT = SBGM_ACT_SILU; __VA_ARGS__; break; }                \
    default: ::sbgm::set_error("unknown activation %d", act); return 1;                            \
  }

// ---- host: tensor-map encoding -----------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn();
// NHWC activation [planes][n][h][w][c] bf16 as a 5-D map, box = (64, bw, bh, bn, 1), element strides (1, s, s, 1, 1)
int encode_act_map(CUtensorMap* map, const void* base, int planes, size_t plane_elems, int n, int h, int w, int c,
                   int box_w, int box_h, int box_n, int stride);
// packed weights [planes][cout][K] bf16 as a 3-D map, box = (64, box_rows, 1)
int encode_weight_map(CUtensorMap* map, const void* base, int planes, size_t plane_elems, int cout, int K, int box_rows);

}  // namespace sbgm
