// conv_tc.cu -- convolution / linear layers as implicit GEMM on the 5th-generation tensor cores.
//
//   D[128 pixels x BLOCK_N channels] (fp32, TMEM)  +=  A[128 x 64] (smem, bf16)  *  B[BLOCK_N x 64]^T (smem, bf16)
//
// * A tiles are fetched straight from the NHWC activation tensor with one 5-D TMA box per filter
//   tap: box = (64 channels, w_tile, h_tile, n_tile, 1 plane) placed at
//   (cb*64, wo0*stride + s - pad, ho0*stride + r - pad, n0, plane).  Out-of-bounds pixels are
//   zero-filled by the TMA unit, which *is* the convolution's zero padding, and strided convolutions
//   use the tensor map's element strides -- no im2col buffer, no index arithmetic on the SMs.
//   The box lands in shared memory as 128 rows x 128 B with the 128-byte swizzle, exactly the
//   K-major canonical layout tcgen05.mma consumes.
// * B tiles come from the pre-packed weight matrix [planes][cout][kh*kw*cin] (K-major).
// * One elected thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16) into a TMEM accumulator; a
//   kStages-deep mbarrier ring decouples TMA from MMA; four epilogue warps drain TMEM with
//   tcgen05.ld and apply bias / residual / activation / time-projection before the NHWC store.
// * BF16X2 ("split") activations and weights carry hi|lo bf16 planes; the three products
//   hi*hi + lo*hi + hi*lo accumulate into the same TMEM tile (fp32-class accuracy, 16-bit operands).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (warp w may only touch TMEM lanes 32*(w%4) .. +31).
#include <cuda.h>

#include "common.cuh"

namespace sbgm {

struct ConvTcParams {
  int n, ho, wo, cout;
  int kh, kw, stride, pad;
  int w_tile, h_tile, n_tile;
  int tiles_w, tiles_h;
  int cin_blocks;
  int act;
  const float* bias;
  const void* residual;
  size_t res_plane;
  const float* tproj;
  int tproj_stride;
  void* out;
  size_t out_plane;
};

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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

// K-major, 128-byte swizzle shared-memory matrix descriptor (see cute/arch/mma_sm100_desc.hpp):
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

// ---- kernel ---------------------------------------------------------------------------------
template <int kSplit, int BLOCK_N, int kStages>
struct ConvTcCfg {
  static constexpr uint32_t kABytes = 128 * 128;               // 128 rows x 64 bf16
  static constexpr uint32_t kBBytes = BLOCK_N * 128;
  static constexpr uint32_t kStageBytes = kSplit * (kABytes + kBBytes);
  static constexpr uint32_t kBarOffset = kStages * kStageBytes;
  static constexpr uint32_t kSmemBytes = kBarOffset + 256 + 1024;  // barriers + alignment slack
};

template <int FMT, int BLOCK_N, int kStages>
__global__ void __launch_bounds__(192, 2)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const ConvTcParams p) {
  constexpr int kSplit = (FMT == SBGM_FMT_BF16X2) ? 2 : 1;
  using Cfg = ConvTcCfg<kSplit, BLOCK_N, kStages>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Cfg::kBarOffset;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * kStages);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 1);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(BLOCK_N) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // tile coordinates
  const int mt = blockIdx.x;
  const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
  const int wo0 = tw * p.w_tile, ho0 = th * p.h_tile, n0 = tn * p.n_tile;
  const int co0 = blockIdx.y * BLOCK_N;
  const int num_kb = p.kh * p.kw * p.cin_blocks;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int tap = kb / p.cin_blocks, cb = kb - tap * p.cin_blocks;
        const int r = tap / p.kw, s = tap - r * p.kw;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
        const uint32_t a_dst = smem_base + stage * Cfg::kStageBytes;
        const uint32_t b_dst = a_dst + kSplit * Cfg::kABytes;
#pragma unroll
        for (int pl = 0; pl < kSplit; ++pl) {
          tma_load_5d(a_dst + pl * Cfg::kABytes, &tmap_a, full_bar(stage), cb * 64, wo0 * p.stride + s - p.pad,
                      ho0 * p.stride + r - p.pad, n0, pl);
          tma_load_3d(b_dst + pl * Cfg::kBBytes, &tmap_b, full_bar(stage), kb * 64, co0, pl);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tcgen05_fence_after();
        const uint32_t a0 = smem_base + stage * Cfg::kStageBytes;
        const uint32_t b0 = a0 + kSplit * Cfg::kABytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // 4 x (K = 16 bf16 = 32 B) per 64-element block
          const uint64_t a_hi = make_smem_desc(a0 + k * 32), b_hi = make_smem_desc(b0 + k * 32);
          umma_bf16(tmem_base, a_hi, b_hi, idesc, (kb | k) != 0);
          if (kSplit == 2) {
            const uint64_t a_lo = make_smem_desc(a0 + Cfg::kABytes + k * 32), b_lo = make_smem_desc(b0 + Cfg::kBBytes + k * 32);
            umma_bf16(tmem_base, a_lo, b_hi, idesc, 1u);
            umma_bf16(tmem_base, a_hi, b_lo, idesc, 1u);
          }
        }
        umma_commit(empty_bar(stage));          // frees the smem slot once these MMAs retire
        if (kb == num_kb - 1) umma_commit(tmem_full_bar);
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ---- epilogue: TMEM -> registers -> bias/residual/act/time -> NHWC global ----
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int w_l = row % p.w_tile, h_l = (row / p.w_tile) % p.h_tile, n_l = row / (p.w_tile * p.h_tile);
    const int n = n0 + n_l, oy = ho0 + h_l, ox = wo0 + w_l;
    const bool valid = (n < p.n) && (oy < p.ho) && (ox < p.wo);
    const size_t pix = (static_cast<size_t>(n) * p.ho + oy) * p.wo + ox;
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c0, r);
      if (valid) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int co = co0 + c0 + g * 8;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += __ldg(p.bias + co + j);
          }
          if (p.residual) {
            float rv[8];
            Act<FMT>::load8(p.residual, p.res_plane, pix * p.cout + co, rv);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += rv[j];
          }
          if (p.act != SBGM_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = apply_act(v[j], p.act);
          }
          if (p.tproj) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += __ldg(p.tproj + static_cast<size_t>(n) * p.tproj_stride + co + j);
          }
          Act<FMT>::store8(p.out, p.out_plane, pix * p.cout + co, v);
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BLOCK_N) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

static int pow2_floor(int v) { int p = 1; while (p * 2 <= v) p *= 2; return p; }
static int pow2_ceil(int v) { int p = 1; while (p < v) p *= 2; return p; }
static int pow2_divisor(int v) { return v & (-v); }

// Pick (w_tile, h_tile, n_tile), product 128, powers of two, covering the output with the least padding.
static void pick_tile(int n, int ho, int wo, int* wt, int* ht, int* nt) {
  if (ho == 1 && n == 1) {  // token matrix [M][C] (Linear layers): plain 128-row tiles, tail masked
    *wt = 128; *ht = 1; *nt = 1;
    return;
  }
  int w_tile = pow2_divisor(wo);
  if (w_tile > 128) w_tile = 128;
  if (w_tile < 8 && wo > w_tile) w_tile = pow2_ceil(wo) > 128 ? 128 : pow2_ceil(wo);
  int rest = 128 / w_tile;
  int h_tile = pow2_divisor(ho);
  if (h_tile > rest) h_tile = rest;
  if (h_tile < rest && ho > h_tile && pow2_divisor(ho) < 4) h_tile = pow2_ceil(ho) > rest ? rest : pow2_ceil(ho);
  *wt = w_tile;
  *ht = h_tile;
  *nt = rest / h_tile;
  (void)n;
  (void)pow2_floor;
}

template <int FMT, int BLOCK_N, int kStages>
static int launch_conv_tc(const CUtensorMap& ta, const CUtensorMap& tb, const ConvTcParams& p, int m_tiles, cudaStream_t st) {
  constexpr int kSplit = (FMT == SBGM_FMT_BF16X2) ? 2 : 1;
  using Cfg = ConvTcCfg<kSplit, BLOCK_N, kStages>;
  auto kern = conv_tc_kernel<FMT, BLOCK_N, kStages>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess) {
      set_error("conv2d_tc: cannot reserve %u bytes of shared memory", Cfg::kSmemBytes);
      return 1;
    }
    configured = true;
  }
  dim3 grid(m_tiles, p.cout / BLOCK_N);
  kern<<<grid, 192, Cfg::kSmemBytes, st>>>(ta, tb, p);
  return check_launch("conv2d_tc");
}

}  // namespace sbgm

using namespace sbgm;

extern "C" int sbgm_conv2d_tc(const void* in, size_t in_plane, const void* weight, size_t w_plane, const float* bias,
                              const void* residual, size_t res_plane, const float* tproj, int tproj_stride,
                              void* out, size_t out_plane, int fmt, int n, int h, int w, int cin, int cout,
                              int kh, int kw, int stride, int pad, int act, void* stream) {
  SBGM_REQUIRE(fmt == SBGM_FMT_BF16 || fmt == SBGM_FMT_BF16X2, "conv2d_tc: format %d is not a tensor-core format", fmt);
  SBGM_REQUIRE(cin % 64 == 0 && cout % 64 == 0, "conv2d_tc: cin=%d and cout=%d must be multiples of 64", cin, cout);
  SBGM_REQUIRE(stride >= 1 && stride <= 8, "conv2d_tc: stride %d unsupported", stride);
  const int ho = (h + 2 * pad - kh) / stride + 1, wo = (w + 2 * pad - kw) / stride + 1;
  SBGM_REQUIRE(ho > 0 && wo > 0, "conv2d_tc: empty output");
  EncodeTiledFn encode = get_encode_fn();
  SBGM_REQUIRE(encode != nullptr, "conv2d_tc: cuTensorMapEncodeTiled unavailable (driver too old?)");
  const int planes = (fmt == SBGM_FMT_BF16X2) ? 2 : 1;

  ConvTcParams p;
  p.n = n; p.ho = ho; p.wo = wo; p.cout = cout;
  p.kh = kh; p.kw = kw; p.stride = stride; p.pad = pad;
  pick_tile(n, ho, wo, &p.w_tile, &p.h_tile, &p.n_tile);
  p.tiles_w = ceil_div(wo, p.w_tile);
  p.tiles_h = ceil_div(ho, p.h_tile);
  const int tiles_n = ceil_div(n, p.n_tile);
  p.cin_blocks = cin / 64;
  p.act = act;
  p.bias = bias; p.residual = residual; p.res_plane = res_plane; p.tproj = tproj; p.tproj_stride = tproj_stride;
  p.out = out; p.out_plane = out_plane;
  SBGM_REQUIRE(p.w_tile * stride <= 256 && p.h_tile * stride <= 256, "conv2d_tc: TMA box too large for stride %d", stride);

  CUtensorMap ta, tb;
  {
    const cuuint64_t dims[5] = {(cuuint64_t)cin, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n, (cuuint64_t)planes};
    const cuuint64_t strides[4] = {(cuuint64_t)cin * 2, (cuuint64_t)w * cin * 2, (cuuint64_t)h * w * cin * 2,
                                   planes == 2 ? (cuuint64_t)in_plane * 2 : (cuuint64_t)n * h * w * cin * 2};
    const cuuint32_t box[5] = {64, (cuuint32_t)(p.w_tile * stride), (cuuint32_t)(p.h_tile * stride), (cuuint32_t)p.n_tile, 1};
    const cuuint32_t estr[5] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1, 1};
    CUresult r = encode(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(in), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SBGM_REQUIRE(r == CUDA_SUCCESS, "conv2d_tc: cuTensorMapEncodeTiled(activations) failed with %d", (int)r);
  }
  const int K = kh * kw * cin;
  const int block_n = (cout % 256 == 0 && planes == 1) ? 256 : (cout % 128 == 0 ? 128 : 64);
  {
    const cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)cout, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)K * 2, planes == 2 ? (cuuint64_t)w_plane * 2 : (cuuint64_t)K * cout * 2};
    const cuuint32_t box[3] = {64, (cuuint32_t)block_n, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(weight), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SBGM_REQUIRE(r == CUDA_SUCCESS, "conv2d_tc: cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
  }
  const int m_tiles = p.tiles_w * p.tiles_h * tiles_n;
  cudaStream_t st = as_stream(stream);
  if (fmt == SBGM_FMT_BF16) {
    if (block_n == 256) return launch_conv_tc<SBGM_FMT_BF16, 256, 4>(ta, tb, p, m_tiles, st);
    if (block_n == 128) return launch_conv_tc<SBGM_FMT_BF16, 128, 3>(ta, tb, p, m_tiles, st);
    return launch_conv_tc<SBGM_FMT_BF16, 64, 4>(ta, tb, p, m_tiles, st);
  }
  if (block_n == 128) return launch_conv_tc<SBGM_FMT_BF16X2, 128, 3>(ta, tb, p, m_tiles, st);
  return launch_conv_tc<SBGM_FMT_BF16X2, 64, 2>(ta, tb, p, m_tiles, st);
}
