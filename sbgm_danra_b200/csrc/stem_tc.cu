// stem_tc.cu -- Encoder.conv1 (8x8, stride 2, pad 3; sbgm/score_unet.py:206-211, :312-315) on the NOISY FIELD channel as one
// persistent tcgen05 kernel with the im2col done in shared memory.
//
// Inside a sampler the conditioning channels are step-invariant: their contribution to conv1 is computed once per call
// (engine.EncoderEngine.stem_partial) and only x -- one fp32 channel -- changes per step.  The earlier path materialised the
// 8x8 stride-2 windows of x as a 64-channel NHWC tensor in HBM (stem_im2col_kernel) and ran a 1x1 implicit GEMM over it
// (conv_tc_kernel): 34 MB written + 34 MB read + 34 MB written per evaluation for 2 GFLOP, ~55 us.  Here a CTA walks tiles of
// 8 x 16 output pixels: builder warps load the tile's 22 x 38 input patch, write its 128 x 64 window matrix straight into the
// K-major 128-byte-swizzle operand layout (16-bit, hi | lo planes in split-bf16), one elected lane issues four K = 16 MMAs
// against the resident 64 x 64 weights, and epilogue warps add the conditioning partial sums and the time projection and
// store NHWC.  HBM traffic: x once (4 B per input pixel) and the output once.
// Warp roles (288 threads): 0..3 = window builders, 4 = TMEM allocation + MMA issue, 5..8 = epilogue.
#include "tc_common.cuh"

namespace sbgm {

constexpr int kStTH = 8, kStTW = 16;                          // output tile: 8 rows x 16 columns = 128 pixels
constexpr int kStPH = 2 * kStTH + 6, kStPW = 2 * kStTW + 6;  // input patch: 22 x 38
constexpr int kStPP = kStPW + 1;                              // patch row pitch in floats

struct StemParams {
  const float* x;          // [n][h][w] fp32
  const void* partial;     // [pn][ho][wo][64] in fmt (pn == 1: broadcast over the batch) or nullptr
  size_t partial_plane;
  int partial_n;
  const float* tproj;      // [n][tproj_stride] or nullptr
  int tproj_stride;
  void* out;               // [n][ho][wo][64] in fmt
  size_t out_plane;
  int n, h, w, ho, wo, tiles_w, tiles_h, total_tiles;
};

template <int FMT>
struct StemCfg {
  static constexpr int kAPl = TcFmt<FMT>::kAPlanes, kBPl = TcFmt<FMT>::kBPlanes;
  static constexpr uint32_t kWBytes = kBPl * 64 * 128;                    // resident weights: [plane][64 co][64 k]
  static constexpr uint32_t kABytes = kAPl * 128 * 128;                   // one window matrix (all planes)
  static constexpr uint32_t kAOffset = kWBytes;
  static constexpr uint32_t kPatchOffset = kAOffset + 2 * kABytes;        // two A buffers
  static constexpr uint32_t kPatchBytes = ((kStPH * kStPP * 4 + 127) / 128) * 128;
  static constexpr uint32_t kBarOffset = kPatchOffset + 2 * kPatchBytes;  // two patch buffers
  static constexpr uint32_t kSmemBytes = kBarOffset + 128 + 1024;
  static constexpr uint32_t kAccCols = 64 * kBPl;
};

template <int FMT>
__global__ void __launch_bounds__(288, 1)
stem_x_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const StemParams p) {
  pdl_grid_sync();
  using Cfg = StemCfg<FMT>;
  constexpr int kAPl = Cfg::kAPl, kBPl = Cfg::kBPl;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t w_base = smem_base, a_base = smem_base + Cfg::kAOffset, bar_base = smem_base + Cfg::kBarOffset;
  auto a_full = [&](int b) { return bar_base + 8u * b; };
  auto a_empty = [&](int b) { return bar_base + 8u * (2 + b); };
  auto acc_full = [&](int b) { return bar_base + 8u * (4 + b); };
  auto acc_empty = [&](int b) { return bar_base + 8u * (6 + b); };
  const uint32_t w_bar = bar_base + 8u * 8, tmem_slot = bar_base + 8u * 9;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_w);
    for (int b = 0; b < 2; ++b) {
      mbar_init(a_full(b), 4);        // one arrival per builder warp
      mbar_init(a_empty(b), 1);       // MMA commit
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), 4);     // one arrival per epilogue warp
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 2 * Cfg::kAccCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
    // ---- builders: patch -> window matrix (thread = output pixel of the tile) ----
    if (threadIdx.x == 0) {
      mbar_expect_tx(w_bar, Cfg::kWBytes);
      for (int pl = 0; pl < kBPl; ++pl) tma_load_3d(w_base + pl * 64 * 128, &tmap_w, w_bar, 0, 0, pl);
    }
    const int row = threadIdx.x;                       // 0..127
    const int py = row / kStTW, px = row % kStTW;
    // A thread's patch values are requested one TILE ahead and all at once (the patch fetch is an L2 round trip that sat on the
    // builders' critical path: ncu, long-scoreboard stalls in front of the patch stores).  Volatile asm with the bounds test as
    // a multiplier: as plain predicated loads the compiler sinks them to their first use, which undoes the prefetch.
    constexpr int kPerThread = (kStPH * kStPW + 127) / 128;
    float pv[kPerThread];
    auto fetch = [&](int tile) {
      const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h, n = tile / (p.tiles_w * p.tiles_h);
      const int iy0 = 2 * th * kStTH - 3, ix0 = 2 * tw * kStTW - 3;
      const float* xs = p.x + static_cast<size_t>(n) * p.h * p.w;
#pragma unroll
      for (int k = 0; k < kPerThread; ++k) {
        const int i = min(static_cast<int>(threadIdx.x) + k * 128, kStPH * kStPW - 1);
        const int r = i / kStPW, c = i - r * kStPW;
        const int iy = iy0 + r, ix = ix0 + c;
        const bool in = iy >= 0 && iy < p.h && ix >= 0 && ix < p.w;
        const float* src = xs + static_cast<size_t>(min(max(iy, 0), p.h - 1)) * p.w + min(max(ix, 0), p.w - 1);
        float v;
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(src));
        pv[k] = in ? v : 0.0f;
      }
    };
    fetch(min(static_cast<int>(blockIdx.x), p.total_tiles - 1));
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      float* patch = reinterpret_cast<float*>(smem_gen + Cfg::kPatchOffset + buf * Cfg::kPatchBytes);
      // the patch buffer `buf` was last read two tiles ago by these same 128 threads: the named barrier below orders it
#pragma unroll
      for (int k = 0; k < kPerThread; ++k) {
        const int i = threadIdx.x + k * 128;
        if (i < kStPH * kStPW) patch[(i / kStPW) * kStPP + (i % kStPW)] = pv[k];
      }
      fetch(min(tile + static_cast<int>(gridDim.x), p.total_tiles - 1));      // in flight while this tile's window matrix is built
      asm volatile("bar.sync 1, 128;" ::: "memory");                    // patch complete (builder warps only)
      mbar_wait(a_empty(buf), ((it >> 1) & 1u) ^ 1u);
      const uint32_t a_row = a_base + buf * Cfg::kABytes + row * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j) {                    // filter row j: 8 consecutive patch values = one 16-byte chunk
        const float* src = patch + (2 * py + j) * kStPP + 2 * px;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = src[e];
        const uint32_t dst = a_row + ((j ^ (row & 7)) << 4);
        {
          const uint4 c = TcFmt<FMT>::pack8(v);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(c.x), "r"(c.y), "r"(c.z), "r"(c.w) : "memory");
        }
        if (kAPl == 2) {
          float lo[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) lo[e] = v[e] - bf16_round(v[e]);
          const uint4 c = pack_bf16x8(lo);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 128 * 128), "r"(c.x), "r"(c.y), "r"(c.z), "r"(c.w) : "memory");
        }
      }
      fence_async_shared();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(buf));
    }
  } else if (warp == 4) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(64, TcFmt<FMT>::kHalf), idesc2 = make_idesc(64 * kBPl, TcFmt<FMT>::kHalf);
      mbar_wait(w_bar, 0);
      const uint64_t w_desc = (static_cast<uint64_t>(kDescHi) << 32) | desc_lo(w_base);
      const uint64_t a_desc0 = (static_cast<uint64_t>(kDescHi) << 32) | desc_lo(a_base);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const uint32_t buf = it & 1u, use = it >> 1;
        mbar_wait(acc_empty(buf), (use & 1u) ^ 1u);
        mbar_wait(a_full(buf), use & 1u);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + buf * Cfg::kAccCols;
        const uint64_t a_d = a_desc0 + buf * (Cfg::kABytes >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k == 0) umma_bf16_first(tmem_d, a_d, w_desc, idesc2); else umma_bf16_acc(tmem_d, a_d + 2 * k, w_desc + 2 * k, idesc2);
          if (kAPl == 2) umma_bf16_acc(tmem_d, a_d + ((128 * 128) >> 4) + 2 * k, w_desc + 2 * k, idesc);
        }
        umma_commit(a_empty(buf));
        umma_commit(acc_full(buf));
      }
    }
    __syncwarp();
  } else {
    // ---- epilogue: acc + conditioning partial sums + time projection -> NHWC ----
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int py = row / kStTW, px = row % kStTW;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1u, use = it >> 1;
      const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h, n = tile / (p.tiles_w * p.tiles_h);
      const int oy = th * kStTH + py, ox = tw * kStTW + px;
      const size_t pix = (static_cast<size_t>(n) * p.ho + oy) * p.wo + ox;
      const size_t ppix = p.partial_n == 1 ? static_cast<size_t>(oy) * p.wo + ox : pix;
      // single-plane formats: the pixel's conditioning partial sums are requested BEFORE the wait for the accumulator (they do
      // not depend on it; ncu showed the L2 round trip behind the wait as long-scoreboard stalls)
      uint4 pa[8];
      if (kAPl == 1 && p.partial) {
        const uint8_t* src = static_cast<const uint8_t*>(p.partial) + ppix * 128;
#pragma unroll
        for (int g = 0; g < 8; ++g)
          asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(pa[g].x), "=r"(pa[g].y), "=r"(pa[g].z), "=r"(pa[g].w) : "l"(src + g * 16));
      }
      mbar_wait(acc_full(buf), use & 1u);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * Cfg::kAccCols;
      uint32_t r0[32], r1[32];
      tmem_ld32(taddr, r0);
      tmem_ld32(taddr + 32, r1);
      if (kBPl == 2) {
        uint32_t t[32];
        tmem_ld32(taddr + 64, t);
        merge_lo<FMT>(r0, t);
        tmem_ld32(taddr + 96, t);
        merge_lo<FMT>(r1, t);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(buf));
      const float* tp = p.tproj ? p.tproj + static_cast<size_t>(n) * p.tproj_stride : nullptr;
      // the conditioning partial sums of the whole pixel are requested before the first store: the output pointer may alias
      // them as far as the compiler knows, so a load-inside-the-loop form serialises eight L2 round trips per tile
      // (stem 49.8 -> 33.7 us per evaluation together with the batched patch loads above; prefetching both a tile ahead
      // bought nothing more: 35.2 us)
      if (p.partial) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float a[8];
          if (kAPl == 1) {
            if (TcFmt<FMT>::kHalf) unpack_f16x8(pa[g], a); else unpack_bf16x8(pa[g], a);
          } else {
            Act<FMT>::load8(p.partial, p.partial_plane, ppix * 64 + g * 8, a);
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            if (g < 4) r0[g * 8 + e] = __float_as_uint(__uint_as_float(r0[g * 8 + e]) + a[e]);
            else r1[(g - 4) * 8 + e] = __float_as_uint(__uint_as_float(r1[(g - 4) * 8 + e]) + a[e]);
          }
        }
      }
      // the row leaves through the 8-lane transposing store of the convolution kernels: every store instruction writes four
      // full 128-byte lines (the lane's own row in 16-byte pieces was 32 quarter-lines per instruction)
      float v[64];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = __uint_as_float(r0[j]);
        v[32 + j] = __uint_as_float(r1[j]);
      }
      if (tp) {
#pragma unroll
        for (int g = 0; g < 16; ++g) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(tp) + g);
          add_f32x2(v[4 * g], v[4 * g + 1], t.x, t.y);
          add_f32x2(v[4 * g + 2], v[4 * g + 3], t.z, t.w);
        }
      }
      store_block64<FMT>(p.out, p.out_plane, 64, v, pix, true, 0, lane);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 2 * Cfg::kAccCols);
}

template <int FMT>
static int launch_stem_x_tc(const CUtensorMap& tw, const StemParams& p, cudaStream_t st) {
  using Cfg = StemCfg<FMT>;
  auto kern = stem_x_tc_kernel<FMT>;
  static bool configured = false;
  static int num_sms = 0;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess) {
      set_error("stem_x_tc: cannot reserve %u bytes of shared memory", Cfg::kSmemBytes);
      return 1;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    configured = true;
  }
  const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  launch_k((kern), grid, 288, Cfg::kSmemBytes, st, tw, p);
  return check_launch("stem_x_tc");
}

}  // namespace sbgm

using namespace sbgm;

extern "C" int sbgm_stem_x_tc(const float* x, const void* w_packed, size_t w_plane, const void* partial, size_t partial_plane, int partial_n,
                              const float* tproj, int tproj_stride, void* out, size_t out_plane, int fmt, int n, int h, int w,
                              void* stream) {
  SBGM_REQUIRE(fmt == SBGM_FMT_BF16 || fmt == SBGM_FMT_BF16X2 || fmt == SBGM_FMT_F16, "stem_x_tc: format %d is not a tensor-core format", fmt);
  SBGM_REQUIRE(h % (2 * kStTH) == 0 && w % (2 * kStTW) == 0, "stem_x_tc: h=%d must be a multiple of %d and w=%d of %d", h, 2 * kStTH, w, 2 * kStTW);
  SBGM_REQUIRE(partial == nullptr || partial_n == 1 || partial_n == n, "stem_x_tc: partial_n=%d must be 1 or n=%d", partial_n, n);
  SBGM_REQUIRE(tproj == nullptr || tproj_stride % 4 == 0, "stem_x_tc: tproj rows must be 16-byte aligned");
  StemParams p;
  p.x = x; p.partial = partial; p.partial_plane = partial_plane; p.partial_n = partial_n; p.tproj = tproj; p.tproj_stride = tproj_stride;
  p.out = out; p.out_plane = out_plane; p.n = n; p.h = h; p.w = w; p.ho = h / 2; p.wo = w / 2;
  p.tiles_w = p.wo / kStTW; p.tiles_h = p.ho / kStTH; p.total_tiles = p.tiles_w * p.tiles_h * n;
  CUtensorMap tw;
  if (encode_weight_map(&tw, w_packed, fmt == SBGM_FMT_BF16 ? 1 : 2, w_plane, 64, 64, 64)) return 1;
  cudaStream_t st = as_stream(stream);
  if (fmt == SBGM_FMT_BF16) return launch_stem_x_tc<SBGM_FMT_BF16>(tw, p, st);
  if (fmt == SBGM_FMT_F16) return launch_stem_x_tc<SBGM_FMT_F16>(tw, p, st);
  return launch_stem_x_tc<SBGM_FMT_BF16X2>(tw, p, st);
}
