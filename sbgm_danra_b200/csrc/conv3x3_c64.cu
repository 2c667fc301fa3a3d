// conv3x3_c64.cu -- persistent tcgen05 kernel for the 64 -> 64 channel 3x3 stride-1 convolutions
// (encoder layer1, decoder block 3 and the final conv_up: 45% of the network's FLOPs, all with a
// K loop of only nine 64-wide blocks, SURVEY.md section 7 "hard parts").
//
// The generic kernel's plain mode (conv_tc.cu) re-fetches a 16 KB activation box and an 8 KB weight box for every
// tap: 24 KB per 128 MMA-cycles, 4x more than the L2->SM path sustains.  Here
//   * one CTA per SM stays resident and walks tiles blockIdx.x, +gridDim.x, ...;
//   * the 9 x 64 x 64 weights (72 KB bf16, 144 KB split-bf16) are loaded ONCE per CTA and stay in smem;
//   * a tile is 16 rows x 8 columns of one image and its input arrives ONCE, as a halo slab of 18 x 10 pixels
//     (zero-filled outside the image) in two 32-channel halves: box = 32 ch x 10 w x 18 h at (w0-1, h0-1), 64-byte
//     pixel rows with the 64-byte swizzle.  Tap (r, s) is the same slab read from pixel (r, s) on: the K-major A
//     descriptor starts at byte (r * 10 + s) * 64 and steps one slab row (640 B) per 8-row group (= one output
//     row).  Neither is a multiple of the 512-byte swizzle atom; the tensor core applies the swizzle to absolute
//     shared-memory address bits (probed on B200: odd 64-byte starts give bit-identical results, the descriptor's
//     base-offset field is ignored), so any 16-byte-aligned start / stride addresses what TMA wrote.
//     11.25 KB per plane and half, 46 KB per tile in split-bf16 instead of 120 KB for three column-shifted slabs:
//     the earlier three-slab version was bound by the L2->SM path (ncu: MMA warp on the "full" barrier 40 %);
//   * two TMEM accumulators (2 x 64 columns) ping-pong between MMA issue and the epilogue warps.
// Warp roles (320 threads): 0 = TMA producer, 1 = TMEM alloc + MMA issue, 2..5 = epilogue group 0 (even tiles,
// accumulator 0), 6..9 = epilogue group 1 (odd tiles, accumulator 1).  ncu showed the kernel epilogue-bound: one
// tile's epilogue (tcgen05.ld, bias, split-bf16 rounding, transpose, store) is ~1.4x its MMA time on a single warp
// per scheduler; two groups alternating tiles give each epilogue two tile periods.
#include "tc_common.cuh"

namespace sbgm {

constexpr int kTH = 16, kTW = 8;                        // output tile (rows x cols) = 128 pixels
constexpr int kSlabW = kTW + 2, kSlabH = kTH + 2;       // halo slab: 10 x 18 pixels
constexpr uint32_t kSlabTxBytes = kSlabH * kSlabW * 64; // 11520 per plane: one 32-channel half of the halo slab
constexpr uint32_t kSlabBytes = 12288;                  // plane pitch in smem (1024-aligned)
// K-major SWIZZLE_64B A operand whose 8-row groups are one slab row (10 pixels x 64 B) apart
constexpr uint32_t kDescHiSlab = ((kSlabW * 64u) >> 4) | (1u << 14) | (4u << 29);
constexpr uint32_t kWTapBytes = 64 * 128;               // 8192 per (tap, plane)

struct C64Params {
  int n, h, w;
  int tiles_w, tiles_h, total_tiles;
  EpilogueParams ep;
  float* gn_partials;   // [n][chunks][8][2] or nullptr (groups of 8 channels)
};

template <int FMT, int kStages>
struct C64Cfg {
  static constexpr int kAPl = TcFmt<FMT>::kAPlanes;      // activation planes (2: split-bf16)
  static constexpr int kBPl = TcFmt<FMT>::kBPlanes;      // weight planes (2: hi|lo, one N = 128 MMA covers both)
  static constexpr uint32_t kWeightBytes = 9 * kBPl * kWTapBytes;
  static constexpr uint32_t kStageBytes = kAPl * kSlabBytes;
  static constexpr uint32_t kBarOffset = kWeightBytes + kStages * kStageBytes;
  static constexpr uint32_t kSmemBytes = kBarOffset + 256 + 1024;
};

template <int FMT, int kStages, int ACT, int PROJ>
__global__ void __launch_bounds__(320, 1)
conv3x3_c64_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const C64Params p) {
  pdl_grid_sync();
  using Cfg = C64Cfg<FMT, kStages>;
  constexpr int kAPl = Cfg::kAPl, kBPl = Cfg::kBPl;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = smem_base;                              // [tap][plane][64 x 128 B]
  const uint32_t slab_base = smem_base + Cfg::kWeightBytes;       // [stage][plane][10 x 16 x 128 B]
  const uint32_t bar_base = smem_base + Cfg::kBarOffset;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  const uint32_t w_bar = bar_base + 8u * (2 * kStages);
  auto acc_full = [&](int a) { return bar_base + 8u * (2 * kStages + 1 + a); };
  auto acc_empty = [&](int a) { return bar_base + 8u * (2 * kStages + 3 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 5);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(w_bar, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(acc_full(a), 1);
      mbar_init(acc_empty(a), 4);     // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  constexpr uint32_t kAccCols = 64u * kBPl;       // two weight planes keep x*w_lo in a second 64-column half
  if (warp == 1) tmem_alloc(tmem_slot, 2 * kAccCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // weights: 9 taps x planes boxes of 64 rows x 64 k
      mbar_expect_tx(w_bar, Cfg::kWeightBytes);
      for (int tap = 0; tap < 9; ++tap)
        for (int pl = 0; pl < kBPl; ++pl)
          tma_load_3d(w_base + (tap * kBPl + pl) * kWTapBytes, &tmap_b, w_bar, tap * 64, 0, pl);
      uint32_t sidx = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h, n = tile / (p.tiles_w * p.tiles_h);
        for (int half = 0; half < 2; ++half, ++sidx) {   // unit = one 32-channel half of the tile's halo slab
          const int stage = sidx % kStages;
          const uint32_t phase = (sidx / kStages) & 1u;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), kAPl * kSlabTxBytes);
          for (int pl = 0; pl < kAPl; ++pl)
            tma_load_5d(slab_base + stage * Cfg::kStageBytes + pl * kSlabBytes, &tmap_a, full_bar(stage), half * 32,
                        tw * kTW - 1, th * kTH - 1, n, pl);
        }
      }
    }
  } else if (warp == 1) {
    // one elected lane runs the whole issue loop: ncu showed the warp-uniform variant (election + predicate vote +
    // descriptor moves to uniform registers for every MMA, ~13.5 instructions each) issue-bound, not data-bound
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(64, TcFmt<FMT>::kHalf), idesc2 = make_idesc(64 * kBPl, TcFmt<FMT>::kHalf);
      mbar_wait(w_bar, 0);
      const uint64_t w_desc = (static_cast<uint64_t>(kDescHi) << 32) | desc_lo(w_base);
      const uint64_t a_desc0 = (static_cast<uint64_t>(kDescHiSlab) << 32) | desc_lo(slab_base);
      uint32_t stage = 0, phase = 0, it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1u, use = it >> 1;
        mbar_wait(acc_empty(acc), (use & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kAccCols;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint64_t a_unit = a_desc0 + stage * (Cfg::kStageBytes >> 4);
#pragma unroll
          for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int s = 0; s < 3; ++s) {
#pragma unroll
              for (int k = 0; k < 2; ++k) {            // two K = 16 steps per 32-channel half
                // tap (r, s): the 128 A rows start at slab pixel (r, s); row groups (output rows) are one slab row apart
                const uint64_t a_d = a_unit + (((r * kSlabW + s) * 64 + k * 32) >> 4);
                const uint64_t b_d = w_desc + ((((r * 3 + s) * kBPl) * kWTapBytes + (half * 2 + k) * 32) >> 4);
                const bool first = (half | r | s | k) == 0;
                // hi|lo weight planes of a tap are adjacent in smem: one N = 128 MMA gives x*w_hi (cols 0..63) and x*w_lo
                // (cols 64..127); split-bf16 adds x_lo*w_hi into cols 0..63.  The epilogue adds the halves.
                if (first) umma_bf16_first(tmem_d, a_d, b_d, idesc2); else umma_bf16_acc(tmem_d, a_d, b_d, idesc2);
                if (kAPl == 2) umma_bf16_acc(tmem_d, a_d + (kSlabBytes >> 4), b_d, idesc);
              }
            }
          }
          umma_commit(empty_bar(stage));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(acc_full(acc));
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const uint32_t group = (warp - 2) >> 2;          // which accumulator buffer / tile parity this warp serves
    const int row = quarter * 32 + lane;
    const int w_l = row % kTW, h_l = row / kTW;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      if ((it & 1u) != group) continue;
      const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h, n = tile / (p.tiles_w * p.tiles_h);
      const int oy = th * kTH + h_l, ox = tw * kTW + w_l;
      const size_t pix = (static_cast<size_t>(n) * p.h + oy) * p.w + ox;
      const uint32_t acc = it & 1u, use = it >> 1;
      float proj_acc[kProjMax];
#pragma unroll
      for (int q = 0; q < kProjMax; ++q) proj_acc[q] = 0.0f;
      mbar_wait(acc_full(acc), use & 1u);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kAccCols;
      uint32_t r0[32], r1[32];
      tmem_ld32(taddr, r0);
      tmem_ld32(taddr + 32, r1);
      if (kBPl == 2) {
        uint32_t t[32];
        tmem_ld32(taddr + 64, t);
#pragma unroll
        for (int j = 0; j < 32; ++j) r0[j] = __float_as_uint(fmaf(__uint_as_float(t[j]), TcFmt<FMT>::kLoScale, __uint_as_float(r0[j])));
        tmem_ld32(taddr + 96, t);
#pragma unroll
        for (int j = 0; j < 32; ++j) r1[j] = __float_as_uint(fmaf(__uint_as_float(t[j]), TcFmt<FMT>::kLoScale, __uint_as_float(r1[j])));
      }
      // the accumulator is in registers: hand the TMEM buffer back before the (long) epilogue math
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(acc));
      if (p.gn_partials) {
        const int chunks = p.tiles_w * p.tiles_h * 4;
        const int chunk = (th * p.tiles_w + tw) * 4 + quarter;
        gn_block64_stats<FMT>(r0, r1, p.ep.bias, 0, true, lane, p.gn_partials + (static_cast<size_t>(n) * chunks + chunk) * 16);
      }
      epilogue_block64<FMT, ACT, PROJ>(p.ep, r0, r1, 0, n, pix, true, lane, proj_acc);
      if (PROJ) epilogue_store_proj(p.ep, pix, proj_acc);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * kAccCols);
}

template <int FMT, int kStages, int ACT, int PROJ>
static int launch_c64_inst(const CUtensorMap& ta, const CUtensorMap& tb, const C64Params& p, cudaStream_t st) {
  using Cfg = C64Cfg<FMT, kStages>;
  auto kern = conv3x3_c64_kernel<FMT, kStages, ACT, PROJ>;
  static bool configured = false;
  static int num_sms = 0;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess) {
      set_error("conv3x3_c64: cannot reserve %u bytes of shared memory", Cfg::kSmemBytes);
      return 1;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    configured = true;
  }
  if (PROJ && cudaMemcpyToSymbolAsync(c_proj_w, p.ep.proj_w, sizeof(float) * kProjN * 64, 0, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
    set_error("conv3x3_c64: projection weight upload failed");
    return 1;
  }
  const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  launch_k((kern), grid, 320, Cfg::kSmemBytes, st, ta, tb, p);
  return check_launch("conv3x3_c64");
}

template <int FMT, int kStages>
static int launch_c64(const CUtensorMap& ta, const CUtensorMap& tb, const C64Params& p, cudaStream_t st) {
  if (p.ep.proj_w) return p.ep.out ? launch_c64_inst<FMT, kStages, SBGM_ACT_NONE, 2>(ta, tb, p, st)
                                   : launch_c64_inst<FMT, kStages, SBGM_ACT_NONE, 1>(ta, tb, p, st);
  if (p.ep.act == SBGM_ACT_NONE) return launch_c64_inst<FMT, kStages, SBGM_ACT_NONE, false>(ta, tb, p, st);
  if (p.ep.act == SBGM_ACT_RELU) return launch_c64_inst<FMT, kStages, SBGM_ACT_RELU, false>(ta, tb, p, st);
  set_error("conv3x3_c64: activation %d not instantiated (none / relu only)", p.ep.act);
  return 1;
}

}  // namespace sbgm

using namespace sbgm;

extern "C" int sbgm_conv3x3_c64(const void* in, size_t in_plane, const void* weight, size_t w_plane, const float* bias,
                                const void* residual, size_t res_plane, const float* tproj, int tproj_stride,
                                void* out, size_t out_plane, int fmt, int n, int h, int w, int act,
                                const float* proj_w, int n_proj, float* proj_out, float* gn_partials, int gn_cpg,
                                void* stream) {
  SBGM_REQUIRE(fmt == SBGM_FMT_BF16 || fmt == SBGM_FMT_BF16X2 || fmt == SBGM_FMT_F16, "conv3x3_c64: format %d is not a tensor-core format", fmt);
  SBGM_REQUIRE(h % kTH == 0 && w % kTW == 0, "conv3x3_c64: h=%d must be a multiple of %d and w=%d of %d", h, kTH, w, kTW);
  SBGM_REQUIRE(proj_w == nullptr || (n_proj == kProjN && proj_out != nullptr && residual == nullptr && tproj == nullptr &&
                                     act == SBGM_ACT_NONE),
               "conv3x3_c64: the projection epilogue needs n_proj == %d and a bias-only epilogue", kProjN);
  SBGM_REQUIRE(gn_partials == nullptr || (gn_cpg == 8 && residual == nullptr && tproj == nullptr && act == SBGM_ACT_NONE),
               "conv3x3_c64: fused GroupNorm statistics need 8 channels per group and a bias-only epilogue");
  const int planes = (fmt == SBGM_FMT_BF16X2) ? 2 : 1, w_planes = (fmt == SBGM_FMT_BF16) ? 1 : 2;
  C64Params p;
  p.n = n; p.h = h; p.w = w;
  p.tiles_w = w / kTW; p.tiles_h = h / kTH; p.total_tiles = p.tiles_w * p.tiles_h * n;
  p.ep.bias = bias; p.ep.residual = residual; p.ep.res_plane = res_plane; p.ep.res_pix_mod = 0; p.ep.staged = 0; p.ep.tproj = tproj; p.ep.tproj_stride = tproj_stride;
  p.ep.act = act; p.ep.cout = 64; p.ep.out = out; p.ep.out_plane = out_plane;
  p.ep.proj_w = proj_w; p.ep.proj_out = proj_out; p.ep.n_proj = n_proj;
  p.gn_partials = gn_partials;
  CUtensorMap ta, tb;
  if (encode_act_map(&ta, in, planes, in_plane, n, h, w, 64, kSlabW, kSlabH, 1, 1, /*box_c=*/32)) return 1;
  if (encode_weight_map(&tb, weight, w_planes, w_plane, 64, 9 * 64, 64)) return 1;
  cudaStream_t st = as_stream(stream);
  if (fmt == SBGM_FMT_BF16) return launch_c64<SBGM_FMT_BF16, 8>(ta, tb, p, st);
  if (fmt == SBGM_FMT_F16) return launch_c64<SBGM_FMT_F16, 6>(ta, tb, p, st);     // 144 KB of weights + 6 x 12 KB slab halves
  return launch_c64<SBGM_FMT_BF16X2, 3>(ta, tb, p, st);
}
