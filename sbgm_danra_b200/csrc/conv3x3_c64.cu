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
//
// UP mode (single-plane formats): the kernel's input is upsample2x(x) of a low-resolution tensor x -- the bilinear resize that
// opens a decoder block (upsample2x_kernel) produced inside the operand stage, so the 4x larger tensor (the final layer: 134 MB
// written and read back per evaluation in fp16x2) never exists.  The TMA warp loads the tile's 10 x 6 low-resolution patch
// (zero-filled outside the image, two buffers); five producer warps (10..14) interpolate the 18 x 10 halo slab from it with
// ATen's rule (horizontal pair first, fp32, one rounding -- upsample2x_kernel's expression -- as packed FMUL2 / FFMA2; weights
// (0, 1) / (1, 0) where a tap lies outside the image = ATen's clamped taps) and write it in the 64-byte-swizzle layout TMA
// would have produced (zero rows / columns outside the image = the convolution's padding).
// The first form of this mode also normalised the patch in the kernel (GroupNorm apply + skip + SiLU of the previous block's
// raw output, from registers prefetched a tile ahead): parity-exact but SLOWER than three launches -- 15 000 warp instructions
// per tile and SM against an issue budget of ~3 900 clocks x 4 schedulers, after the instruction-cache overflow (unrolled:
// 45 % of stall samples "no instruction"), the local-memory arrays and the sunk prefetch loads it first showed were fixed
// (profiles/r02_conv_up_fusion_notes.txt).
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

constexpr int kUpPatchH = kTH / 2 + 2, kUpPatchW = kTW / 2 + 2;          // low-resolution patch of a tile's halo slab: 10 x 6 pixels
constexpr uint32_t kUpPatchTx = kUpPatchH * kUpPatchW * 128;             // 64 channels x 16 bit per pixel
constexpr uint32_t kUpPatchBytes = 8192;                                 // buffer pitch (1024-aligned: 128-byte swizzle)
constexpr int kUpThreads = 160;                                          // five producer warps: one interpolation item per thread

template <int FMT, int kStages, bool UP = false>
struct C64Cfg {
  static constexpr int kAPl = TcFmt<FMT>::kAPlanes;      // activation planes (2: split-bf16)
  static constexpr int kBPl = TcFmt<FMT>::kBPlanes;      // weight planes (2: hi|lo, one N = 128 MMA covers both)
  static constexpr uint32_t kWeightBytes = 9 * kBPl * kWTapBytes;
  static constexpr uint32_t kStageBytes = kAPl * kSlabBytes;
  static constexpr uint32_t kPatchOffset = kWeightBytes + kStages * kStageBytes;
  static constexpr uint32_t kBarOffset = kPatchOffset + (UP ? 2 * kUpPatchBytes : 0u);
  static constexpr uint32_t kSmemBytes = kBarOffset + 256 + 1024;
  static constexpr int kThreads = 320 + (UP ? kUpThreads : 0);
};

template <int FMT, int kStages, int ACT, int PROJ, bool UP = false>
__global__ void __launch_bounds__((C64Cfg<FMT, kStages, UP>::kThreads), 1)
conv3x3_c64_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const C64Params p) {
  pdl_grid_sync();
  using Cfg = C64Cfg<FMT, kStages, UP>;
  static_assert(!UP || TcFmt<FMT>::kAPlanes == 1, "UP mode: single-plane activations only");
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
  auto patch_full = [&](int b) { return bar_base + 8u * (2 * kStages + 6 + b); };       // UP mode
  auto patch_empty = [&](int b) { return bar_base + 8u * (2 * kStages + 8 + b); };
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), UP ? kUpThreads / 32 : 1);      // UP: one arrival per producer warp
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(w_bar, 1);
    if (UP) {
      for (int b = 0; b < 2; ++b) {
        mbar_init(patch_full(b), 1);
        mbar_init(patch_empty(b), kUpThreads / 32);
      }
    }
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
      if (UP) {          // the low-resolution patch of every tile: box = 64 ch x 6 w x 10 h at (4 tw - 1, 8 th - 1)
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
          const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h, n = tile / (p.tiles_w * p.tiles_h);
          const int b = it & 1u;
          mbar_wait(patch_empty(b), ((it >> 1) & 1u) ^ 1u);
          mbar_expect_tx(patch_full(b), kUpPatchTx);
          tma_load_5d(smem_base + Cfg::kPatchOffset + b * kUpPatchBytes, &tmap_a, patch_full(b), 0, tw * (kTW / 2) - 1, th * (kTH / 2) - 1, n, 0);
        }
      }
      uint32_t sidx = 0;
      for (int tile = blockIdx.x; !UP && tile < p.total_tiles; tile += gridDim.x) {
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
  } else if (UP && warp >= 10) {
    // ---- UP mode producers: low-resolution patch -> bilinear 2x -> swizzled halo slab ----
    // 160 items = (slab column sx, channel vector, row half), one per thread.  Item bits: [0:1] 16-byte chunk, [2] sx & 1,
    // [3] 32-channel half, then sx >> 1 (5) and the row half (2): a quarter warp stores 128 contiguous bytes of one plane (two
    // neighbouring slab pixels) and reads 64 contiguous bytes of the patch.
    const int item = threadIdx.x - 320;
    const int chunk = item & 3, half = (item >> 3) & 1, rest = item >> 4;
    const int sx = 2 * (rest % 5) + ((item >> 2) & 1), rh = rest / 5, pc = sx >> 1;
    // fp32 arithmetic, two values per instruction (FMUL2 / FFMA2): eight channels = four 64-bit register pairs.  (Packed 16-bit
    // FMAs would halve the instruction count again but round the horizontal pass to 16 bits: measured 2.9e-4 -> 4.0e-4 rel-L2
    // on this layer's output in fp16x2 against a 1e-3 whole-network gate.)
    struct F8 { unsigned long long v[4]; };
    auto pair = [](float lo, float hi) {
      unsigned long long d;
      asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
      return d;
    };
    auto widen = [&](const uint4& u) {
      float f[8];
      if (TcFmt<FMT>::kHalf) unpack_f16x8(u, f); else unpack_bf16x8(u, f);
      F8 r;
#pragma unroll
      for (int i = 0; i < 4; ++i) r.v[i] = pair(f[2 * i], f[2 * i + 1]);
      return r;
    };
    // w_a * a + w_b * b
    auto blend = [](unsigned long long wa, unsigned long long wb, const F8& a, const F8& b) {
      F8 r;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        unsigned long long t;
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(wb), "l"(b.v[i]));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v[i]) : "l"(wa), "l"(a.v[i]), "l"(t));
      }
      return r;
    };
    auto narrow = [](const F8& a) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(f[2 * i]), "=f"(f[2 * i + 1]) : "l"(a.v[i]));
      return TcFmt<FMT>::pack8(f);
    };
    const unsigned long long kW75 = pair(0.75f, 0.75f), kW25 = pair(0.25f, 0.25f);
    const unsigned long long hx = (sx & 1) ? kW25 : kW75, lx = (sx & 1) ? kW75 : kW25;     // weights of the (left, right) taps
    // the patch is [10 x 6 pixels][128 B] with the 128-byte swizzle: 16-byte chunk c of pixel q sits at chunk c ^ (q & 7)
    const int vch = half * 4 + chunk;
    auto lds_patch = [&](uint32_t base, int q) {
      uint4 u;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(base + q * 128 + ((vch ^ (q & 7)) << 4)));
      return u;
    };
    auto sts = [](uint32_t dst, const uint4& c) {
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(c.x), "r"(c.y), "r"(c.z), "r"(c.w) : "memory");
    };
    uint32_t sidx = 0, it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, sidx += 2, ++it) {
      const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h;
      const uint32_t pb = it & 1u;
      const uint32_t patch = smem_base + Cfg::kPatchOffset + pb * kUpPatchBytes;
      const uint32_t s0 = sidx % kStages, s1 = (sidx + 1) % kStages;
      // A tap outside the image (zero-filled by TMA) is replaced by its neighbour inside: both taps of the pair then hold the
      // edge pixel and 0.75 x + 0.25 x = x exactly -- ATen's clamped taps, without per-row weight selects.
      const int pl = pc + ((tw == 0 && pc == 0) ? 1 : 0), pr = pc + ((tw == p.tiles_w - 1 && pc == kUpPatchW - 2) ? 0 : 1);
      const int row_lo = (th == 0 && rh == 0) ? 1 : 0, row_hi = (th == p.tiles_h - 1 && rh == 1) ? 4 : 5;   // this half's rows inside the image
      auto hrow = [&](int r, uint4& ul, uint4& ur) {
        const int q = (4 * rh + min(max(r, row_lo), row_hi)) * kUpPatchW;
        ul = lds_patch(patch, q + pl);
        ur = lds_patch(patch, q + pr);
      };
      // slab address of row b of this half: pixel q = (8 rh + b) * 10 + sx, chunk ^ ((q >> 1) & 3) = chunk ^ ((b + pc) & 3)
      const uint32_t dst0 = slab_base + (half ? s1 : s0) * Cfg::kStageBytes + (8 * rh * kSlabW + sx) * 64;
      auto dst_of = [&](int b) { return dst0 + b * (kSlabW * 64) + ((chunk ^ ((b + pc) & 3)) << 4); };
      mbar_wait(patch_full(pb), (it >> 1) & 1u);
      mbar_wait(empty_bar(s0), ((sidx / kStages) & 1u) ^ 1u);
      mbar_wait(empty_bar(s1), (((sidx + 1) / kStages) & 1u) ^ 1u);
      // six low-resolution rows (4 * rh ..), blended horizontally; slab rows sy = 8 * rh + b, b = 0..9 (the two halves overlap
      // in rows 8, 9: the first takes 8, the second 9) use the row pair (b >> 1, b >> 1 + 1) with weights (0.75, 0.25) for even
      // b and (0.25, 0.75) for odd b.  Rolled over the five row pairs (code size), the next row's loads ahead of the stores.
      uint4 ul, ur;
      hrow(0, ul, ur);
      F8 prev = blend(hx, lx, widen(ul), widen(ur));
      hrow(1, ul, ur);
#pragma unroll 1
      for (int a = 0; a < 5; ++a) {
        const F8 next = blend(hx, lx, widen(ul), widen(ur));
        hrow(a + 2, ul, ur);                                 // (a = 4: a clamped re-read, unused)
        if (!(rh == 1 && a == 0)) sts(dst_of(2 * a), narrow(blend(kW75, kW25, prev, next)));
        if (!(rh == 0 && a == 4)) sts(dst_of(2 * a + 1), narrow(blend(kW25, kW75, prev, next)));
        prev = next;
      }
      // the convolution's zero padding: slab rows / columns outside the image, written over what the loop stored there
      const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
      if (th == 0 && rh == 0) sts(dst_of(0), zero);
      if (th == p.tiles_h - 1 && rh == 1) sts(dst_of(9), zero);
      if ((sx == 0 && tw == 0) || (sx == kSlabW - 1 && tw == p.tiles_w - 1)) {
#pragma unroll 1
        for (int b = rh; b < 9 + rh; ++b) sts(dst_of(b), zero);
      }
      fence_async_shared();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(full_bar(s0));
        mbar_arrive(full_bar(s1));
        mbar_arrive(patch_empty(pb));
      }
    }
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
        merge_lo<FMT>(r0, t);
        tmem_ld32(taddr + 96, t);
        merge_lo<FMT>(r1, t);
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

template <int FMT, int kStages, int ACT, int PROJ, bool UP = false>
static int launch_c64_inst(const CUtensorMap& ta, const CUtensorMap& tb, const C64Params& p, cudaStream_t st) {
  using Cfg = C64Cfg<FMT, kStages, UP>;
  static_assert(Cfg::kSmemBytes <= 232448, "shared memory budget");
  auto kern = conv3x3_c64_kernel<FMT, kStages, ACT, PROJ, UP>;
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
  launch_k((kern), grid, Cfg::kThreads, Cfg::kSmemBytes, st, ta, tb, p);
  return check_launch("conv3x3_c64");
}

template <int FMT, int kStages, bool UP = false>
static int launch_c64(const CUtensorMap& ta, const CUtensorMap& tb, const C64Params& p, cudaStream_t st) {
  if (p.ep.proj_w) {
    if (!UP && p.ep.out) return launch_c64_inst<FMT, kStages, SBGM_ACT_NONE, 2>(ta, tb, p, st);
    if (p.ep.out) { set_error("conv3x3_c64_up: the projection epilogue keeps no convolution output"); return 1; }
    return launch_c64_inst<FMT, kStages, SBGM_ACT_NONE, 1, UP>(ta, tb, p, st);
  }
  if (p.ep.act == SBGM_ACT_NONE) return launch_c64_inst<FMT, kStages, SBGM_ACT_NONE, false, UP>(ta, tb, p, st);
  if (!UP && p.ep.act == SBGM_ACT_RELU) return launch_c64_inst<FMT, kStages, SBGM_ACT_RELU, false>(ta, tb, p, st);
  set_error("conv3x3_c64: activation %d not instantiated (none / relu only; fused-upsample form: none)", p.ep.act);
  return 1;
}

}  // namespace sbgm

using namespace sbgm;

static int conv3x3_c64_impl(const void* in, size_t in_plane, bool up, const void* weight, size_t w_plane, const float* bias,
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
  if (encode_weight_map(&tb, weight, w_planes, w_plane, 64, 9 * 64, 64)) return 1;
  cudaStream_t st = as_stream(stream);
  if (up) {        // the activation map describes the LOW-resolution tensor: box = 64 ch x 6 x 10 pixels, 128-byte swizzle
    SBGM_REQUIRE(fmt != SBGM_FMT_BF16X2, "conv3x3_c64_up: single-plane formats only (bf16, fp16x2)");
    if (encode_act_map(&ta, in, 1, in_plane, n, h / 2, w / 2, 64, kUpPatchW, kUpPatchH, 1, 1)) return 1;
    if (fmt == SBGM_FMT_BF16) return launch_c64<SBGM_FMT_BF16, 8, true>(ta, tb, p, st);
    return launch_c64<SBGM_FMT_F16, 5, true>(ta, tb, p, st);    // 144 KB of weights + 5 x 12 KB slab halves + 2 x 8 KB patches
  }
  if (encode_act_map(&ta, in, planes, in_plane, n, h, w, 64, kSlabW, kSlabH, 1, 1, /*box_c=*/32)) return 1;
  if (fmt == SBGM_FMT_BF16) return launch_c64<SBGM_FMT_BF16, 8>(ta, tb, p, st);
  if (fmt == SBGM_FMT_F16) return launch_c64<SBGM_FMT_F16, 6>(ta, tb, p, st);     // 144 KB of weights + 6 x 12 KB slab halves
  return launch_c64<SBGM_FMT_BF16X2, 3>(ta, tb, p, st);
}

extern "C" int sbgm_conv3x3_c64(const void* in, size_t in_plane, const void* weight, size_t w_plane, const float* bias,
                                const void* residual, size_t res_plane, const float* tproj, int tproj_stride,
                                void* out, size_t out_plane, int fmt, int n, int h, int w, int act,
                                const float* proj_w, int n_proj, float* proj_out, float* gn_partials, int gn_cpg,
                                void* stream) {
  return conv3x3_c64_impl(in, in_plane, false, weight, w_plane, bias, residual, res_plane, tproj, tproj_stride, out, out_plane, fmt, n, h, w,
                          act, proj_w, n_proj, proj_out, gn_partials, gn_cpg, stream);
}

// The same convolution over upsample2x(x), x = [n][h/2][w/2][64]: the bilinear F.interpolate(scale_factor=2) that opens a decoder
// block and the final layer (score_unet.py:560-570, :652-659) happens inside the operand stage.  h, w: the OUTPUT (high-resolution)
// size.  bf16 and fp16x2 only; bias-only epilogue, optionally with the projection or the GroupNorm statistics.
extern "C" int sbgm_conv3x3_c64_up(const void* x, size_t x_plane, const void* weight, size_t w_plane, const float* bias, void* out,
                                   size_t out_plane, int fmt, int n, int h, int w, int act, const float* proj_w, int n_proj,
                                   float* proj_out, float* gn_partials, int gn_cpg, void* stream) {
  SBGM_REQUIRE(x != nullptr, "conv3x3_c64_up: x is required");
  return conv3x3_c64_impl(x, x_plane, true, weight, w_plane, bias, nullptr, 0, nullptr, 0, out, out_plane, fmt, n, h, w, act, proj_w, n_proj,
                          proj_out, gn_partials, gn_cpg, stream);
}
