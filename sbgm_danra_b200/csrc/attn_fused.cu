// attn_fused.cu -- the attention half of ImageSelfAttention (sbgm/score_unet.py:136-143) as ONE tcgen05 kernel:
//
//     h = x + out_proj( concat_heads( softmax(Q_h K_h^T / sqrt(d)) V_h ) )        Q | K | V = rows of the packed in_proj output
//
// for the single-plane tensor-core formats (float16 = the fp16x2 mode, bfloat16).  It replaces two launches (the mma.sync
// attention core of attention_mma.cu and the out-projection GEMM of conv_tc.cu) and the [tokens][C] round trip between them.
//
// One CTA owns a tile of 128 query tokens.  Maps up to 8 x 8 pixels (S <= 128 keys per image, S | 128) pack 128 / S whole
// images into a tile and attention is block-diagonal inside it (the off-diagonal scores are masked, a 128 x 128 score tile
// is one instruction); a 16 x 16 map (S = 256) gives every image two query tiles that read all 256 keys.  Per 64-channel
// block of the head dimension (one head of d = 64, or two heads of d = 32 that share the loaded tiles):
//   TMA    Q [128 x 64], K [KT x 64], V [KT x 64] boxes of the qkv tensor -> shared memory (128-byte swizzle rows)
//   MMA    S = Q K^T                      M = 128, N = KT, K = d       A, B K-major            -> TMEM columns [0, KT)
//   warps  row-wise softmax on the accumulator (thread = query row = TMEM lane): max pass, exp pass; the un-normalised
//          probabilities go back to shared memory as a K-major operand tile P [128 x KT], the row sums stay in registers
//   MMA    O = P V                        M = 128, N = 64, K = KT      A K-major, B = V as loaded (MN-major)  -> TMEM [KT, KT + 64)
//   warps  O / rowsum -> 16-bit -> the head's columns of the out-projection's A operand [128 x C] in shared memory
// then   MMA  acc = O_all W_out^T      M = 128, N = C (x 2 weight planes in float16: hi | lo * 2^11), K = C, weights streamed by TMA
//        warps  h = acc + bias + x  -> global.
// Warp roles (192 threads): 0 = TMA producer, 1 = TMEM allocation + MMA issue (one elected lane), 2..5 = softmax / epilogue.
#include "tc_common.cuh"

namespace sbgm {

struct AttnFusedParams {
  int tokens;          // B * S
  int s;               // keys per image
  int c, d;            // channels, head dimension (32 or 64)
  float scale_log2e;   // log2(e) / sqrt(d)
  const void* x;       // residual [tokens][c]
  void* out;           // [tokens][c]
  const float* bias;   // out_proj bias [c]
};

constexpr uint32_t kRowBytes = 128;                  // 64 sixteen-bit elements
constexpr uint32_t kTile128 = 128 * kRowBytes;       // 16 KB: a [128 x 64] operand block

// MN-major B operand over a [rows = K][64 = N] tile as TMA wrote it (128-byte swizzle rows): K steps of 16 rows are 2048 B apart
__device__ __forceinline__ constexpr uint32_t idesc_b_mn(uint32_t idesc) { return idesc | (1u << 16); }

template <int FMT, int KT, int C>
struct AttnCfg {
  static constexpr int kWPl = TcFmt<FMT>::kBPlanes;              // weight planes of the out-projection
  static constexpr uint32_t kQBytes = kTile128;
  static constexpr uint32_t kKBytes = KT * kRowBytes;
  static constexpr uint32_t kVBytes = KT * kRowBytes;
  static constexpr uint32_t kQkvBytes = kQBytes + kKBytes + kVBytes;
  static constexpr uint32_t kPBytes = (KT / 64) * kTile128;
  static constexpr uint32_t kOBytes = (C / 64) * kTile128;
  static constexpr int kQkvStages = (2 * kQkvBytes + kPBytes + kOBytes + 2048 <= 227 * 1024) ? 2 : 1;
  static constexpr uint32_t kCoreBytes = kQkvStages * kQkvBytes + kPBytes;     // dead once the last P V product has retired
  static constexpr int kBN = 128;                                                // out-projection columns per weight stage
  static constexpr uint32_t kWStageBytes = kWPl * kBN * kRowBytes;
  static constexpr int kWStages = (kCoreBytes / kWStageBytes) < 4 ? (kCoreBytes / kWStageBytes) : 4;
  static constexpr uint32_t kOOffset = kCoreBytes;
  static constexpr uint32_t kBarOffset = kOOffset + kOBytes;
  static constexpr uint32_t kSmemBytes = kBarOffset + 256 + 1024;
  static constexpr uint32_t kTmemCols = 512;
  static_assert(kWStages >= 1, "no room for a weight stage");
  static_assert(C * kWPl <= 512 && KT + 64 <= 512, "accumulators exceed the 512 TMEM columns");
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

template <int FMT, int KT, int C>
__global__ void __launch_bounds__(192, 1)
attn_fused_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_w, const AttnFusedParams p) {
  pdl_grid_sync();
  using Cfg = AttnCfg<FMT, KT, C>;
  constexpr int kStg = Cfg::kQkvStages, kWStg = Cfg::kWStages, kWPl = Cfg::kWPl;
  constexpr bool kHalf = TcFmt<FMT>::kHalf;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto q_at = [&](int st) { return smem_base + st * Cfg::kQkvBytes; };
  auto k_at = [&](int st) { return q_at(st) + Cfg::kQBytes; };
  auto v_at = [&](int st) { return k_at(st) + Cfg::kKBytes; };
  const uint32_t p_base = smem_base + kStg * Cfg::kQkvBytes;
  const uint32_t o_base = smem_base + Cfg::kOOffset;
  const uint32_t w_base = smem_base;                                   // the weight ring re-uses the core's tiles
  const uint32_t bar_base = smem_base + Cfg::kBarOffset;
  auto qkv_full = [&](int st) { return bar_base + 8u * st; };
  auto qkv_empty = [&](int st) { return bar_base + 8u * (2 + st); };
  const uint32_t s_full = bar_base + 8u * 4, p_full = bar_base + 8u * 5, o_full = bar_base + 8u * 6, o_empty = bar_base + 8u * 7;
  const uint32_t oall_full = bar_base + 8u * 8, core_done = bar_base + 8u * 9, acc_full = bar_base + 8u * 10;
  auto w_full = [&](int st) { return bar_base + 8u * (11 + st); };
  auto w_empty = [&](int st) { return bar_base + 8u * (15 + st); };
  const uint32_t tmem_slot = bar_base + 8u * 19;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_w);
    for (int st = 0; st < 2; ++st) { mbar_init(qkv_full(st), 1); mbar_init(qkv_empty(st), 1); }
    mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_full, 1); mbar_init(o_empty, 4);
    mbar_init(oall_full, 4); mbar_init(core_done, 1); mbar_init(acc_full, 1);
    for (int st = 0; st < 4; ++st) { mbar_init(w_full(st), 1); mbar_init(w_empty(st), 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int q0 = blockIdx.x * 128;                                      // first query token of the tile
  const int key0 = (KT == 256) ? (q0 / 256) * 256 : q0;                 // first key token
  const int nblk = C / 64;                                              // 64-channel blocks of the head dimension
  const int hpb = 64 / p.d;                                             // heads per block (1 or 2)
  constexpr int kNChunks = C / Cfg::kBN;
  constexpr int kKb = C / 64;

  if (warp == 0) {
    if (lane == 0) {
      for (int cb = 0; cb < nblk; ++cb) {
        const int st = cb % kStg;
        const uint32_t use = cb / kStg;
        mbar_wait(qkv_empty(st), (use & 1u) ^ 1u);
        mbar_expect_tx(qkv_full(st), Cfg::kQkvBytes);
        tma_load_3d(q_at(st), &tmap_qkv, qkv_full(st), cb * 64, q0, 0);
#pragma unroll
        for (int r0 = 0; r0 < KT; r0 += 128) {                          // TMA boxes hold at most 256 rows: load in 128-row halves
          tma_load_3d(k_at(st) + r0 * kRowBytes, &tmap_qkv, qkv_full(st), C + cb * 64, key0 + r0, 0);
          tma_load_3d(v_at(st) + r0 * kRowBytes, &tmap_qkv, qkv_full(st), 2 * C + cb * 64, key0 + r0, 0);
        }
      }
      // out-projection weights: [kBN rows x 64 k] boxes per plane, ring over (n chunk, k block); the ring aliases the core's tiles
      mbar_wait(core_done, 0);
      uint32_t stage = 0, phase = 0;
      for (int nc = 0; nc < kNChunks; ++nc) {
        for (int kb = 0; kb < kKb; ++kb) {
          mbar_wait(w_empty(stage), phase ^ 1u);
          mbar_expect_tx(w_full(stage), Cfg::kWStageBytes);
#pragma unroll
          for (int pl = 0; pl < kWPl; ++pl)
            tma_load_3d(w_base + stage * Cfg::kWStageBytes + pl * Cfg::kBN * kRowBytes, &tmap_w, w_full(stage), kb * 64, nc * Cfg::kBN, pl);
          if (++stage == kWStg) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc(KT, kHalf);
      constexpr uint32_t idesc_o = idesc_b_mn(make_idesc(64, kHalf));
      constexpr uint32_t idesc_w = make_idesc(kWPl * Cfg::kBN, kHalf);
      auto kdesc = [&](uint32_t addr) { return (static_cast<uint64_t>(kDescHi) << 32) | desc_lo(addr); };
      const int ksteps = p.d / 16;
      uint32_t it = 0;                                                  // head counter: parity of the per-head barriers
      for (int cb = 0; cb < nblk; ++cb) {
        const int st = cb % kStg;
        mbar_wait(qkv_full(st), (cb / kStg) & 1u);
        tcgen05_fence_after();
        for (int hh = 0; hh < hpb; ++hh, ++it) {
          // S = Q_h K_h^T : the head's d channels are k-steps [hh * ksteps, +ksteps) of the 64-wide tiles
          const uint64_t qd = kdesc(q_at(st)) + 2u * (hh * ksteps), kd = kdesc(k_at(st)) + 2u * (hh * ksteps);
          for (int k = 0; k < ksteps; ++k) {
            if (k == 0) umma_bf16_first(tmem_base, qd, kd, idesc_s);
            else umma_bf16_acc(tmem_base, qd + 2u * k, kd + 2u * k, idesc_s);
          }
          umma_commit(s_full);
          // O_h = P V : wait for the probabilities in shared memory and for the previous head's O to have been read out
          mbar_wait(p_full, it & 1u);
          mbar_wait(o_empty, (it & 1u) ^ 1u);
          tcgen05_fence_after();
          const uint64_t vd = kdesc(v_at(st));
#pragma unroll 1
          for (int kk = 0; kk < KT / 16; ++kk) {
            const uint64_t pd = kdesc(p_base + (kk >> 2) * kTile128) + 2u * (kk & 3);
            if (kk == 0) umma_bf16_first(tmem_base + KT, pd, vd, idesc_o);
            else umma_bf16_acc(tmem_base + KT, pd, vd + kk * 128u, idesc_o);
          }
          umma_commit(o_full);
        }
        umma_commit(qkv_empty(st));                                     // Q / K / V of this block (and P) are free once these MMAs retire
      }
      umma_commit(core_done);
      // out-projection
      mbar_wait(oall_full, 0);
      tcgen05_fence_after();
      uint32_t stage = 0, phase = 0;
      for (int nc = 0; nc < kNChunks; ++nc) {
        const uint32_t acc = tmem_base + nc * (kWPl * Cfg::kBN);
        for (int kb = 0; kb < kKb; ++kb) {
          mbar_wait(w_full(stage), phase);
          tcgen05_fence_after();
          const uint64_t ad = kdesc(o_base + kb * kTile128), bd = kdesc(w_base + stage * Cfg::kWStageBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (kb == 0 && k == 0) umma_bf16_first(acc, ad, bd, idesc_w);
            else umma_bf16_acc(acc, ad + 2u * k, bd + 2u * k, idesc_w);
          }
          umma_commit(w_empty(stage));
          if (++stage == kWStg) { stage = 0; phase ^= 1u; }
        }
      }
      umma_commit(acc_full);
    }
    __syncwarp();
  } else {
    // ---- softmax / epilogue warps: thread = query row = TMEM lane ----
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int tok = q0 + row;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    // keys this row may attend: its own image's block of the tile (all KT keys when the image spans the tile)
    const int kbeg = (KT == 256) ? 0 : (row / p.s) * p.s;
    const int kend = (KT == 256) ? KT : kbeg + p.s;
    // the same for the warp's 32 rows together: which 32-column chunks the warp loads at all (tcgen05.ld is warp-collective)
    const int wbeg = (KT == 256) ? 0 : ((quarter * 32) / p.s) * p.s;
    const int wend = (KT == 256) ? KT : ((quarter * 32 + 31) / p.s) * p.s + p.s;
    uint32_t it = 0;
    for (int cb = 0; cb < nblk; ++cb) {
      for (int hh = 0; hh < hpb; ++hh, ++it) {
        mbar_wait(s_full, it & 1u);
        tcgen05_fence_after();
        float mx = -INFINITY;
#pragma unroll 1
        for (int c0 = 0; c0 < KT; c0 += 32) {
          if (c0 + 32 <= wbeg || c0 >= wend) continue;                  // warp-uniform
          uint32_t r[32];
          tmem_ld32(lane_addr + c0, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = c0 + j;
            if (col >= kbeg && col < kend) mx = fmaxf(mx, __uint_as_float(r[j]));
          }
        }
        float sum = 0.0f;
        const float mxs = mx * p.scale_log2e;
#pragma unroll 1
        for (int c0 = 0; c0 < KT; c0 += 32) {
          uint32_t r[32];
          const bool any = !(c0 + 32 <= wbeg || c0 >= wend);             // warp-uniform
          if (any) tmem_ld32(lane_addr + c0, r);
          // four 16-byte chunks (8 probabilities each) of this row's P tile, k-block c0 / 64
          const uint32_t prow = p_base + (c0 >> 6) * kTile128 + row * kRowBytes;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            float pv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int col = c0 + ch * 8 + e;
              float v = 0.0f;
              if (any && col >= kbeg && col < kend) v = exp2f(fmaf(__uint_as_float(r[ch * 8 + e]), p.scale_log2e, -mxs));
              pv[e] = TcFmt<FMT>::round(v);                             // the sum runs over what the tensor core will multiply
              sum += pv[e];
            }
            const uint4 pk = TcFmt<FMT>::pack8(pv);
            const uint32_t chunk = static_cast<uint32_t>(((c0 & 63) >> 3) + ch);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + ((chunk ^ (row & 7)) << 4)), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
          }
        }
        tcgen05_fence_before();
        fence_async_shared();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        // O_h: [128 x 64] at TMEM columns [KT, KT + 64); this head's d columns start at hh * d
        mbar_wait(o_full, it & 1u);
        tcgen05_fence_after();
        const float inv = 1.0f / sum;
        const uint32_t orow = o_base + cb * kTile128 + row * kRowBytes;
#pragma unroll 1
        for (int c0 = 0; c0 < p.d; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(lane_addr + KT + hh * p.d + c0, r);
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            float ov[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) ov[e] = __uint_as_float(r[ch * 8 + e]) * inv;
            const uint4 pk = TcFmt<FMT>::pack8(ov);
            const uint32_t chunk = static_cast<uint32_t>(((hh * p.d + c0) >> 3) + ch);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(orow + ((chunk ^ (row & 7)) << 4)), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty);
      }
    }
    fence_async_shared();
    __syncwarp();
    if (lane == 0) mbar_arrive(oall_full);
    // ---- out-projection epilogue: h = acc (+ lo plane * 2^-11) + bias + x ----
    mbar_wait(acc_full, 0);
    tcgen05_fence_after();
    const bool valid = tok < p.tokens;
#pragma unroll 1
    for (int c0 = 0; c0 < C; c0 += 32) {
      const int nc = c0 / Cfg::kBN, cin = c0 % Cfg::kBN;
      const uint32_t taddr = lane_addr + nc * (kWPl * Cfg::kBN) + cin;
      uint32_t r[32];
      tmem_ld32(taddr, r);
      if (kWPl == 2) {
        uint32_t t[32];
        tmem_ld32(taddr + Cfg::kBN, t);
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(fmaf(__uint_as_float(t[j]), TcFmt<FMT>::kLoScale, __uint_as_float(r[j])));
      }
      if (valid) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float v[8], xr[8];
          const size_t idx = static_cast<size_t>(tok) * C + c0 + g * 8;
          Act<FMT>::load8(p.x, 0, idx, xr);
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + g * 8)), b1 = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + g * 8) + 1);
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[g * 8 + e]) + bb[e] + xr[e];
          Act<FMT>::store8(p.out, 0, idx, v);
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

template <int FMT, int KT, int C>
static int launch_attn_fused(const CUtensorMap& tq, const CUtensorMap& tw, const AttnFusedParams& p, cudaStream_t st) {
  using Cfg = AttnCfg<FMT, KT, C>;
  auto kern = attn_fused_kernel<FMT, KT, C>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess) {
      set_error("attention_block_core: cannot reserve %u bytes of shared memory", Cfg::kSmemBytes);
      return 1;
    }
    configured = true;
  }
  launch_k((kern), ceil_div(p.tokens, 128), 192, Cfg::kSmemBytes, st, tq, tw, p);
  return check_launch("attention_block_core");
}

}  // namespace sbgm

using namespace sbgm;

// 1 if sbgm_attention_out_proj serves this problem (else the caller runs sbgm_attention + the out-projection GEMM)
extern "C" int sbgm_attention_out_proj_supported(int fmt, int b, int s, int c, int heads) {
  if (fmt != SBGM_FMT_F16 && fmt != SBGM_FMT_BF16) return 0;
  if (heads < 1 || c % heads != 0) return 0;
  const int d = c / heads;
  if (d != 32 && d != 64) return 0;
  if (c != 128 && c != 256) return 0;
  if (s == 256) return 1;
  return (s >= 1 && s <= 128 && 128 % s == 0 && b >= 1) ? 1 : 0;
}

extern "C" int sbgm_attention_out_proj(const void* qkv, const void* x, const void* w_out, size_t w_plane, const float* bias, void* out,
                                       int fmt, int b, int s, int c, int heads, void* stream) {
  SBGM_REQUIRE(sbgm_attention_out_proj_supported(fmt, b, s, c, heads), "attention_out_proj: unsupported problem fmt=%d s=%d c=%d heads=%d", fmt,
               s, c, heads);
  AttnFusedParams p;
  p.tokens = b * s; p.s = s; p.c = c; p.d = c / heads;
  p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(p.d));
  p.x = x; p.out = out; p.bias = bias;
  const int kt = (s == 256) ? 256 : 128;
  CUtensorMap tq, tw;
  // qkv [tokens][3c] as a (k = 3c, rows = tokens) matrix: boxes of 64 columns x 128 rows, 128-byte swizzle
  if (encode_weight_map(&tq, qkv, 1, 0, p.tokens, 3 * c, 128)) return 1;
  if (encode_weight_map(&tw, w_out, fmt == SBGM_FMT_F16 ? 2 : 1, w_plane, c, c, 128)) return 1;
  cudaStream_t st = as_stream(stream);
#define SBGM_AF(F, KTV, CV) return launch_attn_fused<F, KTV, CV>(tq, tw, p, st)
  if (fmt == SBGM_FMT_F16) {
    if (kt == 128 && c == 128) SBGM_AF(SBGM_FMT_F16, 128, 128);
    if (kt == 128 && c == 256) SBGM_AF(SBGM_FMT_F16, 128, 256);
    if (kt == 256 && c == 128) SBGM_AF(SBGM_FMT_F16, 256, 128);
    if (kt == 256 && c == 256) SBGM_AF(SBGM_FMT_F16, 256, 256);
  } else {
    if (kt == 128 && c == 128) SBGM_AF(SBGM_FMT_BF16, 128, 128);
    if (kt == 128 && c == 256) SBGM_AF(SBGM_FMT_BF16, 128, 256);
    if (kt == 256 && c == 128) SBGM_AF(SBGM_FMT_BF16, 256, 128);
    if (kt == 256 && c == 256) SBGM_AF(SBGM_FMT_BF16, 256, 256);
  }
#undef SBGM_AF
  set_error("attention_out_proj: no instantiation");
  return 1;
}
