// pack.cu -- parameter packing for the tensor-core kernels, one launch per weight.
// A training step re-packs all 19 M parameters after every optimizer update (forward layout + the flipped /
// parity-sliced layouts of the data gradients), so the pack must not be a chain of torch permute / cast launches.
//   out[o][t][i] = w[co][ci][taps[t]]   with (o, i) = (co, ci), or (ci, co) when `transpose` (data-gradient weights)
// stored K-major ([o][K = ntaps * I_out]) as bf16 or split-bf16 (hi | lo planes `out_plane` elements apart).
#include "common.cuh"

namespace sbgm {

struct TapList {
  int n;
  int idx[64];
};

// One thread per (o, 8 consecutive i): it walks the tap list, so the fp32 source of a thread (8 x khw contiguous floats in
// the forward orientation) is pulled through L1 once and every output vector is one 16-byte store per plane.
template <int FMT>
__global__ void pack_weight_kernel(const float* __restrict__ w, int cout, int cin, int khw, const TapList taps, int transpose,
                                   void* __restrict__ out, size_t out_plane) {
  pdl_grid_sync();
  const uint32_t oo = transpose ? cin : cout, ii = transpose ? cout : cin;
  const uint32_t ivec = ii >> 3;
  const uint32_t total = oo * ivec;
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < total; v += gridDim.x * blockDim.x) {
    const uint32_t o = v / ivec, i0 = (v - o * ivec) * 8;
    const float* src[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t co = transpose ? i0 + j : o, ci = transpose ? o : i0 + j;
      src[j] = w + (static_cast<size_t>(co) * cin + ci) * khw;
    }
    for (int t = 0; t < taps.n; ++t) {
      const int tap = taps.idx[t];
      float val[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) val[j] = __ldg(src[j] + tap);
      Act<FMT>::store8(out, out_plane, (static_cast<size_t>(o) * taps.n + t) * ii + i0, val);
    }
  }
}

}  // namespace sbgm

using namespace sbgm;

extern "C" int sbgm_pack_weight(const float* w_oihw, int cout, int cin, int khw, const int* taps_host, int ntaps, int transpose,
                                void* out, size_t out_plane, int fmt, void* stream) {
  SBGM_REQUIRE(fmt == SBGM_FMT_BF16 || fmt == SBGM_FMT_BF16X2, "pack_weight: format %d is not a tensor-core format", fmt);
  SBGM_REQUIRE(ntaps >= 1 && ntaps <= 64, "pack_weight: ntaps=%d out of range", ntaps);
  SBGM_REQUIRE((transpose ? cout : cin) % 8 == 0, "pack_weight: inner channel count must be a multiple of 8");
  TapList tl;
  tl.n = ntaps;
  for (int i = 0; i < ntaps; ++i) {
    SBGM_REQUIRE(taps_host[i] >= 0 && taps_host[i] < khw, "pack_weight: tap %d out of range", taps_host[i]);
    tl.idx[i] = taps_host[i];
  }
  const size_t total = static_cast<size_t>(cout) * cin / 8;
  size_t g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  cudaStream_t st = as_stream(stream);
  if (fmt == SBGM_FMT_BF16) launch_k((pack_weight_kernel<SBGM_FMT_BF16>), static_cast<int>(g), 256, 0, st, w_oihw, cout, cin, khw, tl, transpose, out, out_plane);
  else launch_k((pack_weight_kernel<SBGM_FMT_BF16X2>), static_cast<int>(g), 256, 0, st, w_oihw, cout, cin, khw, tl, transpose, out, out_plane);
  return check_launch("pack_weight");
}
