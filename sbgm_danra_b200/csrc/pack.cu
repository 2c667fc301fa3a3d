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

template <int FMT>
__global__ void pack_weight_kernel(const float* __restrict__ w, int cout, int cin, int khw, const TapList taps, int transpose,
                                   void* __restrict__ out, size_t out_plane) {
  pdl_grid_sync();
  const int oo = transpose ? cin : cout, ii = transpose ? cout : cin;
  const int ivec = ii >> 3;
  const size_t total = static_cast<size_t>(oo) * taps.n * ivec;
  for (size_t v = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; v < total; v += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int i0 = static_cast<int>(v % ivec) * 8;
    const int t = static_cast<int>((v / ivec) % taps.n);
    const int o = static_cast<int>(v / (static_cast<size_t>(ivec) * taps.n));
    const int tap = taps.idx[t];
    float val[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int co = transpose ? i0 + j : o, ci = transpose ? o : i0 + j;
      val[j] = __ldg(w + (static_cast<size_t>(co) * cin + ci) * khw + tap);
    }
    Act<FMT>::store8(out, out_plane, v * 8, val);
  }
}

// ---- batched form: every weight of the network in one launch ------------------------------------------------------
// The job table travels as a __grid_constant__ kernel parameter (no device-side table, no host-to-device copy), so the
// launch is capturable in the training step's CUDA graph.  A work item is one 8-channel output vector.
struct PackJobDev {
  const float* w;
  void* out;
  uint32_t out_plane, item_end;
  uint16_t cout, cin;
  uint8_t khw, ntaps, transpose, pad_;
  uint8_t taps[16];
};
constexpr int kPackBatch = 80;
struct PackBatch {
  int njobs;
  PackJobDev jobs[kPackBatch];
};

template <int FMT>
__global__ void pack_batch_kernel(const __grid_constant__ PackBatch b) {
  pdl_grid_sync();
  const uint32_t total = b.jobs[b.njobs - 1].item_end;
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < total; v += gridDim.x * blockDim.x) {
    int lo = 0, hi = b.njobs - 1;                  // first job with item_end > v
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (b.jobs[mid].item_end > v) hi = mid; else lo = mid + 1;
    }
    const PackJobDev& j = b.jobs[lo];
    const uint32_t local = v - (lo > 0 ? b.jobs[lo - 1].item_end : 0u);
    const uint32_t ii = j.transpose ? j.cout : j.cin;
    const uint32_t ivec = ii >> 3;
    const uint32_t i0 = (local % ivec) * 8;
    const uint32_t t = (local / ivec) % j.ntaps;
    const uint32_t o = local / (ivec * j.ntaps);
    const int tap = j.taps[t];
    float val[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const uint32_t co = j.transpose ? i0 + e : o, ci = j.transpose ? o : i0 + e;
      val[e] = __ldg(j.w + (static_cast<size_t>(co) * j.cin + ci) * j.khw + tap);
    }
    Act<FMT>::store8(j.out, j.out_plane, static_cast<size_t>(local) * 8, val);
  }
}

}  // namespace sbgm

using namespace sbgm;

extern "C" int sbgm_pack_weights(const sbgm_pack_job* jobs_host, int njobs, int fmt, void* stream) {
  SBGM_REQUIRE(fmt == SBGM_FMT_BF16 || fmt == SBGM_FMT_BF16X2, "pack_weights: format %d is not a tensor-core format", fmt);
  cudaStream_t st = as_stream(stream);
  for (int base = 0; base < njobs; base += kPackBatch) {
    PackBatch b;
    b.njobs = (njobs - base < kPackBatch) ? njobs - base : kPackBatch;
    uint64_t items = 0;
    for (int k = 0; k < b.njobs; ++k) {
      const sbgm_pack_job& h = jobs_host[base + k];
      SBGM_REQUIRE(h.ntaps >= 1 && h.ntaps <= 16 && h.khw >= 1 && h.khw <= 255, "pack_weights: job %d has %d taps of %d (batched form: <= 16)", base + k, h.ntaps, h.khw);
      SBGM_REQUIRE(h.cout >= 1 && h.cout <= 65535 && h.cin >= 1 && h.cin <= 65535 && (h.transpose ? h.cout : h.cin) % 8 == 0,
                   "pack_weights: job %d: bad channel counts %d, %d", base + k, h.cout, h.cin);
      PackJobDev& d = b.jobs[k];
      d.w = h.w_oihw; d.out = h.out; d.out_plane = static_cast<uint32_t>(h.out_plane);
      d.cout = static_cast<uint16_t>(h.cout); d.cin = static_cast<uint16_t>(h.cin);
      d.khw = static_cast<uint8_t>(h.khw); d.ntaps = static_cast<uint8_t>(h.ntaps); d.transpose = h.transpose ? 1 : 0; d.pad_ = 0;
      for (int t = 0; t < 16; ++t) {
        SBGM_REQUIRE(t >= h.ntaps || (h.taps[t] >= 0 && h.taps[t] < h.khw), "pack_weights: job %d tap out of range", base + k);
        d.taps[t] = static_cast<uint8_t>(t < h.ntaps ? h.taps[t] : 0);
      }
      items += static_cast<uint64_t>(h.cout) * h.cin * h.ntaps / 8;
      SBGM_REQUIRE(items < (1ull << 32), "pack_weights: batch too large");
      d.item_end = static_cast<uint32_t>(items);
    }
    size_t g = (items + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    if (g < 1) g = 1;
    if (fmt == SBGM_FMT_BF16) launch_k((pack_batch_kernel<SBGM_FMT_BF16>), static_cast<int>(g), 256, 0, st, b);
    else launch_k((pack_batch_kernel<SBGM_FMT_BF16X2>), static_cast<int>(g), 256, 0, st, b);
  }
  return check_launch("pack_weights");
}


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
  const size_t total = static_cast<size_t>(cout) * cin * ntaps / 8;
  size_t g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  cudaStream_t st = as_stream(stream);
  if (fmt == SBGM_FMT_BF16) launch_k((pack_weight_kernel<SBGM_FMT_BF16>), static_cast<int>(g), 256, 0, st, w_oihw, cout, cin, khw, tl, transpose, out, out_plane);
  else launch_k((pack_weight_kernel<SBGM_FMT_BF16X2>), static_cast<int>(g), 256, 0, st, w_oihw, cout, cin, khw, tl, transpose, out, out_plane);
  return check_launch("pack_weight");
}
