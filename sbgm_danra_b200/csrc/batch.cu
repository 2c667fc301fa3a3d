// batch.cu -- on-device batch assembly: the host->device boundary of the training / generation loops
// (reference: sbgm/utils.py:405-480 `extract_samples`: one `.to(device).float()` per entry of the dataset's sample dict and a
// torch.cat of the low-resolution conditions; the dataset applied its transforms on the CPU before, sbgm/data_modules.py:727-997
// with the classes of sbgm/special_transforms.py:62-343).
//
// Here the sample dict is staged through ONE pinned buffer and ONE host->device copy (batch.py); this kernel then writes every
// float32 output tensor in a single launch: dtype conversion (the datasets hand out float32 / float64 / float16 / integer
// masks), the channel concatenation of the LR conditions (a source tensor lands at a channel offset inside each destination
// sample) and, optionally, the forward transform of each field (Scale / ZScoreTransform / PrcpLogTransform:
//     v = log ? log(x + eps) : x;   y = (((v - sub) * mul) / div) * post_mul + post_add
// in the reference's float32 operation order), so raw physical fields can cross the bus and be normalised on the device.
// HBM-bound byte work: each element is read once and written once; a block handles 4096 elements of ONE source tensor.
#include "common.cuh"

namespace sbgm {

constexpr int kBatchChunk = 4096;

__device__ __forceinline__ float load_as_float(const void* src, long long i, int dtype) {
  switch (dtype) {
    case SBGM_DT_F32: return static_cast<const float*>(src)[i];
    case SBGM_DT_F64: return static_cast<float>(static_cast<const double*>(src)[i]);
    case SBGM_DT_F16: return __half2float(static_cast<const __half*>(src)[i]);
    case SBGM_DT_BF16: return __bfloat162float(static_cast<const __nv_bfloat16*>(src)[i]);
    case SBGM_DT_I64: return static_cast<float>(static_cast<const long long*>(src)[i]);
    case SBGM_DT_I32: return static_cast<float>(static_cast<const int*>(src)[i]);
    case SBGM_DT_I16: return static_cast<float>(static_cast<const short*>(src)[i]);
    case SBGM_DT_U8: return static_cast<float>(static_cast<const unsigned char*>(src)[i]);
    default: return static_cast<float>(static_cast<const signed char*>(src)[i]);
  }
}

__global__ void __launch_bounds__(256) assemble_batch_kernel(const sbgm_batch_job* __restrict__ jobs, int njobs) {
  pdl_grid_sync();
  int j = 0;
  while (j + 1 < njobs && static_cast<long long>(blockIdx.x) >= jobs[j + 1].first_block) ++j;
  const sbgm_batch_job job = jobs[j];
  const long long base = (static_cast<long long>(blockIdx.x) - job.first_block) * kBatchChunk;
#pragma unroll 4
  for (int k = threadIdx.x; k < kBatchChunk; k += 256) {
    const long long i = base + k;
    if (i >= job.count) break;
    float v = load_as_float(job.src, i, job.dtype);
    if (job.log) v = logf(v + job.eps);
    if (job.transform) v = (((v - job.sub) * job.mul) / job.div) * job.post_mul + job.post_add;
    const long long n = i / job.inner, r = i - n * job.inner;
    job.dst[n * job.dst_stride + job.dst_offset + r] = v;
  }
}

}  // namespace sbgm

using namespace sbgm;

extern "C" {

int sbgm_batch_chunk_elems(void) { return kBatchChunk; }

int sbgm_assemble_batch(const sbgm_batch_job* jobs_dev, int njobs, long long total_blocks, void* stream) {
  if (njobs <= 0 || total_blocks <= 0) return 0;
  SBGM_REQUIRE(jobs_dev != nullptr, "assemble_batch: job table missing");
  SBGM_REQUIRE(total_blocks < (1ll << 31), "assemble_batch: too many blocks");
  launch_k(assemble_batch_kernel, static_cast<unsigned int>(total_blocks), 256, 0, as_stream(stream), jobs_dev, njobs);
  return check_launch("assemble_batch");
}

}  // extern "C"
