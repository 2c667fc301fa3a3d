// optim.cu -- the optimizer update of the DSM training step as ONE launch over every parameter tensor
// (reference: sbgm/training.py:407 `self.optimizer.step()` with the torch.optim.Adam / AdamW built by
// sbgm/training_utils.py:672-698; torch's default foreach path is ~26 launches over 164 tensors = 0.5 ms of a 7 ms step).
//
// The parameters stay where torch put them (the optimizer updates them in place; the training graph re-packs them from
// their live storage), so the kernel walks a table of chunks: (param, grad, exp_avg, exp_avg_sq, count), at most kChunk
// elements of ONE tensor each, one block per chunk.  HBM-bound: 4 reads + 3 writes of 4 B per parameter.
#include "common.cuh"

namespace sbgm {

constexpr int kAdamChunk = 4096;    // elements per block: 256 threads x 4 float4

struct AdamHyper {
  float beta2, one_minus_beta1, one_minus_beta2, eps, weight_decay, decay_mul, step_size, sqrt_bc2;
  int decoupled;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamHyper& h) {
  // torch/optim/adam.py _single_tensor_adam: grad += wd * p (Adam) or p *= 1 - lr * wd (AdamW); m.lerp_(g, 1 - b1);
  // v = b2 * v + (1 - b2) g^2; p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
  if (h.weight_decay != 0.0f) {
    if (h.decoupled) p *= h.decay_mul;
    else g = fmaf(h.weight_decay, p, g);
  }
  m = fmaf(h.one_minus_beta1, g - m, m);
  v = fmaf(h.one_minus_beta2, g * g, h.beta2 * v);
  const float denom = sqrtf(v) / h.sqrt_bc2 + h.eps;
  p = fmaf(-h.step_size, m / denom, p);
}

__global__ void __launch_bounds__(256) adam_step_kernel(const sbgm_adam_chunk* __restrict__ chunks, const AdamHyper h) {
  pdl_grid_sync();
  const sbgm_adam_chunk c = chunks[blockIdx.x];
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(c.param) | reinterpret_cast<uintptr_t>(c.grad) | reinterpret_cast<uintptr_t>(c.exp_avg) |
                        reinterpret_cast<uintptr_t>(c.exp_avg_sq)) & 15) == 0;
  if (vec_ok) {
    const int nvec = c.count >> 2;
    float4* p4 = reinterpret_cast<float4*>(c.param);
    const float4* g4 = reinterpret_cast<const float4*>(c.grad);
    float4* m4 = reinterpret_cast<float4*>(c.exp_avg);
    float4* v4 = reinterpret_cast<float4*>(c.exp_avg_sq);
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
      float4 p = p4[i], m = m4[i], v = v4[i];
      const float4 g = __ldg(g4 + i);
      adam_one(p.x, g.x, m.x, v.x, h);
      adam_one(p.y, g.y, m.y, v.y, h);
      adam_one(p.z, g.z, m.z, v.z, h);
      adam_one(p.w, g.w, m.w, v.w, h);
      p4[i] = p; m4[i] = m; v4[i] = v;
    }
    for (int i = (nvec << 2) + threadIdx.x; i < c.count; i += blockDim.x) adam_one(c.param[i], c.grad[i], c.exp_avg[i], c.exp_avg_sq[i], h);
  } else {
    for (int i = threadIdx.x; i < c.count; i += blockDim.x) adam_one(c.param[i], c.grad[i], c.exp_avg[i], c.exp_avg_sq[i], h);
  }
}

}  // namespace sbgm

using namespace sbgm;

extern "C" {

int sbgm_adam_chunk_elems(void) { return kAdamChunk; }

int sbgm_adam_step(const sbgm_adam_chunk* chunks_dev, int n_chunks, double lr, double beta1, double beta2, float eps, float weight_decay,
                   int decoupled, double bias_correction1, double bias_correction2, void* stream) {
  if (n_chunks <= 0) return 0;
  SBGM_REQUIRE(chunks_dev != nullptr, "adam_step: chunk table missing");
  SBGM_REQUIRE(bias_correction1 > 0.0 && bias_correction2 > 0.0, "adam_step: bias corrections must be positive (step >= 1)");
  AdamHyper h;
  h.decay_mul = static_cast<float>(1.0 - lr * static_cast<double>(weight_decay));
  // torch passes 1 - beta as a double-precision scalar: (float)(1 - 0.999) != 1 - (float)0.999
  h.beta2 = static_cast<float>(beta2); h.one_minus_beta1 = static_cast<float>(1.0 - beta1); h.one_minus_beta2 = static_cast<float>(1.0 - beta2);
  h.eps = eps; h.weight_decay = weight_decay;
  h.step_size = static_cast<float>(lr / bias_correction1);
  h.sqrt_bc2 = static_cast<float>(sqrt(bias_correction2));
  h.decoupled = decoupled;
  launch_k(adam_step_kernel, n_chunks, 256, 0, as_stream(stream), chunks_dev, h);
  return check_launch("adam_step");
}

}  // extern "C"
