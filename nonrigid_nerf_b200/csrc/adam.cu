// One-launch Adam over every trainable tensor of the model (replaces torch.optim.Adam at train.py:656-658,
// stepped at train.py:1608).  PyTorch's fused Adam needs 6 multi-tensor launches (~150 us) for the 137 tensors of
// this model (86 per-image latents, 15 bender tensors, 2 x 18 NeRF tensors); here the parameters live in ONE flat
// fp32 buffer (the nn.Parameters are views into it), the moments in two more, and a block table maps every CUDA
// block to (tensor, offset, count).  Gradients are read in place through a device array of per-tensor pointers
// (autograd hands out one gradient tensor per parameter); a null pointer skips the tensor, like torch.optim.Adam
// skips parameters whose .grad is None.  The per-tensor step counts and the learning rate live on the device so that the
// whole training iteration can be replayed as a CUDA graph.
//
// Arithmetic = torch.optim.Adam (amsgrad=False, weight_decay=0, maximize=False):
//   m = m + (g - m) (1 - b1);  v = b2 v + (1 - b2) g^2;  p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include "adam.cuh"

namespace nrn {

namespace {
constexpr int kAdamThreads = 256;

// torch.optim.Adam counts steps per parameter: a tensor without a gradient does not advance
__global__ void adam_tick_kernel(long long* step, const float* const* grads, int n_tensors) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_tensors && grads[t] != nullptr) step[t] += 1;
}

__global__ void __launch_bounds__(kAdamThreads) adam_kernel(const AdamParams a) {
  const AdamBlock b = a.blocks[blockIdx.x];
  const float* __restrict__ g = a.grads[b.tensor];
  if (g == nullptr) return;
  __shared__ float s_c[2];
  if (threadIdx.x == 0) {
    const double t = static_cast<double>(a.step[b.tensor]);
    const double bc1 = 1.0 - pow(static_cast<double>(a.beta1), t);
    const double bc2 = 1.0 - pow(static_cast<double>(a.beta2), t);
    s_c[0] = static_cast<float>(static_cast<double>(a.lr[0]) / bc1);   // step size
    s_c[1] = static_cast<float>(sqrt(bc2));
  }
  __syncthreads();
  const float step_size = s_c[0], bc2_sqrt = s_c[1];
  const float w1 = 1.0f - a.beta1, w2 = 1.0f - a.beta2;
  g += b.start;
  float* __restrict__ p = a.params + b.flat_off;
  float* __restrict__ m = a.exp_avg + b.flat_off;
  float* __restrict__ v = a.exp_avg_sq + b.flat_off;
  // 8 elements per thread: all loads first (the kernel is a latency-bound stream over 4 arrays)
  constexpr int kPer = kAdamBlockElems / kAdamThreads;
  float gi[kPer], mi[kPer], vi[kPer], pi[kPer];
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    const int i = threadIdx.x + k * kAdamThreads;
    const bool in = i < b.count;
    gi[k] = in ? g[i] : 0.f;
    mi[k] = in ? m[i] : 0.f;
    vi[k] = in ? v[i] : 0.f;
    pi[k] = in ? p[i] : 0.f;
  }
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    const int i = threadIdx.x + k * kAdamThreads;
    if (i < b.count) {
      const float mk = mi[k] + (gi[k] - mi[k]) * w1;
      const float vk = a.beta2 * vi[k] + w2 * gi[k] * gi[k];
      m[i] = mk;
      v[i] = vk;
      p[i] = pi[k] - step_size * (mk / (sqrtf(vk) / bc2_sqrt + a.eps));
    }
  }
}
}  // namespace

cudaError_t launch_adam(const AdamParams& a, int n_tensors, int n_blocks, cudaStream_t st) {
  if (n_tensors <= 0) return cudaSuccess;
  adam_tick_kernel<<<(n_tensors + 127) / 128, 128, 0, st>>>(a.step, a.grads, n_tensors);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || n_blocks <= 0) return e;
  adam_kernel<<<n_blocks, kAdamThreads, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace nrn
