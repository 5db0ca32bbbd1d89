#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nrn {

struct AdamBlock {
  int tensor;        // index into the gradient pointer table
  int start;         // first element of this block inside the tensor
  int count;         // elements handled by this block (<= kAdamBlockElems)
  int flat_off;      // offset of `start` inside the flat parameter / moment buffers
};
constexpr int kAdamBlockElems = 2048;

struct AdamParams {
  float* params;
  float* exp_avg;
  float* exp_avg_sq;
  const float* const* grads;     // device array [n_tensors]; null = skip the tensor this step
  const AdamBlock* blocks;       // device array [n_blocks]
  const float* lr;               // device scalar
  long long* step;               // device array [n_tensors]: steps taken by each tensor (torch counts per parameter)
  float beta1, beta2, eps;
};

cudaError_t launch_adam(const AdamParams& a, int n_tensors, int n_blocks, cudaStream_t st);

}  // namespace nrn
