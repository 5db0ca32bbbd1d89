// Parameter block and launchers of the fused per-ray training loss (loss.cu).
#pragma once
#include <cuda_runtime.h>

namespace nrn {

struct RayLossParams {
  int n, S;
  const float* rgb;     // [n][3] fine (or only) rgb_map
  const float* rgb0;    // [n][3] coarse rgb_map or null
  const float* target;  // [n][3]
  const float* w;       // [n][S] coarse visibility weights (detached) or null
  const float* off;     // [n][S][3] coarse unmasked offsets or null (no offsets term)
  const float* rig;     // [n][S] coarse rigidity mask
  float lam_o, lam_r;
  const float* sched_step;    // device scalar global_step: both regulariser weights are multiplied by 0.01^(1 - step / n_iters); or null
  float sched_n_iters;
  const float* div;           // [n] per-ray divergence regulariser or null
  float lam_div;
  float* u_div;               // [n]
  float* loss;          // [n]
  float* u_rgb;         // [n][3]   gradients per unit upstream gradient
  float* u_rgb0;        // [n][3]
  float* u_off;         // [n][S][3]
  float* u_rig;         // [n][S]
};

cudaError_t launch_ray_loss(const RayLossParams& p, cudaStream_t st);
cudaError_t launch_ray_loss_scale(const float* g, const float* unit, float* out, long long n, int per, cudaStream_t st);

struct RayLossBwdParams {
  int n, S;
  const float* g;
  const float* u[5];   // rgb, rgb0, unmasked offsets, rigidity, divergence (null = absent)
  float* d[5];
};
cudaError_t launch_ray_loss_bwd(const RayLossBwdParams& p, cudaStream_t st);

}  // namespace nrn
