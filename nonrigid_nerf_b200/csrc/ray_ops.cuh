// Parameter blocks and launchers of the per-ray kernels (ray_ops.cu).
#pragma once
#include <cuda_runtime.h>

namespace nrn {

struct CompositeParams {
  const float* raw;     // [n][S][C]
  const float* z;       // [n][S]
  const float* rays_d;  // row stride rays_d_stride floats (8 when pointing into the [n][8] ray table)
  int rays_d_stride;
  const float* noise;   // [n][S] additive sigma noise (already scaled) or null
  int n, S, C, white_bkgd;
  float* rgb;           // [n][3]
  float* disp;          // [n]
  float* acc;           // [n]
  float* depth;         // [n] or null
  float* weights;       // [n][S] or null
  float* alpha;         // [n][S] or null
  // importance resampling (n_imp == 0: off)
  int n_imp;
  const float* u;       // [n][n_imp] or null (deterministic linspace)
  float* z_out;         // [n][S + n_imp]
  float* z_std;         // [n] or null
};

struct CompositeBwdParams {
  const float* raw;
  const float* z;
  const float* rays_d;
  int rays_d_stride;
  const float* noise;
  int n, S, C, white_bkgd;
  const float* d_rgb;   // [n][3]
  const float* d_acc;   // [n] or null
  float* d_raw;         // [n][S][C]
};

cudaError_t launch_sample_coarse(const float* rays, const float* t_rand, int n, int S, int lindisp, float* z_out,
                                 cudaStream_t st);
cudaError_t launch_composite(const CompositeParams& p, cudaStream_t st);
cudaError_t launch_sample_pdf(const float* bins, const float* weights, const float* u, int n, int nb, int n_samp,
                              float* out, cudaStream_t st);
cudaError_t launch_composite_bwd(const CompositeBwdParams& p, cudaStream_t st);
cudaError_t launch_get_rays(const float* c2w, const float* K, int H, int W, float* rays_o, float* rays_d, cudaStream_t st);
cudaError_t launch_ray_batch(const long long* pix, int n, const float* poses, const float* K, const int* image_to_view,
                             const float* images, int H, int W, float* rays_o, float* rays_d, float* target, cudaStream_t st);
cudaError_t launch_pack_rays(const float* o, const float* d, float near, float far, int n, float* rays, cudaStream_t st);
cudaError_t launch_median_index(const float* w, int n, int S, long long* idx, cudaStream_t st);

}  // namespace nrn
