// Parameter block and launchers of the weight-gradient kernels (wgrad.cu) and small utilities.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nrn {

constexpr int kWgScratchFloats = 256 * 256 + 256;  // per-CTA partial: dW tile [256][256] + bias [256]
constexpr int kWgMaxCtas = 192;

struct WgradParams {
  const uint8_t* stash;    // forward activation stash
  const uint8_t* gstash;   // gradient stash written by DGRAD
  float* scratch;          // [kWgMaxCtas][kWgScratchFloats]
  const float* amax;       // loss-scale source (see field_bwd.cu) or null
  int n_tiles;
  int compact;             // 1: stashes hold only the bender images (divergence regulariser), bender jobs only
  long long stash_tile_bytes, gstash_tile_bytes;   // filled in by launch_wgrad
  int n_jobs;
  int job_ids[16];
  int splits[16];
  int* err;
};

cudaError_t launch_wgrad(WgradParams p, bool has_bender, int num_sms, float* nerf_grad, int nerf_n, float* bend_grad,
                         int bend_n, int out_ch, cudaStream_t st);
// amax[0] = max |x[i]| over n floats (device scalar, overwritten)
cudaError_t launch_absmax(const float* x, long long n, float* amax, cudaStream_t st);

}  // namespace nrn
