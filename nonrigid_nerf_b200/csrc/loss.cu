// Per-ray training loss of training_wrapper_class.forward (train.py:208-242), one warp per ray:
//   loss = mean_c (rgb - t)^2 + mean_c (rgb0 - t)^2
//        + lam_o * [ mean_s( w * ||off||^(2 - r) ) + lam_r * mean_s( w * r ) ]
// with w = coarse visibility weights (detached), off = coarse unmasked offsets, r = coarse rigidity mask,
// lam_o = offsets_loss_weight * (1/100)^(1 - step/N_iters), lam_r = rigidity_loss_weight.
// The same pass writes the gradients per unit upstream gradient (the loss is linear in g[ray]):
//   d rgb = 2 (rgb - t) / 3,  d off = lam_o w p ||off||^(p-2) off / S  (p = 2 - r; 0 at off = 0, like torch.pow),
//   d r = lam_o w ( -||off||^p ln||off|| + lam_r ) / S
// so the backward is a broadcast multiply by g (csrc: ray_loss_scale_kernel).
#include "loss.cuh"

namespace nrn {

namespace {
constexpr int kWarpsPerBlock = 4;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
}  // namespace

__global__ void __launch_bounds__(kWarpsPerBlock * 32) ray_loss_kernel(const RayLossParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kWarpsPerBlock + warp;
  if (ray >= p.n) return;
  // increasing schedule of the regularisers (train.py:229, :281), from the device-resident step counter when given
  const float sched = p.sched_step ? powf(0.01f, 1.0f - __ldg(p.sched_step) / p.sched_n_iters) : 1.0f;
  float loss = 0.f;
  if (lane < 3) {
    const float t = p.target[ray * 3 + lane];
    const float d1 = p.rgb[ray * 3 + lane] - t;
    loss += d1 * d1 * (1.0f / 3.0f);
    p.u_rgb[ray * 3 + lane] = d1 * (2.0f / 3.0f);
    if (p.rgb0) {
      const float d0 = p.rgb0[ray * 3 + lane] - t;
      loss += d0 * d0 * (1.0f / 3.0f);
      p.u_rgb0[ray * 3 + lane] = d0 * (2.0f / 3.0f);
    }
  }
  if (p.off) {
    const float lam_o = p.lam_o * sched;
    const float inv_s = 1.0f / static_cast<float>(p.S);
    float acc = 0.f;
    for (int i = lane; i < p.S; i += 32) {
      const long long pt = static_cast<long long>(ray) * p.S + i;
      const float w = p.w[pt], r = p.rig[pt];
      const float ox = p.off[pt * 3], oy = p.off[pt * 3 + 1], oz = p.off[pt * 3 + 2];
      const float nrm = sqrtf(ox * ox + oy * oy + oz * oz);
      const float pw = 2.0f - r;
      float f = 0.f, dn = 0.f, dr = 0.f;   // ||o||^p, d/d||o||, d/dr
      if (nrm > 0.f) {
        f = powf(nrm, pw);
        dn = pw * f / nrm;
        dr = -f * logf(nrm);
      } else if (pw == 0.f) {
        f = 1.0f;                            // torch.pow(0, 0) = 1
      }
      acc += w * (f + p.lam_r * r);
      const float c = lam_o * w * inv_s;
      const float k = nrm > 0.f ? c * dn / nrm : 0.f;   // d||o||/do = o / ||o||, defined as 0 at 0
      p.u_off[pt * 3] = k * ox; p.u_off[pt * 3 + 1] = k * oy; p.u_off[pt * 3 + 2] = k * oz;
      p.u_rig[pt] = c * (dr + p.lam_r);
    }
    loss += lam_o * acc * inv_s;
  }
  loss = warp_sum(loss);
  if (lane == 0) {
    if (p.div) {
      const float c = p.lam_div * sched;
      loss += c * p.div[ray];
      p.u_div[ray] = c;
    }
    p.loss[ray] = loss;
  }
}

// out_k[i] = g[ray of i] * unit_k[i] for all unit arrays of the loss in one launch (grid-stride over the largest array)
__global__ void ray_loss_bwd_kernel(const RayLossBwdParams p) {
  const long long per[5] = {3, 3, 3ll * p.S, p.S, 1};
  const long long total = static_cast<long long>(p.n) * per[2];
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      if (p.u[k] && i < static_cast<long long>(p.n) * per[k]) p.d[k][i] = p.g[i / per[k]] * p.u[k][i];
    }
  }
}

// out[i] = g[i / per] * unit[i]
__global__ void ray_loss_scale_kernel(const float* __restrict__ g, const float* __restrict__ unit, float* __restrict__ out,
                                      long long n, int per) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) out[i] = g[i / per] * unit[i];
}

cudaError_t launch_ray_loss(const RayLossParams& p, cudaStream_t st) {
  if (p.n <= 0) return cudaSuccess;
  ray_loss_kernel<<<(p.n + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0, st>>>(p);
  return cudaGetLastError();
}
cudaError_t launch_ray_loss_bwd(const RayLossBwdParams& p, cudaStream_t st) {
  if (p.n <= 0) return cudaSuccess;
  const long long total = static_cast<long long>(p.n) * 3 * p.S;
  long long blocks = (total + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  ray_loss_bwd_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(p);
  return cudaGetLastError();
}
cudaError_t launch_ray_loss_scale(const float* g, const float* unit, float* out, long long n, int per, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  ray_loss_scale_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(g, unit, out, n, per);
  return cudaGetLastError();
}

}  // namespace nrn
