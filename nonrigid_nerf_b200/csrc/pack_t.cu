// Transposed weight packing for the DGRAD kernel: fp32 nn.Linear tensors -> fp16 chunk-major
// images of W^T (rows = input features, K = output features), layout in nrn_common.cuh.
#include <cuda_fp16.h>
#include "nrn_common.cuh"
#include "pack.cuh"

namespace nrn {

namespace {

__device__ __forceinline__ void decode(int idx, int R, int& k, int& r) {
  const int c = idx / (R * 8);
  const int rem = idx - c * R * 8;
  r = rem >> 3;
  k = c * 8 + (rem & 7);
}

// r = input feature, k = output feature
__global__ void pack_nerf_t_kernel(NerfSrc src, int in_ch, int out_ch, __half* __restrict__ w) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= kNerfTWBytes / 2) return;
  constexpr int nh = kNerfTHeadBytes / 2, nl = kNerfLBytes / 2, ne = kNerfTEBytes / 2;
  int i = idx, k, r;
  float v = 0.f;
  const int ld5 = in_ch + 256;
  if (i < nh) {                           // head^T: K = out_ch padded to 16
    decode(i, 256, k, r);
    v = k < out_ch ? src.w[8][k * 256 + r] : 0.f;
  } else if ((i -= nh) < 2 * nl) {        // L7^T, L6^T
    const int L = 7 - i / nl;
    decode(i % nl, 256, k, r);
    v = src.w[L][k * 256 + r];
  } else if ((i -= 2 * nl) < ne) {        // L5e^T: rows = embedding inputs (in_ch, padded to 64)
    decode(i, 64, k, r);
    v = r < in_ch ? src.w[5][k * ld5 + r] : 0.f;
  } else if ((i -= ne) < nl) {            // L5h^T
    decode(i, 256, k, r);
    v = src.w[5][k * ld5 + in_ch + r];
  } else if ((i -= nl) < 4 * nl) {        // L4^T .. L1^T
    const int L = 4 - i / nl;
    decode(i % nl, 256, k, r);
    v = src.w[L][k * 256 + r];
  } else {                                // L0^T
    i -= 4 * nl;
    decode(i, 64, k, r);
    v = r < in_ch ? src.w[0][k * in_ch + r] : 0.f;
  }
  w[idx] = __float2half_rn(v);
}

__global__ void pack_bender_t_kernel(BenderSrc src, __half* __restrict__ w) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= kBendTWBytes / 2) return;
  constexpr int n4 = kBendTB4Bytes / 2, n3 = kBendTB3Bytes / 2, n2 = kBendTB2Bytes / 2, n1 = kBendTB1Bytes / 2;
  constexpr int ld0 = 3 + kLatent;
  int i = idx, k, r;
  float v = 0.f;
  if (i < n4) {                           // B4^T: rows = 64 hidden, K = 3 outputs (padded to 16)
    decode(i, 64, k, r);
    v = k < 3 ? src.net_w[4][k * 64 + r] : 0.f;
  } else if ((i -= n4) < n3) {            // B3^T
    decode(i, 64, k, r);
    v = src.net_w[3][k * 64 + r];
  } else if ((i -= n3) < n2) {            // B2^T: rows = 96 inputs, K = 80 outputs (64 offset, 1 rigidity, pad)
    decode(i, 96, k, r);
    if (r < 64) { if (k < 64) v = src.net_w[2][k * 64 + r]; }
    else { if (k == 64) v = src.rig_w[2][r - 64]; }
  } else if ((i -= n2) < n1) {            // B1^T: block diagonal
    decode(i, 96, k, r);
    if (r < 64) { if (k < 64) v = src.net_w[1][k * 64 + r]; }
    else { if (k >= 64) v = src.rig_w[1][(k - 64) * 32 + (r - 64)]; }
  } else {                                // B0^T: rows = 48 inputs [xyz_hi xyz_lo latent pad], K = 96 outputs
    i -= n1;
    decode(i, 48, k, r);
    if (r >= 6 && r < 6 + kLatent) { if (k < 64) v = src.net_w[0][k * ld0 + 3 + (r - 6)]; }
    else if (r < 3) v = k < 64 ? src.net_w[0][k * ld0 + r] : src.rig_w[0][(k - 64) * 3 + r];
  }
  w[idx] = __float2half_rn(v);
}

}  // namespace

cudaError_t launch_pack_nerf_t(const NerfSrc& src, int in_ch, int out_ch, void* packed, cudaStream_t st) {
  const int n = kNerfTWBytes / 2;
  pack_nerf_t_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, in_ch, out_ch, reinterpret_cast<__half*>(packed));
  return cudaGetLastError();
}
cudaError_t launch_pack_bender_t(const BenderSrc& src, void* packed, cudaStream_t st) {
  const int n = kBendTWBytes / 2;
  pack_bender_t_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, reinterpret_cast<__half*>(packed));
  return cudaGetLastError();
}

}  // namespace nrn
