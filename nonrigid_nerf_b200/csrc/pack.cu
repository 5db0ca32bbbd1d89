// Weight packing: fp32 nn.Linear tensors ([out][in] row-major, the reference's checkpoint layout,
// run_nerf_helpers.py:218-238 and :411-482) -> fp16 "chunk-major" UMMA operand images laid out in the
// exact order the fused kernels stream them (layout table in nrn_common.cuh).
// Re-run after every optimizer step / load_state_dict (1.07 M parameters: a few microseconds).
#include <cuda_fp16.h>
#include "nrn_common.cuh"
#include "pack.cuh"

namespace nrn {

namespace {

// image element index -> (chunk, row, e) for an image with R rows
__device__ __forceinline__ void decode(int idx, int R, int& k, int& r) {
  const int c = idx / (R * 8);
  const int rem = idx - c * R * 8;
  r = rem >> 3;
  k = c * 8 + (rem & 7);
}

__global__ void pack_nerf_kernel(NerfSrc src, int in_ch, int out_ch, __half* __restrict__ w, float* __restrict__ bias) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  constexpr int n0 = kNerfL0Bytes / 2, nl = kNerfLBytes / 2, n5 = kNerfL5Bytes / 2, nh = kNerfHeadBytes / 2;
  if (idx < kNerfWBytes / 2) {
    int i = idx, k, r;
    float v = 0.f;
    if (i < n0) {  // L0: K = in_ch (63) padded to 64
      decode(i, 256, k, r);
      v = k < in_ch ? src.w[0][r * in_ch + k] : 0.f;
    } else if ((i -= n0) < 4 * nl) {  // L1..L4
      const int L = 1 + i / nl;
      decode(i % nl, 256, k, r);
      v = src.w[L][r * 256 + k];
    } else if ((i -= 4 * nl) < n5) {  // L5: [embedding(in_ch) pad | h(256)]
      decode(i, 256, k, r);
      const int ld = in_ch + 256;
      if (k < 64) v = k < in_ch ? src.w[5][r * ld + k] : 0.f;
      else v = src.w[5][r * ld + in_ch + (k - 64)];
    } else if ((i -= n5) < 2 * nl) {  // L6, L7
      const int L = 6 + i / nl;
      decode(i % nl, 256, k, r);
      v = src.w[L][r * 256 + k];
    } else {  // head, N padded to 16
      i -= 2 * nl;
      decode(i, 16, k, r);
      v = r < out_ch ? src.w[8][r * 256 + k] : 0.f;
    }
    w[idx] = __float2half_rn(v);
  }
  if (idx < kNerfBiasFloats) {
    float b;
    if (idx < 2048) b = src.b[idx >> 8][idx & 255];
    else b = (idx - 2048) < out_ch ? src.b[8][idx - 2048] : 0.f;
    bias[idx] = b;
  }
  (void)nh;
}

__global__ void pack_bender_kernel(BenderSrc src, __half* __restrict__ w, float* __restrict__ bias) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  constexpr int n0 = kBendB0Bytes / 2, n1 = kBendB1Bytes / 2, n2 = kBendB2Bytes / 2, n3 = kBendB3Bytes / 2;
  constexpr int ld0 = 3 + kLatent;
  if (idx < kBendWBytes / 2) {
    int i = idx, k, r;
    float v = 0.f;
    if (i < n0) {  // B0: K = [xyz_hi xyz_lo latent pad] = 48
      decode(i, 96, k, r);
      if (r < 64) {
        if (k < 3) v = src.net_w[0][r * ld0 + k];
        else if (k < 6) v = src.net_w[0][r * ld0 + (k - 3)];
        else if (k < 6 + kLatent) v = src.net_w[0][r * ld0 + 3 + (k - 6)];
      } else {
        if (k < 3) v = src.rig_w[0][(r - 64) * 3 + k];
        else if (k < 6) v = src.rig_w[0][(r - 64) * 3 + (k - 3)];
      }
    } else if ((i -= n0) < n1) {  // B1: block diagonal 64x64 + 32x32
      decode(i, 96, k, r);
      if (r < 64) { if (k < 64) v = src.net_w[1][r * 64 + k]; }
      else { if (k >= 64) v = src.rig_w[1][(r - 64) * 32 + (k - 64)]; }
    } else if ((i -= n1) < n2) {  // B2: offset L2 + rigidity output row
      decode(i, 80, k, r);
      if (r < 64) { if (k < 64) v = src.net_w[2][r * 64 + k]; }
      else if (r == 64) { if (k >= 64) v = src.rig_w[2][k - 64]; }
    } else if ((i -= n2) < n3) {  // B3
      decode(i, 64, k, r);
      v = src.net_w[3][r * 64 + k];
    } else {  // B4: 3 output rows, no bias
      i -= n3;
      decode(i, 16, k, r);
      if (r < 3) v = src.net_w[4][r * 64 + k];
    }
    w[idx] = __float2half_rn(v);
  }
  if (idx < kBendBiasFloats) {
    float b = 0.f;
    if (idx < 96) b = idx < 64 ? src.net_b[0][idx] : src.rig_b[0][idx - 64];
    else if (idx < 192) { const int j = idx - 96; b = j < 64 ? src.net_b[1][j] : src.rig_b[1][j - 64]; }
    else if (idx < 272) { const int j = idx - 192; b = j < 64 ? src.net_b[2][j] : (j == 64 ? src.rig_b[2][0] : 0.f); }
    else b = src.net_b[3][idx - 272];
    bias[idx] = b;
  }
}

}  // namespace

cudaError_t launch_pack_nerf(const NerfSrc& src, int in_ch, int out_ch, void* packed, cudaStream_t st) {
  __half* w = reinterpret_cast<__half*>(packed);
  float* bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(packed) + kNerfWBytes);
  const int n = kNerfWBytes / 2;
  pack_nerf_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, in_ch, out_ch, w, bias);
  return cudaGetLastError();
}
cudaError_t launch_pack_bender(const BenderSrc& src, void* packed, cudaStream_t st) {
  __half* w = reinterpret_cast<__half*>(packed);
  float* bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(packed) + kBendWBytes);
  const int n = kBendWBytes / 2;
  pack_bender_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, w, bias);
  return cudaGetLastError();
}

}  // namespace nrn
