// Shared definitions: packed-weight layout, kernel parameter blocks, launch helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nrn {

// ------------------------------------------------------------------------------------------
// Geometry of the fused field kernels
// ------------------------------------------------------------------------------------------
constexpr int kTileM = 128;                 // points per tile = TMEM lanes = UMMA M
constexpr int kChunkBytes = kTileM * 16;    // one 8-column chunk of a 128-row activation image
constexpr int kHBytes = 32 * kChunkBytes;   // 256-wide hidden activations, 64 KB
constexpr int kEBytes = 8 * kChunkBytes;    // 64-wide positional embedding, 16 KB
constexpr int kSlotBytes = kHBytes + kEBytes;
constexpr int kRingStageBytes = 32768;      // one weight slab: 256 rows x 64 K-columns fp16
constexpr int kRingStages = 2;
constexpr int kFwdThreads = 384;            // 12 warps: producer, mma, tmem, spare, 2 x 4 epilogue

// ------------------------------------------------------------------------------------------
// Packed NeRF weights (fp16, chunk-major images in streaming order), see pack.cu
//   L0   : [ 8 chunks][256 rows][8]  K = 63 (+1 zero pad)                 32 KB   (1 slab)
//   L1-4 : [32 chunks][256 rows][8]                                       128 KB  (4 slabs each)
//   L5   : [40 chunks][256 rows][8]  K = 64 (embedding, padded) + 256     160 KB  (5 slabs)
//   L6-7 : [32 chunks][256 rows][8]                                       128 KB  (4 slabs each)
//   head : [32 chunks][ 16 rows][8]  N = out_ch (<= 16, zero padded)      8 KB    (1 slab)
// followed by fp32 biases: [8][256] + [16]
// ------------------------------------------------------------------------------------------
constexpr int kNerfL0Bytes = 8 * 256 * 16;
constexpr int kNerfLBytes = 32 * 256 * 16;
constexpr int kNerfL5Bytes = 40 * 256 * 16;
constexpr int kNerfHeadBytes = 32 * 16 * 16;
constexpr int kNerfWBytes = kNerfL0Bytes + 6 * kNerfLBytes + kNerfL5Bytes + kNerfHeadBytes;  // 991,232
constexpr int kNerfBiasFloats = 8 * 256 + 16;
constexpr int kNerfPackedBytes = kNerfWBytes + kNerfBiasFloats * 4;

// Packed ray-bender weights (fp16): offset MLP and rigidity MLP fused block-diagonally.
//   B0: N=96 K=48  rows 0-63 offset L0 (cols: xyz_hi 0-2, xyz_lo 3-5, latent 6-37), rows 64-95 rigidity L0
//   B1: N=96 K=96  rows 0-63 offset L1 (cols 0-63),   rows 64-95 rigidity L1 (cols 64-95)
//   B2: N=80 K=96  rows 0-63 offset L2 (cols 0-63),   row 64 rigidity L2 (cols 64-95)
//   B3: N=64 K=64  offset L3
//   B4: N=16 K=64  rows 0-2 offset L4 (no bias)
// followed by fp32 biases: [96] [96] [80] [64]
constexpr int kBendB0Bytes = 6 * 96 * 16;
constexpr int kBendB1Bytes = 12 * 96 * 16;
constexpr int kBendB2Bytes = 12 * 80 * 16;
constexpr int kBendB3Bytes = 8 * 64 * 16;
constexpr int kBendB4Bytes = 8 * 16 * 16;
constexpr int kBendWBytes = kBendB0Bytes + kBendB1Bytes + kBendB2Bytes + kBendB3Bytes + kBendB4Bytes;  // 53,248
constexpr int kBendBiasFloats = 96 + 96 + 80 + 64;
constexpr int kBendPackedBytes = kBendWBytes + kBendBiasFloats * 4;
constexpr int kLatent = 32;

// ------------------------------------------------------------------------------------------
// Kernel parameter blocks
// ------------------------------------------------------------------------------------------
struct FieldFwdParams {
  const float* rays;      // [N][8]  o(3) d(3) near far
  const float* z_vals;    // [N][S]
  const float* pts;       // point mode (rays == null): [N][pts_stride] xyz first, S == 1
  long long pts_stride;
  const float* latents;   // [N][32] (row stride latent_stride floats; 0 = one latent for all rays)
  long long latent_stride;
  int n_rays, S;
  long long P;            // n_rays * S
  int n_tiles;
  const uint8_t* nerf_w;
  const float* nerf_bias;
  const uint8_t* bend_w;
  const float* bend_bias;
  float cutoff, scaling, removal;
  int use_cutoff, use_scaling, use_removal;
  int out_ch;
  float* raw;             // [P][out_ch]
  float* d_init;          // [P][3] or null
  float* d_bent;          // [P][3] or null
  float* d_unmasked;      // [P][3] or null
  float* d_masked;        // [P][3] or null
  float* d_rigid;         // [P]    or null
  int* err;               // device error word (0 = ok)
};

}  // namespace nrn
