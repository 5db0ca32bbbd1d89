// Weight gradients (WGRAD) of the fused field:  dW_l[out][in] = sum_p dY_l[p][out] * X_l[p][in],
// db_l[out] = sum_p dY_l[p][out]  -- what torch.autograd accumulates into the .grad of
// NeRF.pts_linears / output_linear (run_nerf_helpers.py:218-238) and ray_bending.network /
// rigidity_network (run_nerf_helpers.py:411-482).
//
// Both operands are the fp16 chunk-major tile images written by the forward (activation stash) and
// DGRAD (gradient stash) kernels; read MN-major they are exactly the transposed operands WGRAD needs
// (sm100_ptx.cuh), so no transpose pass exists.  The contraction runs over points (K = 128 per tile).
//
// Decomposition: a fixed list of jobs (one per layer, table below); every job is split over a
// contiguous range of tiles per CTA ("split-K"), partial sums go to a scratch buffer and a second
// kernel reduces them in a fixed order (deterministic) while un-padding / un-permuting into the
// reference's parameter layout and dividing out the loss scale.
// This kernel is HBM-bound (each stash byte is used for 256 MACs ~ 4x below the ridge): the warps not
// needed for TMA / MMA issue compute the bias gradient from the same shared-memory stage.
#include "nrn_common.cuh"
#include "sm100_ptx.cuh"
#include "wgrad.cuh"

namespace nrn {

namespace {

constexpr long long kWaitLimitCycles = 1ll << 28;
constexpr int kWgStages = 3;
constexpr int kSubRows = 64;                       // points per pipeline stage
constexpr int kSubChunk = kSubRows * 16;           // 1 KB per chunk of a 64-row sub-image
constexpr int kStageBytes = 64 * kSubChunk;        // up to 32 A chunks + 32 B chunks
constexpr int kWgThreads = 320;                    // producer, mma, 8 bias/drain warps

struct Job {
  int a_off, a_cols;   // dY image in the gradient stash (M = output features)
  int b_off, b_cols;   // X image in the activation stash (N = input features)
  int bias;            // compute column sums of A (bias gradient)
};

// job ids: 0 head, 1..7 = L1..L7 (input h_l), 8 L5e, 9 L0, 10..14 = B4, B3, B2, B1, B0
__device__ __forceinline__ Job job_desc(int j) {
  switch (j) {
    case 0: return {kGsRaw, 16, kStH + 7 * kHBytes, 256, 1};
    case 8: return {kGsY + 5 * kHBytes, 256, kStE, 64, 0};
    case 9: return {kGsY + 0 * kHBytes, 256, kStE, 64, 1};
    case 10: return {kGsYb4, 16, kStHb4, 64, 0};
    case 11: return {kGsYb3, 64, kStHb3, 64, 1};
    case 12: return {kGsYb2, 80, kStHb2, 96, 1};
    case 13: return {kGsYb1, 96, kStHb1, 96, 1};
    case 14: return {kGsYb0, 96, kStBin, 48, 1};
    default: return {kGsY + j * kHBytes, 256, kStH + (j - 1) * kHBytes, 256, 1};  // L_j, j = 1..7
  }
}

struct Shared {
  uint64_t full[kWgStages];
  uint64_t empty[kWgStages];
  uint64_t done;
  uint32_t tmem_base;
  int abort_flag;
};

struct Waiter {
  int* s_abort;
  int* g_err;
  __device__ __forceinline__ bool wait(uint64_t* bar, uint32_t parity, int code) const {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
      if (*reinterpret_cast<volatile int*>(s_abort)) return false;
      if (clock64() - t0 > kWaitLimitCycles) {
        atomicExch(s_abort, code);
        atomicCAS(g_err, 0, code);
        return false;
      }
    }
    return true;
  }
};

// which (job, split) does this CTA own?  splits[] comes from the host (WgradParams)
__device__ __forceinline__ bool locate(const WgradParams& p, int cta, int& job, int& split, int& nsplit) {
  int base = 0;
  for (int j = 0; j < p.n_jobs; ++j) {
    if (cta < base + p.splits[j]) { job = j; split = cta - base; nsplit = p.splits[j]; return true; }
    base += p.splits[j];
  }
  return false;
}

}  // namespace

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Shared* sh = reinterpret_cast<Shared*>(smem + kWgStages * kStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int job_id = 0, split = 0, nsplit = 1;
  const bool have = locate(p, blockIdx.x, job_id, split, nsplit);
  const Job jb = job_desc(job_id);
  const int a_chunks = jb.a_cols / 8, b_chunks = jb.b_cols / 8;
  const int m_halves = jb.a_cols > 128 ? 2 : 1;
  // contiguous tile range of this split
  const int per = (p.n_tiles + nsplit - 1) / nsplit;
  const int t_begin = have ? min(split * per, p.n_tiles) : 0;
  const int t_end = have ? min(t_begin + per, p.n_tiles) : 0;
  const int n_stages_total = (t_end - t_begin) * 2;   // two 64-row sub-stages per tile

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWgStages; ++i) {
      mbar_init(&sh->full[i], 1);
      mbar_init(&sh->empty[i], 1 + 8);   // MMA commit + 8 bias warps
    }
    mbar_init(&sh->done, 1);
    sh->abort_flag = 0;
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&sh->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = sh->tmem_base;
  const Waiter W{&sh->abort_flag, p.err};

  if (warp == 0) {
    // ===================== producer: 64-row sub-images of dY and X =====================
    // all 32 lanes issue bulk copies (one 1 KB chunk each per operand): a single lane issuing up to 64
    // copies per stage would be the bottleneck of this HBM-bound kernel
    uint32_t stage = 0, phase = 0;
    for (int it = 0; it < n_stages_total; ++it) {
      const long long tile = t_begin + (it >> 1);
      const int sub = it & 1;
      if (lane == 0) {
        W.wait(&sh->empty[stage], phase ^ 1u, 101);
        mbar_arrive_expect_tx(&sh->full[stage], (a_chunks + b_chunks) * kSubChunk);
      }
      __syncwarp();
      uint8_t* dst = smem + stage * kStageBytes;
      const uint8_t* ga = p.gstash + tile * kGradTileBytes + jb.a_off + sub * kSubChunk;
      const uint8_t* gb = p.stash + tile * kStashTileBytes + jb.b_off + sub * kSubChunk;
      if (lane < a_chunks) tma_bulk_g2s(dst + lane * kSubChunk, ga + lane * kChunkBytes, kSubChunk, &sh->full[stage]);
      if (lane < b_chunks) tma_bulk_g2s(dst + (32 + lane) * kSubChunk, gb + lane * kChunkBytes, kSubChunk, &sh->full[stage]);
      if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t idesc = umma_instr_desc(128, jb.b_cols, UMMA_F16, UMMA_F16, UMMA_MN_MAJOR, UMMA_MN_MAJOR);
      for (int it = 0; it < n_stages_total; ++it) {
        W.wait(&sh->full[stage], phase, 201);
        tc_fence_after_sync();
        const uint32_t sa = smem_u32(smem + stage * kStageBytes);
        const uint32_t sb = sa + 32 * kSubChunk;
        for (int mh = 0; mh < m_halves; ++mh) {
          // MN-major: SBO = stride between 8-feature chunks, LBO = stride between 8-point groups
          const uint64_t adesc = umma_smem_desc(sa + mh * 16 * kSubChunk, 128, kSubChunk);
          const uint64_t bdesc = umma_smem_desc(sb, 128, kSubChunk);
          for (int k = 0; k < kSubRows / 16; ++k) {
            umma_f16_ss(tmem_base + mh * 256, umma_desc_advance(adesc, k * 256), umma_desc_advance(bdesc, k * 256), idesc,
                        (it | k) ? 1u : 0u);
          }
        }
        umma_commit(&sh->empty[stage]);
        if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
      }
      umma_commit(&sh->done);
    }
  } else {
    // ===================== bias column sums (8 warps), then accumulator drain =====================
    const int t = threadIdx.x - 64;           // 0..255
    const int c = t >> 3, g = t & 7;          // chunk, row group
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    uint32_t stage = 0, phase = 0;
    for (int it = 0; it < n_stages_total; ++it) {
      W.wait(&sh->full[stage], phase, 301);
      if (jb.bias) {   // warp-uniform; lanes whose chunk lies beyond the image contribute zeros
        const bool live = c < a_chunks;
        const uint8_t* src = smem + stage * kStageBytes + c * kSubChunk + g * 8 * 16;
        float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (live) {
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const uint4 w = *reinterpret_cast<const uint4*>(src + r * 16);
            const __half2* h = reinterpret_cast<const __half2*>(&w);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 f = __half22float2(h[q]);
              s[2 * q] += f.x; s[2 * q + 1] += f.y;
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          s[q] += __shfl_xor_sync(0xffffffffu, s[q], 1);
          s[q] += __shfl_xor_sync(0xffffffffu, s[q], 2);
          s[q] += __shfl_xor_sync(0xffffffffu, s[q], 4);
          acc[q] += s[q];
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->empty[stage]);
      if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
    }
    float* part = p.scratch + static_cast<size_t>(blockIdx.x) * kWgScratchFloats;
    if (have && jb.bias && g == 0 && c < a_chunks) {
#pragma unroll
      for (int q = 0; q < 8; ++q) part[65536 + c * 8 + q] = acc[q];
    }
    // drain: warps 2..5 own TMEM lane quarters (warp % 4)
    W.wait(&sh->done, 0, 302);
    tc_fence_after_sync();
    if (have && warp >= 2 && warp < 6) {
      const int q4 = warp & 3;
      for (int mh = 0; mh < m_halves; ++mh) {
        const int m = mh * 128 + q4 * 32 + lane;
        const uint32_t taddr = tmem_base + ((static_cast<uint32_t>(q4) * 32u) << 16) + mh * 256;
        for (int c0 = 0; c0 < jb.b_cols; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
          if (m < jb.a_cols && n_stages_total > 0) {
            float4* dst = reinterpret_cast<float4*>(part + m * 256 + c0);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                   __uint_as_float(v[4 * i + 3]));
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// Deterministic reduction of the split partials into the reference's parameter layout.
// One thread per destination element of the flat gradient buffers:
//   NeRF   : W0[256x63] b0 W1 b1 ... W7 b7 Wout[out_ch x 256] bout
//   bender : net_w0[64x35] net_b0 net_w1 net_b1 net_w2 net_b2 net_w3 net_b3 net_w4[3x64]
//            rig_w0[32x3] rig_b0 rig_w1[32x32] rig_b1 rig_w2[1x32] rig_b2
// ------------------------------------------------------------------------------------------------
namespace {

struct Src {
  int job, m, n, n2;  // element (m, n) of job's partial; n2 >= 0: add a second column; n == -1: bias[m]
};

__device__ __forceinline__ Src nerf_src(int idx, int out_ch, bool& ok) {
  ok = true;
  // layers 0..7
  const int sz0 = 256 * 63 + 256, szl = 256 * 256 + 256, sz5 = 256 * 319 + 256;
  if (idx < sz0) {
    if (idx < 256 * 63) return {9, idx / 63, idx % 63, -1};
    return {9, idx - 256 * 63, -1, -1};
  }
  idx -= sz0;
  for (int l = 1; l < 8; ++l) {
    const int sz = l == 5 ? sz5 : szl;
    if (idx < sz) {
      if (l == 5) {
        if (idx < 256 * 319) {
          const int m = idx / 319, k = idx % 319;
          return k < 63 ? Src{8, m, k, -1} : Src{5, m, k - 63, -1};
        }
        return {5, idx - 256 * 319, -1, -1};
      }
      if (idx < 65536) return {l, idx >> 8, idx & 255, -1};
      return {l, idx - 65536, -1, -1};
    }
    idx -= sz;
  }
  if (idx < out_ch * 256) {
    const int m = idx >> 8;
    ok = m < 4;   // output channel 4 never reaches the loss: zero gradient
    return {0, m, idx & 255, -1};
  }
  idx -= out_ch * 256;
  ok = idx < 4;
  return {0, idx, -1, -1};
}

__device__ __forceinline__ Src bender_src(int idx, bool& ok) {
  ok = true;
  if (idx < 64 * 35) {  // net_w0: xyz columns collect the hi and lo operand columns
    const int m = idx / 35, k = idx % 35;
    return k < 3 ? Src{14, m, k, k + 3} : Src{14, m, 6 + (k - 3), -1};
  }
  idx -= 64 * 35;
  if (idx < 64) return {14, idx, -1, -1};
  idx -= 64;
  if (idx < 4096) return {13, idx >> 6, idx & 63, -1};
  idx -= 4096;
  if (idx < 64) return {13, idx, -1, -1};
  idx -= 64;
  if (idx < 4096) return {12, idx >> 6, idx & 63, -1};
  idx -= 4096;
  if (idx < 64) return {12, idx, -1, -1};
  idx -= 64;
  if (idx < 4096) return {11, idx >> 6, idx & 63, -1};
  idx -= 4096;
  if (idx < 64) return {11, idx, -1, -1};
  idx -= 64;
  if (idx < 192) return {10, idx >> 6, idx & 63, -1};
  idx -= 192;
  if (idx < 96) return {14, 64 + idx / 3, idx % 3, idx % 3 + 3};       // rig_w0
  idx -= 96;
  if (idx < 32) return {14, 64 + idx, -1, -1};                          // rig_b0
  idx -= 32;
  if (idx < 1024) return {13, 64 + (idx >> 5), 64 + (idx & 31), -1};    // rig_w1
  idx -= 1024;
  if (idx < 32) return {13, 64 + idx, -1, -1};                          // rig_b1
  idx -= 32;
  if (idx < 32) return {12, 64, 64 + idx, -1};                          // rig_w2
  idx -= 32;
  return {12, 64, -1, -1};                                               // rig_b2
}

}  // namespace

__global__ void wgrad_reduce_kernel(const WgradParams p, float* __restrict__ nerf_grad, int nerf_n,
                                    float* __restrict__ bend_grad, int bend_n, int out_ch) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nerf_n + bend_n) return;
  bool ok;
  const Src s = idx < nerf_n ? nerf_src(idx, out_ch, ok) : bender_src(idx - nerf_n, ok);
  float scale = 1.0f;
  {
    const float amax = p.amax ? __ldg(p.amax) : 0.f;
    if (amax > 0.f && amax < 3.0e38f) {
      int e;
      frexpf(amax, &e);
      scale = ldexpf(1.0f, min(max(10 - e, -60), 60));
    }
  }
  float sum = 0.f;
  if (ok && s.job < p.n_jobs) {
    int base = 0;
    for (int j = 0; j < s.job; ++j) base += p.splits[j];
    const int per = (p.n_tiles + p.splits[s.job] - 1) / p.splits[s.job];
    for (int sp = 0; sp < p.splits[s.job]; ++sp) {
      if (sp * per >= p.n_tiles) break;   // this split owned no tiles: its scratch is unwritten
      const float* part = p.scratch + static_cast<size_t>(base + sp) * kWgScratchFloats;
      if (s.n < 0) sum += part[65536 + s.m];
      else {
        sum += part[s.m * 256 + s.n];
        if (s.n2 >= 0) sum += part[s.m * 256 + s.n2];
      }
    }
  }
  const float v = sum / scale;
  if (idx < nerf_n) nerf_grad[idx] = v; else bend_grad[idx - nerf_n] = v;
}

// ------------------------------------------------------------------------------------------------
cudaError_t launch_wgrad(WgradParams p, bool has_bender, int num_sms, float* nerf_grad, int nerf_n, float* bend_grad,
                         int bend_n, int out_ch, cudaStream_t st) {
  p.n_jobs = has_bender ? 15 : 10;
  // bytes per tile of every job -> proportional split of the CTAs
  static const int cols[15][2] = {{16, 256}, {256, 256}, {256, 256}, {256, 256}, {256, 256}, {256, 256}, {256, 256}, {256, 256},
                                  {256, 64}, {256, 64},  {16, 64},   {64, 64},   {80, 96},   {96, 96},   {96, 48}};
  long long total = 0;
  for (int j = 0; j < p.n_jobs; ++j) total += cols[j][0] + cols[j][1];
  int used = 0;
  for (int j = 0; j < p.n_jobs; ++j) {
    int s = static_cast<int>((static_cast<long long>(num_sms) * (cols[j][0] + cols[j][1])) / total);
    if (s < 1) s = 1;
    if (s > p.n_tiles) s = p.n_tiles > 0 ? p.n_tiles : 1;
    p.splits[j] = s;
    used += s;
  }
  // hand the remainder to the big jobs (L1..L7)
  for (int j = 1; used < num_sms && p.n_tiles > 0; j = j % 7 + 1) {
    if (p.splits[j] < p.n_tiles) { ++p.splits[j]; ++used; } else if (j == 7) break;
  }
  if (p.n_tiles > 0) {
    const size_t smem = kWgStages * kStageBytes + sizeof(Shared) + 64;
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    wgrad_kernel<<<used, kWgThreads, smem, st>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  const int n = nerf_n + bend_n;
  wgrad_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>(p, nerf_grad, nerf_n, bend_grad, bend_n, out_ch);
  return cudaGetLastError();
}

}  // namespace nrn
