// Weight gradients (WGRAD) of the fused field:  dW_l[out][in] = sum_p dY_l[p][out] * X_l[p][in],
// db_l[out] = sum_p dY_l[p][out]  -- what torch.autograd accumulates into the .grad of
// NeRF.pts_linears / output_linear (run_nerf_helpers.py:218-238) and ray_bending.network /
// rigidity_network (run_nerf_helpers.py:411-482).
//
// Both operands are the fp16 chunk-major tile images written by the forward (activation stash) and
// DGRAD (gradient stash) kernels; read MN-major they are exactly the transposed operands WGRAD needs
// (sm100_ptx.cuh), so no transpose pass exists.  The contraction runs over points (K = 128 per tile).
//
// Decomposition: a fixed list of jobs; a job is a set of up to three "sub-MMAs" over one gradient image block (A)
// and one activation image block (B) per tile (jobs 0-9: one NeRF layer, two 128-row halves of dW; jobs 10-11: the
// five small ray-bender layers, grouped so that their images are one contiguous range of each stash).  A pipeline
// stage holds ONE whole 128-point block, A and B alternating through a 3 x 64 KB ring, so every tile is fetched with
// a handful of 16 KB bulk copies (the first version streamed 64-point halves as 128 one-KB copies per tile and
// topped out at half the HBM bandwidth).  Every job is split over
// contiguous tile ranges ("split-K") proportionally to its bytes; partial sums go to a scratch buffer
// and a second kernel reduces them in a fixed order (deterministic) while un-padding / un-permuting
// into the reference's parameter layout and dividing out the loss scale.
// The kernel is HBM-bound (each stash byte feeds 256 MACs, ~4x below the ridge): the warps not needed
// for TMA / MMA issue compute the bias gradient from the same shared-memory stage.
//
// Compact mode (WgradParams.compact): the same two bender jobs over the tangent / adjoint stashes of
// the divergence regulariser (div.cu), which hold only the bender images.
#include "nrn_common.cuh"
#include "sm100_ptx.cuh"
#include "wgrad.cuh"

namespace nrn {

namespace {

constexpr long long kWaitLimitCycles = 1ll << 28;
constexpr int kWgStages = 3;
constexpr int kStageBytes = 32 * kChunkBytes;      // one block of up to 32 chunk images (128 points x 8 features each)
constexpr int kWgThreads = 320;                    // producer, mma, 8 bias/drain warps

struct Sub {
  int a_chunk, b_chunk, n, tmem_col;  // operand chunk offsets inside the A / B stage, N, accumulator column
  int a_cols;                         // valid rows of the 128-row accumulator
};
struct Job {
  int a_off, a_chunks;   // contiguous range of the gradient stash tile (bytes, chunks)
  int b_off, b_chunks;   // contiguous range of the activation stash tile
  int n_sub;
  Sub sub[3];
  int bias;              // compute column sums over all A chunks
};

// job ids: 0 head, 1..7 = L1..L7 (input h_l), 8 L5e, 9 L0, 10 = {B4,B3,B2}, 11 = {B1,B0}
__device__ __forceinline__ Job job_desc(int j, int compact) {
  Job jb{};
  if (j <= 9) {
    int a_off, a_cols, b_off, b_cols, bias = 1;
    if (j == 0) { a_off = kGsRaw; a_cols = 16; b_off = kStH + 7 * kHBytes; b_cols = 256; }
    else if (j == 8) { a_off = kGsY + 5 * kHBytes; a_cols = 256; b_off = kStE; b_cols = 64; bias = 0; }
    else if (j == 9) { a_off = kGsY; a_cols = 256; b_off = kStE; b_cols = 64; }
    else { a_off = kGsY + j * kHBytes; a_cols = 256; b_off = kStH + (j - 1) * kHBytes; b_cols = 256; }
    jb.a_off = a_off; jb.a_chunks = a_cols / 8; jb.b_off = b_off; jb.b_chunks = b_cols / 8;
    jb.n_sub = a_cols > 128 ? 2 : 1;
    jb.sub[0] = {0, 0, b_cols, 0, a_cols > 128 ? 128 : a_cols};
    jb.sub[1] = {16, 0, b_cols, 256, 128};
    jb.bias = bias;
    return jb;
  }
  // bender jobs: the images of a job are adjacent in both stashes (nrn_common.cuh); in compact mode the stashes hold
  // only the bender images
  const int ga = compact ? kGsYb4 : 0, sa = compact ? kStBin : 0;
  if (j == 10) {
    // A: Yb4 (2 chunks) Yb3 (8) Yb2 (10)      B: Hb2 (12) Hb3 (8) Hb4 (8)
    jb.a_off = kGsYb4 - ga; jb.a_chunks = 20;
    jb.b_off = kStHb2 - sa; jb.b_chunks = 28;
    jb.n_sub = 3;
    jb.sub[0] = {0, 20, 64, 0, 16};      // dYb4 x Hb4
    jb.sub[1] = {2, 12, 64, 64, 64};     // dYb3 x Hb3
    jb.sub[2] = {10, 0, 96, 128, 80};    // dYb2 x Hb2
  } else {
    // A: Yb1 (12) Yb0 (12)                    B: bender input (6) Hb1 (12)
    jb.a_off = kGsYb1 - ga; jb.a_chunks = 24;
    jb.b_off = kStBin - sa; jb.b_chunks = 18;
    jb.n_sub = 2;
    jb.sub[0] = {0, 6, 96, 0, 96};       // dYb1 x Hb1
    jb.sub[1] = {12, 0, 48, 96, 96};     // dYb0 x bender input
  }
  jb.bias = compact ? 0 : 1;   // the tangent chain of the divergence term has no bias
  return jb;
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

// Sum over the warp's 32 lanes of v[j] for each j; lane L returns the total of column L.
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = upper ? v[i] : v[i + n / 2];
      const float keep = upper ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// which (job, split) does this CTA own?  job_ids[] / splits[] come from the host (WgradParams)
__device__ __forceinline__ bool locate(const WgradParams& p, int cta, int& job, int& split, int& nsplit) {
  int base = 0;
  for (int j = 0; j < p.n_jobs; ++j) {
    if (cta < base + p.splits[j]) { job = p.job_ids[j]; split = cta - base; nsplit = p.splits[j]; return true; }
    base += p.splits[j];
  }
  return false;
}

}  // namespace

#ifdef NRN_TRACE
__device__ long long g_wg_prof[192 * 4];   // per CTA: job, tiles, cycles until the last MMA retired, total cycles
#endif

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
#ifdef NRN_TRACE
  const long long prof_t0 = clock64();
#endif
  Shared* sh = reinterpret_cast<Shared*>(smem + kWgStages * kStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int job_id = 0, split = 0, nsplit = 1;
  const bool have = locate(p, blockIdx.x, job_id, split, nsplit);
  const Job jb = job_desc(job_id, p.compact);
  // contiguous tile range of this split
  const int per = (p.n_tiles + nsplit - 1) / nsplit;
  const int t_begin = have ? min(split * per, p.n_tiles) : 0;
  const int t_end = have ? min(t_begin + per, p.n_tiles) : 0;
  const int n_stages_total = (t_end - t_begin) * 2;   // per tile: the A block, then the B block

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
    // ===================== producer: whole image blocks, 16 KB bulk copies =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int it = 0; it < n_stages_total; ++it) {
        const long long tile = t_begin + (it >> 1);
        const bool is_b = (it & 1) != 0;
        const uint8_t* src = is_b ? p.stash + tile * p.stash_tile_bytes + jb.b_off : p.gstash + tile * p.gstash_tile_bytes + jb.a_off;
        const uint32_t bytes = static_cast<uint32_t>(is_b ? jb.b_chunks : jb.a_chunks) * kChunkBytes;
        W.wait(&sh->empty[stage], phase ^ 1u, 101);
        mbar_arrive_expect_tx(&sh->full[stage], bytes);
        uint8_t* dst = smem + stage * kStageBytes;
        for (uint32_t o = 0; o < bytes; o += 16384u) tma_bulk_g2s(dst + o, src + o, bytes - o < 16384u ? bytes - o : 16384u, &sh->full[stage]);
        if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int it = 0; it < n_stages_total; it += 2) {
        const uint32_t sa = stage;
        W.wait(&sh->full[sa], phase, 201);
        if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
        const uint32_t sb = stage;
        W.wait(&sh->full[sb], phase, 202);
        if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
        tc_fence_after_sync();
        const uint32_t a0 = smem_u32(smem + sa * kStageBytes), b0 = smem_u32(smem + sb * kStageBytes);
        for (int s = 0; s < jb.n_sub; ++s) {
          const Sub sb_ = jb.sub[s];
          const uint32_t idesc = umma_instr_desc(128, sb_.n, UMMA_F16, UMMA_F16, UMMA_MN_MAJOR, UMMA_MN_MAJOR);
          // MN-major: SBO = stride between 8-feature chunks, LBO = stride between 8-point groups
          const uint64_t adesc = umma_smem_desc(a0 + sb_.a_chunk * kChunkBytes, 128, kChunkBytes);
          const uint64_t bdesc = umma_smem_desc(b0 + sb_.b_chunk * kChunkBytes, 128, kChunkBytes);
          for (int k = 0; k < kTileM / 16; ++k) {
            umma_f16_ss(tmem_base + sb_.tmem_col, umma_desc_advance(adesc, k * 256), umma_desc_advance(bdesc, k * 256), idesc,
                        (it | k) ? 1u : 0u);
          }
        }
        umma_commit(&sh->empty[sa]);
        umma_commit(&sh->empty[sb]);
      }
      umma_commit(&sh->done);
    }
  } else {
    // ===================== bias column sums (8 warps), then accumulator drain =====================
    // db[feature] = sum over points of dY: warp w owns chunks 4w .. 4w+3 of the A block; lane l reads rows l, l+32,
    // l+64, l+96 of each chunk (consecutive lanes = consecutive 16-byte rows: conflict-free, unlike a per-thread
    // row-group walk, which put all 32 lanes on the same four banks), keeps 4 x 8 partial sums and a warp
    // transpose-reduce leaves the total of (chunk 4w + L/8, feature L%8) in lane L.
    const int bw = warp - 2;                  // 0..7
    float acc = 0.f;
    uint32_t stage = 0, phase = 0;
    for (int it = 0; it < n_stages_total; ++it) {
      W.wait(&sh->full[stage], phase, 301);
      if (jb.bias && (it & 1) == 0) {   // warp-uniform; an A stage
        float s[32];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int c = bw * 4 + cc;
          const uint8_t* src = smem + stage * kStageBytes + c * kChunkBytes + lane * 16;
          float t8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (c < jb.a_chunks) {        // chunks beyond the block hold stale data
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const uint4 w = *reinterpret_cast<const uint4*>(src + r * 512);
              const __half2* h = reinterpret_cast<const __half2*>(&w);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float2 f = __half22float2(h[q]);
                t8[2 * q] += f.x; t8[2 * q + 1] += f.y;
              }
            }
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) s[cc * 8 + q] = t8[q];
        }
        acc += warp_transpose_reduce(s, lane);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->empty[stage]);
      if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
    }
    float* part = p.scratch + static_cast<size_t>(blockIdx.x) * kWgScratchFloats;
    if (have && jb.bias && (bw * 4 + (lane >> 3)) < jb.a_chunks) part[65536 + bw * 32 + lane] = acc;
    // drain the accumulators into this CTA's scratch partial once the last MMA has retired
    W.wait(&sh->done, 0, 302);
    tc_fence_after_sync();
#ifdef NRN_TRACE
    if (threadIdx.x == 64) { g_wg_prof[blockIdx.x * 4 + 0] = job_id; g_wg_prof[blockIdx.x * 4 + 1] = t_end - t_begin; g_wg_prof[blockIdx.x * 4 + 2] = clock64() - prof_t0; }
#endif
    if (have) {
      // all 8 warps drain: warp % 4 selects the TMEM lane quarter, (warp - 2) / 4 the half of the 16-column blocks
      const int q4 = warp & 3, hsel = (warp - 2) >> 2;
      const int m = q4 * 32 + lane;
      for (int s = 0; s < jb.n_sub; ++s) {
        const Sub sb = jb.sub[s];
        const uint32_t taddr = tmem_base + ((static_cast<uint32_t>(q4) * 32u) << 16) + sb.tmem_col;
        // scratch layout: NeRF jobs [(half*128 + m)][256]; bender jobs [sub][m][128]
        float* dst_row = job_id <= 9 ? part + (s * 128 + m) * 256 : part + s * 16384 + m * 128;
        const int nblk = sb.n / 16, nfirst = (nblk + 1) / 2;
        const int b0 = hsel ? nfirst : 0, b1 = hsel ? nblk : nfirst;
        for (int blk = b0; blk < b1; ++blk) {
          const int c0 = blk * 16;
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
          if (m < sb.a_cols && n_stages_total > 0) {
            float4* dst = reinterpret_cast<float4*>(dst_row + c0);
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
#ifdef NRN_TRACE
  if (threadIdx.x == 64) g_wg_prof[blockIdx.x * 4 + 3] = clock64() - prof_t0;
#endif
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}
#ifdef NRN_TRACE
extern "C" void dbg_wgrad_prof_read(long long* out) { cudaMemcpyFromSymbol(out, g_wg_prof, sizeof(long long) * 192 * 4); }
#endif

// ------------------------------------------------------------------------------------------------
// Deterministic reduction of the split partials into the reference's parameter layout.
// One thread per destination element of the flat gradient buffers:
//   NeRF   : W0[256x63] b0 W1 b1 ... W7 b7 Wout[out_ch x 256] bout
//   bender : net_w0[64x35] net_b0 net_w1 net_b1 net_w2 net_b2 net_w3 net_b3 net_w4[3x64]
//            rig_w0[32x3] rig_b0 rig_w1[32x32] rig_b1 rig_w2[1x32] rig_b2
// ------------------------------------------------------------------------------------------------
namespace {

struct Src {
  int job;   // job id
  int off;   // float offset inside the job's partial (weights) or bias index (bias == 1)
  int off2;  // second weight element to add (>= 0) or -1
  int bias;
};

__device__ __forceinline__ Src nerf_w(int job, int m, int n) { return {job, m * 256 + n, -1, 0}; }
__device__ __forceinline__ Src nerf_b(int job, int m) { return {job, m, -1, 1}; }
__device__ __forceinline__ Src bend_w(int job, int sub, int m, int n, int n2 = -1) {
  return {job, sub * 16384 + m * 128 + n, n2 >= 0 ? sub * 16384 + m * 128 + n2 : -1, 0};
}

__device__ __forceinline__ Src nerf_src(int idx, int out_ch, bool& ok) {
  ok = true;
  const int sz0 = 256 * 63 + 256, szl = 256 * 256 + 256, sz5 = 256 * 319 + 256;
  if (idx < sz0) {
    if (idx < 256 * 63) return nerf_w(9, idx / 63, idx % 63);
    return nerf_b(9, idx - 256 * 63);
  }
  idx -= sz0;
  for (int l = 1; l < 8; ++l) {
    const int sz = l == 5 ? sz5 : szl;
    if (idx < sz) {
      if (l == 5) {
        if (idx < 256 * 319) {
          const int m = idx / 319, k = idx % 319;
          return k < 63 ? nerf_w(8, m, k) : nerf_w(5, m, k - 63);
        }
        return nerf_b(5, idx - 256 * 319);
      }
      if (idx < 65536) return nerf_w(l, idx >> 8, idx & 255);
      return nerf_b(l, idx - 65536);
    }
    idx -= sz;
  }
  if (idx < out_ch * 256) {
    const int m = idx >> 8;
    ok = m < 4;   // output channel 4 never reaches the loss: zero gradient
    return nerf_w(0, m, idx & 255);
  }
  idx -= out_ch * 256;
  ok = idx < 4;
  return nerf_b(0, idx);
}

// A-chunk layout of the bender jobs (bias index = column inside the concatenated A images):
//   job 10: Yb4 cols 0-15 | Yb3 cols 16-79 | Yb2 cols 80-159      job 11: Yb1 cols 0-95 | Yb0 cols 96-191
__device__ __forceinline__ Src bender_src(int idx, bool& ok) {
  ok = true;
  if (idx < 64 * 35) {  // net_w0: xyz columns collect the hi and lo operand columns
    const int m = idx / 35, k = idx % 35;
    return k < 3 ? bend_w(11, 1, m, k, k + 3) : bend_w(11, 1, m, 6 + (k - 3));
  }
  idx -= 64 * 35;
  if (idx < 64) return {11, 96 + idx, -1, 1};                                    // net_b0
  idx -= 64;
  if (idx < 4096) return bend_w(11, 0, idx >> 6, idx & 63);                      // net_w1
  idx -= 4096;
  if (idx < 64) return {11, idx, -1, 1};                                         // net_b1
  idx -= 64;
  if (idx < 4096) return bend_w(10, 2, idx >> 6, idx & 63);                      // net_w2
  idx -= 4096;
  if (idx < 64) return {10, 80 + idx, -1, 1};                                    // net_b2
  idx -= 64;
  if (idx < 4096) return bend_w(10, 1, idx >> 6, idx & 63);                      // net_w3
  idx -= 4096;
  if (idx < 64) return {10, 16 + idx, -1, 1};                                    // net_b3
  idx -= 64;
  if (idx < 192) return bend_w(10, 0, idx >> 6, idx & 63);                       // net_w4
  idx -= 192;
  if (idx < 96) return bend_w(11, 1, 64 + idx / 3, idx % 3, idx % 3 + 3);        // rig_w0
  idx -= 96;
  if (idx < 32) return {11, 96 + 64 + idx, -1, 1};                               // rig_b0
  idx -= 32;
  if (idx < 1024) return bend_w(11, 0, 64 + (idx >> 5), 64 + (idx & 31));        // rig_w1
  idx -= 1024;
  if (idx < 32) return {11, 64 + idx, -1, 1};                                    // rig_b1
  idx -= 32;
  if (idx < 32) return bend_w(10, 2, 64, 64 + idx);                              // rig_w2
  idx -= 32;
  return {10, 80 + 64, -1, 1};                                                   // rig_b2
}

}  // namespace

__global__ void wgrad_reduce_kernel(const WgradParams p, const WgradDst dst, int out_ch) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int nerf_n = dst.nerf_n, bend_n = dst.bend_n;
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
  if (ok && !(s.bias && p.compact)) {
    int base = 0, slot = -1;
    for (int j = 0; j < p.n_jobs; ++j) {
      if (p.job_ids[j] == s.job) { slot = j; break; }
      base += p.splits[j];
    }
    if (slot >= 0) {
      const int nsplit = p.splits[slot];
      const int per = (p.n_tiles + nsplit - 1) / nsplit;
      const int n_valid = min(nsplit, (p.n_tiles + per - 1) / per);   // splits beyond that owned no tiles: scratch unwritten
      const float* __restrict__ src = p.scratch + static_cast<size_t>(base) * kWgScratchFloats + (s.bias ? 65536 + s.off : s.off);
      const bool has2 = !s.bias && s.off2 >= 0;
      const int d2 = has2 ? s.off2 - s.off : 0;
      // fixed summation order (split 0, 1, 2, ...) = deterministic gradients; the loads of 8 splits are in flight together
      int sp = 0;
      for (; sp + 8 <= n_valid; sp += 8) {
        float a[8], b[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float* q = src + static_cast<size_t>(sp + i) * kWgScratchFloats;
          a[i] = __ldg(q);
          b[i] = has2 ? __ldg(q + d2) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          sum += a[i];
          if (has2) sum += b[i];
        }
      }
      for (; sp < n_valid; ++sp) {
        const float* q = src + static_cast<size_t>(sp) * kWgScratchFloats;
        sum += __ldg(q);
        if (has2) sum += __ldg(q + d2);
      }
    }
  }
  const float v = sum / scale;
  float* out;
  bool acc;
  if (idx < nerf_n) {
    const int n_pts = nerf_n - out_ch * 257;
    out = (idx >= n_pts && dst.nerf_head) ? dst.nerf_head + (idx - n_pts) : dst.nerf + idx;
    acc = dst.acc_nerf != 0;
  } else {
    out = dst.bend + (idx - nerf_n);
    acc = dst.acc_bend != 0;
  }
  *out = acc ? *out + v : v;
}

// ------------------------------------------------------------------------------------------------
cudaError_t launch_wgrad(WgradParams p, bool has_bender, int num_sms, const WgradDst& dst, int out_ch, cudaStream_t st) {
  // relative cost of one tile of every job = 2 KB chunks it moves; the head job (a 4 KB gradient block alternating
  // with a 64 KB activation block keeps fewer bytes in flight) and the three-MMA bender job stream a little slower
  // per byte (scripts/trace_wgrad.py), hence their surcharge
  static const int kJobChunks[12] = {46, 64, 64, 64, 64, 64, 64, 64, 32 + 8, 32 + 8, 54, 24 + 18};
  int first = 0, last = has_bender ? 12 : 10;
  if (p.compact) { first = 10; last = 12; p.stash_tile_bytes = kTanTileBytes; p.gstash_tile_bytes = kAdjTileBytes; }
  else { p.stash_tile_bytes = kStashTileBytes; p.gstash_tile_bytes = kGradTileBytes; }
  p.n_jobs = last - first;
  // Every CTA streams at the same bytes/clk (the kernel is HBM-bound), so the launch ends when the CTA with the most
  // bytes ends: start with one CTA per job and hand each further CTA to the job whose CTAs currently carry the most
  // (chunks per tile x tiles per CTA).
  int used = 0;
  for (int j = first; j < last; ++j) { p.job_ids[j - first] = j; p.splits[j - first] = 1; ++used; }
  const int tiles = p.n_tiles > 0 ? p.n_tiles : 1;
  while (used < num_sms) {
    int best = -1;
    long long best_load = -1;
    for (int j = first; j < last; ++j) {
      const int sp = p.splits[j - first];
      if (sp >= tiles) continue;   // a CTA needs at least one tile
      const long long load = static_cast<long long>(kJobChunks[j]) * ((tiles + sp - 1) / sp);
      if (load > best_load) { best_load = load; best = j; }
    }
    if (best < 0) break;
    ++p.splits[best - first];
    ++used;
  }
  if (p.n_tiles > 0) {
    const size_t smem = kWgStages * kStageBytes + sizeof(Shared) + 64;
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    wgrad_kernel<<<used, kWgThreads, smem, st>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  const int n = dst.nerf_n + dst.bend_n;
  if (n > 0) wgrad_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>(p, dst, out_ch);
  return cudaGetLastError();
}

}  // namespace nrn
