// Fused point-wise field evaluation, forward:  (ray, z) -> bent point -> PE -> 8x256 MLP -> raw.
//
// Replaces, for one coarse or fine pass, the reference call chain
//   run_network (train.py:57-105) -> batchify (train.py:27-54) -> NeRF.forward
//   (run_nerf_helpers.py:240-314) -> ray_bending.forward (run_nerf_helpers.py:507-584)
//   -> Embedder.embed (run_nerf_helpers.py:149-150)
// with ONE persistent sm_100a kernel.  No [P,95] / [P,256] tensor ever touches HBM.
//
// Work decomposition
//   tile   = 128 consecutive sample points (= 128 TMEM lanes = UMMA M)
//   CTA    = 1 per SM, persistent; processes tile PAIRS (slot 0 / slot 1) so that the epilogue of
//            one slot overlaps the tensor-core work of the other
//   warps  : 0 weight producer (bulk TMA ring), 1 MMA issuer (tcgen05.mma, one lane),
//            2 TMEM allocator, 3 idle, 4-7 epilogue warpgroup of slot 0, 8-11 of slot 1
//   a "step" = one dense layer for one slot:  D[128 x N] (TMEM, fp32) = A[128 x K] (smem, fp16)
//            . W[N x K]^T (smem ring, fp16); the epilogue warpgroup drains D, applies
//            bias/ReLU/(bend + positional encoding), and writes the next A operand in place.
//   steps  : B0..B4 (ray bender, offset + rigidity MLPs fused block-diagonally), L0..L7, head
//
// Shared memory (per CTA): 2 x (H 64 KB + E 16 KB) activations + 2 x 32 KB weight ring + barriers.
// Tensor memory: 512 columns, 256 per slot.
#include "nrn_common.cuh"
#include "sm100_ptx.cuh"

namespace nrn {

namespace {

constexpr long long kWaitLimitCycles = 1ll << 28;  // ~0.14 s: protocol bug => error flag, not a hang

struct Shared {
  uint64_t w_full[kRingStages];
  uint64_t w_empty[kRingStages];
  uint64_t a_ready[2];
  uint64_t d_full[2];
  uint32_t tmem_base;
  int abort_flag;
};

struct StepShape {
  uint32_t N, nslabs, slab_bytes, k16;
};

// step index: 0-4 bender B0..B4, 5-12 NeRF L0..L7, 13 head
__device__ __forceinline__ StepShape step_shape(int step) {
  switch (step) {
    case 0: return {96u, 1u, (uint32_t)kBendB0Bytes, 3u};
    case 1: return {96u, 1u, (uint32_t)kBendB1Bytes, 6u};
    case 2: return {80u, 1u, (uint32_t)kBendB2Bytes, 6u};
    case 3: return {64u, 1u, (uint32_t)kBendB3Bytes, 4u};
    case 4: return {16u, 1u, (uint32_t)kBendB4Bytes, 4u};
    case 5: return {256u, 1u, 32768u, 4u};
    case 10: return {256u, 5u, 32768u, 4u};
    case 13: return {16u, 1u, (uint32_t)kNerfHeadBytes, 16u};
    default: return {256u, 4u, 32768u, 4u};
  }
}
// byte offset (inside a slot's activation region: H at 0, E at kHBytes) of the A operand of slab j
__device__ __forceinline__ uint32_t a_operand_offset(int step, uint32_t j) {
  if (step == 0 || step == 5) return kHBytes;                       // bender input / embedding live in E
  if (step == 10) return j == 0 ? kHBytes : (j - 1) * 8 * kChunkBytes;  // skip: [embedding | h]
  if (step < 5 || step == 13) return 0;
  return j * 8 * kChunkBytes;
}

// Developer profile (NRN_DEBUG_MODE=9): cycles CTA 0's roles spend waiting, read with nrn_debug_profile_fwd1().
//   [0] issuer total  [1] issuer: a_ready  [3] issuer: w_full  [4] producer: w_empty  [5] epilogue WG (slot 0) total
//   [6] epilogue: d_full  [10] slabs issued  [11] tile pairs
__device__ unsigned long long g_fwd1_prof[16];

struct Waiter {
  int* s_abort;
  int* g_err;
  __device__ __forceinline__ bool wait(uint64_t* bar, uint32_t parity, int code, unsigned long long* acc = nullptr) const {
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
    if (acc) *acc += static_cast<unsigned long long>(clock64() - t0);
    return true;
  }
};

__device__ __forceinline__ float relu_h(float v) { return fminf(fmaxf(v, 0.f), 65504.f); }

// Drain NCOLS accumulator columns (multiple of 32), add bias, ReLU, convert to fp16 and store them
// as chunks [chunk0, chunk0 + NCOLS/8) of this thread's row in a chunk-major activation image.
template <int NCOLS>
__device__ __forceinline__ void epi_bias_relu_store(uint32_t taddr, const float* __restrict__ bias,
                                                    uint8_t* dst_row) {
  // software pipeline: the TMEM load of chunk c+1 and the bias loads of chunk c are in flight while
  // chunk c-1 / c is being processed (tcgen05.wait::ld only ever waits for a load issued a block ago)
  constexpr int NC = NCOLS / 32;
  uint32_t v[2][32];
  tmem_ld32(taddr, v[0]);
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    float4 b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = __ldg(reinterpret_cast<const float4*>(bias + c * 32 + i * 4));
    tmem_ld_wait();
    if (c + 1 < NC) tmem_ld32(taddr + (c + 1) * 32, v[(c + 1) & 1]);
    const uint32_t(&w)[32] = v[c & 1];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b0 = b[2 * q], b1 = b[2 * q + 1];
      uint4 pk;
      // one cvt.rn.relu.satfinite.f16x2 per two outputs: ReLU, clamp to fp16 range and pack in a single instruction
      pk.x = pack_h2_relu_sat(__uint_as_float(w[q * 8 + 0]) + b0.x, __uint_as_float(w[q * 8 + 1]) + b0.y);
      pk.y = pack_h2_relu_sat(__uint_as_float(w[q * 8 + 2]) + b0.z, __uint_as_float(w[q * 8 + 3]) + b0.w);
      pk.z = pack_h2_relu_sat(__uint_as_float(w[q * 8 + 4]) + b1.x, __uint_as_float(w[q * 8 + 5]) + b1.y);
      pk.w = pack_h2_relu_sat(__uint_as_float(w[q * 8 + 6]) + b1.z, __uint_as_float(w[q * 8 + 7]) + b1.w);
      *reinterpret_cast<uint4*>(dst_row + (c * 4 + q) * kChunkBytes) = pk;
    }
  }
}

// Positional encoding of one point (Embedder.embed, run_nerf_helpers.py:149-150 with the settings of
// get_embedder :157-164: raw xyz first, then per octave sin(2^k xyz), cos(2^k xyz); k = 0..9).
// Written as fp16 chunks 0..7 of the row (63 features + one zero pad column).
// sin/cos: the argument 2^k * x is reduced EXACTLY to [-0.5, 0.5) turns (x / 2pi carried as a
// two-float value), then evaluated with MUFU (abs err ~4e-7), well below fp16 resolution.
__device__ __forceinline__ void write_pe(const float (&x)[3], uint8_t* dst_row) {
  float f[64];
  f[0] = x[0]; f[1] = x[1]; f[2] = x[2];
  const float kInv2PiHi = 0.15915494f;      // fl(1/2pi)
  const float kInv2PiLo = 6.4206199e-09f;   // 1/2pi - fl(1/2pi)
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float thi = x[d] * kInv2PiHi;
    const float tlo = fmaf(x[d], kInv2PiLo, fmaf(x[d], kInv2PiHi, -thi));
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      const float sc = static_cast<float>(1 << k);
      const float a = thi * sc;
      const float ph = (a - rintf(a)) + tlo * sc;
      const float ang = ph * 6.2831853071795865f;
      f[3 + 6 * k + d] = __sinf(ang);
      f[3 + 6 * k + 3 + d] = __cosf(ang);
    }
  }
  f[63] = 1.f;  // pad column: its weight column is zero (forward unaffected); WGRAD reads it as the bias input
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint4 pk;
    pk.x = pack_h2(f[c * 8 + 0], f[c * 8 + 1]);
    pk.y = pack_h2(f[c * 8 + 2], f[c * 8 + 3]);
    pk.z = pack_h2(f[c * 8 + 4], f[c * 8 + 5]);
    pk.w = pack_h2(f[c * 8 + 6], f[c * 8 + 7]);
    *reinterpret_cast<uint4*>(dst_row + c * kChunkBytes) = pk;
  }
}

}  // namespace

template <bool HAS_BENDER>
__global__ void __launch_bounds__(kFwdThreads, 1) field_fwd_kernel(const FieldFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* act = smem;                                  // 2 slots x (H | E)
  uint8_t* ring = smem + 2 * kSlotBytes;                // kRingStages x 32 KB
  Shared* sh = reinterpret_cast<Shared*>(ring + kRingStages * kRingStageBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_pairs = (p.n_tiles + 1) >> 1;
  constexpr int kFirstStep = HAS_BENDER ? 0 : 5;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kRingStages; ++i) {
      mbar_init(&sh->w_full[i], 1);
      mbar_init(&sh->w_empty[i], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sh->a_ready[s], 128);
      mbar_init(&sh->d_full[s], 1);
    }
    sh->abort_flag = 0;
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&sh->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = sh->tmem_base;
  const Waiter W{&sh->abort_flag, p.err};
  const bool prof = p.debug_mode == 9 && blockIdx.x == 0;

  if (warp == 0) {
    // ===================== weight producer: global -> smem ring (bulk TMA) =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      unsigned long long pw = 0;
      for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        uint32_t gb = 0, gn = 0;
#pragma unroll 1
        for (int step = kFirstStep; step < 14; ++step) {
          const StepShape s = step_shape(step);
          const uint8_t* src = step < 5 ? p.bend_w + gb : p.nerf_w + gn;
          for (int slot = 0; slot < 2; ++slot) {
            for (uint32_t j = 0; j < s.nslabs; ++j) {
              W.wait(&sh->w_empty[stage], phase ^ 1u, 101, prof ? &pw : nullptr);
              uint8_t* dst = ring + stage * kRingStageBytes;
              mbar_arrive_expect_tx(&sh->w_full[stage], s.slab_bytes);
              const uint8_t* g = src + j * s.slab_bytes;
              for (uint32_t off = 0; off < s.slab_bytes; off += 16384u) {
                const uint32_t n = s.slab_bytes - off < 16384u ? s.slab_bytes - off : 16384u;
                tma_bulk_g2s(dst + off, g + off, n, &sh->w_full[stage]);
              }
              if (++stage == kRingStages) { stage = 0; phase ^= 1u; }
            }
          }
          if (step < 5) gb += s.nslabs * s.slab_bytes; else gn += s.nslabs * s.slab_bytes;
        }
      }
      if (prof) g_fwd1_prof[4] = pw;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      uint32_t aph[2] = {0u, 0u};
      unsigned long long w_a = 0, w_w = 0, n_slabs = 0, n_done = 0;
      const long long t_start = clock64();
      for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        ++n_done;
#pragma unroll 1
        for (int step = kFirstStep; step < 14; ++step) {
          const StepShape s = step_shape(step);
          const uint32_t idesc = umma_instr_desc(kTileM, s.N, UMMA_F16, UMMA_F16, UMMA_K_MAJOR, UMMA_K_MAJOR);
          for (int slot = 0; slot < 2; ++slot) {
            W.wait(&sh->a_ready[slot], aph[slot], 201, prof ? &w_a : nullptr);
            aph[slot] ^= 1u;
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + slot * 256;
            const uint32_t a_base = smem_u32(act + slot * kSlotBytes);
            for (uint32_t j = 0; j < s.nslabs; ++j) {
              ++n_slabs;
              W.wait(&sh->w_full[stage], phase, 202, prof ? &w_w : nullptr);
              tc_fence_after_sync();
              const uint64_t adesc = umma_smem_desc(a_base + a_operand_offset(step, j), kChunkBytes, 128);
              const uint64_t bdesc = umma_smem_desc(smem_u32(ring + stage * kRingStageBytes), s.N * 16, 128);
              for (uint32_t k = 0; k < s.k16 && p.debug_mode != 2; ++k) {
                umma_f16_ss(d_tmem, umma_desc_advance(adesc, k * 2 * kChunkBytes),
                            umma_desc_advance(bdesc, k * 2 * s.N * 16), idesc, (j | k) ? 1u : 0u);
              }
              umma_commit(&sh->w_empty[stage]);  // slab free once these MMAs retire
              if (++stage == kRingStages) { stage = 0; phase ^= 1u; }
            }
            umma_commit(&sh->d_full[slot]);      // accumulator complete
          }
        }
      }
      if (prof) {
        g_fwd1_prof[0] = static_cast<unsigned long long>(clock64() - t_start);
        g_fwd1_prof[1] = w_a; g_fwd1_prof[3] = w_w; g_fwd1_prof[10] = n_slabs; g_fwd1_prof[11] = n_done;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warpgroups =====================
    const int slot = (warp - 4) >> 2;
    const int row = ((warp & 3) << 5) | lane;
    uint8_t* Hs = act + slot * kSlotBytes;
    uint8_t* Es = Hs + kHBytes;
    uint8_t* h_row = Hs + row * 16;
    uint8_t* e_row = Es + row * 16;
    const uint32_t taddr = tmem_base + ((static_cast<uint32_t>(warp & 3) * 32u) << 16) + slot * 256;
    uint32_t dph = 0;
    unsigned long long w_d = 0;
    const bool tprof = prof && slot == 0 && (threadIdx.x & 127) == 0;
    const long long t_start = clock64();
    auto signal_ready = [&]() {
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(&sh->a_ready[slot]);
    };
    auto wait_acc = [&](int code) {
      W.wait(&sh->d_full[slot], dph, code, tprof ? &w_d : nullptr);
      dph ^= 1u;
      tc_fence_after_sync();
    };

    for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      const long long pt = (static_cast<long long>(pair) * 2 + slot) * kTileM + row;
      const bool valid = pt < p.P;
      // Training stash: every finished activation image (a contiguous chunk-major block of shared memory) is
      // written to this tile's stash block with bulk TMA stores issued by one thread of the warpgroup; the
      // epilogue threads spend no load/store slots on it.  stash_begin(): the previous store must have finished
      // READING shared memory before any image is overwritten.
      uint8_t* st = p.stash ? p.stash + (static_cast<long long>(pair) * 2 + slot) * kStashTileBytes : nullptr;
      const bool wg_leader = (threadIdx.x & 127) == 0;
      auto stash_begin = [&]() {
        if (st) {
          if (wg_leader) tma_bulk_wait_read<0>();
          asm volatile("bar.sync %0, 128;" ::"r"(1 + slot) : "memory");
        }
      };
      auto stash_store = [&](uint32_t stash_off, const uint8_t* img, uint32_t bytes) {
        if (st) {
          fence_proxy_async_smem();
          asm volatile("bar.sync %0, 128;" ::"r"(1 + slot) : "memory");
          if (wg_leader) {
            for (uint32_t o = 0; o < bytes; o += 16384u) tma_bulk_s2g(st + stash_off + o, img + o, bytes - o < 16384u ? bytes - o : 16384u);
            tma_bulk_commit();
          }
        }
      };
      float x[3] = {0.f, 0.f, 0.f};
      long long ray = 0;
      if (valid) {
        ray = pt / p.S;
        if (p.pts) {
          const float* q = p.pts + pt * p.pts_stride;  // point mode: NeRF.forward(x) reads x[:, :3]
          x[0] = __ldg(q + 0); x[1] = __ldg(q + 1); x[2] = __ldg(q + 2);
        } else {
          const float z = __ldg(p.z_vals + pt);
          const float* r = p.rays + ray * 8;
          // pts = rays_o + rays_d * z  (train.py:871-873), multiply then add like the reference
          x[0] = __fadd_rn(__ldg(r + 0), __fmul_rn(__ldg(r + 3), z));
          x[1] = __fadd_rn(__ldg(r + 1), __fmul_rn(__ldg(r + 4), z));
          x[2] = __fadd_rn(__ldg(r + 2), __fmul_rn(__ldg(r + 5), z));
        }
        if (p.d_init) {
          p.d_init[pt * 3 + 0] = x[0]; p.d_init[pt * 3 + 1] = x[1]; p.d_init[pt * 3 + 2] = x[2];
        }
      }
      float rigidity = 0.f;
      if (HAS_BENDER) {
        // ---- bender input row: [xyz_hi(3) xyz_lo(3) latent(32) 0(10)] fp16, chunks 0..5 of E ----
        stash_begin();
        float in[48];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const float hi = __half2float(__float2half_rn(x[d]));
          in[d] = hi;
          in[3 + d] = x[d] - hi;
        }
        const float* lat = p.latents + ray * p.latent_stride;
#pragma unroll
        for (int i = 0; i < kLatent; ++i) in[6 + i] = valid ? __ldg(lat + i) : 0.f;
#pragma unroll
        for (int i = 38; i < 48; ++i) in[i] = 0.f;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          uint4 pk;
          pk.x = pack_h2(in[c * 8 + 0], in[c * 8 + 1]);
          pk.y = pack_h2(in[c * 8 + 2], in[c * 8 + 3]);
          pk.z = pack_h2(in[c * 8 + 4], in[c * 8 + 5]);
          pk.w = pack_h2(in[c * 8 + 6], in[c * 8 + 7]);
          *reinterpret_cast<uint4*>(e_row + c * kChunkBytes) = pk;
        }
        stash_store(kStBin, Es, 6 * kChunkBytes);
        signal_ready();
        // ---- B0, B1: 96 hidden units (64 offset | 32 rigidity) ----
        wait_acc(301);
        stash_begin();
        epi_bias_relu_store<96>(taddr, p.bend_bias, h_row);
        stash_store(kStHb1, Hs, 12 * kChunkBytes);
        signal_ready();
        wait_acc(302);
        stash_begin();
        epi_bias_relu_store<96>(taddr, p.bend_bias + 96, h_row);
        stash_store(kStHb2, Hs, 12 * kChunkBytes);
        signal_ready();
        // ---- B2: 64 offset hidden + rigidity output (column 64) ----
        wait_acc(303);
        stash_begin();
        epi_bias_relu_store<64>(taddr, p.bend_bias + 192, h_row);
        stash_store(kStHb3, Hs, 8 * kChunkBytes);
        {
          uint32_t v[16];
          tmem_ld16(taddr + 64, v);
          tmem_ld_wait();
          const float rr = __uint_as_float(v[0]) + __ldg(p.bend_bias + 192 + 64);
          rigidity = (tanhf(rr) + 1.0f) * 0.5f;   // run_nerf_helpers.py:559-561
          if (p.use_cutoff && rigidity <= p.cutoff) rigidity = 0.f;  // :563-564
        }
        signal_ready();
        // ---- B3 ----
        wait_acc(304);
        stash_begin();
        epi_bias_relu_store<64>(taddr, p.bend_bias + 272, h_row);
        stash_store(kStHb4, Hs, 8 * kChunkBytes);
        signal_ready();
        // ---- B4: offsets; bend; positional encoding of the bent point -> E ----
        wait_acc(305);
        {
          uint32_t v[16];
          tmem_ld16(taddr, v);
          tmem_ld_wait();
          float un[3], ma[3];
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            un[d] = __uint_as_float(v[d]);
            ma[d] = __fmul_rn(rigidity, un[d]);              // :567
            if (p.use_scaling) ma[d] = __fmul_rn(ma[d], p.scaling);  // :568-569
          }
          if (valid) {
            if (p.d_unmasked) { p.d_unmasked[pt * 3 + 0] = un[0]; p.d_unmasked[pt * 3 + 1] = un[1]; p.d_unmasked[pt * 3 + 2] = un[2]; }
            if (p.d_masked) { p.d_masked[pt * 3 + 0] = ma[0]; p.d_masked[pt * 3 + 1] = ma[1]; p.d_masked[pt * 3 + 2] = ma[2]; }
            if (p.d_rigid) p.d_rigid[pt] = rigidity;
          }
#pragma unroll
          for (int d = 0; d < 3; ++d) x[d] = __fadd_rn(x[d], ma[d]);  // :570
        }
      }
      if (valid && p.d_bent) {
        p.d_bent[pt * 3 + 0] = x[0]; p.d_bent[pt * 3 + 1] = x[1]; p.d_bent[pt * 3 + 2] = x[2];
      }
      stash_begin();
      write_pe(x, e_row);
      stash_store(kStE, Es, kEBytes);
      signal_ready();
      // ---- L0 .. L7 ----
#pragma unroll 1
      for (int L = 0; L < 8; ++L) {
        wait_acc(310 + L);
        stash_begin();
        if (p.debug_mode != 1) epi_bias_relu_store<256>(taddr, p.nerf_bias + L * 256, h_row);
        stash_store(kStH + L * kHBytes, Hs, kHBytes);
        signal_ready();
      }
      // ---- head: raw = output_linear(h) (run_nerf_helpers.py:306) ----
      wait_acc(320);
      {
        uint32_t v[16];
        tmem_ld16(taddr, v);
        tmem_ld_wait();
        if (valid) {
          float o[5];
#pragma unroll
          for (int c = 0; c < 5; ++c) o[c] = __uint_as_float(v[c]) + __ldg(p.nerf_bias + 2048 + c);
          // test-time non-rigid object removal (run_nerf_helpers.py:309-310)
          if (HAS_BENDER && p.use_removal && rigidity >= p.removal) o[3] *= 0.f;
          float* dst = p.raw + pt * p.out_ch;
          for (int c = 0; c < p.out_ch; ++c) dst[c] = o[c];
        }
      }
      // the next a_ready arrival is the next pair's prologue (which also means TMEM is drained)
    }
    if (p.stash && (threadIdx.x & 127) == 0) tma_bulk_wait<0>();   // all stash stores complete before the CTA exits
    if (tprof) { g_fwd1_prof[5] = static_cast<unsigned long long>(clock64() - t_start); g_fwd1_prof[6] = w_d; }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
extern "C" int nrn_debug_profile_fwd1(unsigned long long* out16) {
  return cudaMemcpyFromSymbol(out16, g_fwd1_prof, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : -2;
}

size_t field_fwd_smem_bytes() { return 2 * kSlotBytes + kRingStages * kRingStageBytes + sizeof(Shared) + 64; }

cudaError_t launch_field_fwd(const FieldFwdParams& p, bool has_bender, int num_sms, cudaStream_t stream) {
  const size_t smem = field_fwd_smem_bytes();
  const int n_pairs = (p.n_tiles + 1) / 2;
  if (n_pairs <= 0) return cudaSuccess;
  const int grid = n_pairs < num_sms ? n_pairs : num_sms;
  cudaError_t e;
  if (has_bender) {
    e = cudaFuncSetAttribute(field_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    field_fwd_kernel<true><<<grid, kFwdThreads, smem, stream>>>(p);
  } else {
    e = cudaFuncSetAttribute(field_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    field_fwd_kernel<false><<<grid, kFwdThreads, smem, stream>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace nrn
