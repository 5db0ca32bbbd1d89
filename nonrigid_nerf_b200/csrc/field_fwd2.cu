// Fused point-wise field evaluation, forward, CTA-pair version:  (ray, z) -> bent point -> PE -> 8x256 MLP -> raw.
//
// Same contract and arithmetic as field_fwd.cu (reference chain: run_network train.py:57-105 -> batchify
// train.py:27-54 -> NeRF.forward run_nerf_helpers.py:240-314 -> ray_bending.forward :507-584 -> Embedder.embed
// :149-150); what changes is how the machine is kept busy.  Measurements on B200 (tests/cuda/umma_rate_probe.cu,
// umma_interf_probe.cu, scripts/trace_field.py) showed that
//   * the tensor pipe retires one 128 x 256 x 16 MMA per 128 cycles from this operand layout no matter what
//     the other warps do, but accepts new MMAs only as fast as it retires them (the issuing thread blocks),
//   * the single-CTA kernel was bound by the per-slot dependency chain
//       MMA block (ring-latency bound: a 2 x 32 KB ring cannot cover the ~0.8 us slab reload)
//       -> epilogue (one warpgroup, ~1.7 us per 128 x 256 tile) -> handshake -> next MMA block.
// This version (opt-in, NRN_PAIR=1; bit-identical output, tests/test_field_forward_gpu.py) attacks every link of that
// chain.  Measured outcome on B200: 8.4 ms vs 8.0 ms of field_fwd.cu for 65,536 x 128 points without bender, 9.9 vs
// 9.8 ms with, 0.394 vs 0.400 ms per training step -- a wash, because with a ring of exactly one layer slot 0's next
// block waits for the reload that slot 1's use of the stage releases.  Kept as the starting
// point for the deeper-ring version and as the repo's reference for the cluster / cta_group::2 protocol:
//   * CTA pairs (cluster of 2, tcgen05 cta_group::2, M = 256): a CTA holds only its N/2 rows of every weight
//     slab, so the 64 KB ring holds a whole layer.  A slab is loaded ONCE per tile group, used by slot 0 and by
//     slot 1, and recycled when both have consumed it: no ring-latency stalls, half the weight traffic.
//   * one MMA-issuing thread per slot (their bookkeeping overlaps the other slot's MMAs),
//   * two epilogue warpgroups per slot (each drains half of the accumulator columns): 16 epilogue warps.
//
// Work decomposition
//   tile    = 128 consecutive sample points (= 128 TMEM lanes)
//   group   = 4 tiles = {slot 0, slot 1} x {leader CTA, peer CTA}; tile (slot, rank) = (group*2 + slot)*2 + rank
//   cluster = 2 CTAs on 2 SMs, persistent over groups
//   warps   : 0 weight producer (bulk TMA), 1 issuer of slot 0 (leader) / landing relay (peer), 2 TMEM allocator,
//             3 issuer of slot 1 (leader), 4-11 epilogue of slot 0 (two warpgroups), 12-19 epilogue of slot 1
//   issue steps: B0..B4 (ray bender), L0..L4, L5 embedding part, L5 hidden part, L6, L7, head
// Shared memory per CTA: 2 x (H 64 KB + E 16 KB) activations + 4 x 16 KB weight ring.  TMEM: 2 x 256 columns.
#include "nrn_common.cuh"
#include "sm100_ptx.cuh"

namespace nrn {

#ifdef NRN_TRACE
// developer instrumentation (make EXTRA=-DNRN_TRACE): time stamps of the handshake chain of cluster 0, third work
// group.  Every traced thread owns a log region and writes with plain stores (no atomics: tracing must not stall).
// roles: 0/1 issuer of slot 0/1, 2/3 epilogue leader of slot 0/1
__device__ unsigned long long g_trace2_buf[2 * 4 * 256];
__device__ __forceinline__ void trace2_ev(bool on, uint32_t role, uint32_t& cnt, uint32_t ev, uint32_t step, uint32_t slot, uint32_t j) {
  if (!on || cnt >= 256u) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  g_trace2_buf[((blockIdx.x & 1) * 4 + role) * 256 + cnt++] = (t << 20) | ((unsigned long long)(blockIdx.x & 1) << 19) | (ev << 12) | (step << 6) | (slot << 4) | j;
}
#define TRACE(role, ev, step, slot, j) trace2_ev(trace_on, role, trace_cnt, ev, step, slot, j)
#else
#define TRACE(role, ev, step, slot, j)
#endif

namespace {

constexpr long long kWaitLimitCycles = 1ll << 28;  // ~0.14 s: protocol bug => error flag, not a hang
constexpr int kThreads = 640;
constexpr int kStages = 4;
constexpr int kStageBytes = kRingStageBytes / 2;    // each CTA holds N/2 rows of a slab
constexpr int kEpiThreads = 256;                    // per slot

struct Shared {
  uint64_t w_full[kStages];    // this CTA's half of the slab has landed (TMA tx)
  uint64_t w_empty[kStages];   // both slots' MMAs on the slab have retired (2 multicast commits)
  uint64_t w_peer[kStages];    // leader only: the peer CTA's half has landed (remote arrive by the peer's relay)
  uint64_t a_ready[2];         // leader only: both CTAs' operand images of the slot are complete (2 arrives)
  uint64_t d_full[2];          // accumulator of the slot complete (multicast commit)
  uint32_t turn;               // leader only: number of issue steps slot 0's issuer has issued (slot 1 follows)
  uint32_t tmem_base;
  int abort_flag;
};

struct StepShape {
  uint32_t N, nslabs, slab_bytes, k16;
  bool first, last;   // first / last issue step of a layer: wait for the A operand / hand the accumulator over
};

// The skip layer L5 = [embedding | h] is issued as two steps accumulating into the same tile so that no step needs
// more than the four ring stages.
constexpr int kNumIssueSteps = 15;
__device__ __forceinline__ StepShape step_shape(int step) {
  switch (step) {
    case 0: return {96u, 1u, (uint32_t)kBendB0Bytes, 3u, true, true};
    case 1: return {96u, 1u, (uint32_t)kBendB1Bytes, 6u, true, true};
    case 2: return {80u, 1u, (uint32_t)kBendB2Bytes, 6u, true, true};
    case 3: return {64u, 1u, (uint32_t)kBendB3Bytes, 4u, true, true};
    case 4: return {16u, 1u, (uint32_t)kBendB4Bytes, 4u, true, true};
    case 5: return {256u, 1u, 32768u, 4u, true, true};
    case 10: return {256u, 1u, 32768u, 4u, true, false};
    case 11: return {256u, 4u, 32768u, 4u, false, true};
    case 14: return {16u, 1u, (uint32_t)kNerfHeadBytes, 16u, true, true};
    default: return {256u, 4u, 32768u, 4u, true, true};
  }
}
// byte offset (inside a slot's activation region: H at 0, E at kHBytes) of the A operand of slab j
__device__ __forceinline__ uint32_t a_operand_offset(int step, uint32_t j) {
  if (step == 0 || step == 5 || step == 10) return kHBytes;   // bender input / embedding live in E
  if (step < 5 || step == 14) return 0;
  return j * 8 * kChunkBytes;
}

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

// Drain 16 * NB accumulator columns starting at taddr, add bias, ReLU, convert to fp16 and store them as chunks
// [0, 2 * NB) relative to dst_row (this thread's row of a chunk-major activation image).  The TMEM load of block
// c + 1 and the bias loads of block c are in flight while block c is converted.
template <int NB>
__device__ __forceinline__ void epi_bias_relu_store(uint32_t taddr, const float* __restrict__ bias, uint8_t* dst_row) {
  uint32_t v[2][16];
  tmem_ld16(taddr, v[0]);
#pragma unroll
  for (int c = 0; c < NB; ++c) {
    float4 b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = __ldg(reinterpret_cast<const float4*>(bias + c * 16 + i * 4));
    tmem_ld_wait();
    if (c + 1 < NB) tmem_ld16(taddr + (c + 1) * 16, v[(c + 1) & 1]);
    const uint32_t(&w)[16] = v[c & 1];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float4 b0 = b[2 * q], b1 = b[2 * q + 1];
      uint4 pk;
      // one cvt.rn.relu.satfinite.f16x2 per two outputs: ReLU, clamp to fp16 range and pack in a single instruction
      pk.x = pack_h2_relu_sat(__uint_as_float(w[q * 8 + 0]) + b0.x, __uint_as_float(w[q * 8 + 1]) + b0.y);
      pk.y = pack_h2_relu_sat(__uint_as_float(w[q * 8 + 2]) + b0.z, __uint_as_float(w[q * 8 + 3]) + b0.w);
      pk.z = pack_h2_relu_sat(__uint_as_float(w[q * 8 + 4]) + b1.x, __uint_as_float(w[q * 8 + 5]) + b1.y);
      pk.w = pack_h2_relu_sat(__uint_as_float(w[q * 8 + 6]) + b1.z, __uint_as_float(w[q * 8 + 7]) + b1.w);
      *reinterpret_cast<uint4*>(dst_row + (c * 2 + q) * kChunkBytes) = pk;
    }
  }
}

// One feature of the positional encoding (Embedder.embed, run_nerf_helpers.py:149-150 with the settings of
// get_embedder :157-164): raw xyz first, then per octave sin(2^k xyz), cos(2^k xyz); k = 0..9; column 63 is the
// constant 1 the weight-gradient kernel reads as the bias input (its forward weight column is zero).
// sin/cos: the argument 2^k * x is reduced EXACTLY to [-0.5, 0.5) turns (x / 2pi carried as a two-float value),
// then evaluated with MUFU (abs err ~4e-7), well below fp16 resolution.
template <int F>
__device__ __forceinline__ float pe_feature(const float (&x)[3], const float (&thi)[3], const float (&tlo)[3]) {
  if (F < 3) return x[F];
  if (F == 63) return 1.f;
  constexpr int k = (F - 3) / 6, r = (F - 3) % 6, d = r % 3;
  const float sc = static_cast<float>(1 << k);
  const float a = thi[d] * sc;
  const float ang = ((a - rintf(a)) + tlo[d] * sc) * 6.2831853071795865f;
  return r < 3 ? __sinf(ang) : __cosf(ang);
}
template <int F0>
__device__ __forceinline__ uint4 pe_chunk(const float (&x)[3], const float (&thi)[3], const float (&tlo)[3]) {
  uint4 pk;
  pk.x = pack_h2(pe_feature<F0 + 0>(x, thi, tlo), pe_feature<F0 + 1>(x, thi, tlo));
  pk.y = pack_h2(pe_feature<F0 + 2>(x, thi, tlo), pe_feature<F0 + 3>(x, thi, tlo));
  pk.z = pack_h2(pe_feature<F0 + 4>(x, thi, tlo), pe_feature<F0 + 5>(x, thi, tlo));
  pk.w = pack_h2(pe_feature<F0 + 6>(x, thi, tlo), pe_feature<F0 + 7>(x, thi, tlo));
  return pk;
}
// chunks 4*HALF .. 4*HALF+3 of this thread's row of the embedding image
template <int HALF>
__device__ __forceinline__ void write_pe_half(const float (&x)[3], uint8_t* e_row) {
  const float kInv2PiHi = 0.15915494f;      // fl(1/2pi)
  const float kInv2PiLo = 6.4206199e-09f;   // 1/2pi - fl(1/2pi)
  float thi[3], tlo[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    thi[d] = x[d] * kInv2PiHi;
    tlo[d] = fmaf(x[d], kInv2PiLo, fmaf(x[d], kInv2PiHi, -thi[d]));
  }
  *reinterpret_cast<uint4*>(e_row + (4 * HALF + 0) * kChunkBytes) = pe_chunk<32 * HALF + 0>(x, thi, tlo);
  *reinterpret_cast<uint4*>(e_row + (4 * HALF + 1) * kChunkBytes) = pe_chunk<32 * HALF + 8>(x, thi, tlo);
  *reinterpret_cast<uint4*>(e_row + (4 * HALF + 2) * kChunkBytes) = pe_chunk<32 * HALF + 16>(x, thi, tlo);
  *reinterpret_cast<uint4*>(e_row + (4 * HALF + 3) * kChunkBytes) = pe_chunk<32 * HALF + 24>(x, thi, tlo);
}

// bender input row: [xyz_hi(3) xyz_lo(3) latent(32) 0(10)] fp16 = 6 chunks; this half writes chunks 3*HALF..3*HALF+2
template <int HALF>
__device__ __forceinline__ void write_bender_input_half(const float (&x)[3], const float* __restrict__ lat, bool valid,
                                                        uint8_t* e_row) {
  float in[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    const int e = 24 * HALF + i;
    if (e < 3) {
      in[i] = __half2float(__float2half_rn(x[e]));
    } else if (e < 6) {
      in[i] = x[e - 3] - __half2float(__float2half_rn(x[e - 3]));
    } else if (e < 6 + kLatent) {
      in[i] = valid ? __ldg(lat + (e - 6)) : 0.f;
    } else {
      in[i] = 0.f;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    uint4 pk;
    pk.x = pack_h2(in[c * 8 + 0], in[c * 8 + 1]);
    pk.y = pack_h2(in[c * 8 + 2], in[c * 8 + 3]);
    pk.z = pack_h2(in[c * 8 + 4], in[c * 8 + 5]);
    pk.w = pack_h2(in[c * 8 + 6], in[c * 8 + 7]);
    *reinterpret_cast<uint4*>(e_row + (3 * HALF + c) * kChunkBytes) = pk;
  }
}

}  // namespace

template <bool HAS_BENDER>
__global__ void __launch_bounds__(kThreads, 1) field_fwd2_kernel(const FieldFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* act = smem;                                  // 2 slots x (H | E)
  uint8_t* ring = smem + 2 * kSlotBytes;                // kStages x kStageBytes
  Shared* sh = reinterpret_cast<Shared*>(ring + kStages * kStageBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();              // 0 = leader (issues the MMAs)
  const int n_groups = (p.n_tiles + 3) / 4;
  const int group0 = blockIdx.x >> 1, group_stride = gridDim.x >> 1;
  constexpr int kFirstStep = HAS_BENDER ? 0 : 5;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&sh->w_full[i], 1);
      mbar_init(&sh->w_empty[i], 2);    // one commit from each slot's issuer
      mbar_init(&sh->w_peer[i], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sh->a_ready[s], 2);    // one elected arrive per CTA
      mbar_init(&sh->d_full[s], 1);
    }
    sh->turn = 0;
    sh->abort_flag = 0;
    fence_mbar_init();
  }
  cluster_sync_all();                   // the peer's barriers exist before anything arrives on them
  if (warp == 2) {
    tmem_alloc2(&sh->tmem_base, 512);
    tmem_relinquish2();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = sh->tmem_base;
  const Waiter W{&sh->abort_flag, p.err};

  if (warp == 0) {
    // ===================== weight producer: this CTA's half of every slab, once per tile group =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int group = group0; group < n_groups; group += group_stride) {
        uint32_t gb = 0, gn = 0;
#pragma unroll 1
        for (int step = kFirstStep; step < kNumIssueSteps; ++step) {
          const StepShape s = step_shape(step);
          const uint8_t* src = step < 5 ? p.bend_w + gb : p.nerf_w + gn;
          // rows [rank * N/2, (rank+1) * N/2) of every 8-column chunk: [chunk][N rows][16 B] -> [chunk][N/2 rows][16 B]
          const uint32_t hb = s.N * 8u, nch = s.slab_bytes / (s.N * 16u);
          for (uint32_t j = 0; j < s.nslabs; ++j) {
            W.wait(&sh->w_empty[stage], phase ^ 1u, 101);
            uint8_t* dst = ring + stage * kStageBytes;
            mbar_arrive_expect_tx(&sh->w_full[stage], s.slab_bytes / 2);
            const uint8_t* g = src + j * s.slab_bytes + rank * hb;
            for (uint32_t c = 0; c < nch; ++c) tma_bulk_g2s(dst + c * hb, g + c * 2u * hb, hb, &sh->w_full[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          if (step < 5) gb += s.nslabs * s.slab_bytes; else gn += s.nslabs * s.slab_bytes;
        }
      }
    }
  } else if (rank == 1 && warp == 1) {
    // ===================== peer relay: "my half of the slab has landed" -> leader =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int group = group0; group < n_groups; group += group_stride) {
#pragma unroll 1
        for (int step = kFirstStep; step < kNumIssueSteps; ++step) {
          const uint32_t n = step_shape(step).nslabs;
          for (uint32_t i = 0; i < n; ++i) {
            W.wait(&sh->w_full[stage], phase, 401);
            mbar_arrive_cluster(&sh->w_peer[stage], 0);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (rank == 0 && (warp == 1 || warp == 3)) {
    // ===================== MMA issuers: warp 1 for slot 0, warp 3 for slot 1 =====================
    // Both walk the same slab sequence; a stage is recycled once both have committed it.  For every issue step
    // slot 1's block follows slot 0's: the slots then alternate on the tensor pipe and -- the point -- their
    // epilogues alternate on the TMEM read path (64 B/clk per SM: draining a 128 x 256 fp32 accumulator takes as
    // long as computing it).  Left alone the two slots fall into lock step, where both the MMA blocks and the
    // drains share their unit and neither overlaps the other.
    if (lane == 0) {
      const int slot = warp == 1 ? 0 : 1;
      uint32_t stage = 0, phase = 0, aph = 0, steps_issued = 0;
      volatile uint32_t* turn = &sh->turn;
      const uint32_t d_tmem = tmem_base + slot * 256;
      const uint32_t a_base = smem_u32(act + slot * kSlotBytes);
      const uint32_t ring_base = smem_u32(ring);
      [[maybe_unused]] uint32_t trace_cnt = 0;
      for (int group = group0; group < n_groups; group += group_stride) {
        [[maybe_unused]] const bool trace_on = blockIdx.x < 2 && group == group0 + 2 * group_stride;
#pragma unroll 1
        for (int step = kFirstStep; step < kNumIssueSteps; ++step) {
          const StepShape s = step_shape(step);
          const uint32_t idesc = umma_instr_desc(2 * kTileM, s.N, UMMA_F16, UMMA_F16, UMMA_K_MAJOR, UMMA_K_MAJOR);
          const uint32_t b_lbo = s.N * 8;        // N/2 rows of B held by this CTA x 16 B
          if (s.first) {
            W.wait(&sh->a_ready[slot], aph, 201);
            aph ^= 1u;
            TRACE(slot, 1, step, slot, 0);
            tc_fence_after_sync();
          }
          if (slot == 1 && *turn <= steps_issued) {
            const long long t0 = clock64();
            while (*turn <= steps_issued) {
              if (*reinterpret_cast<volatile int*>(&sh->abort_flag)) break;
              if (clock64() - t0 > kWaitLimitCycles) { atomicExch(&sh->abort_flag, 204); atomicCAS(p.err, 0, 204); break; }
            }
          }
          TRACE(slot, 6, step, slot, 0);
          for (uint32_t j = 0; j < s.nslabs; ++j) {
            W.wait(&sh->w_full[stage], phase, 202);
            W.wait(&sh->w_peer[stage], phase, 203);
            TRACE(slot, 2, step, slot, j);
            tc_fence_after_sync();
            uint64_t ad = umma_smem_desc(a_base + a_operand_offset(step, j), kChunkBytes, 128);
            uint64_t bd = umma_smem_desc(ring_base + stage * kStageBytes, b_lbo, 128);
            const uint64_t a_step = (2 * kChunkBytes) >> 4, b_step = (2 * b_lbo) >> 4;
            if (p.debug_mode != 2) {
              for (uint32_t k = 0; k < s.k16; ++k) {
                umma_f16_ss2(d_tmem, ad, bd, idesc, (!s.first || (j | k)) ? 1u : 0u);
                ad += a_step;
                bd += b_step;
              }
            }
            umma_commit2(&sh->w_empty[stage]);   // free (in both CTAs) once both slots' MMAs on it have retired
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          if (s.last) { umma_commit2(&sh->d_full[slot]); TRACE(slot, 3, step, slot, 0); }   // accumulator complete
          ++steps_issued;
          if (slot == 0) *turn = steps_issued;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: 2 warpgroups per slot, each draining half of the columns =====================
    const int ew = warp - 4;
    const int slot = ew >> 3;
    const int half = (ew >> 2) & 1;
    const int row = ((warp & 3) << 5) | lane;
    uint8_t* Hs = act + slot * kSlotBytes;
    uint8_t* Es = Hs + kHBytes;
    uint8_t* h_row = Hs + row * 16;
    uint8_t* e_row = Es + row * 16;
    const uint32_t taddr = tmem_base + ((static_cast<uint32_t>(warp & 3) * 32u) << 16) + slot * 256;
    const bool slot_leader = (ew & 7) == 0 && lane == 0;
    const bool writer = half == 0;          // per-point outputs are written once
    uint32_t dph = 0;
    [[maybe_unused]] uint32_t trace_cnt = 0;
    auto wait_acc = [&](int code) {
      W.wait(&sh->d_full[slot], dph, code);
      dph ^= 1u;
      tc_fence_after_sync();
    };
    auto slot_barrier = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "n"(kEpiThreads) : "memory"); };

    for (int group = group0; group < n_groups; group += group_stride) {
      [[maybe_unused]] const bool trace_on = blockIdx.x < 2 && group == group0 + 2 * group_stride && slot_leader;
      const long long tile = (static_cast<long long>(group) * 2 + slot) * 2 + rank;
      const long long pt = tile * kTileM + row;
      const bool valid = pt < p.P;
      // Training stash: every finished activation image (a contiguous chunk-major block of shared memory) is
      // written to this tile's stash block with bulk TMA stores issued by one thread of the slot.
      // stash_begin(): the previous store must have finished READING shared memory before an image is overwritten.
      uint8_t* st = p.stash ? p.stash + tile * kStashTileBytes : nullptr;
      auto stash_begin = [&]() {
        if (st) {
          if (slot_leader) tma_bulk_wait_read<0>();
          slot_barrier();
        }
      };
      // publish(): the slot's operand image is complete.  All 256 threads make their shared-memory writes visible
      // to the async proxy and meet; one thread stores the image to the stash (bytes > 0) and arrives on the
      // leader CTA's barrier.
      auto publish = [&](uint32_t stash_off, const uint8_t* img, uint32_t bytes) {
        fence_proxy_async_smem();
        tc_fence_before_sync();
        slot_barrier();
        if (slot_leader) {
          if (st && bytes) {
            for (uint32_t o = 0; o < bytes; o += 16384u) tma_bulk_s2g(st + stash_off + o, img + o, bytes - o < 16384u ? bytes - o : 16384u);
            tma_bulk_commit();
          }
          mbar_arrive_cluster(&sh->a_ready[slot], 0);
          TRACE(2 + slot, 5, 0, slot, 0);
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
        if (p.d_init && writer) {
          p.d_init[pt * 3 + 0] = x[0]; p.d_init[pt * 3 + 1] = x[1]; p.d_init[pt * 3 + 2] = x[2];
        }
      }
      float rigidity = 0.f;
      if (HAS_BENDER) {
        // ---- bender input row -> chunks 0..5 of E ----
        stash_begin();
        const float* lat = p.latents + ray * p.latent_stride;
        if (half) write_bender_input_half<1>(x, lat, valid, e_row); else write_bender_input_half<0>(x, lat, valid, e_row);
        publish(kStBin, Es, 6 * kChunkBytes);
        // ---- B0, B1: 96 hidden units (64 offset | 32 rigidity); each half drains 48 columns ----
        wait_acc(301);
        stash_begin();
        epi_bias_relu_store<3>(taddr + 48 * half, p.bend_bias + 48 * half, h_row + 6 * half * kChunkBytes);
        publish(kStHb1, Hs, 12 * kChunkBytes);
        wait_acc(302);
        stash_begin();
        epi_bias_relu_store<3>(taddr + 48 * half, p.bend_bias + 96 + 48 * half, h_row + 6 * half * kChunkBytes);
        publish(kStHb2, Hs, 12 * kChunkBytes);
        // ---- B2: 64 offset hidden (32 per half) + rigidity output (column 64, read by both halves) ----
        wait_acc(303);
        stash_begin();
        epi_bias_relu_store<2>(taddr + 32 * half, p.bend_bias + 192 + 32 * half, h_row + 4 * half * kChunkBytes);
        {
          uint32_t v[8];
          tmem_ld8(taddr + 64, v);
          tmem_ld_wait();
          const float rr = __uint_as_float(v[0]) + __ldg(p.bend_bias + 192 + 64);
          rigidity = (tanhf(rr) + 1.0f) * 0.5f;   // run_nerf_helpers.py:559-561
          if (p.use_cutoff && rigidity <= p.cutoff) rigidity = 0.f;  // :563-564
        }
        publish(kStHb3, Hs, 8 * kChunkBytes);
        // ---- B3 ----
        wait_acc(304);
        stash_begin();
        epi_bias_relu_store<2>(taddr + 32 * half, p.bend_bias + 272 + 32 * half, h_row + 4 * half * kChunkBytes);
        publish(kStHb4, Hs, 8 * kChunkBytes);
        // ---- B4: offsets; bend (both halves compute the bent point, one writes the details) ----
        wait_acc(305);
        {
          uint32_t v[8];
          tmem_ld8(taddr, v);
          tmem_ld_wait();
          float un[3], ma[3];
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            un[d] = __uint_as_float(v[d]);
            ma[d] = __fmul_rn(rigidity, un[d]);              // :567
            if (p.use_scaling) ma[d] = __fmul_rn(ma[d], p.scaling);  // :568-569
          }
          if (valid && writer) {
            if (p.d_unmasked) { p.d_unmasked[pt * 3 + 0] = un[0]; p.d_unmasked[pt * 3 + 1] = un[1]; p.d_unmasked[pt * 3 + 2] = un[2]; }
            if (p.d_masked) { p.d_masked[pt * 3 + 0] = ma[0]; p.d_masked[pt * 3 + 1] = ma[1]; p.d_masked[pt * 3 + 2] = ma[2]; }
            if (p.d_rigid) p.d_rigid[pt] = rigidity;
          }
#pragma unroll
          for (int d = 0; d < 3; ++d) x[d] = __fadd_rn(x[d], ma[d]);  // :570
        }
      }
      if (valid && writer && p.d_bent) {
        p.d_bent[pt * 3 + 0] = x[0]; p.d_bent[pt * 3 + 1] = x[1]; p.d_bent[pt * 3 + 2] = x[2];
      }
      // ---- positional encoding of the (bent) point -> E; each half writes four of the eight chunks ----
      stash_begin();
      if (half) write_pe_half<1>(x, e_row); else write_pe_half<0>(x, e_row);
      publish(kStE, Es, kEBytes);
      // ---- L0 .. L7: each half drains 128 of the 256 columns ----
#pragma unroll 1
      for (int L = 0; L < 8; ++L) {
        wait_acc(310 + L);
        TRACE(2 + slot, 4, 5 + L, slot, 0);
        stash_begin();
        if (p.debug_mode != 1) epi_bias_relu_store<8>(taddr + 128 * half, p.nerf_bias + L * 256 + 128 * half, h_row + 16 * half * kChunkBytes);
        publish(kStH + L * kHBytes, Hs, kHBytes);
      }
      // ---- head: raw = output_linear(h) (run_nerf_helpers.py:306) ----
      wait_acc(320);
      if (writer) {
        uint32_t v[8];
        tmem_ld8(taddr, v);
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
      // the slot's next a_ready arrival is the next group's first image: it is published after a slot_barrier(),
      // i.e. after every thread of the slot has drained this group's head accumulator.
    }
    if (p.stash && slot_leader) tma_bulk_wait<0>();   // all stash stores complete before the CTA exits
  }

  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();   // no CTA exits (or frees TMEM) while its peer can still reach it
  if (warp == 2) tmem_dealloc2(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
template <bool HAS_BENDER>
static cudaError_t launch_variant(const FieldFwdParams& p, int num_sms, cudaStream_t stream) {
  const size_t smem = 2 * kSlotBytes + kStages * kStageBytes + sizeof(Shared) + 64;
  const int n_groups = (p.n_tiles + 3) / 4;
  const int max_groups = num_sms / 2;
  const int grid = 2 * (n_groups < max_groups ? n_groups : max_groups);
  auto kern = field_fwd2_kernel<HAS_BENDER>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

cudaError_t launch_field_fwd2(const FieldFwdParams& p, bool has_bender, int num_sms, cudaStream_t stream) {
  if (p.n_tiles <= 0) return cudaSuccess;
  return has_bender ? launch_variant<true>(p, num_sms, stream) : launch_variant<false>(p, num_sms, stream);
}

#ifdef NRN_TRACE
extern "C" int dbg_trace2_read(unsigned long long* out, int max_n) {
  if (max_n < 2 * 4 * 256) return -1;
  cudaMemcpyFromSymbol(out, g_trace2_buf, sizeof(unsigned long long) * 2 * 4 * 256);
  return 2 * 4 * 256;
}
extern "C" void dbg_trace2_reset() {
  void* ptr = nullptr;
  cudaGetSymbolAddress(&ptr, g_trace2_buf);
  cudaMemset(ptr, 0, sizeof(unsigned long long) * 2 * 4 * 256);
}
#endif

}  // namespace nrn
