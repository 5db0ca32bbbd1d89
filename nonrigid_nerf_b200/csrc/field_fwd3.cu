// Fused point-wise field evaluation, forward -- "shared-slab" schedule.
//
// Same computation and bit-identical results as field_fwd.cu (see its header for the reference call chain it replaces:
// run_network / batchify train.py:27-105, NeRF.forward run_nerf_helpers.py:240-314, ray_bending.forward :507-584,
// Embedder.embed :149-150), different machine.  What the measurements of round 1 said about field_fwd.cu:
//   * each of the two tile slots of a CTA streams its own copy of the layer weights from L2: 256 KB per layer and CTA,
//     ~30 B/clk per SM, 4,600 of the chip's ~6,300 B/clk of L2 bandwidth -- the weight stream, not the tensor pipe, set
//     the 2.2 us "MMA block" (the MMAs alone need 1.04 us);
//   * a layer's epilogue (TMEM drain, bias, ReLU, fp16 pack, 1.6 us) cannot start before the layer's last MMA and the next
//     layer cannot start before the epilogue's last store: the two slots hid only part of that chain.
// Here:
//   * ONE weight stream per CTA: every 16 KB piece of a layer ([K = 64] x [128 output rows]) is consumed by BOTH slots
//     before the ring stage is released -- half the L2 traffic per tile;
//   * a layer's 256 outputs are computed as two halves of 128 (two accumulators of 128 TMEM columns per slot): while the
//     tensor pipe works on half 1, the epilogue of half 0 drains / packs, and the next layer starts on the K range that
//     half 0 produced while half 1 is still being packed.  Four epilogue warpgroups: (slot, half).
//
// Work decomposition
//   tile   = 128 consecutive sample points; a CTA (1 per SM, persistent) processes tile PAIRS (slot 0 / slot 1) in lock step
//   warps  : 0 weight producer (bulk TMA ring, 4 x 16 KB), 1 MMA issuer (one lane), 2 TMEM allocator, 3 idle,
//            4-7 / 8-11 epilogue of output half 0 for slot 0 / 1 ("primary": also ray set-up, bender steps, positional
//            encoding, head), 12-15 / 16-19 epilogue of output half 1 for slot 0 / 1
//   piece  = one ring stage: [<= 8 K-chunks][N rows][8] fp16 of one layer, with the MMAs both slots issue on it
//   order  : bender B0 B1(2 pieces) B2 B3 B4, then per NeRF layer  h0: k-pieces 0..3 | h1: k-pieces 0..3,  head
//   barriers: w_full/w_empty per stage; a_ready[kh] "A columns of K half kh written (and accumulator half kh drained)" with
//            256 arrivals (both slots' warpgroups of that half); d_full[nh] "accumulator half nh complete" (commit);
//            a_free "the MMAs of half 1 that read A columns 0..127 are done" (commit): the half-0 epilogue overwrites
//            those columns in place.
//
// Shared memory: 2 x (H 64 KB + E 16 KB) activations + 4 x 16 KB ring.  Tensor memory: 512 columns = 2 slots x 2 halves x 128.
#include "nrn_common.cuh"
#include "sm100_ptx.cuh"

namespace nrn {

namespace {

constexpr long long kWaitLimitCycles3 = 1ll << 28;
constexpr int kStages3 = 4;
constexpr int kStageBytes3 = 16384;
constexpr int kFwd3Threads = 640;     // 20 warps

enum : int { BAR_NONE = -1, BAR_READY0 = 0, BAR_READY1 = 1 };
enum : int { COMMIT_NONE = 0, COMMIT_D0 = 1, COMMIT_D1 = 2, COMMIT_AFREE = 4 };

struct Piece {
  uint32_t src_off;     // byte offset into the bender image (bender == 1) or the split-order NeRF image
  uint16_t bytes_div16; // piece size / 16
  uint16_t n;           // UMMA N (rows of the piece)
  uint32_t a_off;       // byte offset of the A operand inside a slot's activation region (H at 0, E at kHBytes)
  uint8_t k16;          // K / 16 of the piece
  uint8_t bender;       // source image
  uint8_t acc_col;      // accumulator column inside the slot's 256 (0 or 128)
  uint8_t first;        // 1: the piece's first MMA overwrites the accumulator
  int8_t wait0;         // wait a_ready[0] before the piece
  int8_t wait1;         // wait a_ready[1] before the piece
  uint8_t commit;       // COMMIT_* bits after the piece
  uint8_t pad;
};
constexpr int kMaxPieces = 96;
struct Schedule {
  int n;
  Piece p[kMaxPieces];
};
__constant__ Schedule c_sched[2];   // [0] without bender, [1] with bender

struct Shared3 {
  uint64_t w_full[kStages3];
  uint64_t w_empty[kStages3];
  uint64_t a_ready[2];
  uint64_t d_full[2];
  uint64_t a_free;
  uint32_t tmem_base;
  int abort_flag;
};

// Developer profile (NRN_DEBUG_MODE=9): cycles CTA 0's roles spend in each kind of wait, read with nrn_debug_profile().
//   [0] issuer total  [1] issuer: a_ready[0]  [2] issuer: a_ready[1]  [3] issuer: w_full  [4] producer: w_empty
//   [5] primary WG (slot 0) total  [6] primary: d_full[0]  [7] primary: a_free  [8] half-1 WG (slot 0) total  [9] half-1: d_full[1]
//   [10] pieces issued  [11] tile pairs
__device__ unsigned long long g_fwd3_prof[16];

struct Waiter3 {
  int* s_abort;
  int* g_err;
  __device__ __forceinline__ bool wait(uint64_t* bar, uint32_t parity, int code, unsigned long long* acc = nullptr) const {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
      if (*reinterpret_cast<volatile int*>(s_abort)) return false;
      if (clock64() - t0 > kWaitLimitCycles3) {
        atomicExch(s_abort, code);
        atomicCAS(g_err, 0, code);
        return false;
      }
    }
    if (acc) *acc += static_cast<unsigned long long>(clock64() - t0);
    return true;
  }
};

// Drain NCOLS accumulator columns (multiple of 32), add bias, ReLU, convert to fp16 and store them as chunks
// [0, NCOLS/8) relative to dst_row (the caller offsets dst_row / bias / taddr to the half it owns).
template <int NCOLS>
__device__ __forceinline__ void epi3_bias_relu_store(uint32_t taddr, const float* __restrict__ bias, uint8_t* dst_row) {
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
      pk.x = pack_h2_relu_sat(__uint_as_float(w[q * 8 + 0]) + b0.x, __uint_as_float(w[q * 8 + 1]) + b0.y);
      pk.y = pack_h2_relu_sat(__uint_as_float(w[q * 8 + 2]) + b0.z, __uint_as_float(w[q * 8 + 3]) + b0.w);
      pk.z = pack_h2_relu_sat(__uint_as_float(w[q * 8 + 4]) + b1.x, __uint_as_float(w[q * 8 + 5]) + b1.y);
      pk.w = pack_h2_relu_sat(__uint_as_float(w[q * 8 + 6]) + b1.z, __uint_as_float(w[q * 8 + 7]) + b1.w);
      *reinterpret_cast<uint4*>(dst_row + (c * 4 + q) * kChunkBytes) = pk;
    }
  }
}

// Positional encoding of one point: identical to field_fwd.cu's write_pe (Embedder.embed, run_nerf_helpers.py:149-150)
__device__ __forceinline__ void write_pe3(const float (&x)[3], uint8_t* dst_row) {
  float f[64];
  f[0] = x[0]; f[1] = x[1]; f[2] = x[2];
  const float kInv2PiHi = 0.15915494f;
  const float kInv2PiLo = 6.4206199e-09f;
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
  f[63] = 1.f;
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
__global__ void __launch_bounds__(kFwd3Threads, 1) field_fwd3_kernel(const FieldFwdParams p, const uint8_t* __restrict__ nerf_split) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* act = smem;                                  // 2 slots x (H | E)
  uint8_t* ring = smem + 2 * kSlotBytes;                // kStages3 x 16 KB
  Shared3* sh = reinterpret_cast<Shared3*>(ring + kStages3 * kStageBytes3);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_pairs = (p.n_tiles + 1) >> 1;
  const Schedule& sched = c_sched[HAS_BENDER ? 1 : 0];

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages3; ++i) {
      mbar_init(&sh->w_full[i], 1);
      mbar_init(&sh->w_empty[i], 1);
    }
    mbar_init(&sh->a_ready[0], 256);
    mbar_init(&sh->a_ready[1], 256);
    mbar_init(&sh->d_full[0], 1);
    mbar_init(&sh->d_full[1], 1);
    mbar_init(&sh->a_free, 1);
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
  const Waiter3 W{&sh->abort_flag, p.err};
  const bool prof = p.debug_mode == 9 && blockIdx.x == 0;

  if (warp == 0) {
    // ===================== weight producer: global -> smem ring (bulk TMA) =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      unsigned long long pw = 0;
      unsigned long long* const ppw = prof ? &pw : nullptr;
      for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
#pragma unroll 1
        for (int i = 0; i < sched.n; ++i) {
          const Piece pc = sched.p[i];
          const uint32_t bytes = static_cast<uint32_t>(pc.bytes_div16) * 16u;
          const uint8_t* src = (pc.bender ? p.bend_w : nerf_split) + pc.src_off;
          W.wait(&sh->w_empty[stage], phase ^ 1u, 101, ppw);
          mbar_arrive_expect_tx(&sh->w_full[stage], bytes);
          tma_bulk_g2s(ring + stage * kStageBytes3, src, bytes, &sh->w_full[stage]);
          if (++stage == kStages3) { stage = 0; phase ^= 1u; }
        }
      }
      if (prof) g_fwd3_prof[4] = pw;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      uint32_t aph[2] = {0u, 0u};
      const uint32_t a_base0 = smem_u32(act), a_base1 = smem_u32(act + kSlotBytes);
      unsigned long long w_a0 = 0, w_a1 = 0, w_w = 0, n_pieces = 0, n_pairs_done = 0;
      const long long t_start = clock64();
      for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        ++n_pairs_done;
#pragma unroll 1
        for (int i = 0; i < sched.n; ++i) {
          const Piece pc = sched.p[i];
          ++n_pieces;
          if (pc.wait0) { W.wait(&sh->a_ready[0], aph[0], 201, prof ? &w_a0 : nullptr); aph[0] ^= 1u; }
          if (pc.wait1) { W.wait(&sh->a_ready[1], aph[1], 203, prof ? &w_a1 : nullptr); aph[1] ^= 1u; }
          W.wait(&sh->w_full[stage], phase, 202, prof ? &w_w : nullptr);
          tc_fence_after_sync();
          const uint32_t n = pc.n;
          const uint32_t idesc = umma_instr_desc(kTileM, n, UMMA_F16, UMMA_F16, UMMA_K_MAJOR, UMMA_K_MAJOR);
          const uint64_t bdesc = umma_smem_desc(smem_u32(ring + stage * kStageBytes3), n * 16, 128);
          const uint64_t adesc0 = umma_smem_desc(a_base0 + pc.a_off, kChunkBytes, 128);
          const uint64_t adesc1 = umma_smem_desc(a_base1 + pc.a_off, kChunkBytes, 128);
          const uint32_t d0 = tmem_base + pc.acc_col, d1 = tmem_base + 256 + pc.acc_col;
          if (p.debug_mode != 2) {
            for (uint32_t k = 0; k < pc.k16; ++k)
              umma_f16_ss(d0, umma_desc_advance(adesc0, k * 2 * kChunkBytes), umma_desc_advance(bdesc, k * 2 * n * 16), idesc,
                          (k | (pc.first ^ 1u)) ? 1u : 0u);
            for (uint32_t k = 0; k < pc.k16; ++k)
              umma_f16_ss(d1, umma_desc_advance(adesc1, k * 2 * kChunkBytes), umma_desc_advance(bdesc, k * 2 * n * 16), idesc,
                          (k | (pc.first ^ 1u)) ? 1u : 0u);
          }
          umma_commit(&sh->w_empty[stage]);
          if (pc.commit & COMMIT_AFREE) umma_commit(&sh->a_free);
          if (pc.commit & COMMIT_D0) umma_commit(&sh->d_full[0]);
          if (pc.commit & COMMIT_D1) umma_commit(&sh->d_full[1]);
          if (++stage == kStages3) { stage = 0; phase ^= 1u; }
        }
      }
      if (prof) {
        g_fwd3_prof[0] = static_cast<unsigned long long>(clock64() - t_start);
        g_fwd3_prof[1] = w_a0; g_fwd3_prof[2] = w_a1; g_fwd3_prof[3] = w_w; g_fwd3_prof[10] = n_pieces; g_fwd3_prof[11] = n_pairs_done;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warpgroups: (slot, output half) =====================
    const int wg = (warp - 4) >> 2;
    const int slot = wg & 1;
    const int half = wg >> 1;
    const int row = ((warp & 3) << 5) | lane;
    uint8_t* Hs = act + slot * kSlotBytes;
    uint8_t* Es = Hs + kHBytes;
    uint8_t* h_row = Hs + row * 16;
    uint8_t* e_row = Es + row * 16;
    const uint32_t taddr = tmem_base + ((static_cast<uint32_t>(warp & 3) * 32u) << 16) + slot * 256 + half * 128;
    const bool wg_leader = (threadIdx.x & 127) == 0;
    const int bar_id = 1 + wg;               // named barrier of this warpgroup
    uint32_t dph = 0, fph = 0;
    unsigned long long w_d = 0, w_f = 0;
    const bool tprof = prof && slot == 0 && (threadIdx.x & 127) == 0;
    const long long t_start = clock64();
    auto signal_ready = [&]() {
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(&sh->a_ready[half]);
    };
    auto wait_acc = [&](int code) {
      W.wait(&sh->d_full[half], dph, code, tprof ? &w_d : nullptr);
      dph ^= 1u;
      tc_fence_after_sync();
    };

    for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      const long long tile = static_cast<long long>(pair) * 2 + slot;
      const long long pt = tile * kTileM + row;
      const bool valid = pt < p.P;
      uint8_t* st = p.stash ? p.stash + tile * kStashTileBytes : nullptr;
      auto stash_begin = [&]() {
        if (st) {
          if (wg_leader) tma_bulk_wait_read<0>();
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        }
      };
      auto stash_store = [&](uint32_t stash_off, const uint8_t* img, uint32_t bytes) {
        if (st) {
          fence_proxy_async_smem();
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          if (wg_leader) {
            for (uint32_t o = 0; o < bytes; o += 16384u) tma_bulk_s2g(st + stash_off + o, img + o, bytes - o < 16384u ? bytes - o : 16384u);
            tma_bulk_commit();
          }
        }
      };

      if (half == 1) {
        // ---- output half 1 of L0 .. L7: columns 128..255 ----
#pragma unroll 1
        for (int L = 0; L < 8; ++L) {
          wait_acc(330 + L);
          stash_begin();
          if (p.debug_mode != 1) epi3_bias_relu_store<128>(taddr, p.nerf_bias + L * 256 + 128, h_row + 16 * kChunkBytes);
          stash_store(kStH + L * kHBytes + 16 * kChunkBytes, Hs + 16 * kChunkBytes, 16 * kChunkBytes);
          signal_ready();
        }
        continue;
      }

      // ---- primary warpgroup: ray set-up, bender, positional encoding, half 0 of every layer, head ----
      float x[3] = {0.f, 0.f, 0.f};
      long long ray = 0;
      if (valid) {
        ray = pt / p.S;
        if (p.pts) {
          const float* q = p.pts + pt * p.pts_stride;
          x[0] = __ldg(q + 0); x[1] = __ldg(q + 1); x[2] = __ldg(q + 2);
        } else {
          const float z = __ldg(p.z_vals + pt);
          const float* r = p.rays + ray * 8;
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
        wait_acc(301);
        stash_begin();
        epi3_bias_relu_store<96>(taddr, p.bend_bias, h_row);
        stash_store(kStHb1, Hs, 12 * kChunkBytes);
        signal_ready();
        wait_acc(302);
        stash_begin();
        epi3_bias_relu_store<96>(taddr, p.bend_bias + 96, h_row);
        stash_store(kStHb2, Hs, 12 * kChunkBytes);
        signal_ready();
        wait_acc(303);
        stash_begin();
        epi3_bias_relu_store<64>(taddr, p.bend_bias + 192, h_row);
        stash_store(kStHb3, Hs, 8 * kChunkBytes);
        {
          uint32_t v[16];
          tmem_ld16(taddr + 64, v);
          tmem_ld_wait();
          const float rr = __uint_as_float(v[0]) + __ldg(p.bend_bias + 192 + 64);
          rigidity = (tanhf(rr) + 1.0f) * 0.5f;
          if (p.use_cutoff && rigidity <= p.cutoff) rigidity = 0.f;
        }
        signal_ready();
        wait_acc(304);
        stash_begin();
        epi3_bias_relu_store<64>(taddr, p.bend_bias + 272, h_row);
        stash_store(kStHb4, Hs, 8 * kChunkBytes);
        signal_ready();
        wait_acc(305);
        {
          uint32_t v[16];
          tmem_ld16(taddr, v);
          tmem_ld_wait();
          float un[3], ma[3];
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            un[d] = __uint_as_float(v[d]);
            ma[d] = __fmul_rn(rigidity, un[d]);
            if (p.use_scaling) ma[d] = __fmul_rn(ma[d], p.scaling);
          }
          if (valid) {
            if (p.d_unmasked) { p.d_unmasked[pt * 3 + 0] = un[0]; p.d_unmasked[pt * 3 + 1] = un[1]; p.d_unmasked[pt * 3 + 2] = un[2]; }
            if (p.d_masked) { p.d_masked[pt * 3 + 0] = ma[0]; p.d_masked[pt * 3 + 1] = ma[1]; p.d_masked[pt * 3 + 2] = ma[2]; }
            if (p.d_rigid) p.d_rigid[pt] = rigidity;
          }
#pragma unroll
          for (int d = 0; d < 3; ++d) x[d] = __fadd_rn(x[d], ma[d]);
        }
      }
      if (valid && p.d_bent) {
        p.d_bent[pt * 3 + 0] = x[0]; p.d_bent[pt * 3 + 1] = x[1]; p.d_bent[pt * 3 + 2] = x[2];
      }
      stash_begin();
      write_pe3(x, e_row);
      stash_store(kStE, Es, kEBytes);
      signal_ready();
      // ---- output half 0 of L0 .. L7: columns 0..127, written in place over the A operand once half 1's MMAs on
      //      those columns are done (a_free) ----
#pragma unroll 1
      for (int L = 0; L < 8; ++L) {
        wait_acc(310 + L);
        W.wait(&sh->a_free, fph, 340 + L, tprof ? &w_f : nullptr);
        fph ^= 1u;
        stash_begin();
        if (p.debug_mode != 1) epi3_bias_relu_store<128>(taddr, p.nerf_bias + L * 256, h_row);
        stash_store(kStH + L * kHBytes, Hs, 16 * kChunkBytes);
        signal_ready();
      }
      // ---- head ----
      wait_acc(320);
      {
        uint32_t v[16];
        tmem_ld16(taddr, v);
        tmem_ld_wait();
        if (valid) {
          float o[5];
#pragma unroll
          for (int c = 0; c < 5; ++c) o[c] = __uint_as_float(v[c]) + __ldg(p.nerf_bias + 2048 + c);
          if (HAS_BENDER && p.use_removal && rigidity >= p.removal) o[3] *= 0.f;
          float* dst = p.raw + pt * p.out_ch;
          for (int c = 0; c < p.out_ch; ++c) dst[c] = o[c];
        }
      }
    }
    if (p.stash && (threadIdx.x & 127) == 0) tma_bulk_wait<0>();
    if (tprof) {
      const unsigned long long tot = static_cast<unsigned long long>(clock64() - t_start);
      if (half == 0) { g_fwd3_prof[5] = tot; g_fwd3_prof[6] = w_d; g_fwd3_prof[7] = w_f; }
      else { g_fwd3_prof[8] = tot; g_fwd3_prof[9] = w_d; }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// host: the piece schedule (built once per process, copied to constant memory per device)
// ------------------------------------------------------------------------------------------------
namespace {

void add_piece(Schedule& s, uint32_t src_off, uint32_t bytes, int n, int k16, uint32_t a_off, bool bender, int acc_col, bool first,
               bool wait0, bool wait1, int commit) {
  Piece& q = s.p[s.n++];
  q.src_off = src_off; q.bytes_div16 = static_cast<uint16_t>(bytes / 16); q.n = static_cast<uint16_t>(n); q.a_off = a_off;
  q.k16 = static_cast<uint8_t>(k16); q.bender = bender ? 1 : 0; q.acc_col = static_cast<uint8_t>(acc_col); q.first = first ? 1 : 0;
  q.wait0 = wait0 ? 1 : 0; q.wait1 = wait1 ? 1 : 0; q.commit = static_cast<uint8_t>(commit); q.pad = 0;
}

Schedule build_schedule(bool has_bender) {
  Schedule s{};
  s.n = 0;
  const uint32_t E = kHBytes;                 // A operand offsets: H at 0, E behind it
  if (has_bender) {
    uint32_t off = 0;
    add_piece(s, off, kBendB0Bytes, 96, 3, E, true, 0, true, true, false, COMMIT_D0); off += kBendB0Bytes;
    // B1: K = 96 in two pieces (8 + 4 chunks of 96 rows)
    add_piece(s, off, 8 * 96 * 16, 96, 4, 0, true, 0, true, true, false, COMMIT_NONE);
    add_piece(s, off + 8 * 96 * 16, 4 * 96 * 16, 96, 2, 8 * kChunkBytes, true, 0, false, false, false, COMMIT_D0); off += kBendB1Bytes;
    add_piece(s, off, kBendB2Bytes, 80, 6, 0, true, 0, true, true, false, COMMIT_D0); off += kBendB2Bytes;
    add_piece(s, off, kBendB3Bytes, 64, 4, 0, true, 0, true, true, false, COMMIT_D0); off += kBendB3Bytes;
    add_piece(s, off, kBendB4Bytes, 16, 4, 0, true, 0, true, true, false, COMMIT_D0);
  }
  // NeRF layers in split order: per layer, per output half, K pieces of 64
  uint32_t off = 0;
  for (int L = 0; L < 8; ++L) {
    const int nk = L == 0 ? 1 : (L == 5 ? 5 : 4);
    for (int nh = 0; nh < 2; ++nh) {
      for (int ks = 0; ks < nk; ++ks) {
        uint32_t a_off;
        bool w0 = false, w1 = false;
        int kh;    // which half of the previous layer's output this K piece reads (0, 1) or -1: the embedding
        if (L == 0) { a_off = E; kh = -1; }
        else if (L == 5) { a_off = ks == 0 ? E : (ks - 1) * 8 * kChunkBytes; kh = ks == 0 ? -1 : (ks - 1) / 2; }
        else { a_off = ks * 8 * kChunkBytes; kh = ks / 2; }
        if (nh == 0) {
          if (L == 0) w0 = ks == 0;                                   // positional encoding written
          else if (L == 5) { w0 = ks == 0; w1 = ks == 3; }           // the embedding piece reads nothing new, but it OVERWRITES
                                                                     // accumulator half 0: L4's half-0 epilogue must have drained it
          else { w0 = ks == 0; w1 = ks == 2; }
        }
        int commit = COMMIT_NONE;
        if (ks == nk - 1) commit |= nh == 0 ? COMMIT_D0 : COMMIT_D1;
        // a_free: the last piece of half 1 that reads A columns 0..127 (K pieces of kh <= 0)
        if (nh == 1) {
          bool last_low = false;
          if (L == 0) last_low = true;
          else if (L == 5) last_low = ks == 2;
          else last_low = ks == 1;
          if (last_low) commit |= COMMIT_AFREE;
        }
        (void)kh;
        add_piece(s, off, 8 * 128 * 16, 128, 4, a_off, false, nh * 128, ks == 0, w0, w1, commit);
        off += 8 * 128 * 16;
      }
    }
  }
  // head: N = 16, K = 256 (8 KB); needs both halves of L7
  add_piece(s, off, kNerfHeadBytes, 16, 16, 0, false, 0, true, true, true, COMMIT_D0);
  return s;
}

bool g_sched_uploaded[64] = {};

}  // namespace

extern "C" int nrn_debug_profile_fwd3(unsigned long long* out16) {
  return cudaMemcpyFromSymbol(out16, g_fwd3_prof, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : -2;
}

size_t field_fwd3_smem_bytes() { return 2 * kSlotBytes + kStages3 * kStageBytes3 + sizeof(Shared3) + 64; }

cudaError_t launch_field_fwd3(const FieldFwdParams& p, bool has_bender, int num_sms, cudaStream_t stream) {
  const size_t smem = field_fwd3_smem_bytes();
  const int n_pairs = (p.n_tiles + 1) / 2;
  if (n_pairs <= 0) return cudaSuccess;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && !g_sched_uploaded[dev]) {
    // first use on this device (never inside a stream capture: ops.py warms the kernels up before capturing)
    Schedule both[2] = {build_schedule(false), build_schedule(true)};
    e = cudaMemcpyToSymbol(c_sched, both, sizeof(both));
    if (e != cudaSuccess) return e;
    g_sched_uploaded[dev] = true;
  }
  const int grid = n_pairs < num_sms ? n_pairs : num_sms;
  const uint8_t* split = p.nerf_w + (kNerfSOffset - 0);   // nerf_w points at the packed buffer's start
  if (has_bender) {
    e = cudaFuncSetAttribute(field_fwd3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    field_fwd3_kernel<true><<<grid, kFwd3Threads, smem, stream>>>(p, split);
  } else {
    e = cudaFuncSetAttribute(field_fwd3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    field_fwd3_kernel<false><<<grid, kFwd3Threads, smem, stream>>>(p, split);
  }
  return cudaGetLastError();
}

}  // namespace nrn
