// Fused point-wise field evaluation, backward data-gradient chain (DGRAD) -- "shared-slab" schedule.
//
// Same arithmetic, same stash formats and bit-identical results as field_bwd.cu (see its header: what torch.autograd derives
// for NeRF.forward + ray_bending.forward + Embedder.embed, run_nerf_helpers.py:240-314, :507-584, :149-150, given dL/draw;
// SURVEY.md appendix C), on the machine of field_fwd3.cu:
//   * ONE stream of W^T pieces (<= 16 KB: [K = 64] x [128 input features]) per CTA, consumed by both tile slots;
//   * every 256-wide gradient dh_l is produced as two halves of 128 columns in two accumulators per slot, so the epilogue of
//     half 0 (TMEM drain, ReLU mask from the forward stash, fp16 pack, gradient-stash store) overlaps the MMAs of half 1, and
//     the next layer starts on the K range half 0 produced.  Four epilogue warpgroups: (slot, half).
//   step  0      head^T   dh8 = d_raw . Wout          -> dY7 = dh8 * [h8 > 0]
//   steps 1,2    L7^T, L6^T                           -> dY6, dY5
//   step  3      L5e^T    dE  = dY5 . W5[:, :63]      -> PE backward -> d(bent xyz)          (64 columns, primary warpgroup)
//   steps 4..8   L5h^T, L4^T..L1^T                    -> dY4..dY0
//   step  9      L0^T     dE                          -> PE backward; bend backward -> dYb4 (primary warpgroup)
//   steps 10..14 B4^T..B0^T (bender)                  -> dYb3..dYb0, per-ray latent gradient (primary warpgroup)
// Barriers as in field_fwd3.cu: a_ready[kh] (256 arrivals: both slots' warpgroups of that half), d_full[nh] and a_free
// (tcgen05.commit).  The gradient operand is overwritten in place: half 0 of a step may only be stored once the MMAs of half 1
// that read K columns 0..127 are done (a_free).
#include "nrn_common.cuh"
#include "sm100_ptx.cuh"

namespace nrn {

namespace {

constexpr long long kWaitLimitB3 = 1ll << 28;
constexpr int kStagesB3 = 6;
constexpr int kStageBytesB3 = 16384;
constexpr int kBwd3Threads = 640;

enum : int { CB_NONE = 0, CB_D0 = 1, CB_D1 = 2, CB_AFREE = 4 };

struct PieceB {
  uint32_t src_off;
  uint16_t bytes_div16;
  uint16_t n;
  uint32_t a_off;       // byte offset of the A operand inside a slot's 64 KB gradient image
  uint8_t k16;
  uint8_t bender;
  uint8_t acc_col;
  uint8_t first;
  int8_t wait0;
  int8_t wait1;
  uint8_t commit;
  uint8_t pad;
};
constexpr int kMaxPiecesB = 112;
struct ScheduleB {
  int n;
  PieceB p[kMaxPiecesB];
};
__constant__ ScheduleB c_sched_b[2];   // [0] without bender, [1] with bender

struct SharedB3 {
  uint64_t w_full[kStagesB3];
  uint64_t w_empty[kStagesB3];
  uint64_t a_ready[2];
  uint64_t d_full[2];
  uint64_t a_free;
  uint32_t tmem_base;
  int abort_flag;
};

struct WaiterB3 {
  int* s_abort;
  int* g_err;
  __device__ __forceinline__ bool wait(uint64_t* bar, uint32_t parity, int code) const {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
      if (*reinterpret_cast<volatile int*>(s_abort)) return false;
      if (clock64() - t0 > kWaitLimitB3) {
        atomicExch(s_abort, code);
        atomicCAS(g_err, 0, code);
        return false;
      }
    }
    return true;
  }
};

__device__ __forceinline__ float clamp_h3(float v) { return fminf(fmaxf(v, -65504.f), 65504.f); }

// dY = dh * [h > 0] for NCOLS accumulator columns (see field_bwd.cu: epi_mask_store); the caller offsets taddr / mask_row /
// dst_row to the columns it owns
template <int NCOLS>
__device__ __forceinline__ void epi3_mask_store(uint32_t taddr, const uint8_t* __restrict__ mask_row, uint8_t* dst_row) {
  constexpr int NC = NCOLS / 32;
  uint32_t v[2][32];
  tmem_ld32(taddr, v[0]);
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    uint4 m[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) m[q] = __ldg(reinterpret_cast<const uint4*>(mask_row + (c * 4 + q) * kChunkBytes));
    tmem_ld_wait();
    if (c + 1 < NC) tmem_ld32(taddr + (c + 1) * 32, v[(c + 1) & 1]);
    const uint32_t(&w)[32] = v[c & 1];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t mw[4] = {m[q].x, m[q].y, m[q].z, m[q].w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t g2 = pack_h2_sat(__uint_as_float(w[q * 8 + 2 * j]), __uint_as_float(w[q * 8 + 2 * j + 1]));
        const __half2 hm = __hgt2(*reinterpret_cast<const __half2*>(&mw[j]), __float2half2_rn(0.f));
        const __half2 r2 = __hmul2(*reinterpret_cast<const __half2*>(&g2), hm);
        o[j] = *reinterpret_cast<const uint32_t*>(&r2);
      }
      *reinterpret_cast<uint4*>(dst_row + (c * 4 + q) * kChunkBytes) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

__device__ __forceinline__ float h3_lo(uint32_t w) { return __half2float(__ushort_as_half(static_cast<unsigned short>(w & 0xffffu))); }
__device__ __forceinline__ float h3_hi(uint32_t w) { return __half2float(__ushort_as_half(static_cast<unsigned short>(w >> 16))); }

// Backward of the positional encoding (identical to field_bwd.cu: pe_backward)
__device__ __forceinline__ void pe_backward3(uint32_t taddr, const uint8_t* __restrict__ e_row, float (&dx)[3]) {
  float de[64];
  {
    uint32_t v[32];
    tmem_ld32(taddr, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) de[i] = __uint_as_float(v[i]);
    tmem_ld32(taddr + 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) de[32 + i] = __uint_as_float(v[i]);
  }
  float e[64];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 w = __ldg(reinterpret_cast<const uint4*>(e_row + c * kChunkBytes));
    e[c * 8 + 0] = h3_lo(w.x); e[c * 8 + 1] = h3_hi(w.x); e[c * 8 + 2] = h3_lo(w.y); e[c * 8 + 3] = h3_hi(w.y);
    e[c * 8 + 4] = h3_lo(w.z); e[c * 8 + 5] = h3_hi(w.z); e[c * 8 + 6] = h3_lo(w.w); e[c * 8 + 7] = h3_hi(w.w);
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float acc = de[d];
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      const float f = static_cast<float>(1 << k);
      const float s = e[3 + 6 * k + d], c = e[3 + 6 * k + 3 + d];
      acc += f * (de[3 + 6 * k + d] * c - de[3 + 6 * k + 3 + d] * s);
    }
    dx[d] += acc;
  }
}

__device__ __forceinline__ float warp_transpose_reduce3(float (&v)[32], int lane) {
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

}  // namespace

template <bool HAS_BENDER>
__global__ void __launch_bounds__(kBwd3Threads, 1) field_bwd3_kernel(const FieldBwdParams p, const uint8_t* __restrict__ nerf_ts) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* act = smem;                       // 2 slots x 64 KB gradient operand
  uint8_t* ring = smem + 2 * kHBytes;        // kStagesB3 x 16 KB
  SharedB3* sh = reinterpret_cast<SharedB3*>(ring + kStagesB3 * kStageBytesB3);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_pairs = (p.n_tiles + 1) >> 1;
  const ScheduleB& sched = c_sched_b[HAS_BENDER ? 1 : 0];

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStagesB3; ++i) {
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
  const WaiterB3 W{&sh->abort_flag, p.err};

  if (warp == 0) {
    // ===================== weight producer (W^T pieces) =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
#pragma unroll 1
        for (int i = 0; i < sched.n; ++i) {
          const PieceB pc = sched.p[i];
          const uint32_t bytes = static_cast<uint32_t>(pc.bytes_div16) * 16u;
          const uint8_t* src = (pc.bender ? p.bend_wT : nerf_ts) + pc.src_off;
          W.wait(&sh->w_empty[stage], phase ^ 1u, 101);
          mbar_arrive_expect_tx(&sh->w_full[stage], bytes);
          tma_bulk_g2s(ring + stage * kStageBytesB3, src, bytes, &sh->w_full[stage]);
          if (++stage == kStagesB3) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      uint32_t aph[2] = {0u, 0u};
      const uint32_t a_base0 = smem_u32(act), a_base1 = smem_u32(act + kHBytes);
      for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
#pragma unroll 1
        for (int i = 0; i < sched.n; ++i) {
          const PieceB pc = sched.p[i];
          if (pc.wait0) { W.wait(&sh->a_ready[0], aph[0], 201); aph[0] ^= 1u; }
          if (pc.wait1) { W.wait(&sh->a_ready[1], aph[1], 203); aph[1] ^= 1u; }
          W.wait(&sh->w_full[stage], phase, 202);
          tc_fence_after_sync();
          const uint32_t n = pc.n;
          const uint32_t idesc = umma_instr_desc(kTileM, n, UMMA_F16, UMMA_F16, UMMA_K_MAJOR, UMMA_K_MAJOR);
          const uint64_t bdesc = umma_smem_desc(smem_u32(ring + stage * kStageBytesB3), n * 16, 128);
          const uint64_t adesc0 = umma_smem_desc(a_base0 + pc.a_off, kChunkBytes, 128);
          const uint64_t adesc1 = umma_smem_desc(a_base1 + pc.a_off, kChunkBytes, 128);
          const uint32_t d0 = tmem_base + pc.acc_col, d1 = tmem_base + 256 + pc.acc_col;
          for (uint32_t k = 0; k < pc.k16; ++k)
            umma_f16_ss(d0, umma_desc_advance(adesc0, k * 2 * kChunkBytes), umma_desc_advance(bdesc, k * 2 * n * 16), idesc,
                        (k | (pc.first ^ 1u)) ? 1u : 0u);
          for (uint32_t k = 0; k < pc.k16; ++k)
            umma_f16_ss(d1, umma_desc_advance(adesc1, k * 2 * kChunkBytes), umma_desc_advance(bdesc, k * 2 * n * 16), idesc,
                        (k | (pc.first ^ 1u)) ? 1u : 0u);
          umma_commit(&sh->w_empty[stage]);
          if (pc.commit & CB_AFREE) umma_commit(&sh->a_free);
          if (pc.commit & CB_D0) umma_commit(&sh->d_full[0]);
          if (pc.commit & CB_D1) umma_commit(&sh->d_full[1]);
          if (++stage == kStagesB3) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warpgroups: (slot, output half) =====================
    const int wg = (warp - 4) >> 2;
    const int slot = wg & 1;
    const int half = wg >> 1;
    const int row = ((warp & 3) << 5) | lane;
    uint8_t* a_img = act + slot * kHBytes;
    uint8_t* a_row = a_img + row * 16;
    const uint32_t taddr = tmem_base + ((static_cast<uint32_t>(warp & 3) * 32u) << 16) + slot * 256 + half * 128;
    const bool wg_leader = (threadIdx.x & 127) == 0;
    const int bar_id = 1 + wg;
    uint32_t dph = 0, fph = 0;
    auto signal_ready = [&]() {
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(&sh->a_ready[half]);
    };
    auto wait_acc = [&](int code) {
      W.wait(&sh->d_full[half], dph, code);
      dph ^= 1u;
      tc_fence_after_sync();
    };
    float scale = 1.0f;
    {
      const float amax = p.amax ? __ldg(p.amax) : 0.f;
      if (amax > 0.f && amax < 3.0e38f) {
        int e;
        frexpf(amax, &e);
        scale = ldexpf(1.0f, min(max(10 - e, -60), 60));
      }
    }
    const float inv_scale = 1.0f / scale;

    for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      const long long tile = static_cast<long long>(pair) * 2 + slot;
      const long long pt = tile * kTileM + row;
      const bool valid = pt < p.P;
      const uint8_t* st_tile = p.stash + tile * kStashTileBytes;
      const uint8_t* st = st_tile + row * 16;
      uint8_t* gs = p.gstash + tile * kGradTileBytes;
      auto stash_begin = [&]() {
        if (wg_leader) tma_bulk_wait_read<0>();
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      };
      // store [bytes] of this slot's gradient image, starting at img_off, to the gradient stash at off
      auto stash_store = [&](uint32_t off, uint32_t img_off, uint32_t bytes) {
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (wg_leader) {
          for (uint32_t o = 0; o < bytes; o += 16384u) tma_bulk_s2g(gs + off + o, a_img + img_off + o, bytes - o < 16384u ? bytes - o : 16384u);
          tma_bulk_commit();
        }
      };
      // this warpgroup's half of the forward image the NEXT masked step reads, pulled into L2 ahead of the per-thread loads
      auto prefetch_half = [&](uint32_t off) {
        if (wg_leader) {
          const uint8_t* q = st_tile + off + half * 16 * kChunkBytes;
          tma_prefetch_l2(q, 16384u);
          tma_prefetch_l2(q + 16384u, 16384u);
        }
      };
      auto prefetch = [&](uint32_t off, uint32_t bytes) {
        if (wg_leader) {
          for (uint32_t o = 0; o < bytes; o += 16384u) tma_prefetch_l2(st_tile + off + o, bytes - o < 16384u ? bytes - o : 16384u);
        }
      };

      if (half == 1) {
        // ---- columns 128..255 of dY7, dY6, dY5 and of dY4 .. dY0 ----
        prefetch_half(kStH + 7 * kHBytes);
#pragma unroll 1
        for (int s = 0; s < 8; ++s) {
          const int l = s < 3 ? 7 - s : 7 - s;      // s = 0..2 -> dY7..dY5, s = 3..7 -> dY4..dY0
          wait_acc(330 + s);
          if (l > 0) prefetch_half(kStH + (l - 1) * kHBytes);
          stash_begin();
          epi3_mask_store<128>(taddr, st + kStH + l * kHBytes + 16 * kChunkBytes, a_row + 16 * kChunkBytes);
          stash_store(kGsY + l * kHBytes + 16 * kChunkBytes, 16 * kChunkBytes, 16 * kChunkBytes);
          signal_ready();
        }
        continue;
      }

      // ---- primary warpgroup ----
      prefetch_half(kStH + 7 * kHBytes);
      {
        float g[4] = {0.f, 0.f, 0.f, 0.f};
        if (valid) {
          const float* q = p.d_raw + pt * p.out_ch;
#pragma unroll
          for (int c = 0; c < 4; ++c) g[c] = clamp_h3(__ldg(q + c) * scale);
        }
        const uint4 c0 = make_uint4(pack_h2(g[0], g[1]), pack_h2(g[2], g[3]), 0u, 0u);
        const uint4 zz = make_uint4(0u, 0u, 0u, 0u);
        stash_begin();
        *reinterpret_cast<uint4*>(a_row) = c0;
        *reinterpret_cast<uint4*>(a_row + kChunkBytes) = zz;
        stash_store(kGsRaw, 0, 2 * kChunkBytes);
      }
      signal_ready();
      float dx[3] = {0.f, 0.f, 0.f};
      // ---- head^T, L7^T, L6^T : columns 0..127 of dY7, dY6, dY5 ----
#pragma unroll 1
      for (int s = 0; s < 3; ++s) {
        wait_acc(300 + s);
        W.wait(&sh->a_free, fph, 340 + s);
        fph ^= 1u;
        if (s < 2) prefetch_half(kStH + (6 - s) * kHBytes); else prefetch(kStE, kEBytes);
        stash_begin();
        epi3_mask_store<128>(taddr, st + kStH + (7 - s) * kHBytes, a_row);
        stash_store(kGsY + (7 - s) * kHBytes, 0, 16 * kChunkBytes);
        signal_ready();
      }
      // ---- L5e^T: gradient into the skip-connected embedding ----
      wait_acc(303);
      prefetch_half(kStH + 4 * kHBytes);
      pe_backward3(taddr, st + kStE, dx);
      signal_ready();   // A operand (dY5) untouched; accumulator half 0 drained
      // ---- L5h^T, L4^T .. L1^T : columns 0..127 of dY4 .. dY0 ----
#pragma unroll 1
      for (int s = 0; s < 5; ++s) {
        wait_acc(304 + s);
        W.wait(&sh->a_free, fph, 344 + s);
        fph ^= 1u;
        if (s < 4) prefetch_half(kStH + (3 - s) * kHBytes); else prefetch(kStE, kEBytes);
        stash_begin();
        epi3_mask_store<128>(taddr, st + kStH + (4 - s) * kHBytes, a_row);
        stash_store(kGsY + (4 - s) * kHBytes, 0, 16 * kChunkBytes);
        signal_ready();
      }
      // ---- L0^T: gradient into the embedding; then through the bend ----
      wait_acc(309);
      if (HAS_BENDER) prefetch(kStHb4, 8 * kChunkBytes);
      pe_backward3(taddr, st + kStE, dx);
      if (!HAS_BENDER) continue;

      float rig = 0.f, drpre = 0.f;
      {
        float un[3] = {0.f, 0.f, 0.f}, dun[3], dm[3];
        float up_r = 0.f, up_u[3] = {0.f, 0.f, 0.f};
        if (valid) {
          rig = __ldg(p.rigidity + pt);
#pragma unroll
          for (int d = 0; d < 3; ++d) un[d] = __ldg(p.unmasked + pt * 3 + d);
          if (p.d_rigid_up) up_r = __ldg(p.d_rigid_up + pt) * scale;
          if (p.d_unmasked_up) {
#pragma unroll
            for (int d = 0; d < 3; ++d) up_u[d] = __ldg(p.d_unmasked_up + pt * 3 + d) * scale;
          }
        }
        float dr = up_r;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          dm[d] = p.use_scaling ? dx[d] * p.scaling : dx[d];
          dun[d] = rig * dm[d] + up_u[d];
          dr += un[d] * dm[d];
        }
        drpre = dr * 2.0f * rig * (1.0f - rig);
        if (p.use_cutoff && rig <= p.cutoff) drpre = 0.f;
        if (!valid) { dun[0] = dun[1] = dun[2] = 0.f; drpre = 0.f; }
        const uint4 c0 = make_uint4(pack_h2(clamp_h3(dun[0]), clamp_h3(dun[1])), pack_h2(clamp_h3(dun[2]), 0.f), 0u, 0u);
        const uint4 zz = make_uint4(0u, 0u, 0u, 0u);
        stash_begin();
        *reinterpret_cast<uint4*>(a_row) = c0;
        *reinterpret_cast<uint4*>(a_row + kChunkBytes) = zz;
        stash_store(kGsYb4, 0, 2 * kChunkBytes);
      }
      signal_ready();
      // ---- B4^T -> dYb3 ----
      wait_acc(310);
      prefetch(kStHb3, 8 * kChunkBytes);
      stash_begin();
      epi3_mask_store<64>(taddr, st + kStHb4, a_row);
      stash_store(kGsYb3, 0, 8 * kChunkBytes);
      signal_ready();
      // ---- B3^T -> dYb2 = [dh * mask (64) | d rigidity pre-activation | 0 (15)] ----
      wait_acc(311);
      prefetch(kStHb2, 12 * kChunkBytes);
      stash_begin();
      epi3_mask_store<64>(taddr, st + kStHb3, a_row);
      {
        const uint4 c8 = make_uint4(pack_h2(clamp_h3(drpre), 0.f), 0u, 0u, 0u);
        const uint4 zz = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(a_row + 8 * kChunkBytes) = c8;
        *reinterpret_cast<uint4*>(a_row + 9 * kChunkBytes) = zz;
      }
      stash_store(kGsYb2, 0, 10 * kChunkBytes);
      signal_ready();
      // ---- B2^T -> dYb1, B1^T -> dYb0 ----
      wait_acc(312);
      prefetch(kStHb1, 12 * kChunkBytes);
      stash_begin();
      epi3_mask_store<96>(taddr, st + kStHb2, a_row);
      stash_store(kGsYb1, 0, 12 * kChunkBytes);
      signal_ready();
      wait_acc(313);
      stash_begin();
      epi3_mask_store<96>(taddr, st + kStHb1, a_row);
      stash_store(kGsYb0, 0, 12 * kChunkBytes);
      signal_ready();
      // ---- B0^T: d(bender input); columns 6..37 are the latent code -> per-ray reduction ----
      wait_acc(314);
      {
        float dl[32];
        {
          uint32_t v[32];
          tmem_ld32(taddr, v);
          uint32_t w[16];
          tmem_ld16(taddr + 32, w);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 26; ++i) dl[i] = __uint_as_float(v[6 + i]) * inv_scale;
#pragma unroll
          for (int i = 0; i < 6; ++i) dl[26 + i] = __uint_as_float(w[i]) * inv_scale;
        }
        const long long ray = valid ? pt / p.S : -1;
        const long long ray0 = __shfl_sync(0xffffffffu, ray, 0);
        if (__all_sync(0xffffffffu, ray == ray0 && valid)) {
          const float tot = warp_transpose_reduce3(dl, lane);
          atomicAdd(p.d_latents + ray0 * kLatent + lane, tot);
        } else if (valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i) atomicAdd(p.d_latents + ray * kLatent + i, dl[i]);
        }
      }
    }
    if ((threadIdx.x & 127) == 0) tma_bulk_wait<0>();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
namespace {

void add_piece_b(ScheduleB& s, uint32_t src_off, uint32_t bytes, int n, int k16, uint32_t a_off, bool bender, int acc_col, bool first,
                 bool wait0, bool wait1, int commit) {
  PieceB& q = s.p[s.n++];
  q.src_off = src_off; q.bytes_div16 = static_cast<uint16_t>(bytes / 16); q.n = static_cast<uint16_t>(n); q.a_off = a_off;
  q.k16 = static_cast<uint8_t>(k16); q.bender = bender ? 1 : 0; q.acc_col = static_cast<uint8_t>(acc_col); q.first = first ? 1 : 0;
  q.wait0 = wait0 ? 1 : 0; q.wait1 = wait1 ? 1 : 0; q.commit = static_cast<uint8_t>(commit); q.pad = 0;
}

// one 256 -> 256 step: (half, K piece of 64) pieces.  `fresh_a`: the A operand was written by the previous step's epilogues
// (wait for its K halves); otherwise A is the image the previous step already read (L5h^T after L5e^T): only the half-0
// accumulator has to be free.
void add_main_step(ScheduleB& s, uint32_t& off, bool fresh_a) {
  for (int nh = 0; nh < 2; ++nh) {
    for (int ks = 0; ks < 4; ++ks) {
      bool w0 = false, w1 = false;
      if (nh == 0) { w0 = ks == 0; w1 = fresh_a && ks == 2; }
      int commit = CB_NONE;
      if (ks == 3) commit |= nh == 0 ? CB_D0 : CB_D1;
      if (nh == 1 && ks == 1) commit |= CB_AFREE;
      add_piece_b(s, off, 8 * 128 * 16, 128, 4, ks * 8 * kChunkBytes, false, nh * 128, ks == 0, w0, w1, commit);
      off += 8 * 128 * 16;
    }
  }
}
// a 256 -> 64 step into the embedding (L5e^T, L0^T): two K pieces of the unchanged [32 chunks][64 rows][8] image
void add_embed_step(ScheduleB& s, uint32_t& off) {
  add_piece_b(s, off, 16384, 64, 8, 0, false, 0, true, true, true, CB_NONE);
  add_piece_b(s, off + 16384, 16384, 64, 8, 16 * kChunkBytes, false, 0, false, false, false, CB_D0);
  off += 32768;
}

ScheduleB build_schedule_b(bool has_bender) {
  ScheduleB s{};
  s.n = 0;
  uint32_t off = 0;
  // step 0, head^T: A = d_raw image (K = 16), two pieces [2 chunks][128 rows][8]
  add_piece_b(s, off, 2 * 128 * 16, 128, 1, 0, false, 0, true, true, false, CB_D0); off += 2 * 128 * 16;
  add_piece_b(s, off, 2 * 128 * 16, 128, 1, 0, false, 128, true, false, false, CB_D1 | CB_AFREE); off += 2 * 128 * 16;
  add_main_step(s, off, true);    // L7^T
  add_main_step(s, off, true);    // L6^T
  add_embed_step(s, off);         // L5e^T
  add_main_step(s, off, false);   // L5h^T (same A as L5e^T)
  for (int i = 0; i < 4; ++i) add_main_step(s, off, true);   // L4^T .. L1^T
  add_embed_step(s, off);         // L0^T
  if (has_bender) {
    uint32_t b = 0;
    add_piece_b(s, b, kBendTB4Bytes, 64, 1, 0, true, 0, true, true, false, CB_D0); b += kBendTB4Bytes;
    add_piece_b(s, b, kBendTB3Bytes, 64, 4, 0, true, 0, true, true, false, CB_D0); b += kBendTB3Bytes;
    add_piece_b(s, b, kBendTB2Bytes, 96, 5, 0, true, 0, true, true, false, CB_D0); b += kBendTB2Bytes;
    // B1^T: K = 96 in two pieces (8 + 4 chunks of 96 rows)
    add_piece_b(s, b, 8 * 96 * 16, 96, 4, 0, true, 0, true, true, false, CB_NONE);
    add_piece_b(s, b + 8 * 96 * 16, 4 * 96 * 16, 96, 2, 8 * kChunkBytes, true, 0, false, false, false, CB_D0); b += kBendTB1Bytes;
    add_piece_b(s, b, kBendTB0Bytes, 48, 6, 0, true, 0, true, true, false, CB_D0);
  }
  return s;
}

bool g_sched_b_uploaded[64] = {};

}  // namespace

cudaError_t launch_field_bwd3(const FieldBwdParams& p, const uint8_t* nerf_packed_base, bool has_bender, int num_sms, cudaStream_t stream) {
  const size_t smem = 2 * kHBytes + kStagesB3 * kStageBytesB3 + sizeof(SharedB3) + 64;
  const int n_pairs = (p.n_tiles + 1) / 2;
  if (n_pairs <= 0) return cudaSuccess;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && !g_sched_b_uploaded[dev]) {
    ScheduleB both[2] = {build_schedule_b(false), build_schedule_b(true)};
    e = cudaMemcpyToSymbol(c_sched_b, both, sizeof(both));
    if (e != cudaSuccess) return e;
    g_sched_b_uploaded[dev] = true;
  }
  const int grid = n_pairs < num_sms ? n_pairs : num_sms;
  const uint8_t* ts = nerf_packed_base + kNerfTSOffset;
  if (has_bender) {
    e = cudaFuncSetAttribute(field_bwd3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    field_bwd3_kernel<true><<<grid, kBwd3Threads, smem, stream>>>(p, ts);
  } else {
    e = cudaFuncSetAttribute(field_bwd3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    field_bwd3_kernel<false><<<grid, kBwd3Threads, smem, stream>>>(p, ts);
  }
  return cudaGetLastError();
}

}  // namespace nrn
