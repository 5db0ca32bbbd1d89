// Gradient all-reduce fused into the optimizer launch, over NVLink peer memory (no NCCL on the data path).
//
// Replaces what DataParallel's backward does on GPU 0 (reduce_add of every replica's gradients, train.py:290-297,
// :1606-1608) followed by torch.optim.Adam.step: one process per GPU, every rank owns a "window" -- one cudaMalloc'ed
// buffer opened by all other ranks of the node through CUDA IPC:
//
//   window = [ flags: 3 x kPeerMaxRanks u32 | row slots 2 x slot_bytes | gradient arena (total floats) ]
//
// The optimizer's gradient arena (every p.grad is a view of it) lives INSIDE the window, so the backward kernels
// write their weight gradients where the peers can read them.  One optimizer step is three launches, all plain
// kernels (capturable in a CUDA graph, re-entrant through a device-resident epoch counter):
//
//   peer_tick_kernel   (1 block)  epoch += 1; per-tensor Adam step counts += 1; system fence; ARRIVE[my rank] = epoch
//                                 stored into every peer's window (st.release.sys over NVLink)
//   peer_adam_kernel   (n blocks) wait until ARRIVE[r] >= epoch for all r (spin on LOCAL memory); then for its 2048
//                                 elements: g = sum_r arena_r[i] read from the peers' windows in rank order (every rank
//                                 adds the same numbers in the same order: bit-identical sums, replicated Adam stays
//                                 replicated without a broadcast); Adam update of the local parameters / moments;
//                                 g kept in a local buffer; last block: DONE[my rank] = epoch to every peer
//   peer_finish_kernel (n blocks) wait until DONE[r] >= epoch for all r (nobody reads this rank's arena any more), then
//                                 arena = g (p.grad holds the reduced gradient, like after an in-place all-reduce)
//
// The transfer is the reduction's operand fetch: no staging copy, no second pass over the gradients, 7/8 of the bytes
// cross the NVSwitch exactly once per reader.  4.3 MB per rank: at 8 ranks each GPU pulls 30 MB.
//
// peer_gather_rows: all-gather of small per-ray tensors (the per-ray losses the training loop prints) through the same
// windows: publish into slot[epoch & 1], ARRIVE, then every rank copies the peers' slots.  Double buffering makes a
// second barrier unnecessary (a rank can only reach publish(e+2) after every peer finished collect(e)).
#include "peer.cuh"
#include "adam.cuh"

namespace nrn {

namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer(const float* p) {
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

constexpr long long kPeerWaitCycles = 1ll << 32;   // ~2 s: a missing peer becomes an error word, not a hung GPU

// spin until flags[r] has reached `epoch` for every rank (wrap-safe comparison); thread 0 of a block
__device__ __forceinline__ bool wait_all(const uint32_t* flags, int world, uint32_t epoch, int* err, int code) {
  for (int r = 0; r < world; ++r) {
    const long long t0 = clock64();
    while (static_cast<int32_t>(ld_acquire_sys(flags + r) - epoch) < 0) {
      if (clock64() - t0 > kPeerWaitCycles) {
        atomicCAS(err, 0, code);
        return false;
      }
      __nanosleep(64);
    }
  }
  return true;
}

__global__ void peer_tick_kernel(const PeerCtx c, uint32_t* epoch_word, long long* step, int n_tensors, int which) {
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0) {
    s_epoch = epoch_word[which] + 1u;
    epoch_word[which] = s_epoch;
  }
  __syncthreads();
  if (step) {
    for (int t = threadIdx.x; t < n_tensors; t += blockDim.x) step[t] += 1;
  }
  __threadfence_system();   // everything earlier kernels of this stream wrote into the window is visible system-wide
  __syncthreads();
  if (threadIdx.x < c.world) {
    uint32_t* peer_flags = reinterpret_cast<uint32_t*>(c.window[threadIdx.x]);
    st_release_sys(peer_flags + which * kPeerMaxRanks + c.rank, s_epoch);
  }
}

__global__ void __launch_bounds__(256) peer_adam_kernel(const PeerCtx c, const AdamParams a, const uint32_t* epoch_word,
                                                        float* __restrict__ gsum, unsigned int* done_counter, int* err) {
  __shared__ float s_c[2];
  __shared__ int s_ok;
  const AdamBlock b = a.blocks[blockIdx.x];
  const uint32_t epoch = epoch_word[kPeerArrive];
  if (threadIdx.x == 0) {
    const uint32_t* flags = reinterpret_cast<const uint32_t*>(c.window[c.rank]) + kPeerArrive * kPeerMaxRanks;
    s_ok = wait_all(flags, c.world, epoch, err, 901) ? 1 : 0;
    const double t = static_cast<double>(a.step[b.tensor]);
    const double bc1 = 1.0 - pow(static_cast<double>(a.beta1), t);
    const double bc2 = 1.0 - pow(static_cast<double>(a.beta2), t);
    s_c[0] = static_cast<float>(static_cast<double>(a.lr[0]) / bc1);
    s_c[1] = static_cast<float>(sqrt(bc2));
  }
  __syncthreads();
  if (s_ok) {
    const float step_size = s_c[0], bc2_sqrt = s_c[1];
    const float w1 = 1.0f - a.beta1, w2 = 1.0f - a.beta2;
    float* __restrict__ p = a.params + b.flat_off;
    float* __restrict__ m = a.exp_avg + b.flat_off;
    float* __restrict__ v = a.exp_avg_sq + b.flat_off;
    constexpr int kPer = kAdamBlockElems / 256;
    float gi[kPer], mi[kPer], vi[kPer], pi[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) gi[k] = 0.f;
    // rank order 0, 1, 2, ...: identical on every rank -> identical sums; all loads of one peer are in flight together
    for (int r = 0; r < c.world; ++r) {
      const float* __restrict__ src = reinterpret_cast<const float*>(c.window[r] + c.arena_off) + b.flat_off;
      float t[kPer];
#pragma unroll
      for (int k = 0; k < kPer; ++k) {
        const int i = threadIdx.x + k * 256;
        t[k] = i < b.count ? ld_peer(src + i) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < kPer; ++k) gi[k] += t[k];
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int i = threadIdx.x + k * 256;
      const bool in = i < b.count;
      mi[k] = in ? m[i] : 0.f;
      vi[k] = in ? v[i] : 0.f;
      pi[k] = in ? p[i] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < b.count) {
        const float mk = mi[k] + (gi[k] - mi[k]) * w1;
        const float vk = a.beta2 * vi[k] + w2 * gi[k] * gi[k];
        m[i] = mk;
        v[i] = vk;
        p[i] = pi[k] - step_size * (mk / (sqrtf(vk) / bc2_sqrt + a.eps));
        gsum[b.flat_off + i] = gi[k];
      }
    }
  }
  // the last block to finish tells every peer that this rank no longer reads their arenas
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(done_counter, 1u);
    if (prev == gridDim.x - 1) {
      *done_counter = 0u;
      __threadfence_system();
      for (int r = 0; r < c.world; ++r) {
        uint32_t* peer_flags = reinterpret_cast<uint32_t*>(c.window[r]);
        st_release_sys(peer_flags + kPeerDone * kPeerMaxRanks + c.rank, epoch);
      }
    }
  }
}

__global__ void __launch_bounds__(256) peer_finish_kernel(const PeerCtx c, const uint32_t* epoch_word, const float* __restrict__ gsum,
                                                          long long total, int* err) {
  __shared__ int s_ok;
  if (threadIdx.x == 0) {
    const uint32_t* flags = reinterpret_cast<const uint32_t*>(c.window[c.rank]) + kPeerDone * kPeerMaxRanks;
    s_ok = wait_all(flags, c.world, epoch_word[kPeerArrive], err, 902) ? 1 : 0;
  }
  __syncthreads();
  if (!s_ok) return;
  float* __restrict__ arena = reinterpret_cast<float*>(c.window[c.rank] + c.arena_off);
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) arena[i] = gsum[i];
}

// ---- row all-gather -------------------------------------------------------------------------------------------
__global__ void peer_publish_kernel(const PeerCtx c, const uint32_t* epoch_word, const float* __restrict__ local, int n) {
  // epoch of the gather about to be announced = stored epoch + 1 (the tick kernel that follows increments it)
  const uint32_t next = epoch_word[kPeerGather] + 1u;
  float* slot = reinterpret_cast<float*>(c.window[c.rank] + c.slot_off + static_cast<size_t>(next & 1u) * c.slot_bytes);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) slot[i] = local[i];
}

__global__ void peer_collect_kernel(const PeerCtx c, const uint32_t* epoch_word, float* __restrict__ out, int n_per_rank, int* err) {
  __shared__ int s_ok;
  const uint32_t epoch = epoch_word[kPeerGather];
  if (threadIdx.x == 0) {
    const uint32_t* flags = reinterpret_cast<const uint32_t*>(c.window[c.rank]) + kPeerGather * kPeerMaxRanks;
    s_ok = wait_all(flags, c.world, epoch, err, 903) ? 1 : 0;
  }
  __syncthreads();
  if (!s_ok) return;
  const int total = n_per_rank * c.world;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / n_per_rank, j = i - r * n_per_rank;
    const float* slot = reinterpret_cast<const float*>(c.window[r] + c.slot_off + static_cast<size_t>(epoch & 1u) * c.slot_bytes);
    out[i] = ld_peer(slot + j);
  }
}

}  // namespace

cudaError_t launch_peer_reduce_adam(const PeerCtx& c, const AdamParams& a, int n_tensors, int n_blocks, long long total,
                                    uint32_t* epoch_word, float* gsum, unsigned int* done_counter, int* err, cudaStream_t st) {
  if (n_blocks <= 0) return cudaSuccess;
  peer_tick_kernel<<<1, 256, 0, st>>>(c, epoch_word, a.step, n_tensors, kPeerArrive);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  peer_adam_kernel<<<n_blocks, 256, 0, st>>>(c, a, epoch_word, gsum, done_counter, err);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int blocks = static_cast<int>((total + 256 * 8 - 1) / (256 * 8));
  peer_finish_kernel<<<blocks, 256, 0, st>>>(c, epoch_word, gsum, total, err);
  return cudaGetLastError();
}

cudaError_t launch_peer_gather(const PeerCtx& c, uint32_t* epoch_word, const float* local, int n_per_rank, float* out, int* err,
                               cudaStream_t st) {
  if (n_per_rank <= 0) return cudaSuccess;
  const int blocks = (n_per_rank + 255) / 256 < 64 ? (n_per_rank + 255) / 256 : 64;
  peer_publish_kernel<<<blocks, 256, 0, st>>>(c, epoch_word, local, n_per_rank);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  peer_tick_kernel<<<1, 256, 0, st>>>(c, epoch_word, nullptr, 0, kPeerGather);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int total = n_per_rank * c.world;
  const int cb = (total + 255) / 256 < 128 ? (total + 255) / 256 : 128;
  peer_collect_kernel<<<cb, 256, 0, st>>>(c, epoch_word, out, n_per_rank, err);
  return cudaGetLastError();
}

}  // namespace nrn
