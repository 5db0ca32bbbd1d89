// Interference probe: issue rate of back-to-back tcgen05.mma (M=128, N=256, K=16, fp16 SS, no-swizzle chunk-major
// operands) while other warps of the same CTA do what the field kernels' epilogues/producers do:
//   bit 0: tcgen05.ld of the other accumulator (4 warps)      bit 1: 16-byte shared-memory stores (same 4 warps)
//   bit 2: bulk-TMA global->shared writes (weight ring)        bit 3: tcgen05.commit after every 4 MMAs
//   bit 4: a second warpgroup doing the same ld/st             bit 5: fp32 ALU work in the warpgroups (bias+ReLU+pack)
//   bit 7: tcgen05.fence::after_thread_sync before every 4 MMAs   bit 8: a (satisfied) mbarrier wait before every 4 MMAs
//   bit 6: NO MMAs (the issuer just idles for the same time): baseline speed of the other warps
// Also reports how fast the warpgroup loop ran (cycles per 32-column block per warp): does the MMA slow the epilogue?
// Build: make -C tests/cuda umma_interf_probe     Run (on a B200): tests/cuda/umma_interf_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../nonrigid_nerf_b200/csrc/sm100_ptx.cuh"
using namespace nrn;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int kA = 65536, kB = 65536, kS = 65536, kT = 32768;

template <bool PAIR>
__global__ void __launch_bounds__(384, 1) interf_kernel(int mode, int iters, const uint8_t* __restrict__ gsrc, long long* cycles, int* err, float* sink, int* wg_iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, tbar[2], cbar;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < (kA + kB) / 4; i += blockDim.x) {
    uint32_t h = i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 13;
    reinterpret_cast<uint32_t*>(smem)[i] = (h & 0x03ff03ffu) | 0x2c002c00u;
  }
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&tbar[0], 1); mbar_init(&tbar[1], 1); mbar_init(&cbar, 1); done = 0; fence_mbar_init(); }
  fence_proxy_async_smem();
  if (PAIR) cluster_sync_all();
  if (warp == 0) {
    if (PAIR) { tmem_alloc2(&tmem_base_s, 512); tmem_relinquish2(); } else { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  if (warp == 1) {
    if (lane == 0 && rank == 1) {
      mbar_wait(&bar, 0, err, 8);   // multicast commit from the leader
      done = 1;
    }
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = umma_instr_desc(PAIR ? 256 : 128, 256, UMMA_F16, UMMA_F16, UMMA_K_MAJOR, UMMA_K_MAJOR);
      const uint32_t b_lbo = PAIR ? 2048 : 4096;
      const uint64_t adesc = umma_smem_desc(smem_u32(smem), 2048, 128);
      const uint64_t bdesc = umma_smem_desc(smem_u32(smem + kA), b_lbo, 128);
      uint64_t ad[16], bd[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) { ad[j] = umma_desc_advance(adesc, j * 4096); bd[j] = umma_desc_advance(bdesc, (j & 7) * 2 * b_lbo); }
      const long long t0 = clock64();
      if (mode & 64) {
        while (clock64() - t0 < 128ll * iters) {}
        if (PAIR) { umma_commit2(&bar); mbar_wait(&bar, 0, err, 7); }
      } else {
        for (int it = 0; it < iters; it += 16) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if ((j & 3) == 0) {
              if (mode & 256) { mbar_wait(&tbar[1], 1, err, 11); }        // wait on an already-complete phase (a fresh barrier passes parity 1)
              if (mode & 128) tc_fence_after_sync();
            }
            if (PAIR) umma_f16_ss2(tmem_base, ad[j], bd[j], idesc, 1u); else umma_f16_ss(tmem_base, ad[j], bd[j], idesc, 1u);
            if ((mode & 8) && (j & 3) == 3) { if (PAIR) umma_commit2(&cbar); else umma_commit(&cbar); }
          }
        }
        if (PAIR) umma_commit2(&bar); else umma_commit(&bar);
        mbar_wait(&bar, 0, err, 7);
      }
      cycles[blockIdx.x] = clock64() - t0;
      done = 1;
    }
  } else if (warp == 2) {
    if ((mode & 4) && lane == 0) {
      uint32_t ph[2] = {0u, 0u};
      int s = 0;
      bool first[2] = {true, true};
      while (!done) {
        if (!first[s]) { if (!mbar_wait(&tbar[s], ph[s], err, 9)) break; ph[s] ^= 1u; }
        first[s] = false;
        mbar_arrive_expect_tx(&tbar[s], 16384u);
        tma_bulk_g2s(smem + kA + kB + kS + s * 16384, gsrc + ((blockIdx.x * 7 + s) & 63) * 16384, 16384u, &tbar[s]);
        s ^= 1;
      }
      for (int q = 0; q < 2; ++q) if (!first[q]) mbar_wait(&tbar[q], ph[q], err, 10);
    }
  } else if (warp >= 4 && (warp < 8 || (mode & 16))) {
    const int row = ((warp & 3) << 5) | lane;
    uint8_t* dst = smem + kA + kB + (warp >= 8 ? 32768 : 0) + row * 16;
    const uint32_t taddr = tmem_base + ((static_cast<uint32_t>(warp & 3) * 32u) << 16) + 256;
    float acc = 0.f;
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = i * 3 + lane;
    int c = 0;
    while (!done) {
      if (mode & 1) { tmem_ld32(taddr + (c & 7) * 32, v); tmem_ld_wait(); }
      if (mode & 32) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(fmaxf(__uint_as_float(v[i]) + 0.25f, 0.f));
      }
      if (mode & 2) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(dst + ((c * 4 + q) & 15) * 2048) = make_uint4(v[q * 8] ^ v[q * 8 + 1], v[q * 8 + 2] ^ v[q * 8 + 3], v[q * 8 + 4] ^ v[q * 8 + 5], v[q * 8 + 6] ^ v[q * 8 + 7]);
      }
      acc += __uint_as_float(v[c & 31]);
      ++c;
    }
    if (acc == 123.456f) sink[threadIdx.x] = acc;
    if (lane == 0 && blockIdx.x == 0) wg_iters[warp] = c;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == 0) { if (PAIR) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
}

int main() {
  const int iters = 8192, grid = 148;
  long long* dc; int* de; uint8_t* gsrc; float* sink; int* wgi;
  CK(cudaMalloc(&wgi, 64 * 4));
  CK(cudaMalloc(&dc, sizeof(long long) * grid)); CK(cudaMalloc(&de, 4)); CK(cudaMalloc(&gsrc, 64 * 16384)); CK(cudaMalloc(&sink, 4096));
  CK(cudaMemset(gsrc, 0x2c, 64 * 16384));
  const size_t smem = kA + kB + kS + kT + 1024;
  CK(cudaFuncSetAttribute(interf_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(interf_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int modes[] = {0, 8, 128, 256, 128 | 256 | 8, 128 | 256 | 8 | 32 | 16 | 3};
  for (int pair = 0; pair < 2; ++pair)
  for (int mode : modes) {
    CK(cudaMemset(dc, 0, sizeof(long long) * grid)); CK(cudaMemset(de, 0, 4));
    CK(cudaMemset(wgi, 0, 64 * 4));
    for (int rep = 0; rep < 2; ++rep) {
      cudaLaunchConfig_t lc = {};
      lc.gridDim = dim3(grid); lc.blockDim = dim3(384); lc.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = pair ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      lc.attrs = at; lc.numAttrs = 1;
      if (pair) CK(cudaLaunchKernelEx(&lc, interf_kernel<true>, mode, iters, (const uint8_t*)gsrc, dc, de, sink, wgi));
      else CK(cudaLaunchKernelEx(&lc, interf_kernel<false>, mode, iters, (const uint8_t*)gsrc, dc, de, sink, wgi));
      CK(cudaDeviceSynchronize());
    }
    int hw[64]; CK(cudaMemcpy(hw, wgi, 64 * 4, cudaMemcpyDeviceToHost));
    std::vector<long long> c(grid); int err;
    CK(cudaMemcpy(c.data(), dc, sizeof(long long) * grid, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&err, de, 4, cudaMemcpyDeviceToHost));
    long long mx = 0; for (auto v : c) mx = v > mx ? v : mx;
    printf("cta_group::%d mode %3d [%s%s%s%s%s%s%s%s%s]: %7.1f cycles/MMA | warpgroup loop: %7.1f cycles per 32-column block (warp 4), %7.1f (warp 8) err=%d\n", pair + 1, mode, (mode & 128) ? "fence4 " : "", (mode & 256) ? "wait4 " : "",
           (mode & 64) ? "NO-MMA " : "", (mode & 1) ? "ld " : "", (mode & 2) ? "sts " : "", (mode & 4) ? "tma " : "", (mode & 8) ? "commit4 " : "",
           (mode & 16) ? "2wg " : "", (mode & 32) ? "alu " : "", (double)mx / iters, hw[4] ? (double)mx / hw[4] : 0.0, hw[8] ? (double)mx / hw[8] : 0.0, err);
  }
  return 0;
}
