// Issue-rate probe: how many SM cycles does one tcgen05.mma (M=128|256, N=256, K=16, fp16, shared-memory
// operands in the no-swizzle chunk-major layout of the field kernels) take when issued back to back?
// All SMs run it at once (realistic clocks/power).  Variants: cta_group::1 vs cta_group::2 (CTA pair, each CTA
// holding half of B), one or two accumulators, commit granularity.
// Build: make -C tests/cuda umma_rate_probe     Run (on a B200): tests/cuda/umma_rate_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../nonrigid_nerf_b200/csrc/sm100_ptx.cuh"
using namespace nrn;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

struct Cfg { int pair, n, iters, per_commit, two_acc, a_bytes, b_bytes; };

template <bool PAIR>
__global__ void __launch_bounds__(128, 1) rate_kernel(Cfg cfg, long long* cycles, int* err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // pseudo-random small fp16 operands
  for (uint32_t i = threadIdx.x; i < (uint32_t)(cfg.a_bytes + cfg.b_bytes) / 4; i += blockDim.x) {
    uint32_t h = i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 13;
    reinterpret_cast<uint32_t*>(smem)[i] = (h & 0x03ff03ffu) | 0x2c002c00u;   // halves in [1/16, 1/8)
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (PAIR) cluster_sync_all();
  if (warp == 0) {
    if (PAIR) { tmem_alloc2(&tmem_base_s, 512); tmem_relinquish2(); } else { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const int C = PAIR ? 2 : 1;
  if (warp == 1 && lane == 0 && rank == 0) {
    const uint32_t idesc = umma_instr_desc(128 * C, cfg.n, UMMA_F16, UMMA_F16, UMMA_K_MAJOR, UMMA_K_MAJOR);
    const uint32_t b_lbo = (cfg.n / C) * 16;
    const uint32_t a_k = cfg.a_bytes / (2 * 2048), b_k = cfg.b_bytes / (2 * b_lbo);   // K-steps available in each buffer
    const uint64_t adesc = umma_smem_desc(smem_u32(smem), 2048, 128);
    const uint64_t bdesc = umma_smem_desc(smem_u32(smem + cfg.a_bytes), b_lbo, 128);
    uint32_t ph = 0;
    const long long t0 = clock64();
    // 16 MMAs per iteration with precomputed descriptors: the issue loop itself must not be the limiter
    uint64_t ad[16], bd[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      ad[j] = umma_desc_advance(adesc, (j % a_k) * 2 * 2048);
      bd[j] = umma_desc_advance(bdesc, (j % b_k) * 2 * b_lbo);
    }
    for (int it = 0; it < cfg.iters; it += 16) {
      const uint32_t d = tmem_base + ((cfg.two_acc && ((it >> 4) & 1)) ? 256u : 0u);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (PAIR) umma_f16_ss2(d, ad[j], bd[j], idesc, 1u); else umma_f16_ss(d, ad[j], bd[j], idesc, 1u);
      }
    }
    if (PAIR) umma_commit2(&bar); else umma_commit(&bar);
    if (!mbar_wait(&bar, ph, err, 7)) {}
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  } else if (PAIR && rank == 1 && warp == 1 && lane == 0) {
    mbar_wait(&bar, 0, err, 8);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == 0) { if (PAIR) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
}

static void run(int pair, int n, int iters, int two_acc, int b_kb, int grid) {
  Cfg cfg{pair, n, iters, 16, two_acc, 65536, b_kb * 1024};
  long long* dc; int* de;
  CK(cudaMalloc(&dc, sizeof(long long) * grid)); CK(cudaMalloc(&de, 4));
  CK(cudaMemset(dc, 0, sizeof(long long) * grid)); CK(cudaMemset(de, 0, 4));
  const size_t smem = cfg.a_bytes + cfg.b_bytes + 1024;
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(grid); lc.blockDim = dim3(128); lc.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = pair ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = 1;
  float ms = 0;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    if (pair) { CK(cudaFuncSetAttribute(rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); CK(cudaLaunchKernelEx(&lc, rate_kernel<true>, cfg, dc, de)); }
    else { CK(cudaFuncSetAttribute(rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); CK(cudaLaunchKernelEx(&lc, rate_kernel<false>, cfg, dc, de)); }
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, e0, e1);
  }
  std::vector<long long> c(grid); int err;
  CK(cudaMemcpy(c.data(), dc, sizeof(long long) * grid, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&err, de, 4, cudaMemcpyDeviceToHost));
  long long mx = 0; for (auto v : c) mx = v > mx ? v : mx;
  const double per = (double)mx / iters;
  const double flop = 2.0 * 128 * n * 16 * (double)iters * grid;   // every CTA contributes 128 rows
  printf("cta_group::%d N=%3d two_acc=%d B=%2dKB grid=%3d: %7.1f cycles/MMA  (%.3f ms, %.0f TFLOP/s) err=%d\n", pair ? 2 : 1, n, two_acc, b_kb, grid, per, ms,
         flop / ms / 1e9, err);
  cudaFree(dc); cudaFree(de);
}

int main() {
  const int iters = 16384;
  run(0, 256, iters, 0, 64, 148);
  run(0, 256, iters, 1, 64, 148);
  run(0, 256, iters, 1, 64, 37);
  run(1, 256, iters, 0, 32, 148);
  run(1, 256, iters, 1, 32, 148);
  run(0, 128, iters, 1, 32, 148);
  run(0, 96, iters, 1, 32, 148);
  run(0, 64, iters, 1, 32, 148);
  run(0, 16, iters, 1, 32, 148);
  run(1, 128, iters, 1, 32, 148);
  run(1, 64, iters, 1, 32, 148);
  run(1, 16, iters, 1, 32, 148);
  return 0;
}
