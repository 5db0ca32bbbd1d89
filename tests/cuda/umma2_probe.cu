// Known-answer probe for the 2-CTA (cta_group::2) tcgen05 path planned for round 2:
// a CTA pair computes D[256 x N] = A[256 x K] . B[N x K]^T with M = 256 spread over both CTAs' TMEM,
// each CTA holding its own 128 rows of A and HALF of B (N/2 rows) in shared memory.
// Checks: cluster launch, tcgen05.alloc/dealloc.cta_group::2, the MMA itself, multicast commit to both
// CTAs' mbarriers, remote mbarrier arrive (mapa), per-CTA accumulator drain.
// Build: make -C tests/cuda umma2_probe     Run (on a B200): tests/cuda/umma2_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>
#include "../../nonrigid_nerf_b200/csrc/sm100_ptx.cuh"

using namespace nrn;

#define CK(x)                                                                     \
  do {                                                                            \
    cudaError_t e_ = (x);                                                         \
    if (e_ != cudaSuccess) {                                                      \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                    \
    }                                                                             \
  } while (0)

struct Cfg {
  int N, K, remote_bar;   // remote_bar: the peer's bulk copies complete_tx on the LEADER's mbarrier (no relay needed)
  uint32_t a_bytes, b_bytes;  // per CTA
  uint32_t idesc;
};

__device__ __forceinline__ void remote_arrive(uint64_t* local_bar, uint32_t target_cta) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local_bar)), "r"(target_cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe2_kernel(Cfg cfg, const uint8_t* __restrict__ a_img, const uint8_t* __restrict__ b_img, float* __restrict__ d_out,
              int* __restrict__ err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_load, bar_peer, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const uint32_t rank = cluster_ctarank();
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((cfg.a_bytes + 1023u) & ~1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_peer, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  cluster_sync_all();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;

  if (threadIdx.x == 0) {
    uint32_t bar_addr = smem_u32(&bar_load);
    if (cfg.remote_bar) {
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(bar_addr) : "r"(smem_u32(&bar_load)), "r"(0u));
      if (rank == 0) mbar_arrive_expect_tx(&bar_load, 2u * (cfg.a_bytes + cfg.b_bytes));
    } else {
      mbar_arrive_expect_tx(&bar_load, cfg.a_bytes + cfg.b_bytes);
    }
    const uint8_t* ga = a_img + static_cast<size_t>(rank) * cfg.a_bytes;
    const uint8_t* gb = b_img + static_cast<size_t>(rank) * cfg.b_bytes;
    auto copy = [&](uint8_t* dst, const uint8_t* src, uint32_t n) {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
                   "l"(src), "r"(n), "r"(bar_addr)
                   : "memory");
    };
    for (uint32_t off = 0; off < cfg.a_bytes; off += 16384u) copy(sa + off, ga + off, cfg.a_bytes - off < 16384u ? cfg.a_bytes - off : 16384u);
    for (uint32_t off = 0; off < cfg.b_bytes; off += 16384u) copy(sb + off, gb + off, cfg.b_bytes - off < 16384u ? cfg.b_bytes - off : 16384u);
  }
  if (warp == 1 && lane == 0) {
    bool ok = (cfg.remote_bar && rank == 1) ? true : mbar_wait(&bar_load, 0, err, 11);
    if (rank == 1) {
      if (ok && !cfg.remote_bar) remote_arrive(&bar_peer, 0);   // tell the leader that this CTA's operands have landed
    } else {
      ok = ok && (cfg.remote_bar || mbar_wait(&bar_peer, 0, err, 12));
      tc_fence_after_sync();
      if (ok) {
        const uint32_t half_rows = cfg.N / 2;
        const uint64_t adesc0 = umma_smem_desc(smem_u32(sa), 128 * 16, 128);
        const uint64_t bdesc0 = umma_smem_desc(smem_u32(sb), half_rows * 16, 128);
        for (int k = 0; k < cfg.K / 16; ++k) {
          const uint64_t ad = umma_desc_advance(adesc0, k * 2 * 128 * 16), bd = umma_desc_advance(bdesc0, k * 2 * half_rows * 16);
          const uint32_t acc = k ? 1u : 0u;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_base),
              "l"(ad), "l"(bd), "r"(cfg.idesc), "r"(acc)
              : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                         smem_u32(&bar_mma)),
                     "h"(static_cast<uint16_t>(3))
                     : "memory");
      }
    }
  }
  const bool ok2 = mbar_wait(&bar_mma, 0, err, 13 + rank);
  tc_fence_after_sync();
  if (ok2) {
    const int row = rank * 128 + warp * 32 + lane;
    for (int c0 = 0; c0 < cfg.N; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_addr(tmem_base, warp * 32, c0), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) d_out[row * cfg.N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

static uint16_t f2h(float f) { __half h = __float2half_rn(f); uint16_t u; memcpy(&u, &h, 2); return u; }
static float h2f(uint16_t u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }
static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }
// chunk-major image of rows [r0, r0+R) of a logical [*][C] matrix
static void image(const std::vector<uint16_t>& m, int r0, int R, int C, std::vector<uint16_t>& out) {
  const size_t base = out.size();
  out.resize(base + (size_t)R * C);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c) out[base + (size_t)(c / 8) * R * 8 + (size_t)r * 8 + (c % 8)] = m[(size_t)(r0 + r) * C + c];
}

static int run_case(int N, int K, int remote_bar) {
  const int M = 256;
  std::vector<uint16_t> A((size_t)M * K), B((size_t)N * K);
  for (auto& v : A) v = f2h(frand());
  for (auto& v : B) v = f2h(frand());
  std::vector<uint16_t> a_img, b_img;
  image(A, 0, 128, K, a_img); image(A, 128, 128, K, a_img);
  image(B, 0, N / 2, K, b_img); image(B, N / 2, N / 2, K, b_img);
  Cfg cfg{};
  cfg.N = N; cfg.K = K; cfg.remote_bar = remote_bar;
  cfg.a_bytes = 128 * K * 2; cfg.b_bytes = (N / 2) * K * 2;
  cfg.idesc = umma_instr_desc(256, N, UMMA_F16, UMMA_F16, UMMA_K_MAJOR, UMMA_K_MAJOR);
  uint8_t *da, *db; float* dd; int* de;
  CK(cudaMalloc(&da, a_img.size() * 2)); CK(cudaMalloc(&db, b_img.size() * 2));
  CK(cudaMalloc(&dd, sizeof(float) * M * N)); CK(cudaMalloc(&de, 4));
  CK(cudaMemcpy(da, a_img.data(), a_img.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b_img.data(), b_img.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dd, 0xff, sizeof(float) * M * N)); CK(cudaMemset(de, 0, 4));
  const size_t smem = ((cfg.a_bytes + 1023u) & ~1023u) + cfg.b_bytes + 1024;
  CK(cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe2_kernel<<<2, 128, smem>>>(cfg, da, db, dd, de);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[N=%d K=%d] KERNEL ERROR: %s\n", N, K, cudaGetErrorString(e)); return 2; }
  std::vector<float> D((size_t)M * N); int err;
  CK(cudaMemcpy(D.data(), dd, sizeof(float) * M * N, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&err, de, 4, cudaMemcpyDeviceToHost));
  int bad = 0, bad_top = 0; double maxerr = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += (double)h2f(A[(size_t)m * K + k]) * h2f(B[(size_t)n * K + k]);
      const double d = fabs(acc - D[(size_t)m * N + n]);
      if (!(d <= 1e-2 + 1e-3 * fabs(acc))) { ++bad; if (m < 128) ++bad_top; }
      if (d > maxerr || d != d) maxerr = d;
    }
  printf("[cta_group::2 M=256 N=%3d K=%3d remote_bar=%d] err=%d bad=%d/%d (rows<128: %d) maxerr=%.3e -> %s\n", N, K, remote_bar, err, bad, M * N, bad_top, maxerr,
         (bad == 0 && err == 0) ? "PASS" : "FAIL");
  cudaFree(da); cudaFree(db); cudaFree(dd); cudaFree(de);
  return (bad == 0 && err == 0) ? 0 : 1;
}

int main() {
  srand(4321);
  int fails = 0;
  const int cases[][2] = {{256, 64}, {256, 256}, {96, 96}, {80, 96}, {64, 64}, {16, 256}, {48, 96}};
  for (int rb = 0; rb < 2; ++rb)
    for (auto& c : cases) {
      const int r = run_case(c[0], c[1], rb);
      if (r == 2) { printf("aborting after kernel error\n"); return 2; }
      fails += r;
    }
  printf("2-CTA probe done: %d failing cases\n", fails);
  return fails ? 1 : 0;
}
