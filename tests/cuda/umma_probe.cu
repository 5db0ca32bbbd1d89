// Standalone known-answer probe for the tcgen05 building blocks in csrc/sm100_ptx.cuh.
// Build: make -C tests/cuda     Run (on a B200): tests/cuda/umma_probe
// One CTA computes D[M x N] = A . B^T from "chunk-major" shared-memory images (see sm100_ptx.cuh)
// for several operand interpretations (K-major / MN-major, fp16 / bf16, many N and K), and the
// host checks every element against an fp32 computation on the same rounded inputs.
// It also times a long back-to-back MMA stream (cycles per tcgen05.mma) and a TMEM drain.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>
#include "../../nonrigid_nerf_b200/csrc/sm100_ptx.cuh"

using namespace nrn;

struct ProbeCfg {
  int M, N, K;                 // logical MMA problem (M = 128)
  uint32_t a_bytes, b_bytes;   // image sizes
  uint32_t a_lbo, a_sbo, a_kstep;
  uint32_t b_lbo, b_sbo, b_kstep;
  uint32_t idesc;
  int reps;                    // re-issue the whole K loop this many times (timing)
};

#define CK(x)                                                                     \
  do {                                                                            \
    cudaError_t e_ = (x);                                                         \
    if (e_ != cudaSuccess) {                                                      \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                    \
    }                                                                             \
  } while (0)

__global__ void __launch_bounds__(128, 1)
probe_kernel(ProbeCfg cfg, const uint8_t* __restrict__ a_img, const uint8_t* __restrict__ b_img,
             float* __restrict__ d_out, long long* __restrict__ cycles, int* __restrict__ err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((cfg.a_bytes + 1023u) & ~1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_load, cfg.a_bytes + cfg.b_bytes);
    // bulk copies are limited in size per instruction; issue in 16 KB pieces
    for (uint32_t off = 0; off < cfg.a_bytes; off += 16384u) {
      uint32_t n = cfg.a_bytes - off < 16384u ? cfg.a_bytes - off : 16384u;
      tma_bulk_g2s(sa + off, a_img + off, n, &bar_load);
    }
    for (uint32_t off = 0; off < cfg.b_bytes; off += 16384u) {
      uint32_t n = cfg.b_bytes - off < 16384u ? cfg.b_bytes - off : 16384u;
      tma_bulk_g2s(sb + off, b_img + off, n, &bar_load);
    }
  }
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    bool ok = mbar_wait(&bar_load, 0, err, 11);
    tc_fence_after_sync();
    if (ok && lane == 0) {
      const uint64_t adesc0 = umma_smem_desc(smem_u32(sa), cfg.a_lbo, cfg.a_sbo);
      const uint64_t bdesc0 = umma_smem_desc(smem_u32(sb), cfg.b_lbo, cfg.b_sbo);
      t0 = clock64();
      for (int r = 0; r < cfg.reps; ++r) {
        for (int k = 0; k < cfg.K / 16; ++k) {
          umma_f16_ss(tmem_base, umma_desc_advance(adesc0, k * cfg.a_kstep),
                      umma_desc_advance(bdesc0, k * cfg.b_kstep), cfg.idesc, (r | k) ? 1u : 0u);
        }
      }
      umma_commit(&bar_mma);
    }
    __syncwarp();
  }
  bool ok2 = mbar_wait(&bar_mma, 0, err, 12);
  tc_fence_after_sync();
  t1 = clock64();
  if (warp == 1 && lane == 0) cycles[0] = t1 - t0;
  long long t2 = clock64();
  if (ok2) {
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < cfg.N; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_addr(tmem_base, warp * 32, c0), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) d_out[row * cfg.N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  long long t3 = clock64();
  if (threadIdx.x == 0) cycles[1] = t3 - t2;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------
static uint16_t f2h(float f) { __half h = __float2half_rn(f); uint16_t u; memcpy(&u, &h, 2); return u; }
static float h2f(uint16_t u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }
static uint16_t f2b(float f) { __nv_bfloat16 h = __float2bfloat16_rn(f); uint16_t u; memcpy(&u, &h, 2); return u; }
static float b2f(uint16_t u) { __nv_bfloat16 h; memcpy(&h, &u, 2); return __bfloat162float(h); }

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

// chunk-major image of a logical [R][C] matrix: img[c/8][r][c%8]
static std::vector<uint16_t> image(const std::vector<uint16_t>& m, int R, int C) {
  std::vector<uint16_t> img((size_t)R * C);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c) img[(size_t)(c / 8) * R * 8 + (size_t)r * 8 + (c % 8)] = m[(size_t)r * C + c];
  return img;
}

struct Case {
  std::string name;
  int N, K;
  bool bf16;
  bool mn_major;  // operands stored [K rows][MN cols] and consumed MN-major (WGRAD shape)
  bool swap_lbo_sbo;
  int reps;
};

static int run_case(const Case& cs) {
  const int M = 128, N = cs.N, K = cs.K;
  std::vector<uint16_t> A, B;  // logical storage
  std::vector<float> Af, Bf;   // rounded values, logical [M][K], [N][K]
  Af.resize((size_t)M * K);
  Bf.resize((size_t)N * K);
  auto enc = [&](float f) { return cs.bf16 ? f2b(f) : f2h(f); };
  auto dec = [&](uint16_t u) { return cs.bf16 ? b2f(u) : h2f(u); };
  ProbeCfg cfg{};
  cfg.M = M; cfg.N = N; cfg.K = K; cfg.reps = cs.reps;
  std::vector<uint16_t> a_img, b_img;
  if (!cs.mn_major) {
    A.resize((size_t)M * K); B.resize((size_t)N * K);
    for (size_t i = 0; i < A.size(); ++i) { A[i] = enc(frand()); Af[i] = dec(A[i]); }
    for (size_t i = 0; i < B.size(); ++i) { B[i] = enc(frand()); Bf[i] = dec(B[i]); }
    a_img = image(A, M, K);
    b_img = image(B, N, K);
    cfg.a_lbo = M * 16; cfg.a_sbo = 128; cfg.a_kstep = 2 * M * 16;
    cfg.b_lbo = N * 16; cfg.b_sbo = 128; cfg.b_kstep = 2 * N * 16;
    cfg.idesc = umma_instr_desc(M, N, cs.bf16, cs.bf16, UMMA_K_MAJOR, UMMA_K_MAJOR);
  } else {
    // storage: X1[K points][M feats], X2[K points][N feats]; D[m][n] = sum_p X1[p][m] X2[p][n]
    A.resize((size_t)K * M); B.resize((size_t)K * N);
    for (int p = 0; p < K; ++p)
      for (int m = 0; m < M; ++m) { uint16_t u = enc(frand()); A[(size_t)p * M + m] = u; Af[(size_t)m * K + p] = dec(u); }
    for (int p = 0; p < K; ++p)
      for (int n = 0; n < N; ++n) { uint16_t u = enc(frand()); B[(size_t)p * N + n] = u; Bf[(size_t)n * K + p] = dec(u); }
    a_img = image(A, K, M);  // rows = points, chunks over features
    b_img = image(B, K, N);
    cfg.a_sbo = K * 16; cfg.a_lbo = 128; cfg.a_kstep = 256;
    cfg.b_sbo = K * 16; cfg.b_lbo = 128; cfg.b_kstep = 256;
    cfg.idesc = umma_instr_desc(M, N, cs.bf16, cs.bf16, UMMA_MN_MAJOR, UMMA_MN_MAJOR);
  }
  if (cs.swap_lbo_sbo) { std::swap(cfg.a_lbo, cfg.a_sbo); std::swap(cfg.b_lbo, cfg.b_sbo); }
  cfg.a_bytes = (uint32_t)a_img.size() * 2;
  cfg.b_bytes = (uint32_t)b_img.size() * 2;

  uint8_t *da, *db; float* dd; long long* dc; int* de;
  CK(cudaMalloc(&da, cfg.a_bytes)); CK(cudaMalloc(&db, cfg.b_bytes));
  CK(cudaMalloc(&dd, sizeof(float) * M * N)); CK(cudaMalloc(&dc, 16)); CK(cudaMalloc(&de, 4));
  CK(cudaMemcpy(da, a_img.data(), cfg.a_bytes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b_img.data(), cfg.b_bytes, cudaMemcpyHostToDevice));
  CK(cudaMemset(dd, 0xff, sizeof(float) * M * N)); CK(cudaMemset(dc, 0, 16)); CK(cudaMemset(de, 0, 4));
  size_t smem = ((cfg.a_bytes + 1023u) & ~1023u) + cfg.b_bytes + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  probe_kernel<<<1, 128, smem>>>(cfg, da, db, dd, dc, de);
  CK(cudaEventRecord(e1));
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[%s] KERNEL ERROR: %s\n", cs.name.c_str(), cudaGetErrorString(e)); return 2; }
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<float> D((size_t)M * N); long long cyc[2]; int err;
  CK(cudaMemcpy(D.data(), dd, sizeof(float) * M * N, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(cyc, dc, 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&err, de, 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0; int bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += (double)Af[(size_t)m * K + k] * Bf[(size_t)n * K + k];
      acc *= cs.reps;
      double d = fabs(acc - D[(size_t)m * N + n]);
      if (!(d <= 1e-2 * cs.reps + 1e-3 * fabs(acc))) ++bad;
      if (d > maxerr || d != d) maxerr = d;
      if (fabs(acc) > maxref) maxref = fabs(acc);
    }
  int nmma = cs.reps * (K / 16);
  printf("[%-34s] N=%3d K=%3d %s %s err=%d bad=%6d/%d maxerr=%.3e (ref max %.2f) mma_cycles=%lld (%.1f/mma) drain_cycles=%lld  %.3f ms -> %s\n",
         cs.name.c_str(), N, K, cs.bf16 ? "bf16" : "fp16", cs.mn_major ? "MN-major" : "K-major ", err, bad, M * N,
         maxerr, maxref, cyc[0], (double)cyc[0] / nmma, cyc[1], ms, (bad == 0 && err == 0) ? "PASS" : "FAIL");
  cudaFree(da); cudaFree(db); cudaFree(dd); cudaFree(dc); cudaFree(de);
  return (bad == 0 && err == 0) ? 0 : 1;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device: %s sm_%d%d SMs=%d smem/block optin=%zu\n", p.name, p.major, p.minor, p.multiProcessorCount,
         p.sharedMemPerBlockOptin);
  srand(1234);
  std::vector<Case> cases = {
      {"kmajor_fp16_N256_K64", 256, 64, false, false, false, 1},
      {"kmajor_fp16_N256_K256", 256, 256, false, false, false, 1},
      {"kmajor_fp16_N128_K320", 128, 320, false, false, false, 1},
      {"kmajor_bf16_N256_K256", 256, 256, true, false, false, 1},
      {"kmajor_fp16_N96_K48", 96, 48, false, false, false, 1},
      {"kmajor_fp16_N96_K96", 96, 96, false, false, false, 1},
      {"kmajor_fp16_N80_K96", 80, 96, false, false, false, 1},
      {"kmajor_fp16_N64_K64", 64, 64, false, false, false, 1},
      {"kmajor_fp16_N16_K64", 16, 64, false, false, false, 1},
      {"kmajor_fp16_N16_K256", 16, 256, false, false, false, 1},
      {"kmajor_fp16_N64_K256", 64, 256, false, false, false, 1},
      {"kmajor_fp16_N128_K256", 128, 256, false, false, false, 1},
      {"mnmajor_bf16_N256_K128", 256, 128, true, true, false, 1},
      {"mnmajor_bf16_N64_K128", 64, 128, true, true, false, 1},
      {"mnmajor_fp16_N128_K64", 128, 64, false, true, false, 1},
      {"timing_kmajor_fp16_N256_K256_x64", 256, 256, false, false, false, 64},
      {"timing_kmajor_fp16_N128_K256_x64", 128, 256, false, false, false, 64},
      {"timing_kmajor_fp16_N64_K64_x256", 64, 64, false, false, false, 256},
  };
  int fails = 0;
  for (auto& c : cases) {
    int r = run_case(c);
    if (r == 2) { printf("aborting after kernel error\n"); return 2; }
    if (r && c.name.rfind("SWAPPED", 0) != 0) ++fails;
  }
  printf("probe done: %d failing (non-SWAPPED) cases\n", fails);
  return fails ? 1 : 0;
}
