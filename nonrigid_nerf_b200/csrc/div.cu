// Divergence regulariser of the ray-bending offset field, forward and backward, without autograd.
//
// Reference: training_wrapper_class.forward (train.py:245-286) -> compute_divergence_loss /
// divergence_approx (run_nerf_helpers.py:22-116): for every COARSE sample point the Hutchinson
// estimate  d = e^T J e,  J = d(masked offsets)/d(xyz),  e ~ N(0, I), is formed with
// autograd.grad(create_graph=True) and the loss  mean_s( w d^2 )  is differentiated again w.r.t. the
// bender weights (a double backward).  Here the same quantities are computed in closed form:
//
//   masked = r * off,   off = W4 relu(W3 relu(W2 relu(W1 relu(W0 [x, l] + b0) ..))),  r = (tanh(c3)+1)/2
//   J e    = r * (W4 D4 W3 D3 W2 D2 W1 D1 W0[:, :3] e)  +  off * r'(c3) * (R2 E2 R1 E1 R0 e)
//          = r * tau_off + off * tau_r                       (D_i, E_i: ReLU masks of the primal pass)
//   d      = r * alpha + beta * tau_r,   alpha = e . tau_off,  beta = e . off,  tau_r = 2 r (1 - r) tau_c
//
// Forward kernel: the tangent chain t_i = D_i (W_{i-1} t_{i-1}) (no bias), per point, fp32 SIMT with
// the weights in shared memory; masks come from the coarse pass's activation stash; the tangent
// activations are written as fp16 chunk-major images (same layout as the bender part of the stash).
// Backward kernel: given G = dL/dd per point, the adjoint chain of the tangent pass -> fp16 adjoint
// images; the weight gradients  sum_p abar_i t_{i-1}^T  are then formed by the WGRAD kernel (compact
// mode) exactly like the primal bender layers.  ReLU masks are piecewise constant, so no gradient
// flows into them (autograd's double backward gives the same zeros).  The dependence on the PRIMAL
// quantities r and off is returned as gradients w.r.t. the coarse pass's `rigidity_mask` and
// `unmasked_offsets` outputs and continues through the ordinary field backward:
//   dL/d off = G tau_r e,     dL/d r = G (alpha + 2 beta tau_c (1 - 2 r))
#include <cuda_fp16.h>
#include "nrn_common.cuh"
#include "div.cuh"

namespace nrn {

namespace {

constexpr int kDivThreads = 128;   // one tile per block
// shared-memory weight table (fp32)
constexpr int kW0x = 0;                   // [64][4]  (3 used)
constexpr int kW1 = kW0x + 64 * 4;        // [64][64]
constexpr int kW2 = kW1 + 4096;
constexpr int kW3 = kW2 + 4096;
constexpr int kW4 = kW3 + 4096;           // [3][64]
constexpr int kR0 = kW4 + 192;            // [32][4]
constexpr int kR1 = kR0 + 128;            // [32][32]
constexpr int kR2 = kR1 + 1024;           // [32]
constexpr int kWTotal = kR2 + 32;         // 13,920 floats
constexpr int kActFloats = 64 * kDivThreads;   // per-thread activation column [k][thread]

// compact tile layouts (bytes): tangent stash and adjoint stash
constexpr int kTE = 0, kT1 = 6 * kChunkBytes, kT2 = 18 * kChunkBytes, kT3 = 30 * kChunkBytes, kT4 = 38 * kChunkBytes;
constexpr int kA4 = 0, kA3 = 2 * kChunkBytes, kA2 = 10 * kChunkBytes, kA1 = 20 * kChunkBytes, kA0 = 32 * kChunkBytes;

__device__ __forceinline__ void load_weights(float* sw, const DivParams& p) {
  for (int i = threadIdx.x; i < 64 * 4; i += blockDim.x) sw[kW0x + i] = (i & 3) < 3 ? p.net_w[0][(i >> 2) * 35 + (i & 3)] : 0.f;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) {
    sw[kW1 + i] = p.net_w[1][i];
    sw[kW2 + i] = p.net_w[2][i];
    sw[kW3 + i] = p.net_w[3][i];
  }
  for (int i = threadIdx.x; i < 192; i += blockDim.x) sw[kW4 + i] = p.net_w[4][i];
  for (int i = threadIdx.x; i < 32 * 4; i += blockDim.x) sw[kR0 + i] = (i & 3) < 3 ? p.rig_w[0][(i >> 2) * 3 + (i & 3)] : 0.f;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sw[kR1 + i] = p.rig_w[1][i];
  for (int i = threadIdx.x; i < 32; i += blockDim.x) sw[kR2 + i] = p.rig_w[2][i];
}

// bit j = (stashed fp16 activation j > 0), over `nchunks` consecutive chunks of this thread's row
__device__ __forceinline__ unsigned long long read_mask(const uint8_t* row, int nchunks) {
  unsigned long long m = 0ull;
  for (int c = 0; c < nchunks; ++c) {
    const uint4 w = __ldg(reinterpret_cast<const uint4*>(row + c * kChunkBytes));
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (ww[q] & 0x7fffu) m |= 1ull << (c * 8 + 2 * q);
      if (ww[q] & 0x7fff0000u) m |= 1ull << (c * 8 + 2 * q + 1);
    }
  }
  return m;
}

__device__ __forceinline__ uint32_t pk(float a, float b) {
  __half2 h = __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
  return *reinterpret_cast<uint32_t*>(&h);
}

// out[j] = sum_k W[j][k] in[k]  (W row-major [OUT][IN] in smem, `in` in registers), out -> act column
template <int IN>
__device__ __forceinline__ void matvec(const float* W, const float (&in)[IN], float* act, int OUT) {
#pragma unroll 1
  for (int j0 = 0; j0 < OUT; j0 += 4) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int k = 0; k < IN; k += 4) {
      const float4 w0 = *reinterpret_cast<const float4*>(W + (j0 + 0) * IN + k);
      const float4 w1 = *reinterpret_cast<const float4*>(W + (j0 + 1) * IN + k);
      const float4 w2 = *reinterpret_cast<const float4*>(W + (j0 + 2) * IN + k);
      const float4 w3 = *reinterpret_cast<const float4*>(W + (j0 + 3) * IN + k);
      a0 += w0.x * in[k] + w0.y * in[k + 1] + w0.z * in[k + 2] + w0.w * in[k + 3];
      a1 += w1.x * in[k] + w1.y * in[k + 1] + w1.z * in[k + 2] + w1.w * in[k + 3];
      a2 += w2.x * in[k] + w2.y * in[k + 1] + w2.z * in[k + 2] + w2.w * in[k + 3];
      a3 += w3.x * in[k] + w3.y * in[k + 1] + w3.z * in[k + 2] + w3.w * in[k + 3];
    }
    act[(j0 + 0) * kDivThreads] = a0; act[(j0 + 1) * kDivThreads] = a1;
    act[(j0 + 2) * kDivThreads] = a2; act[(j0 + 3) * kDivThreads] = a3;
  }
}
// out[k] = sum_j W[j][k] in[j]  (transposed product; W row-major [NJ][NK] in smem)
template <int NJ>
__device__ __forceinline__ void matvec_t(const float* W, const float (&in)[NJ], float* act, int NK) {
#pragma unroll 1
  for (int k0 = 0; k0 < NK; k0 += 4) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float4 w = *reinterpret_cast<const float4*>(W + j * NK + k0);
      a0 += w.x * in[j]; a1 += w.y * in[j]; a2 += w.z * in[j]; a3 += w.w * in[j];
    }
    act[(k0 + 0) * kDivThreads] = a0; act[(k0 + 1) * kDivThreads] = a1;
    act[(k0 + 2) * kDivThreads] = a2; act[(k0 + 3) * kDivThreads] = a3;
  }
}
// masked copy act column -> registers, and fp16 image chunks [chunk0, chunk0 + N/8) of this row
template <int N>
__device__ __forceinline__ void take(const float* act, unsigned long long mask, float scale, float (&out)[N], uint8_t* img_row) {
#pragma unroll
  for (int j = 0; j < N; ++j) out[j] = ((mask >> j) & 1ull) ? act[j * kDivThreads] * scale : 0.f;
#pragma unroll
  for (int c = 0; c < N / 8; ++c) {
    const uint4 v = make_uint4(pk(out[c * 8], out[c * 8 + 1]), pk(out[c * 8 + 2], out[c * 8 + 3]), pk(out[c * 8 + 4], out[c * 8 + 5]),
                               pk(out[c * 8 + 6], out[c * 8 + 7]));
    *reinterpret_cast<uint4*>(img_row + c * kChunkBytes) = v;
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDivThreads) div_fwd_kernel(const DivParams p) {
  extern __shared__ __align__(16) float sm[];
  float* sw = sm;
  float* act = sm + kWTotal + threadIdx.x;   // column of this thread: act[k * 128]
  load_weights(sw, p);
  __syncthreads();
  const long long tile = blockIdx.x;
  const int row = threadIdx.x;
  const long long pt = tile * kTileM + row;
  const bool valid = pt < p.P;
  const uint8_t* st = p.stash + tile * kStashTileBytes + row * 16;
  uint8_t* tn = p.tan + tile * kTanTileBytes + row * 16;

  float e[4] = {0.f, 0.f, 0.f, 0.f}, off[3] = {0.f, 0.f, 0.f};
  float r = 0.f, w = 0.f;
  if (valid) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { e[d] = p.e[pt * 3 + d]; off[d] = p.unmasked[pt * 3 + d]; }
    r = p.rigidity[pt];
    w = p.w[pt];
    if (p.w_is_alpha) w = 1.0f - expf(-fmaxf(w, 0.f));
  }
  // e image (bender-input layout: hi columns 0-2, lo columns 3-5)
  {
    float hi[3], lo[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) { hi[d] = __half2float(__float2half_rn(e[d])); lo[d] = e[d] - hi[d]; }
    *reinterpret_cast<uint4*>(tn + kTE) = make_uint4(pk(hi[0], hi[1]), pk(hi[2], lo[0]), pk(lo[1], lo[2]), 0u);
    const uint4 zz = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int c = 1; c < 6; ++c) *reinterpret_cast<uint4*>(tn + kTE + c * kChunkBytes) = zz;
  }
  const unsigned long long m1 = read_mask(st + kStHb1, 8), e1 = read_mask(st + kStHb1 + 8 * kChunkBytes, 4);
  const unsigned long long m2 = read_mask(st + kStHb2, 8), e2 = read_mask(st + kStHb2 + 8 * kChunkBytes, 4);
  const unsigned long long m3 = read_mask(st + kStHb3, 8), m4 = read_mask(st + kStHb4, 8);

  float t[64], s[32];
  // layer 0: t1 = D1 (W0[:, :3] e), s1 = E1 (R0 e)
  matvec<4>(sw + kW0x, e, act, 64);
  take<64>(act, m1, 1.0f, t, tn + kT1);
  matvec<4>(sw + kR0, e, act, 32);
  take<32>(act, e1, 1.0f, s, tn + kT1 + 8 * kChunkBytes);
  // layer 1
  matvec<64>(sw + kW1, t, act, 64);
  take<64>(act, m2, 1.0f, t, tn + kT2);
  matvec<32>(sw + kR1, s, act, 32);
  take<32>(act, e2, 1.0f, s, tn + kT2 + 8 * kChunkBytes);
  // layers 2, 3
  matvec<64>(sw + kW2, t, act, 64);
  take<64>(act, m3, 1.0f, t, tn + kT3);
  matvec<64>(sw + kW3, t, act, 64);
  take<64>(act, m4, 1.0f, t, tn + kT4);
  // outputs
  float tau_off[3] = {0.f, 0.f, 0.f}, tau_c = 0.f;
#pragma unroll
  for (int k = 0; k < 64; ++k) {
    tau_off[0] += sw[kW4 + k] * t[k]; tau_off[1] += sw[kW4 + 64 + k] * t[k]; tau_off[2] += sw[kW4 + 128 + k] * t[k];
  }
#pragma unroll
  for (int k = 0; k < 32; ++k) tau_c += sw[kR2 + k] * s[k];
  const float alpha = e[0] * tau_off[0] + e[1] * tau_off[1] + e[2] * tau_off[2];
  const float beta = e[0] * off[0] + e[1] * off[1] + e[2] * off[2];
  const float tau_r = 2.0f * r * (1.0f - r) * tau_c;
  const float d = r * alpha + beta * tau_r;
  if (valid) {
    p.d[pt] = d; p.adot[pt] = alpha; p.beta[pt] = beta; p.tauc[pt] = tau_c;
    atomicAdd(p.loss + pt / p.S, w * d * d / static_cast<float>(p.S));   // mean over the ray's samples of w |d|^2
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDivThreads) div_bwd_kernel(const DivParams p) {
  extern __shared__ __align__(16) float sm[];
  float* sw = sm;
  float* act = sm + kWTotal + threadIdx.x;
  load_weights(sw, p);
  __syncthreads();
  const long long tile = blockIdx.x;
  const int row = threadIdx.x;
  const long long pt = tile * kTileM + row;
  const bool valid = pt < p.P;
  const uint8_t* st = p.stash + tile * kStashTileBytes + row * 16;
  uint8_t* ad = p.adj + tile * kAdjTileBytes + row * 16;

  float scale = 1.0f;
  {
    const float amax = p.amax ? __ldg(p.amax) : 0.f;
    if (amax > 0.f && amax < 3.0e38f) {
      int ex;
      frexpf(amax, &ex);
      scale = ldexpf(1.0f, min(max(10 - ex, -60), 60));
    }
  }
  float e[3] = {0.f, 0.f, 0.f};
  float r = 0.f, G = 0.f, alpha = 0.f, beta = 0.f, tau_c = 0.f;
  if (valid) {
#pragma unroll
    for (int d = 0; d < 3; ++d) e[d] = p.e[pt * 3 + d];
    r = p.rigidity[pt]; G = p.G[pt]; alpha = p.adot[pt]; beta = p.beta[pt]; tau_c = p.tauc[pt];
    const float rp = 2.0f * r * (1.0f - r);
    const float tau_r = rp * tau_c;
#pragma unroll
    for (int d = 0; d < 3; ++d) p.d_unmasked[pt * 3 + d] = G * tau_r * e[d];
    p.d_rigid[pt] = G * (alpha + 2.0f * beta * tau_c * (1.0f - 2.0f * r));
  }
  const float Gs = G * scale;
  const float rp = 2.0f * r * (1.0f - r);
  float tb_off[4] = {Gs * r * e[0], Gs * r * e[1], Gs * r * e[2], 0.f};   // adjoint of tau_off
  const float tb_c = Gs * beta * rp;                                      // adjoint of tau_c
  const uint4 zz = make_uint4(0u, 0u, 0u, 0u);
  *reinterpret_cast<uint4*>(ad + kA4) = make_uint4(pk(tb_off[0], tb_off[1]), pk(tb_off[2], 0.f), 0u, 0u);
  *reinterpret_cast<uint4*>(ad + kA4 + kChunkBytes) = zz;

  const unsigned long long m1 = read_mask(st + kStHb1, 8), e1 = read_mask(st + kStHb1 + 8 * kChunkBytes, 4);
  const unsigned long long m2 = read_mask(st + kStHb2, 8), e2 = read_mask(st + kStHb2 + 8 * kChunkBytes, 4);
  const unsigned long long m3 = read_mask(st + kStHb3, 8), m4 = read_mask(st + kStHb4, 8);

  float a[64], q[32];
  // tbar4 = W4^T taubar_off ; abar4 = D4 tbar4
#pragma unroll 1
  for (int k = 0; k < 64; ++k)
    act[k * kDivThreads] = sw[kW4 + k] * tb_off[0] + sw[kW4 + 64 + k] * tb_off[1] + sw[kW4 + 128 + k] * tb_off[2];
  take<64>(act, m4, 1.0f, a, ad + kA3);
  // abar3 = D3 (W3^T abar4); rigidity output adjoint rides in column 64 of the same image
  matvec_t<64>(sw + kW3, a, act, 64);
  take<64>(act, m3, 1.0f, a, ad + kA2);
  *reinterpret_cast<uint4*>(ad + kA2 + 8 * kChunkBytes) = make_uint4(pk(tb_c, 0.f), 0u, 0u, 0u);
  *reinterpret_cast<uint4*>(ad + kA2 + 9 * kChunkBytes) = zz;
  // abar2 = D2 (W2^T abar3), qbar2 = E2 (R2^T taubar_c)
  matvec_t<64>(sw + kW2, a, act, 64);
  take<64>(act, m2, 1.0f, a, ad + kA1);
#pragma unroll 1
  for (int k = 0; k < 32; ++k) act[k * kDivThreads] = sw[kR2 + k] * tb_c;
  take<32>(act, e2, 1.0f, q, ad + kA1 + 8 * kChunkBytes);
  // abar1 = D1 (W1^T abar2), qbar1 = E1 (R1^T qbar2)
  matvec_t<64>(sw + kW1, a, act, 64);
  take<64>(act, m1, 1.0f, a, ad + kA0);
  matvec_t<32>(sw + kR1, q, act, 32);
  take<32>(act, e1, 1.0f, q, ad + kA0 + 8 * kChunkBytes);
}

// ------------------------------------------------------------------------------------------------
static size_t div_smem() { return sizeof(float) * (kWTotal + kActFloats); }

namespace {
__global__ void div_G_kernel(const DivParams p, const float* __restrict__ g_ray, float* __restrict__ G, float* __restrict__ amax) {
  const long long pt = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  float v = 0.f;
  if (pt < p.P) {
    float w = p.w[pt];
    if (p.w_is_alpha) w = 1.0f - expf(-fmaxf(w, 0.f));
    v = g_ray[pt / p.S] * (2.0f / static_cast<float>(p.S)) * w * p.d[pt];
    G[pt] = v;
  }
  float m = fabsf(v);
  if (!(m < 3.0e38f)) m = 0.f;   // ignore inf / nan: the scale must stay finite
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(amax), __float_as_int(m));   // non-negative floats order like ints
}
}  // namespace

cudaError_t launch_div_G(const DivParams& p, const float* g_ray, float* G, float* amax, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(amax, 0, sizeof(float), st);
  if (e != cudaSuccess || p.P <= 0) return e;
  div_G_kernel<<<static_cast<unsigned>((p.P + 255) / 256), 256, 0, st>>>(p, g_ray, G, amax);
  return cudaGetLastError();
}

cudaError_t launch_div_fwd(const DivParams& p, cudaStream_t st) {
  const long long tiles = (p.P + kTileM - 1) / kTileM;
  if (tiles <= 0) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(div_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)div_smem());
  if (e != cudaSuccess) return e;
  div_fwd_kernel<<<static_cast<unsigned>(tiles), kDivThreads, div_smem(), st>>>(p);
  return cudaGetLastError();
}
cudaError_t launch_div_bwd(const DivParams& p, cudaStream_t st) {
  const long long tiles = (p.P + kTileM - 1) / kTileM;
  if (tiles <= 0) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(div_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)div_smem());
  if (e != cudaSuccess) return e;
  div_bwd_kernel<<<static_cast<unsigned>(tiles), kDivThreads, div_smem(), st>>>(p);
  return cudaGetLastError();
}

}  // namespace nrn
