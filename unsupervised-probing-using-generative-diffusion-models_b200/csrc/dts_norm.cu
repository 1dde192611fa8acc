// DiffusionTS normalisation layers, forward and backward in one pass each:
//   AdaLayerNorm  y = LayerNorm(x) * (1 + scale[t]) + shift[t]   (diffusionts_model_utils.py:187-202; no affine in the LN)
//   nn.LayerNorm  y = LayerNorm(x) * w + b                       (ln2 of every block, diffusionts_transformer.py:215, 287)
// Both are  y = xhat * gamma + beta  with per-channel vectors (all rows of a launch share the diffusion step, so the
// AdaLN modulation is a [d] vector).  The backward feeds the Langevin refinement gradient (DiffusionTS.py:384-399):
//   dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma      (weights are constants: no dgamma / dbeta)
// One warp per row, d in {32,64,96,128,192,256,384,512,1024}; the library path was 3 launches forward and 4 backward per layer.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_helpers.cuh"

namespace {


__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NL>
__global__ void dts_ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                  long long rows, int D, float* __restrict__ y, float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  constexpr int n = NL;
  float v[NL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < n; ++i) { v[i] = x[r * D + lane + 32 * i]; s += v[i]; }
  const float mean = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < n; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
  const float rstd = rsqrtf(warp_sum(q) / D + 1e-5f);
#pragma unroll
  for (int i = 0; i < n; ++i) {
    const int c = lane + 32 * i;
    y[r * D + c] = fmaf((v[i] - mean) * rstd, gamma[c], beta[c]);
  }
  if (stats && lane == 0) { stats[2 * r] = mean; stats[2 * r + 1] = rstd; }
}

// The same forward, emitting the split operand [hi | lo | hi | 1 1 0..] (fp16, row pitch 3D + 8) of the dense layer that
// consumes y (fx_encoder.gemm3): the fp32 y is optional -- when only a linear layer reads it, it never reaches HBM.
// A lane owns pairs of neighbouring channels (c = 2 lane + 64 i) so that the fp16 halves are stored as 32-bit words.
template <int NP>
__global__ void dts_ln_fwd_a3_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                     long long rows, int D, float* __restrict__ y, float* __restrict__ stats,
                                     __half* __restrict__ a3) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  float2 v[NP];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NP; ++i) { v[i] = *reinterpret_cast<const float2*>(x + r * D + 2 * lane + 64 * i); s += v[i].x + v[i].y; }
  const float mean = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NP; ++i) { const float d0 = v[i].x - mean, d1 = v[i].y - mean; q = fmaf(d0, d0, q); q = fmaf(d1, d1, q); }
  const float rstd = rsqrtf(warp_sum(q) / D + 1e-5f);
  __half* row = a3 + r * (3LL * D + 8);
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int c = 2 * lane + 64 * i;
    const float2 g = *reinterpret_cast<const float2*>(gamma + c), b = *reinterpret_cast<const float2*>(beta + c);
    const float o0 = fmaf((v[i].x - mean) * rstd, g.x, b.x), o1 = fmaf((v[i].y - mean) * rstd, g.y, b.y);
    if (y) *reinterpret_cast<float2*>(y + r * D + c) = make_float2(o0, o1);
    uint32_t hi, lo;
    tc::split_f16x2(o0, o1, hi, lo);
    *reinterpret_cast<uint32_t*>(row + c) = hi;
    *reinterpret_cast<uint32_t*>(row + D + c) = lo;
    *reinterpret_cast<uint32_t*>(row + 2 * D + c) = hi;
  }
  if (lane == 0) *reinterpret_cast<uint4*>(row + 3 * D) = make_uint4(0x3C003C00u, 0u, 0u, 0u);   // bias columns: 1, 1, 0 x 6
  if (stats && lane == 0) { stats[2 * r] = mean; stats[2 * r + 1] = rstd; }
}

template <int NL>
__global__ void dts_ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma,
                                  const float* __restrict__ stats, long long rows, int D, float* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  constexpr int n = NL;
  const float mean = stats[2 * r], rstd = stats[2 * r + 1];
  float xh[NL], g[NL];
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int i = 0; i < n; ++i) {
    const int c = lane + 32 * i;
    xh[i] = (x[r * D + c] - mean) * rstd;
    g[i] = dy[r * D + c] * gamma[c];
    sg += g[i];
    sgx = fmaf(g[i], xh[i], sgx);
  }
  const float mg = warp_sum(sg) / D, mgx = warp_sum(sgx) / D;
#pragma unroll
  for (int i = 0; i < n; ++i) dx[r * D + lane + 32 * i] = rstd * (g[i] - mg - xh[i] * mgx);
}

}  // namespace

cudaError_t upd_launch_dts_layernorm(const float* x, const float* gamma, const float* beta, long long rows, int D, float* y,
                                     float* stats, void* a3, cudaStream_t stream) {
  const int wpb = 8;
  const unsigned grid = (unsigned)((rows + wpb - 1) / wpb);
  if (a3) {
    switch (D) {
#define UPD_LN(NP) case 64 * NP: dts_ln_fwd_a3_kernel<NP><<<grid, wpb * 32, 0, stream>>>(x, gamma, beta, rows, D, y, stats, (__half*)a3); break;
      UPD_LN(1) UPD_LN(2) UPD_LN(3) UPD_LN(4) UPD_LN(6) UPD_LN(8) UPD_LN(16)
#undef UPD_LN
      default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
  }
  if (!y) return cudaErrorInvalidValue;
  switch (D) {
#define UPD_LN(NL) case 32 * NL: dts_ln_fwd_kernel<NL><<<grid, wpb * 32, 0, stream>>>(x, gamma, beta, rows, D, y, stats); break;
    UPD_LN(1) UPD_LN(2) UPD_LN(3) UPD_LN(4) UPD_LN(6) UPD_LN(8) UPD_LN(12) UPD_LN(16) UPD_LN(32)
#undef UPD_LN
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t upd_launch_dts_layernorm_bwd(const float* x, const float* dy, const float* gamma, const float* stats,
                                         long long rows, int D, float* dx, cudaStream_t stream) {
  const int wpb = 8;
  const unsigned grid = (unsigned)((rows + wpb - 1) / wpb);
  switch (D) {
#define UPD_LN(NL) case 32 * NL: dts_ln_bwd_kernel<NL><<<grid, wpb * 32, 0, stream>>>(x, dy, gamma, stats, rows, D, dx); break;
    UPD_LN(1) UPD_LN(2) UPD_LN(3) UPD_LN(4) UPD_LN(6) UPD_LN(8) UPD_LN(12) UPD_LN(16) UPD_LN(32)
#undef UPD_LN
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
